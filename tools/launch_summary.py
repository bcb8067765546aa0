"""Summarises an `ncu --csv` launch list (gpu__time_duration.sum [+ dram__bytes_*]) per kernel name and writes
profiles/conv_traffic.json when DRAM byte metrics are present.  Usage: python tools/launch_summary.py list.csv [frames]"""
import collections, csv, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
path = sys.argv[1]; frames = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
rows = list(csv.reader(open(path)))
start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr = rows[start]; ki, mi, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.defaultdict(lambda: collections.defaultdict(float)); cnt = collections.Counter()
scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
for r in rows[start + 1:]:
    if len(r) <= vi: continue
    n = r[ki].split("(")[0].replace("void ", "").replace("fvc::", "")
    agg[n][r[mi]] += float(r[vi].replace(",", "")) * scale.get(r[ui], 1.0)
    if r[mi] == "gpu__time_duration.sum": cnt[n] += 1
tot = sum(v["gpu__time_duration.sum"] for v in agg.values())
print("%d launches, %.1f us total (cold-cache, serialised: compare shares)" % (sum(cnt.values()), tot))
conv_t = conv_b = 0.0
for n, v in sorted(agg.items(), key=lambda kv: -kv[1]["gpu__time_duration.sum"]):
    b = v.get("dram__bytes_read.sum", 0) + v.get("dram__bytes_write.sum", 0)
    print("%-44s n=%4d %10.1f us %5.1f%%  dram %8.1f MB" % (n[:44], cnt[n], v["gpu__time_duration.sum"], 100 * v["gpu__time_duration.sum"] / tot, b / 1e6))
    if n.startswith("k_conv_tc"): conv_t += v["gpu__time_duration.sum"]; conv_b += b
print("k_conv_tc share of listed time: %.1f%%" % (100 * conv_t / tot))
if conv_b > 0:
    out = {"dram_bytes_per_frame": conv_b / frames, "frames": frames, "source": os.path.basename(path),
           "kernel": "k_conv_tc (all instantiations)", "conv_us_per_frame_under_ncu": conv_t / frames}
    json.dump(out, open(os.path.join(ROOT, "profiles", "conv_traffic.json"), "w"), indent=1)
    print(out)
