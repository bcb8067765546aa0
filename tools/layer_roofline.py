"""Per-layer table for one 1088x1920 P-frame: shape, algorithmic GFLOP, measured CUDA-event time (a file written by
tools/layer_ab.py / layer_times.py on a B200), algorithmic TFLOP/s and the share of the measured sustained bf16
peak counting the three MMAs issued per product.  Pure host script (no GPU):
    python tools/layer_roofline.py profiles/r01_layer_times_final.txt > profiles/r01_layer_roofline.txt"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from fastvideocodec_b200.synthetic import init_state_dict

H, W = 1088, 1920
PEAK = 1383.4
try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("bf16_tflops_sustained", PEAK)
except Exception:
    pass
sd = init_state_dict(0)
COL = int(sys.argv[2]) if len(sys.argv) > 2 else 1   # which timing column of the input file


def out_res(name):
    """(Hout, Wout, stride, transposed) of a convolution of the DVC P-frame path at 1088x1920."""
    p = name.split(".")
    if p[0] == "opticFlow":
        s = 2 ** (3 - int(p[2]))
        return H // s, W // s, 1, 0
    if p[0] == "mvEncoder":
        k = int(p[1][4:]); s = 2 ** ((k + 1) // 2)
        return H // s, W // s, 2 if k % 2 else 1, 0
    if p[0] == "mvDecoder":
        k = int(p[1][6:]); s = 2 ** (4 - (k + 1) // 2)
        return H // s, W // s, 2 if k % 2 else 1, k % 2      # deconv2/4/6/8 are stride-1 nn.Conv2d in the reference
    if p[0] == "warpnet":
        s = {"feature_ext": 1, "conv0": 1, "conv1": 2, "conv2": 4, "conv3": 4, "conv4": 2, "conv5": 1, "conv6": 1}[p[1]]
        return H // s, W // s, 1, 0
    if p[0] == "resEncoder":
        k = int(p[1].replace("#norm", "")[-1]); s = 2 ** k
        return H // s, W // s, (1 if "#norm" in name else 2), 0
    if p[0] == "resDecoder":
        k = int(p[1].replace("#norm", "")[-1]); s = 2 ** (4 - k)
        return H // s, W // s, (1 if "#norm" in name else 2), (0 if "#norm" in name else 1)
    if p[0] == "respriorEncoder":
        k = int(p[1][-1]); s = 16 * 2 ** (k - 1)
        return H // s, W // s, 1 if k == 1 else 2, 0
    if p[0] == "respriorDecoder":
        k = int(p[1][-1]); s = 16 * 2 ** max(0, 2 - k)
        return H // s, W // s, 2 if k < 3 else 1, 1
    raise KeyError(name)


rows = []
for line in open(sys.argv[1]):
    f = line.split()
    if len(f) < 2 or f[0] in ("layer", "TOTAL"):
        continue
    if f[0].startswith("@") or "#taps" in f[0]:   # frame kernels / tap sums: no MMA work
        continue
    try:
        name, ms = f[0], float(f[COL])
    except (ValueError, IndexError):
        continue
    if ms != ms:                              # "nan": the layer does not exist in this variant (fused away)
        continue
    ho, wo, st, tr = out_res(name)
    if "#norm" in name:                       # GDN norm = 1x1 convolution C -> C of the squared activations
        C = sd[name.replace("#norm", "") + ".beta"].numel(); cin = cout = C; k = 1
    else:
        w = sd[name + ".weight"]
        cin, cout = (w.shape[0], w.shape[1]) if tr else (w.shape[1], w.shape[0]); k = w.shape[2]
    gflop = 2.0 * cin * cout * k * k * ho * wo / (st * st if tr else 1) / 1e9
    rows.append((name, "%dx%d %s%dx%d s%d -> %dx%d" % (cin, cout, "T" if tr else "", k, k, st, ho, wo), gflop, ms))
tot_g = sum(r[2] for r in rows); tot_ms = sum(r[3] for r in rows)
print("%-34s %-34s %9s %8s %9s %7s" % ("layer", "shape", "GFLOP", "ms", "TFLOP/s", "3x/peak"))
for n, sh, g, ms in rows:
    print("%-34s %-34s %9.2f %8.4f %9.1f %6.1f%%" % (n, sh, g, ms, g / ms, 300.0 * g / ms / PEAK))
print("%-34s %-34s %9.2f %8.4f %9.1f %6.1f%%" % ("TOTAL", "", tot_g, tot_ms, tot_g / tot_ms, 300.0 * tot_g / tot_ms / PEAK))
