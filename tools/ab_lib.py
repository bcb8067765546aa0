"""Same-box A/B of two builds of libfvc_b200.so: per-layer CUDA-event times of one 1080p P-frame, alternating the
libraries (FVC_LIB_PATH) so that box-to-box clock differences cancel.
    python tools/ab_lib.py tools/_bin/libfvc_r01.so fastvideocodec_b200/libfvc_b200.so [rounds]"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
libs = [os.path.abspath(p) for p in sys.argv[1:3]]
rounds = int(sys.argv[3]) if len(sys.argv) > 3 else 2
res = {l: [] for l in libs}
for r in range(rounds):
    for l in libs:
        env = dict(os.environ, FVC_LIB_PATH=l)
        out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "layer_times.py")], env=env, capture_output=True, text=True)
        d = {}
        for ln in out.stdout.strip().splitlines():
            f = ln.split()
            if len(f) == 2:
                d[f[0]] = d.get(f[0], 0.0) + float(f[1])
        if not d:
            print(out.stderr[-2000:])
        res[l].append(d)
names = list(res[libs[1]][0].keys())
print("%-36s %12s %12s %8s" % ("layer", os.path.basename(libs[0])[-12:], os.path.basename(libs[1])[-12:], "ratio"))
ta = tb = 0.0
for n in names:
    a = [d[n] for d in res[libs[0]] if n in d]
    b = [d[n] for d in res[libs[1]] if n in d]
    if not a or not b:
        continue
    a, b = min(a), min(b)
    if not n.startswith("@"):
        ta += a; tb += b
    print("%-36s %12.4f %12.4f %8.3f" % (n, a, b, b / a))
print("%-36s %12.4f %12.4f %8.3f" % ("TOTAL conv", ta, tb, tb / ta))
