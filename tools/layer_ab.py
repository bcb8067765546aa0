"""A/B of tile-shape knobs: per-layer conv times of one 1080p P-frame under several env settings (one process each)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
variants = [v for v in sys.argv[1:]] or [""]
res = {}
for v in variants:
    env = dict(os.environ)
    for kv in v.split(","):
        if kv: env[kv.split("=")[0]] = kv.split("=")[1]
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "layer_times.py")], env=env, capture_output=True, text=True).stdout
    res[v] = {l.split()[0]: float(l.split()[1]) for l in out.strip().splitlines() if len(l.split()) == 2}
names = list(res[variants[0]].keys())
print("%-36s" % "layer" + "".join("%14s" % (v[-13:] or "default") for v in variants))
for n in names:
    print("%-36s" % n + "".join("%14.4f" % res[v].get(n, float("nan")) for v in variants))
print("%-36s" % "TOTAL" + "".join("%14.4f" % sum(res[v].values()) for v in variants))
