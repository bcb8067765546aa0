"""Same-box A/B of whole-frame throughput (the number bench.py reports as `value`) under env variants, alternating
subprocesses: python tools/fps_ab.py "" FVC_GDN_FUSED=0 ... [--rounds 2] [--steps 6]"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, json, torch
sys.path.insert(0, %r)
from fastvideocodec_b200 import VideoCompressor
from fastvideocodec_b200._lib import check, lib, ptr, stream_ptr
from fastvideocodec_b200.synthetic import init_state_dict, synthetic_gop
steps = int(sys.argv[1]); H, W, GOP = 1088, 1920, 10
dev = torch.device("cuda")
m = VideoCompressor(precision=os.environ.get("FVC_PRECISION", "exact")); m.load_state_dict(init_state_dict(0)); m = m.to(dev).eval()
fr = synthetic_gop(H, W, gop=GOP, gop_id=0).to(dev)
ctx = m._context(1, H, W, dev); rec = torch.empty((2, 1, 3, H, W), device=dev); sc = torch.empty((GOP - 1, 7), device=dev)
def gop():
    prev = fr[0]
    for i in range(1, GOP):
        out = rec[i & 1]
        check(lib().fvc_pframe_forward(ctx.handle, ptr(fr[i]), ptr(prev), ptr(out), ptr(sc[i - 1]), stream_ptr()), "fwd")
        prev = out
for _ in range(3): gop()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps): gop()
e1.record(); torch.cuda.synchronize()
print(json.dumps({"fps": steps * (GOP - 1) / (e0.elapsed_time(e1) * 1e-3), "bpp": float(sc[:, 6].mean())}))
''' % ROOT
args = sys.argv[1:]
rounds, steps = 2, 6
if "--rounds" in args: i = args.index("--rounds"); rounds = int(args[i + 1]); del args[i:i + 2]
if "--steps" in args: i = args.index("--steps"); steps = int(args[i + 1]); del args[i:i + 2]
variants = args or [""]
res = {v: [] for v in variants}
for r in range(rounds):
    for v in variants:
        env = dict(os.environ)
        for kv in v.split(","):
            if kv: env[kv.split("=")[0]] = kv.split("=")[1]
        out = subprocess.run([sys.executable, "-c", CHILD, str(steps)], env=env, capture_output=True, text=True)
        try:
            res[v].append(json.loads(out.stdout.strip().splitlines()[-1])["fps"])
        except Exception:
            print(out.stderr[-1500:])
for v in variants:
    print("%-40s %s  best %.2f" % (v or "default", " ".join("%.2f" % f for f in res[v]), max(res[v]) if res[v] else float("nan")))
