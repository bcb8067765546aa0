"""Per-layer convolution times of one 1080p P-frame (CUDA events around every conv launch)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
os.environ["FVC_PROFILE"] = "1"
import torch
from fastvideocodec_b200 import VideoCompressor
from fastvideocodec_b200._lib import lib
from fastvideocodec_b200.synthetic import init_state_dict, synthetic_gop
H, W = int(os.environ.get("H", 1088)), int(os.environ.get("W", 1920))
dev = torch.device("cuda")
m = VideoCompressor(); m.load_state_dict(init_state_dict(0)); m = m.to(dev).eval()
fr = synthetic_gop(H, W, gop=2, gop_id=0)[:, 0].to(dev)
FL = {}  # GFLOP per layer at this size
def gf(cin, cout, k, h, w): return 2.0 * cin * cout * k * k * h * w / 1e9
with torch.no_grad():
    for _ in range(3):
        m(fr[1:2], fr[0:1])
txt = lib().fvc_ctx_profile_text(m._last_ctx.handle).decode()
rows = [(l.split()[0], float(l.split()[1])) for l in txt.strip().splitlines()]
tot = sum(t for _, t in rows)
groups = {}
for n, t in rows:
    g = n.split(".")[0] if not n.startswith("opticFlow") else "opticFlow.L" + n.split(".")[2]
    groups[g] = groups.get(g, 0) + t
print("total conv ms %.3f" % tot)
for g, t in groups.items(): print("%-22s %8.3f ms  %5.1f%%" % (g, t, 100 * t / tot))
print("--- top layers")
for n, t in sorted(rows, key=lambda r: -r[1])[:24]: print("%-40s %8.3f ms" % (n, t))
