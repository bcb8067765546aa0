"""Runs FRAMES (default 1) P-frame forwards at H x W (default 1088x1920) through the C ABI - the command
line profiled by ncu for profiles/ (launch list and the --set full capture of the dominant kernel)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from fastvideocodec_b200 import VideoCompressor
from fastvideocodec_b200.synthetic import init_state_dict, synthetic_gop
H, W, n = int(os.environ.get("H", 1088)), int(os.environ.get("W", 1920)), int(os.environ.get("FRAMES", 1))
dev = torch.device("cuda")
m = VideoCompressor(); m.load_state_dict(init_state_dict(0)); m = m.to(dev).eval()
fr = synthetic_gop(H, W, gop=2, gop_id=0)[:, 0].to(dev)
with torch.no_grad():
    for _ in range(n):
        out = m(fr[1:2], fr[0:1])
torch.cuda.synchronize()
print("bpp %.5f launches %d" % (float(out[7]), m.launch_count()))
