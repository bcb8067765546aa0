import os, sys
sys.path.insert(0, "/root/repo"); os.environ["FVC_PROFILE"] = "1"
import torch
from fastvideocodec_b200 import VideoCompressor
from fastvideocodec_b200._lib import lib
from fastvideocodec_b200.synthetic import init_state_dict, synthetic_gop
dev = torch.device("cuda")
m = VideoCompressor(); m.load_state_dict(init_state_dict(0)); m = m.to(dev).eval()
fr = synthetic_gop(1088, 1920, gop=2, gop_id=0)[:, 0].to(dev)
with torch.no_grad():
    for _ in range(3): m(fr[1:2], fr[0:1])
txt = lib().fvc_ctx_profile_text(m._last_ctx.handle).decode()
for l in txt.strip().splitlines(): print(l)
