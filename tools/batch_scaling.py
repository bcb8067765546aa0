import os, sys, time
sys.path.insert(0, "/root/repo")
import torch
from fastvideocodec_b200 import VideoCompressor
from fastvideocodec_b200._lib import check, lib, ptr, stream_ptr
from fastvideocodec_b200.synthetic import init_state_dict, synthetic_gop
dev = torch.device("cuda")
m = VideoCompressor(); m.load_state_dict(init_state_dict(0)); m = m.to(dev).eval()
for B in (1, 2, 4):
    fr = synthetic_gop(1088, 1920, gop=10, gop_id=0, batch=B).to(dev)   # [G,B,3,H,W]
    ctx = m._context(B, 1088, 1920, dev)
    rec = torch.empty((2, B, 3, 1088, 1920), device=dev); scal = torch.empty((9, 7), device=dev)
    def gop():
        prev = fr[0]
        for i in range(1, 10):
            out = rec[i & 1]
            check(lib().fvc_pframe_forward(ctx.handle, ptr(fr[i]), ptr(prev), ptr(out), ptr(scal[i-1]), stream_ptr()), "f")
            prev = out
    for _ in range(2): gop()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): gop()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print("B=%d: %.2f P-frames/s (%.2f ms per frame-batch)" % (B, 3 * 9 * B / (ms * 1e-3), ms / 27))
    m.release()
