"""Does a CUDA graph of the whole GOP (9 P-frames, ~820 kernel nodes) run faster than the eager launches?
Same box, alternating: eager loop vs graph replay."""
import os, sys, json, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from fastvideocodec_b200 import VideoCompressor
from fastvideocodec_b200._lib import check, lib, ptr
import ctypes as C
from fastvideocodec_b200.synthetic import init_state_dict, synthetic_gop
H, W, GOP = 1088, 1920, 10
dev = torch.device("cuda")
m = VideoCompressor(); m.load_state_dict(init_state_dict(0)); m = m.to(dev).eval()
fr = synthetic_gop(H, W, gop=GOP, gop_id=0).to(dev)
ctx = m._context(1, H, W, dev); rec = torch.empty((2, 1, 3, H, W), device=dev); sc = torch.empty((GOP - 1, 7), device=dev)
def gop():
    s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    prev = fr[0]
    for i in range(1, GOP):
        out = rec[i & 1]
        check(lib().fvc_pframe_forward(ctx.handle, ptr(fr[i]), ptr(prev), ptr(out), ptr(sc[i - 1]), s), "fwd")
        prev = out
for _ in range(3): gop()
torch.cuda.synchronize()
ref = sc.clone()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    gop()
g.replay(); torch.cuda.synchronize()
assert torch.equal(sc, ref), "graph replay differs"
def timeit(fn, n=6):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return n * (GOP - 1) / (e0.elapsed_time(e1) * 1e-3)
for r in range(3):
    print(json.dumps({"eager_fps": timeit(gop), "graph_fps": timeit(g.replay)}))
