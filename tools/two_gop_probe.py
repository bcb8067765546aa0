"""Probe: K independent GOPs in flight on one GPU (one context + one CUDA stream each, frames issued round-robin from one
host thread) against one GOP at a time.  The P-frame chain of a GOP is sequential and has phases that cannot fill 148
SMs (SpyNet's coarse levels, the 17x30 hyper-prior layers, the tails of persistent kernels); a second GOP's kernels can
run there.   python tools/two_gop_probe.py [K=2] [steps=6]"""
import json, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from fastvideocodec_b200 import VideoCompressor
from fastvideocodec_b200._lib import check, lib, ptr
from fastvideocodec_b200.synthetic import init_state_dict, synthetic_gop
import ctypes as C

K = int(sys.argv[1]) if len(sys.argv) > 1 else 2
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
H, W, GOP = 1088, 1920, 10
dev = torch.device("cuda")
sd = init_state_dict(0)


def run(k):
    ms, ctxs, frs, recs, scs, streams = [], [], [], [], [], []
    for j in range(k):
        m = VideoCompressor(precision=os.environ.get("FVC_PRECISION", "exact")); m.load_state_dict(sd); m = m.to(dev).eval()
        ms.append(m); ctxs.append(m._context(1, H, W, dev))
        frs.append(synthetic_gop(H, W, gop=GOP, gop_id=j).to(dev))
        recs.append(torch.empty((2, 1, 3, H, W), device=dev)); scs.append(torch.empty((GOP - 1, 7), device=dev))
        streams.append(torch.cuda.Stream())
    torch.cuda.synchronize()

    def gops():
        for i in range(1, GOP):
            for j in range(k):
                prev = frs[j][0] if i == 1 else recs[j][(i - 1) & 1]
                check(lib().fvc_pframe_forward(ctxs[j].handle, ptr(frs[j][i]), ptr(prev), ptr(recs[j][i & 1]), ptr(scs[j][i - 1]),
                                               C.c_void_p(streams[j].cuda_stream)), "fwd")
    for _ in range(3): gops()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in streams: s.wait_event(e0)
    for _ in range(steps): gops()
    for s in streams: torch.cuda.current_stream().wait_stream(s)
    e1.record(); torch.cuda.synchronize()
    fps = k * steps * (GOP - 1) / (e0.elapsed_time(e1) * 1e-3)
    bpp = [float(sc[:, 6].mean()) for sc in scs]
    for m in ms: m.release()
    return fps, bpp


for k in (1, K, 1, K):
    fps, bpp = run(k)
    print(json.dumps({"probe": "gops_in_flight", "k": k, "fps": round(fps, 2), "bpp": [round(b, 6) for b in bpp]}), flush=True)
