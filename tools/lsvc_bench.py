"""LSVC throughput at 1088x1920: one tree GOP forward (I-frame + P P-frames) per step through
fastvideocodec_b200.lsvc.LSVC, CUDA-event timed.  Usage: python tools/lsvc_bench.py [P=6] [steps=5]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from fastvideocodec_b200.lsvc import LSVC
from fastvideocodec_b200.synthetic import init_state_dict, synthetic_gop
P = int(sys.argv[1]) if len(sys.argv) > 1 else 6
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
dev = torch.device("cuda")
out = []
for name in ("LSVC-128", "LSVC-L-128"):
    m = LSVC(name); m.load_state_dict(init_state_dict(0)); m = m.to(dev).eval()
    x = synthetic_gop(1088, 1920, gop=P + 1, gop_id=0)[:, 0].to(dev)
    with torch.no_grad():
        for _ in range(2): r = m(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps): r = m(x)
        e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    out.append({"model": name, "p_frames_per_gop": P, "ms_per_gop": ms, "p_frames_per_s": P * 1000.0 / ms,
                "bpp": float(r[7]), "rec_mse": float(r[3])})
    m.release(); del m; torch.cuda.empty_cache()
for o in out: print(json.dumps(o))
