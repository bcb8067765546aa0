"""Throughput of the device intra codec (fvc_iframe_forward) at 1088x1920: estimated-bits mode and with real entropy
coding.   python tools/iframe_bench.py [steps=30]"""
import json, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from fastvideocodec_b200 import VideoCompressor
from fastvideocodec_b200.synthetic import init_state_dict, synthetic_gop

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
dev = torch.device("cuda")
m = VideoCompressor(); m.load_state_dict(init_state_dict(0)); m = m.to(dev).eval()
x = synthetic_gop(1088, 1920, gop=2, gop_id=0)[1:2, 0].to(dev)
for real in (False, True):
    m.calrealbits = real
    with torch.no_grad():
        for _ in range(3): out = m.iframe_forward(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps): out = m.iframe_forward(x)
        e1.record(); torch.cuda.synchronize()
    print(json.dumps({"probe": "iframe_forward", "size": "1088x1920", "calrealbits": real,
                      "frames_per_s": round(steps / (e0.elapsed_time(e1) * 1e-3), 1), "ms": round(e0.elapsed_time(e1) / steps, 3),
                      "bpp": round(float(out[4]), 5), "psnr_db": round(float(10 * torch.log10(1 / out[1])), 3)}), flush=True)
