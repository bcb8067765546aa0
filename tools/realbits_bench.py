"""P-frame throughput at 1088x1920 with calrealbits (real entropy coding of the three latents on the GPU, net.py:123-195)
against the estimated-bits default, plus the per-kernel CUDA-event times of the coder.
    python tools/realbits_bench.py [steps=4]"""
import json, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
os.environ.setdefault("FVC_PROFILE", "1")
from fastvideocodec_b200 import VideoCompressor
from fastvideocodec_b200._lib import lib
from fastvideocodec_b200.synthetic import init_state_dict, synthetic_gop

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
H, W, GOP = 1088, 1920, 10
dev = torch.device("cuda")
m = VideoCompressor(); m.load_state_dict(init_state_dict(0)); m = m.to(dev).eval()
fr = synthetic_gop(H, W, gop=GOP, gop_id=0)[:, 0].to(dev)
for real in (False, True):
    m.calrealbits = real
    def gop():
        prev = fr[0:1]
        for i in range(1, GOP):
            out = m(fr[i:i + 1], prev)
            prev = out[0]
        return out
    with torch.no_grad():
        gop(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps): out = gop()
        e1.record(); torch.cuda.synchronize()
    txt = lib().fvc_ctx_profile_text(m._last_ctx.handle).decode()
    coder = {l.split()[0]: float(l.split()[1]) for l in txt.splitlines() if any(t in l for t in ("rans_encode", "entropy_model"))}
    print(json.dumps({"probe": "pframe_realbits", "calrealbits": real, "fps": round(steps * (GOP - 1) / (e0.elapsed_time(e1) * 1e-3), 2),
                      "bpp": round(float(out[7]), 5), "coder_ms": coder}), flush=True)
