"""Measures the systematic bias of tcgen05 fp32 accumulation: all-positive inputs/weights so every
partial sum is positive; compare TC and SIMT against float64."""
import os, sys, math, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch, torch.nn.functional as F
from fastvideocodec_b200 import ops
dev = torch.device("cuda")
g = torch.Generator().manual_seed(0)
for (cin, cout, k, H, W) in [(64, 32, 7, 48, 64), (32, 64, 7, 48, 64), (64, 64, 3, 48, 64), (128, 128, 3, 48, 64), (64, 64, 1, 48, 64)]:
    for mode in ("pos", "neg", "mixed"):
        x = torch.rand((1, cin, H, W), generator=g) + 0.5
        w = torch.rand((cout, cin, k, k), generator=g) * 0.1 + 0.01
        if mode == "mixed":
            w = w * torch.sign(torch.randn(w.shape, generator=g))
        if mode == "neg":      # all partial sums negative: tells truncation toward zero from truncation toward -inf
            w = -w
        b = torch.zeros(cout)
        want = F.conv2d(x.double(), w.double(), None, padding=k // 2)
        sl = (slice(None), slice(None), slice(k, H - k), slice(k, W - k))
        res = {}
        for name, impl in (("simt", ops.IMPL_SIMT), ("tc", ops.IMPL_TC)):
            got = ops.conv2d(x.to(dev), w.to(dev), b.to(dev), 1, 0, impl).cpu().double()
            rel = ((got - want) / want.abs().clamp(min=1e-3))[sl]
            if mode == "mixed":    # signed towards-zero measure: (|got| - |want|) / |want|
                rel = ((got.abs() - want.abs()) / want.abs().clamp(min=1e-3))[sl]
            res[name] = {"mean_rel": float(rel.mean()), "rms_rel": float(rel.pow(2).mean().sqrt())}
        nmma = k * k * (cin // 16) * 3
        print(json.dumps({"cin": cin, "cout": cout, "k": k, "mode": mode, "mma_chain": nmma, **{f"{n}_{m}": v for n, d in res.items() for m, v in d.items()}}))
