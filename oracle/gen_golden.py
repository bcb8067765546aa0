"""Generate golden vectors by running the UNMODIFIED reference.  TEST INFRASTRUCTURE ONLY.

Run in the build container (needs /root/reference):

    python -m oracle.gen_golden

Writes tests/golden/pframe_64.npz   : one P-frame, 64x64, all intermediates + outputs
       tests/golden/pframe_128.npz  : one P-frame, 128x128 (non-trivial 4-level pyramid)
       tests/golden/gop_64.npz      : closed-loop 4-frame GOP at 64x64 (3 P-frames)
       tests/golden/ops.npz         : op-level known answers (warp, up-samplers, GDN, bit estimators)
Weights come from ``fastvideocodec_b200.synthetic.init_state_dict(seed=0)`` loaded (strict) into
the reference model; frames from ``synthetic_gop``.  The fixtures hold inputs and outputs, so
the GPU box (no reference there) can check against them.
"""
from __future__ import annotations

import math
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from fastvideocodec_b200.synthetic import init_state_dict, synthetic_gop  # noqa: E402
from oracle import ref_shim  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def _np(d):
    return {k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in d.items()}


def gen_pframe(model, size, gop_id):
    frames = synthetic_gop(size, size, gop=2, gop_id=gop_id)[:, 0]
    ref, cur = frames[0:1], frames[1:2]
    out, cap = ref_shim.run_reference_with_capture(model, cur, ref)
    d = dict(cur=cur, ref=ref, clipped=out[0])
    names = ["mse", "warploss", "interloss", "bpp_feature", "bpp_z", "bpp_mv", "bpp"]
    for n, v in zip(names, out[1:]):
        d[n] = v
    d.update(cap)
    return _np(d)


def gen_gop(model, size, n, gop_id):
    frames = synthetic_gop(size, size, gop=n, gop_id=gop_id)[:, 0]
    x_prev = frames[0:1]
    rows, recs = [], []
    with torch.no_grad():
        for i in range(1, n):
            out = model(frames[i:i + 1], x_prev)
            x_prev = out[0].detach()
            psnr = 10.0 * torch.log(1 / out[1]) / math.log(10.0)
            rows.append([float(v) for v in out[1:]] + [float(psnr)])
            recs.append(x_prev)
    return _np(dict(frames=frames, recon=torch.cat(recs, 0), rows=np.asarray(rows, dtype=np.float64)))


def gen_ops(model, refnet):
    import torch.nn.functional as F
    from DVC.subnet import endecoder as endec
    from DVC.subnet.GDN import GDN
    g = torch.Generator().manual_seed(77)
    d = {}
    img = torch.rand((2, 3, 12, 20), generator=g)
    flow = (torch.rand((2, 2, 12, 20), generator=g) - 0.5) * 9.0
    d["warp_img"], d["warp_flow"] = img, flow
    d["warp_out"] = endec.flow_warp(img, flow)
    d["warp_zero_flow_out"] = endec.flow_warp(img, torch.zeros_like(flow))
    x = torch.randn((1, 5, 6, 10), generator=g)
    d["up_in"] = x
    d["up_half_pixel"] = endec.bilinearupsacling(x)
    d["up_align_corners"] = endec.bilinearupsacling2(x)
    d["pool_out"] = F.avg_pool2d(x, kernel_size=2, stride=2)
    gd = GDN(8)
    with torch.no_grad():
        gd.beta.copy_(torch.rand(8, generator=g) + 0.5)
        gd.gamma.copy_(torch.rand(8, 8, generator=g) * 0.3)
        gd.gamma[0, 1] = 1e-7   # below gamma_bound: exercises LowerBound
        gd.beta[2] = 1e-4       # below beta_bound
    xg = torch.randn((1, 8, 4, 6), generator=g) * 2
    d["gdn_beta"], d["gdn_gamma"], d["gdn_in"] = gd.beta.detach(), gd.gamma.detach(), xg
    with torch.no_grad():
        d["gdn_out"] = gd(xg)
        gd.inverse = True
        d["igdn_out"] = gd(xg)
    # bit estimators on a grid of integers incl. large magnitudes (softplus/sigmoid tails)
    q = torch.arange(-40, 41, dtype=torch.float32).view(1, 1, 1, -1).repeat(1, 64, 1, 1)
    with torch.no_grad():
        d["be_q"] = q
        d["be_cdf_hi"] = model.bitEstimator_z(q + 0.5)
        d["be_cdf_lo"] = model.bitEstimator_z(q - 0.5)
    sig = torch.exp(torch.randn((1, 4, 3, 81), generator=g) * 2.0)
    ql = torch.arange(-40, 41, dtype=torch.float32).view(1, 1, 1, -1).repeat(1, 4, 3, 1)
    lap = torch.distributions.laplace.Laplace(torch.zeros_like(sig), sig.clamp(1e-5, 1e10))
    d["lap_q"], d["lap_sigma"] = ql, sig
    d["lap_prob"] = lap.cdf(ql + 0.5) - lap.cdf(ql - 0.5)
    return _np(d)


def main():
    os.makedirs(GOLD, exist_ok=True)
    torch.manual_seed(0)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    sd = init_state_dict(seed=0)
    model = ref_shim.build_reference_model(sd)
    refnet = ref_shim.load_reference()
    np.savez_compressed(os.path.join(GOLD, "pframe_64.npz"), **gen_pframe(model, 64, 0))
    np.savez_compressed(os.path.join(GOLD, "pframe_128.npz"), **gen_pframe(model, 128, 1))
    np.savez_compressed(os.path.join(GOLD, "gop_64.npz"), **gen_gop(model, 64, 4, 2))
    np.savez_compressed(os.path.join(GOLD, "ops.npz"), **gen_ops(model, refnet))
    for f in sorted(os.listdir(GOLD)):
        print(f, os.path.getsize(os.path.join(GOLD, f)))


if __name__ == "__main__":
    main()
