"""Diagnostic (GPU): per-tensor errors of the L=6 pyramid golden for both engines."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from fastvideocodec_b200 import VideoCompressor
from fastvideocodec_b200.synthetic import init_state_dict
dev = torch.device("cuda")
with np.load(os.path.join(ROOT, "tests", "golden", "pframe_L6_256.npz")) as z:
    g = {k: torch.from_numpy(z[k]) for k in z.files}
for L in (6,):
    m = VideoCompressor(spynet_levels=L); m.load_state_dict(init_state_dict(0, spynet_levels=L)); m = m.to(dev).eval()
    for impl in (0, 1):
        m.impl = impl
        with torch.no_grad():
            out = m(g["cur"].to(dev), g["ref"].to(dev))
        r = {"impl": impl}
        for n in ("estmv", "mvfeature", "mv_hat", "feature", "z", "sigma"):
            a = m.get_intermediate(n).cpu()
            r[n] = (float((a - g[n]).abs().max()), float((a - g[n]).abs().mean()))
        for n in ("quant_mv", "z_hat", "feat_hat"):
            a = m.get_intermediate(n).cpu()
            r[n] = int((a != g[n]).sum())
        d = (m.get_intermediate("quant_mv").cpu() != g["quant_mv"])
        pre = g["mvfeature"][d]
        r["mv_flip_frac_dist"] = [round(float(abs((v - v.floor()) - 0.5)), 6) for v in pre]
        e = (m.get_intermediate("estmv").cpu() - g["estmv"]).abs()
        r["estmv_err_rows"] = [round(float(e[0, :, i * 32:(i + 1) * 32].max()), 6) for i in range(8)]
        print(r)
