"""Generates tests/golden/lsvc_64.npz by running the UNMODIFIED reference ``models.LSVC`` (CPU fp32) on seeded
inputs (import-time shims only, oracle/ref_shim.py).  Run here (needs /root/reference):

    python -m oracle.gen_golden_lsvc
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from fastvideocodec_b200.synthetic import init_state_dict, synthetic_gop  # noqa: E402
from oracle import ref_shim  # noqa: E402


def main():
    refmodels = ref_shim.load_reference_models()
    sd = init_state_dict(0)
    out = {}
    x = synthetic_gop(64, 64, gop=5, gop_id=4)[:, 0]          # I-frame + 4 P-frames
    out["x"] = x.numpy()
    for tag, name in (("tree", "LSVC-128"), ("chain", "LSVC-L-128"), ("onehop", "LSVC-O-128")):
        with ref_shim._cwd(ref_shim.REF_ROOT):
            m = refmodels.LSVC(name, use_split=False)
        m.load_state_dict(sd, strict=True)
        m.eval()
        with torch.no_grad():
            res = m(x.clone())
        names = ["com", "mc", "warped", "rec_loss", "warp_loss", "mc_loss", "bpp_res", "bpp"]
        for n, v in zip(names, res):
            out["%s_%s" % (tag, n)] = v.detach().numpy()
    path = os.path.join(ROOT, "tests", "golden", "lsvc_64.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
