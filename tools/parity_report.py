"""Parity report (GPU): per-tensor error/mismatch statistics of the CUDA path against the reference
golden vectors and the CPU oracle, for both convolution engines.  Prints JSON lines (no asserts)."""
import json, math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from fastvideocodec_b200 import VideoCompressor
from fastvideocodec_b200.synthetic import init_state_dict, synthetic_gop
from oracle import dvc_oracle as O

dev = torch.device("cuda")
sd = init_state_dict(0)
model = VideoCompressor(); model.load_state_dict(sd); model = model.to(dev).eval()
tag = os.environ.get("FVC_SPLIT", "fp16")

def gold(name):
    with np.load(os.path.join(ROOT, "tests", "golden", name)) as z:
        return {k: torch.from_numpy(z[k]) for k in z.files}

def psnr(m): return 10 * math.log10(1 / float(m))

def frame_report(label, g, impl):
    model.impl = impl
    with torch.no_grad():
        out = model(g["cur"].to(dev), g["ref"].to(dev))
    r = {"split": tag, "test": label, "engine": "tc" if impl else "simt"}
    for n in ("quant_mv", "z_hat", "feat_hat"):
        a = model.get_intermediate(n).cpu()
        r[n + "_mismatch"] = int((a != g[n]).sum()); r[n + "_n"] = a.numel()
    for n in ("estmv", "mvfeature", "mv_hat", "prediction", "feature", "z", "sigma", "recon_res"):
        a = model.get_intermediate(n).cpu()
        r[n + "_maxabs"] = float((a - g[n]).abs().max()); r[n + "_meanabs"] = float((a - g[n]).abs().mean())
    r["clipped_maxabs"] = float((out[0].cpu() - g["clipped"]).abs().max())
    r["bpp_rel"] = abs(float(out[7]) - float(g["bpp"])) / float(g["bpp"])
    r["psnr_db"] = abs(psnr(out[1]) - psnr(g["mse"]))
    print(json.dumps(r)); sys.stdout.flush()

for impl in (0, 1):
    frame_report("golden64", gold("pframe_64.npz"), impl)
    frame_report("golden128", gold("pframe_128.npz"), impl)

# config 1: 256x256 open-loop frame vs oracle (captures) and closed-loop GOP
fr = synthetic_gop(256, 256, gop=10, gop_id=0)[:, 0]
with torch.no_grad():
    o, cap = O.pframe_forward(sd, fr[1:2], fr[0:1], capture=True)
g = dict(cap); g.update(cur=fr[1:2], ref=fr[0:1], clipped=o[0], bpp=o[7], mse=o[1])
for impl in (0, 1):
    frame_report("cfg1_256_open", g, impl)
rows, rec = O.gop_forward(sd, fr)
for impl in (0, 1):
    model.impl = impl
    grec, sc = model.gop_forward_host(fr.unsqueeze(1).contiguous())
    r = {"split": tag, "test": "cfg1_256_gop10_closed", "engine": "tc" if impl else "simt",
         "recon_maxabs": float((grec[:, 0] - rec).abs().max()),
         "bpp_rel_max": max(abs(float(sc[i, 6]) - rows[i][0]) / rows[i][0] for i in range(9)),
         "psnr_db_max": max(abs(psnr(sc[i, 0]) - rows[i][1]) for i in range(9))}
    print(json.dumps(r)); sys.stdout.flush()

# HD: engine vs engine
fh = synthetic_gop(1088, 1920, gop=2, gop_id=3)[:, 0].to(dev)
res = {}
for impl in (0, 1):
    model.impl = impl
    with torch.no_grad():
        out = model(fh[1:2], fh[0:1])
    res[impl] = (out, {n: model.get_intermediate(n) for n in ("quant_mv", "feat_hat", "z_hat", "mvfeature", "feature", "estmv")})
r = {"split": tag, "test": "hd_simt_vs_tc"}
for n in ("quant_mv", "feat_hat", "z_hat"):
    r[n + "_diff"] = int((res[0][1][n] != res[1][1][n]).sum()); r[n + "_n"] = res[0][1][n].numel()
for n in ("mvfeature", "feature", "estmv"):
    r[n + "_meanabs"] = float((res[0][1][n] - res[1][1][n]).abs().mean())
r["bpp_rel"] = abs(float(res[0][0][7]) - float(res[1][0][7])) / float(res[0][0][7])
r["psnr_db"] = abs(psnr(res[0][0][1]) - psnr(res[1][0][1]))
print(json.dumps(r))
