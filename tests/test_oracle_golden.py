"""Pins the CPU oracle (oracle/dvc_oracle.py) against outputs of the unmodified reference.

The golden vectors were produced by oracle/gen_golden.py importing /root/reference; these tests
run everywhere (no GPU, no reference needed).  ``test_live_reference_*`` additionally re-runs the
reference when the checkout is present (build container only).
"""
import math
import os

import pytest
import torch

from oracle import dvc_oracle as O
from oracle import ref_shim

INTER = ["estmv", "mvfeature", "quant_mv", "mv_hat", "warpframe", "prediction", "feature", "z", "z_hat",
         "sigma", "feat_hat", "recon_res", "recon"]
SCALARS = ["mse", "warploss", "interloss", "bpp_feature", "bpp_z", "bpp_mv", "bpp"]


def _check_pframe(sd, gold):
    out, cap = O.pframe_forward(sd, gold["cur"], gold["ref"], capture=True)
    for name in INTER:
        a, b = cap[name], gold[name]
        assert a.shape == b.shape, name
        if name in ("quant_mv", "z_hat", "feat_hat"):
            mism = (a != b).float().mean().item()
            assert mism <= 1e-4, (name, mism)
        else:
            err = (a - b).abs().max().item()
            scale = max(1.0, b.abs().max().item())
            assert err <= 2e-4 * scale, (name, err, scale)
    assert (out[0] - gold["clipped"]).abs().max().item() <= 1e-4
    for i, name in enumerate(SCALARS, start=1):
        a, b = float(out[i]), float(gold[name])
        assert abs(a - b) <= 1e-4 * max(abs(b), 1e-3), (name, a, b)


def test_oracle_matches_reference_pframe_64(state_dict, golden_pframe_64):
    _check_pframe(state_dict, golden_pframe_64)


def test_oracle_matches_reference_pframe_128(state_dict, golden_pframe_128):
    _check_pframe(state_dict, golden_pframe_128)


def test_oracle_matches_reference_gop(state_dict, golden_gop_64):
    rows, rec = O.gop_forward(state_dict, golden_gop_64["frames"])
    g = golden_gop_64["rows"]  # [mse, warploss, interloss, bpp_f, bpp_z, bpp_mv, bpp, psnr]
    assert (rec - golden_gop_64["recon"]).abs().max().item() <= 1e-2
    for i, (bpp, psnr, mse) in enumerate(rows):
        assert abs(bpp - g[i, 6].item()) <= 0.005 * g[i, 6].item()
        assert abs(psnr - g[i, 7].item()) <= 0.02


def test_oracle_ops(golden_ops):
    g = golden_ops
    assert (O.flow_warp(g["warp_img"], g["warp_flow"]) - g["warp_out"]).abs().max() <= 1e-5
    z = O.flow_warp(g["warp_img"], torch.zeros_like(g["warp_flow"]))
    assert (z - g["warp_zero_flow_out"]).abs().max() <= 1e-5
    # zero flow is NOT the identity
    assert (z - g["warp_img"]).abs().max() > 1e-3
    assert (O.upsample2x_bilinear(g["up_in"], False) - g["up_half_pixel"]).abs().max() <= 1e-6
    assert (O.upsample2x_bilinear(g["up_in"], True) - g["up_align_corners"]).abs().max() <= 1e-5
    assert torch.equal(O.avg_pool2(g["up_in"]), g["pool_out"])
    sd = {"g.beta": g["gdn_beta"], "g.gamma": g["gdn_gamma"]}
    assert (O.gdn(sd, "g", g["gdn_in"]) - g["gdn_out"]).abs().max() <= 1e-5
    assert (O.gdn(sd, "g", g["gdn_in"], inverse=True) - g["igdn_out"]).abs().max() <= 1e-4


def test_oracle_bit_estimators(state_dict, golden_ops):
    g = golden_ops
    hi = O.bit_estimator_cdf(state_dict, "bitEstimator_z", g["be_q"] + 0.5)
    lo = O.bit_estimator_cdf(state_dict, "bitEstimator_z", g["be_q"] - 0.5)
    assert (hi - g["be_cdf_hi"]).abs().max() <= 1e-6
    assert (lo - g["be_cdf_lo"]).abs().max() <= 1e-6
    _, prob = O.laplace_bits(g["lap_q"], g["lap_sigma"])
    assert (prob - g["lap_prob"]).abs().max() <= 1e-6


@pytest.mark.skipif(not ref_shim.available(), reason="reference checkout not present")
def test_live_reference_real_spynet_weights():
    """Same check against the live reference with its real SpyNet .npy weights (|w|max ~ 5)."""
    from fastvideocodec_b200.synthetic import synthetic_gop
    torch.manual_seed(0)
    model = ref_shim.build_reference_model(None)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    fr = synthetic_gop(64, 64, gop=2, gop_id=5)[:, 0]
    out_ref, cap_ref = ref_shim.run_reference_with_capture(model, fr[1:2], fr[0:1])
    out, cap = O.pframe_forward(sd, fr[1:2], fr[0:1], capture=True)
    for name in ("estmv", "mv_hat", "prediction", "recon"):
        err = (cap[name] - cap_ref[name]).abs().max().item()
        assert err <= 2e-4 * max(1.0, cap_ref[name].abs().max().item()), (name, err)
    for name in ("quant_mv", "z_hat", "feat_hat"):
        assert (cap[name] != cap_ref[name]).float().mean().item() <= 1e-4
    assert abs(float(out[7]) - float(out_ref[7])) <= 1e-4 * float(out_ref[7])
    assert abs(10 * math.log10(1 / float(out[1])) - 10 * math.log10(1 / float(out_ref[1]))) <= 0.02


def test_oracle_lsvc_matches_reference(state_dict):
    """the oracle's LSVC.forward restatement (tree and chain GOP graphs) against outputs of the
    unmodified reference models.LSVC('LSVC-128' / 'LSVC-L-128') — tests/golden/lsvc_64.npz (oracle/gen_golden_lsvc.py)."""
    from conftest import load_golden
    from fastvideocodec_b200.lsvc import graph_from_batch, refidx_from_graph
    gold = load_golden("lsvc_64.npz")
    x = gold["x"]
    for tag, linear, onehop in (("tree", False, False), ("chain", True, False), ("onehop", False, True)):
        g, layers, parents = graph_from_batch(4, isLinear=linear, isOnehop=onehop)
        out = O.lsvc_forward(state_dict, x, layers, parents, refidx_from_graph(g, 4))
        for i, n in enumerate(["com", "mc", "warped"]):
            assert (out[i] - gold["%s_%s" % (tag, n)]).abs().max().item() <= 2e-5, (tag, n)
        for i, n in enumerate(["rec_loss", "warp_loss", "mc_loss", "bpp_res", "bpp"], start=3):
            a, b = float(out[i]), float(gold["%s_%s" % (tag, n)])
            assert abs(a - b) <= 1e-5 * abs(b), (tag, n, a, b)


def test_lsvc_graph_helpers():
    """generate_graph / graph_from_batch / refidx_from_graph (models.py:683-728, 923-949)."""
    from fastvideocodec_b200.lsvc import graph_from_batch, refidx_from_graph
    g, layers, parents = graph_from_batch(6)
    assert layers == [[1, 4], [2, 3, 5, 6]] and refidx_from_graph(g, 6) == [0, 1, 1, 0, 4, 4]
    g, layers, parents = graph_from_batch(4)
    assert refidx_from_graph(g, 4) == [0, 1, 1, 0]
    g, layers, parents = graph_from_batch(9, isLinear=True)
    assert refidx_from_graph(g, 9) == list(range(9)) and layers[:3] == [[1], [2], [3]]
    g, layers, parents = graph_from_batch(14)
    assert len(layers) == 3 and parents[14] == 12 and refidx_from_graph(g, 14)[7] == 0
    g, layers, parents = graph_from_batch(5, isOnehop=True)
    assert refidx_from_graph(g, 5) == [0] * 5


# ------------------------------------------------------------------------------------------------
# round 2 goldens (oracle/gen_golden_r2.py)
# ------------------------------------------------------------------------------------------------
def test_oracle_matches_reference_with_real_spynet_weights(real_state_dict, golden_pframe_real_128):
    """The reference run with its own pretrained SpyNet weights (committed as tests/golden/spynet_real.npz)."""
    assert max(v.abs().max().item() for k, v in real_state_dict.items() if k.startswith("opticFlow.")) > 3.0
    _check_pframe(real_state_dict, golden_pframe_real_128)


def test_oracle_matches_reference_deeper_pyramid_L6(golden_pframe_L6_256):
    """levels=6 (configs[3] "deeper flow pyramid") against the reference class with ME_Spynet.L patched to 6."""
    from fastvideocodec_b200.synthetic import init_state_dict
    g = golden_pframe_L6_256
    sd = init_state_dict(0, spynet_levels=6, spynet_gain=1.8)
    out, cap = O.pframe_forward(sd, g["cur"], g["ref"], levels=6, capture=True)
    for name in ("estmv", "mv_hat", "mvfeature", "feature", "z", "sigma"):
        err = (cap[name] - g[name]).abs().max().item()
        assert err <= 2e-4 * max(1.0, g[name].abs().max().item()), (name, err)
    for name in ("quant_mv", "z_hat", "feat_hat"):
        assert (cap[name] != g[name]).float().mean().item() <= 1e-4, name
    assert (out[0] - g["clipped"]).abs().max().item() <= 1e-4
    for i, name in enumerate(SCALARS, start=1):
        assert abs(float(out[i]) - float(g[name])) <= 1e-4 * max(abs(float(g[name])), 1e-3), name
    # and it differs from the 4-level pyramid (the extra levels are really used)
    out4 = O.pframe_forward({k: v for k, v in sd.items()}, g["cur"], g["ref"], levels=4)
    assert abs(float(out4[7]) - float(out[7])) > 1e-6


def test_oracle_matches_reference_hd_gop_first_frame(state_dict, golden_hd_gop10):
    """First (open-loop) P-frame of the 1088x1920 GOP bench.py times: the oracle against the unmodified reference's
    scalars, quantised latents and clipped frame (tests/golden/hd_gop10.npz).  ~15 s of host time."""
    from fastvideocodec_b200.synthetic import synthetic_gop
    g = golden_hd_gop10
    frames = synthetic_gop(1088, 1920, gop=2, gop_id=int(g["gop_id"]))[:, 0]
    out, cap = O.pframe_forward(state_dict, frames[1:2], frames[0:1], capture=True)
    for name in ("quant_mv", "z_hat", "feat_hat"):
        assert (cap[name] != g["f1_" + name].float()).float().mean().item() <= 1e-4, name
    want = torch.from_numpy(g["f1_clipped_u16"].numpy().astype("float32")) / 65535.0
    assert ((out[0] - want).abs() > 2e-5).float().mean().item() <= 1e-3     # u16 steps; tie flips perturb locally
    for i in range(7):
        assert abs(float(out[1 + i]) - g["rows"][0, i].item()) <= 1e-4 * abs(g["rows"][0, i].item()), i


def test_oracle_iframe_composition_matches_reference_modules(state_dict):
    """O.iframe_forward composes the residual branch (net.py:86-105) on the frame itself; its pieces are
    checked here against the live reference's own modules (resEncoder, respriorEncoder, respriorDecoder, resDecoder)."""
    if not os.path.isdir(ref_shim.REF_ROOT):
        pytest.skip("reference checkout not present")
    from fastvideocodec_b200.synthetic import synthetic_gop
    model = ref_shim.build_reference_model(state_dict)
    x = synthetic_gop(64, 128, gop=2, gop_id=9)[1:2, 0].contiguous()
    out, cap = O.iframe_forward(state_dict, x, capture=True)
    with torch.no_grad():
        feature = model.resEncoder(x)
        z = model.respriorEncoder(feature)
        sigma = model.respriorDecoder(torch.round(z))
        recon = model.resDecoder(torch.round(feature))
    for got, ref, name in ((cap["feature"], feature, "feature"), (cap["z"], z, "z"), (cap["sigma"], sigma, "sigma"),
                           (cap["recon_res"], recon, "recon")):
        err = (got - ref).abs().max().item()
        assert err <= 2e-5 * max(1.0, ref.abs().max().item()), (name, err)
    assert torch.equal(out[0], cap["recon_res"].clamp(0, 1))
    assert abs(float(out[4]) - float(out[2]) - float(out[3])) <= 1e-6 * float(out[4])
