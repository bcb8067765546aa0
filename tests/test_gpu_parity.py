"""GPU parity tests: every CUDA kernel / the whole P-frame path, called through the C ABI, against
the CPU oracle on identical seeded inputs and against the committed golden vectors produced by the
unmodified reference.  Gates are the north-star's: quantised latents bit-exact except <= 1e-4 of
elements, reconstructed frames <= 1e-2 max-abs, bpp within 0.5 %, PSNR within 0.02 dB."""
import math
import os

import pytest
import torch

from oracle import dvc_oracle as O

pytestmark = pytest.mark.gpu

LATENTS = ("quant_mv", "z_hat", "feat_hat")
PREQUANT = {"quant_mv": "mvfeature", "z_hat": "z", "feat_hat": "feature"}
TIE_TOL = 1.8e-4  # |x - (k + 1/2)| of the REFERENCE pre-quantisation value for a mismatch to count as a rounding
                  # tie, in units of max(1, rms of the tensor) (fp32 rounding noise scales with the magnitudes).
                  # = 3x the largest pre-round error measured on the tcgen05 engine (mvfeature at 1080p: 2.0e-4
                  # max-abs at rms 3.4, profiles/r01_parity_fp16_shortchain.jsonl); since q = rint(pre) bit-exactly,
                  # a flip further from the tie than this means the pre-round value itself is off by more than that.


def check_latent(name, got, want, prequant, tie_rule=True, extra=0):
    """North-star gate for quantised latents: bit-exact except <= 1e-4 of the elements, and every
    exception must be a rounding tie: it differs by exactly 1 and the reference's own pre-round value
    lies within TIE_TOL * max(1, rms(x)) of k + 1/2 (fp32 evaluation order alone moves such values across the tie; the
    unmodified reference on a GPU flips them against its own CPU run as well).  For tensors so small
    that 1e-4 of the elements is less than one element the count bound is 2 elements: one tie in
    8192 is already 1.2e-4, so the fraction is only meaningful for the HD-sized tensors, where it is
    enforced as stated."""
    diff = got != want
    n_bad = int(diff.sum())
    if n_bad == 0:
        return 0.0
    # `extra`: free-running feat_hat / z_hat only: every upstream quant_mv flip moves `feature` (and `z`) in its
    # neighbourhood by ~1e-3 and may push elements there across their tie (second-order flips); the strict
    # per-tensor bound for these two is enforced in the teacher-forced pass
    assert n_bad <= max(2, int(1e-4 * got.numel())) + extra, (name, n_bad, got.numel(), extra)
    assert float((got - want)[diff].abs().max()) == 1.0, (name, "mismatch by more than one level")
    if not tie_rule:
        return n_bad / got.numel()
    frac = prequant[diff] - torch.floor(prequant[diff])
    tol = TIE_TOL * max(1.0, float(prequant.pow(2).mean().sqrt()))
    assert float((frac - 0.5).abs().max()) <= tol, (name, "mismatch away from a rounding tie",
                                                    float((frac - 0.5).abs().max()), tol)
    return n_bad / got.numel()


def _impls():
    """Both convolution engines; the tcgen05 engine is listed once the library ships it (version >= 200)."""
    from fastvideocodec_b200 import _lib
    out = [("simt", _lib.IMPL_SIMT)]
    try:
        if _lib.lib().fvc_version() >= 200:
            out.append(("tc", _lib.IMPL_TC))
    except Exception:
        pass
    return out


def _psnr(mse):
    return 10.0 * math.log10(1.0 / float(mse))


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def model(dev, state_dict):
    from fastvideocodec_b200 import VideoCompressor
    m = VideoCompressor()
    m.load_state_dict(state_dict)
    return m.to(dev).eval()


# ------------------------------------------------------------------------------------------------
# op level
# ------------------------------------------------------------------------------------------------
def test_avg_pool2_bit_exact(dev):
    from fastvideocodec_b200 import ops
    g = torch.Generator().manual_seed(1)
    for shape in ((1, 3, 64, 64), (2, 3, 136, 240), (1, 1, 2, 2)):
        x = torch.rand(shape, generator=g)
        assert torch.equal(ops.avg_pool2(x.to(dev)).cpu(), O.avg_pool2(x))


@pytest.mark.parametrize("ac", [False, True])
def test_upsample2x(dev, ac):
    from fastvideocodec_b200 import ops
    g = torch.Generator().manual_seed(2)
    for shape in ((1, 2, 8, 8), (2, 2, 17, 30), (1, 64, 5, 3), (1, 1, 1, 1)):
        x = torch.randn(shape, generator=g) * 3
        got = ops.upsample2x_bilinear(x.to(dev), ac, 2.0).cpu()
        want = O.upsample2x_bilinear(x, ac) * 2.0
        assert (got - want).abs().max().item() <= 2e-6 * max(1.0, want.abs().max().item())


def test_flow_warp(dev, golden_ops):
    from fastvideocodec_b200 import ops
    g = golden_ops
    got = ops.flow_warp(g["warp_img"].to(dev), g["warp_flow"].to(dev)).cpu()
    assert (got - g["warp_out"]).abs().max().item() <= 1e-5
    z = ops.flow_warp(g["warp_img"].to(dev), torch.zeros_like(g["warp_flow"]).to(dev)).cpu()
    assert (z - g["warp_zero_flow_out"]).abs().max().item() <= 1e-5
    gen = torch.Generator().manual_seed(3)
    img = torch.rand((2, 3, 68, 120), generator=gen)
    flow = torch.randn((2, 2, 68, 120), generator=gen) * 20  # many samples clamp at the border
    got = ops.flow_warp(img.to(dev), flow.to(dev)).cpu()
    assert (got - O.flow_warp(img, flow)).abs().max().item() <= 2e-4


CONV_CASES = [
    # cin, cout, k, stride, transposed, act, H, W
    (8, 32, 7, 1, 0, 1, 24, 40),
    (32, 64, 7, 1, 0, 1, 17, 30),
    (64, 32, 7, 1, 0, 1, 16, 24),
    (16, 2, 7, 1, 0, 0, 20, 20),
    (2, 128, 3, 2, 0, 2, 32, 48),
    (128, 128, 3, 1, 0, 2, 17, 30),
    (128, 128, 3, 2, 0, 2, 16, 32),
    (128, 128, 3, 2, 1, 2, 9, 15),
    (128, 2, 3, 1, 0, 0, 16, 16),
    (6, 64, 3, 1, 0, 1, 16, 40),
    (64, 64, 3, 1, 0, 0, 33, 21),
    (64, 3, 3, 1, 0, 0, 16, 16),
    (3, 64, 5, 2, 0, 0, 32, 64),
    (64, 96, 5, 2, 0, 0, 8, 16),
    (96, 64, 5, 2, 1, 0, 4, 8),
    (64, 3, 5, 2, 1, 0, 8, 12),
    (96, 64, 3, 1, 0, 1, 4, 8),
    (64, 64, 5, 2, 1, 1, 1, 2),
    (64, 96, 3, 1, 1, 3, 4, 7),
]


@pytest.mark.parametrize("impl_name,impl", _impls())
@pytest.mark.parametrize("case", CONV_CASES)
def test_conv2d_matches_oracle(dev, impl_name, impl, case):
    import torch.nn.functional as F
    from fastvideocodec_b200 import ops
    cin, cout, k, stride, transposed, act, H, W = case
    g = torch.Generator().manual_seed(100 + cin * 7 + cout + k + stride)
    x = torch.randn((2, cin, H, W), generator=g)
    wshape = (cin, cout, k, k) if transposed else (cout, cin, k, k)
    w = torch.randn(wshape, generator=g) / math.sqrt(cin * k * k)
    b = torch.randn((cout,), generator=g) * 0.1
    if transposed:
        want = F.conv_transpose2d(x, w, b, stride=stride, padding=k // 2, output_padding=stride - 1)
        got = ops.conv_transpose2d(x.to(dev), w.to(dev), b.to(dev), stride, act, impl).cpu()
    else:
        want = F.conv2d(x, w, b, stride=stride, padding=k // 2)
        got = ops.conv2d(x.to(dev), w.to(dev), b.to(dev), stride, act, impl).cpu()
    want = {0: lambda t: t, 1: torch.relu, 2: lambda t: F.leaky_relu(t, 0.1), 3: torch.exp}[act](want)
    assert got.shape == want.shape
    err = (got - want).abs().max().item()
    assert err <= 1e-4 * max(1.0, want.abs().max().item()), (impl_name, case, err)


@pytest.mark.parametrize("inverse", [False, True])
def test_gdn(dev, golden_ops, inverse):
    from fastvideocodec_b200 import ops
    g = golden_ops
    got = ops.gdn(g["gdn_in"].to(dev), g["gdn_beta"].to(dev), g["gdn_gamma"].to(dev), inverse).cpu()
    want = g["igdn_out"] if inverse else g["gdn_out"]
    assert (got - want).abs().max().item() <= 1e-4 * max(1.0, want.abs().max().item())


def test_quant_bits_factorized(dev, state_dict):
    from fastvideocodec_b200 import ops
    g = torch.Generator().manual_seed(5)
    for prefix, C in (("bitEstimator_z", 64), ("bitEstimator_mv", 128)):
        x = torch.randn((2, C, 17, 30), generator=g) * 4
        x[0, 0, 0, :8] = torch.tensor([0.5, 1.5, 2.5, -0.5, -1.5, 3.5, 40.0, -60.0])  # ties: half to even
        params = []
        for i in (1, 2, 3, 4):
            for p in ("h", "b", "a"):
                if not (i == 4 and p == "a"):
                    params.append(state_dict[f"{prefix}.f{i}.{p}"].to(dev))
        q, bits = ops.quant_bits_factorized(x.to(dev), params)
        want_q = torch.round(x)
        want_bits, _ = O.factorized_bits(state_dict, prefix, want_q)
        assert torch.equal(q.cpu(), want_q)
        assert abs(float(bits) - float(want_bits)) <= 1e-5 * float(want_bits)
    # empty input (edge case): zero bits
    q, bits = ops.quant_bits_factorized(torch.empty((0, 128, 4, 4), device=dev), params)
    assert float(bits) == 0.0 and q.numel() == 0


def test_quant_bits_laplace(dev, golden_ops):
    from fastvideocodec_b200 import ops
    g = torch.Generator().manual_seed(6)
    x = torch.randn((1, 96, 68, 120), generator=g) * 3
    sigma = torch.exp(torch.randn((1, 96, 68, 120), generator=g) * 2)
    sigma[0, 0, 0, :4] = torch.tensor([0.0, 1e-7, 1e12, 1e-5])  # clamp(1e-5, 1e10)
    q, bits = ops.quant_bits_laplace(x.to(dev), sigma.to(dev))
    want_bits, _ = O.laplace_bits(torch.round(x), sigma)
    assert torch.equal(q.cpu(), torch.round(x))
    assert abs(float(bits) - float(want_bits)) <= 1e-5 * float(want_bits)
    # known answers from the reference's torch.distributions.Laplace
    q2, bits2 = ops.quant_bits_laplace(golden_ops["lap_q"].to(dev), golden_ops["lap_sigma"].to(dev))
    want2 = O.clamp_log2_bits(golden_ops["lap_prob"])
    assert abs(float(bits2) - float(want2)) <= 1e-5 * float(want2)


def test_recon_losses(dev):
    from fastvideocodec_b200 import ops
    g = torch.Generator().manual_seed(7)
    cur, pred, warp = (torch.rand((1, 3, 64, 128), generator=g) for _ in range(3))
    res = torch.randn((1, 3, 64, 128), generator=g) * 0.3
    clipped, means = ops.recon_losses(cur.to(dev), pred.to(dev), warp.to(dev), res.to(dev))
    rec = pred + res
    assert torch.equal(clipped.cpu(), rec.clamp(0, 1))
    want = [torch.mean((rec - cur) ** 2), torch.mean((warp - cur) ** 2), torch.mean((pred - cur) ** 2)]
    for a, b in zip(means.cpu().tolist(), want):
        assert abs(a - float(b)) <= 1e-5 * float(b)


def test_compressai_style_likelihoods(dev):
    """entropy_models.py boundary (EntropyBottleneck / GaussianConditional); parity unpinned, checked
    against the oracle's restatement of the published CompressAI algorithm."""
    from fastvideocodec_b200 import ops
    g = torch.Generator().manual_seed(8)
    C = 16
    filters = (1, 3, 3, 3, 3, 1)
    mats = [torch.randn((C, filters[i + 1], filters[i]), generator=g) * 0.5 for i in range(5)]
    bias = [torch.randn((C, filters[i + 1], 1), generator=g) * 0.5 for i in range(5)]
    facs = [torch.randn((C, filters[i + 1], 1), generator=g) * 0.5 for i in range(4)]
    med = torch.randn((C,), generator=g)
    x = torch.randn((2, C, 9, 11), generator=g) * 5
    xh, lik, bits = ops.eb_forward(x.to(dev), ops.pack_eb_params(mats, bias, facs).to(dev), med.to(dev))
    wxh, wlik = O.eb_forward(mats, bias, facs, med, x)
    assert torch.equal(xh.cpu(), wxh)
    assert (lik.cpu() - wlik).abs().max().item() <= 2e-6
    assert abs(float(bits) - float(O.clamp_log2_bits(wlik))) <= 1e-4 * float(O.clamp_log2_bits(wlik))
    scales = torch.exp(torch.randn(x.shape, generator=g))
    means = torch.randn(x.shape, generator=g)
    xh, lik, bits = ops.gaussian_forward(x.to(dev), scales.to(dev), means.to(dev))
    wxh, wlik = O.gaussian_forward(x, scales, means)
    assert (xh.cpu() - wxh).abs().max().item() <= 1e-6
    assert (lik.cpu() - wlik).abs().max().item() <= 2e-6
    assert abs(float(bits) - float(O.clamp_log2_bits(wlik))) <= 1e-4 * float(O.clamp_log2_bits(wlik))


# ------------------------------------------------------------------------------------------------
# whole P-frame against the reference's golden vectors
# ------------------------------------------------------------------------------------------------
def _flip_mask(model, gold, H, W, report=None):
    """Free-running latent gates + the receptive-field mask around feat_hat flips (_mask_feat_flips).  The tie rule is applied here to quant_mv only, the one quantiser with nothing quantised upstream:
    a quant_mv flip moves `feature` in its neighbourhood by ~1e-3, so free-running feat_hat / z_hat flips may be
    second-order; their tie rule is checked in the teacher-forced pass (_check_against, part B), where the
    reference's quant_mv is forced upstream and our own pre-round `feature` / `z` are rounded."""
    mask = torch.zeros((H, W), dtype=torch.bool)
    n_mv = 0
    for name, scale in (("quant_mv", 16), ("feat_hat", 16), ("z_hat", 64)):
        a = model.get_intermediate(name).cpu()
        frac = check_latent(name, a, gold[name], gold[PREQUANT[name]], tie_rule=(name == "quant_mv"),
                            extra=0 if name == "quant_mv" else n_mv)
        if name == "quant_mv":
            n_mv = int((a != gold[name]).sum())
        if report is not None:
            report[name] = frac
        if name == "feat_hat":
            _mask_feat_flips(mask, a != gold[name])
    return mask


def _mask_feat_flips(mask, diff):
    """Receptive field of a flipped feat_hat element in the reconstructed frame: resDecoder = four k5 s2 transposed
    convs, output pixels [16y - 30, 16y + 30] (measured on the oracle: |d recon| ~ 0.1 inside, exactly 0 outside);
    masked with margin: +-48 px around the latent's centre.  Flipped quant_mv elements are NOT masked: one level of
    one of the 128 mv channels moves the decoded frame by <= 2e-3 (measured on the oracle, 6 channels), well inside
    the 1e-2 gate; z_hat only feeds sigma (rate, not reconstruction)."""
    for (_, _, y, x) in diff.nonzero().tolist():
        cy, cx = y * 16 + 8, x * 16 + 8
        mask[max(0, cy - 48):cy + 48, max(0, cx - 48):cx + 48] = True


INTERMEDIATES = ("estmv", "mvfeature", "mv_hat", "warpframe", "prediction", "feature", "z", "sigma", "recon_res")
SCALAR_NAMES = ["mse", "warploss", "interloss", "bpp_feature", "bpp_z", "bpp_mv", "bpp"]


def _check_against(model, gold, dev, masked_free_run=False, inter_tol=5e-4):
    """All north-star gates of one P-frame against reference / oracle tensors `gold`.

    A. free running: quantised latents (count, +-1, tie rule), the tensors upstream of every quantiser (estmv,
       mvfeature) element-wise WITHOUT any mask, bpp 0.5 %, PSNR 0.02 dB, the other scalars 0.5 %.
    B. teacher forced (the reference's own quantised latents replace ours right after the three quantisers,
       fvc_ctx_force_latents): EVERY intermediate and the reconstructed frame element-wise, no mask at all:
       intermediates <= inter_tol * scale, recon <= 1e-2 (north star) -- in fact <= 1e-3.
       (One latent that flips at a rounding tie moves decoded pixels by ~0.1 with random-init decoders, SURVEY 7.2-1
       caveat (i); forcing removes the flips instead of hiding their neighbourhood.)
    C. masked_free_run (HD and larger): additionally the free-running reconstruction <= 1e-2 outside the +-48 px
       receptive field of flipped feat_hat elements, and the mask may not cover more than half of the frame.
    """
    cur, ref = gold["cur"].to(dev), gold["ref"].to(dev)
    B, _, H, W = cur.shape
    report = {}
    model.force_latents(B, H, W, dev)                      # free running
    with torch.no_grad():
        out = model(cur, ref)
    torch.cuda.synchronize()
    mask = _flip_mask(model, gold, H, W, report)
    for name in ("estmv", "mvfeature"):
        err = (model.get_intermediate(name).cpu() - gold[name]).abs().max().item()
        report[name] = err
        assert err <= inter_tol * max(1.0, gold[name].abs().max().item()), (name, err)
    for i, n in enumerate(SCALAR_NAMES, start=1):
        a, b = float(out[i]), float(gold[n])
        assert abs(a - b) <= 0.005 * abs(b), (n, a, b)
    assert abs(_psnr(out[1]) - _psnr(gold["mse"])) <= 0.02
    if masked_free_run:
        report["masked_fraction"] = mask.float().mean().item()
        assert report["masked_fraction"] <= 0.5, "more than half of the frame is masked"
        d = (out[0].cpu() - gold["clipped"]).abs().masked_fill(mask, 0.0)
        report["recon_free_masked"] = d.max().item()
        assert report["recon_free_masked"] <= 1e-2
    # ---- B: teacher forced ------------------------------------------------------------------------------
    forced = [gold[n].to(dev).float().contiguous() for n in LATENTS]   # quant_mv, z_hat, feat_hat
    model.force_latents(B, H, W, dev, *forced)
    try:
        with torch.no_grad():
            fout = model(cur, ref)
        torch.cuda.synchronize()
        for name in INTERMEDIATES:
            err = (model.get_intermediate(name).cpu() - gold[name]).abs().max().item()
            report["forced_" + name] = err
            assert err <= inter_tol * max(1.0, gold[name].abs().max().item()), (name, err)
        # our own quantisers under forcing: round(pre-round tensor) against the reference's latents, tie rule on
        for name in ("feat_hat", "z_hat"):
            q = torch.round(model.get_intermediate(PREQUANT[name]).cpu())
            report["forced_" + name] = check_latent(name, q, gold[name], gold[PREQUANT[name]])
        err = (fout[0].cpu() - gold["clipped"]).abs().max().item()
        report["forced_recon"] = err
        assert err <= 1e-3, err                      # north star: 1e-2
        for i, n in enumerate(SCALAR_NAMES[:3], start=1):
            a, b = float(fout[i]), float(gold[n])
            assert abs(a - b) <= 1e-3 * abs(b), (n, a, b)
    finally:
        model.force_latents(B, H, W, dev)
    return report


def _assert_full_size_latent_rates(rep, gold):
    """North-star fraction gate at full size (>= 1080p): quant_mv free running, feat_hat / z_hat of our own quantisers
    with the reference's quant_mv forced upstream: <= 1e-4 of the elements each.  Free-running feat_hat / z_hat may in
    addition carry the second-order flips induced by the quant_mv flips (at most one per flipped mv element)."""
    assert rep["quant_mv"] <= 1e-4, rep["quant_mv"]
    assert rep["forced_feat_hat"] <= 1e-4 and rep["forced_z_hat"] <= 1e-4, (rep["forced_feat_hat"], rep["forced_z_hat"])
    n_mv = rep["quant_mv"] * gold["quant_mv"].numel()
    for name in ("feat_hat", "z_hat"):
        assert rep[name] <= 1e-4 + n_mv / gold[name].numel(), (name, rep[name])


def _gold_from_oracle(sd, cur, ref, levels=4):
    with torch.no_grad():
        o, cap = O.pframe_forward(sd, cur, ref, levels=levels, capture=True)
    gold = dict(cap)
    gold.update(cur=cur, ref=ref, clipped=o[0], mse=o[1], warploss=o[2], interloss=o[3], bpp_feature=o[4], bpp_z=o[5],
                bpp_mv=o[6], bpp=o[7])
    return gold


@pytest.mark.parametrize("impl_name,impl", _impls())
def test_pframe_matches_reference_golden_64(model, golden_pframe_64, dev, impl_name, impl):
    model.impl = impl
    _check_against(model, golden_pframe_64, dev)


@pytest.mark.parametrize("impl_name,impl", _impls())
def test_pframe_matches_reference_golden_128(model, golden_pframe_128, dev, impl_name, impl):
    model.impl = impl
    _check_against(model, golden_pframe_128, dev)


@pytest.mark.parametrize("impl_name,impl", _impls())
def test_gop_closed_loop_matches_reference_golden(model, golden_gop_64, dev, impl_name, impl):
    """parallel_compression, 'DVC-pretrained' branch, closed loop over 3 P-frames (models.py:368-383)
    against the reference's own closed-loop run: the metric-level gates (bpp 0.5 %, PSNR 0.02 dB per frame).
    Element-level recon parity in closed loop is checked step by step in
    test_gop_closed_loop_stepwise_vs_oracle: once one latent flips at a rounding tie the next reference
    frame differs locally by ~0.1 and a max-abs gate on later frames is meaningless."""
    from fastvideocodec_b200 import parallel_compression
    model.impl = impl
    model.r = 1024
    data = golden_gop_64["frames"].to(dev)
    with torch.no_grad():
        out = parallel_compression(None, model, data)
    rows = golden_gop_64["rows"]
    assert (out[0].cpu() - golden_gop_64["recon"]).abs().mean().item() <= 2e-3
    assert abs(out[3] - rows[:, 6].mean().item()) <= 0.005 * rows[:, 6].mean().item()
    assert abs(out[5] - rows[:, 7].mean().item()) <= 0.02
    for got, want in zip(out[6], rows[:, 7].tolist()):
        assert abs(got - want) <= 0.02
    # the host-buffer GOP entry point is the same closed loop: identical numbers, bit for bit
    rec, sc = model.gop_forward_host(golden_gop_64["frames"].unsqueeze(1).contiguous().pin_memory())
    assert torch.equal(rec[:, 0], out[0].cpu())
    assert (sc[:, 6].double() - rows[:, 6]).abs().max().item() <= 0.005 * rows[:, 6].max().item()


def test_gop_closed_loop_stepwise_vs_oracle(model, state_dict, dev):
    """Closed loop, element level: frame i is coded from OUR previous reconstruction, and the oracle is
    evaluated on exactly the same inputs, so every step is an open-loop comparison with all four gates
    (tie-aware latents, masked recon <= 1e-2, bpp 0.5 %, PSNR 0.02 dB) while x_prev is carried as in
    models.py:372-375."""
    from fastvideocodec_b200.synthetic import synthetic_gop
    model.impl = _impls()[-1][1]
    frames = synthetic_gop(128, 192, gop=5, gop_id=2)[:, 0]
    prev = frames[0:1]
    for i in range(1, 5):
        _check_against(model, _gold_from_oracle(state_dict, frames[i:i + 1], prev), dev)
        with torch.no_grad():
            prev = model(frames[i:i + 1].to(dev), prev.to(dev))[0].cpu()


@pytest.mark.parametrize("impl_name,impl", _impls())
def test_config1_256x256_gop10_vs_oracle(model, state_dict, dev, impl_name, impl):
    """BASELINE config 1: 256x256, GOP 10 (9 P-frames), closed loop, against the CPU oracle."""
    from fastvideocodec_b200.synthetic import synthetic_gop
    model.impl = impl
    frames = synthetic_gop(256, 256, gop=10, gop_id=0)[:, 0]
    rows, rec = O.gop_forward(state_dict, frames)
    got_rec, sc = model.gop_forward_host(frames.unsqueeze(1).contiguous())
    # open-loop per frame would be tighter; closed loop compounds through x_prev, gates still hold
    bpp = sum(r[0] for r in rows) / len(rows)
    psnr = sum(r[1] for r in rows) / len(rows)
    got_bpp = float(sc[:, 6].mean())
    got_psnr = sum(_psnr(m) for m in sc[:, 0].tolist()) / len(rows)
    assert abs(got_bpp - bpp) <= 0.005 * bpp
    assert abs(got_psnr - psnr) <= 0.02
    # per frame: closed loop amplifies a single tie flip of frame t into frames t+1.. (x_prev differs locally by
    # ~0.1 with random-init decoders), so the per-frame gates are twice the GOP-level ones; element-level
    # closed-loop parity is test_gop_closed_loop_stepwise_vs_oracle
    for i, r in enumerate(rows):
        assert abs(float(sc[i, 6]) - r[0]) <= 0.01 * r[0]
        assert abs(_psnr(sc[i, 0]) - r[1]) <= 0.04


# ------------------------------------------------------------------------------------------------
# full-size properties (no oracle needed)
# ------------------------------------------------------------------------------------------------
def test_hd_frame_properties(model, dev):
    """1088x1920: deterministic, finite, engines agree, quantised latents are integers."""
    from fastvideocodec_b200 import ops
    from fastvideocodec_b200.synthetic import synthetic_gop
    frames = synthetic_gop(1088, 1920, gop=2, gop_id=3)[:, 0].to(dev)
    res = {}
    for name, impl in _impls():
        model.impl = impl
        with torch.no_grad():
            a = model(frames[1:2], frames[0:1])
            q1 = model.get_intermediate("quant_mv")
            b = model(frames[1:2], frames[0:1])
        assert torch.equal(a[0], b[0]) and all(float(x) == float(y) for x, y in zip(a[1:], b[1:]))
        assert all(math.isfinite(float(v)) for v in a[1:])
        assert torch.equal(q1, torch.round(q1))
        assert float(a[0].min()) >= 0.0 and float(a[0].max()) <= 1.0
        assert abs(float(a[7]) - (float(a[4]) + float(a[5]) + float(a[6]))) <= 1e-6 * float(a[7])
        res[name] = (a, q1, model.get_intermediate("feat_hat"))
    if len(res) == 2:
        (a, qa, fa), (b, qb, fb) = res["simt"], res["tc"]
        # full size: the fraction gate as the north star states it (1.04 M and 0.78 M elements)
        assert (qa != qb).float().mean().item() <= 1e-4
        assert (fa != fb).float().mean().item() <= 1e-4
        assert float((qa - qb).abs().max()) <= 1.0 and float((fa - fb).abs().max()) <= 1.0
        assert abs(float(a[7]) - float(b[7])) <= 0.005 * float(a[7])
        assert abs(_psnr(a[1]) - _psnr(b[1])) <= 0.02


def test_hd_frame_matches_oracle(model, state_dict, dev):
    """BASELINE configs[1] at full size (1088x1920): one open-loop P-frame of the tcgen05 engine against
    the CPU oracle (a few seconds of host time), all four north-star gates."""
    from fastvideocodec_b200 import _lib
    from fastvideocodec_b200.synthetic import synthetic_gop
    frames = synthetic_gop(1088, 1920, gop=2, gop_id=5)[:, 0]
    gold = _gold_from_oracle(state_dict, frames[1:2], frames[0:1])
    model.impl = _impls()[-1][1]
    rep = _check_against(model, gold, dev, masked_free_run=True)
    _assert_full_size_latent_rates(rep, gold)
    print("HD open-loop parity report:", {k: float("%.3g" % v) for k, v in rep.items()})


def test_batch_equals_independent_views(model, dev):
    """Multiview shape (views folded into batch, train_multiview.py:232-233): B=2 == two B=1 runs."""
    from fastvideocodec_b200.synthetic import synthetic_gop
    fr = synthetic_gop(128, 192, gop=2, gop_id=9, batch=2).to(dev)  # [2, B=2, 3, H, W]
    with torch.no_grad():
        both = model(fr[1], fr[0])
        one = [model(fr[1, v:v + 1], fr[0, v:v + 1]) for v in range(2)]
    for v in range(2):
        assert (both[0][v:v + 1] - one[v][0]).abs().max().item() <= 1e-6
    bpp_mean = 0.5 * (float(one[0][7]) + float(one[1][7]))
    assert abs(float(both[7]) - bpp_mean) <= 1e-5 * bpp_mean


def test_config4_4k_frame_properties(model, dev):
    """BASELINE configs[3]: 3840x2160 padded to 2176 rows (H, W multiples of 64): one P-frame, both engines agree
    within the north-star gates (latent fraction <= 1e-4, +-1 only; bpp 0.5 %; PSNR 0.02 dB), deterministic."""
    from fastvideocodec_b200.synthetic import synthetic_gop
    frames = synthetic_gop(2176, 3840, gop=2, gop_id=7)[:, 0].to(dev)
    res = {}
    for name, impl in _impls():
        model.impl = impl
        with torch.no_grad():
            a = model(frames[1:2], frames[0:1])
        res[name] = (a, model.get_intermediate("quant_mv"), model.get_intermediate("feat_hat"))
        assert all(math.isfinite(float(v)) for v in a[1:])
        if name == "tc":
            with torch.no_grad():
                b = model(frames[1:2], frames[0:1])
            assert torch.equal(a[0], b[0]) and float(a[7]) == float(b[7])
    if len(res) == 2:
        (a, qa, fa), (b, qb, fb) = res["simt"], res["tc"]
        assert (qa != qb).float().mean().item() <= 1e-4 and float((qa - qb).abs().max()) <= 1.0
        assert (fa != fb).float().mean().item() <= 1e-4 and float((fa - fb).abs().max()) <= 1.0
        assert abs(float(a[7]) - float(b[7])) <= 0.005 * float(a[7])
        assert abs(_psnr(a[1]) - _psnr(b[1])) <= 0.02
    model.release()


@pytest.mark.parametrize("levels", [4, 6])
def test_config4_4k_frame_matches_oracle(dev, levels):
    """BASELINE configs[3] at full size against the CPU oracle (about 1 minute of host time per case): one open-loop
    P-frame at 2176x3840 with every gate of _check_against (free-running latents / metrics, masked free-running
    recon, teacher-forced element-wise comparison of every intermediate).  levels=4 is the reference's pyramid;
    levels=6 is the "deeper flow pyramid" of configs[3] (reference class with ME_Spynet.L patched, endecoder.py:318;
    the oracle's levels=6 path is pinned against that class by tests/golden/pframe_L6_256.npz)."""
    from fastvideocodec_b200 import VideoCompressor
    from fastvideocodec_b200.synthetic import init_state_dict, synthetic_gop
    sd = init_state_dict(0, spynet_levels=levels, spynet_gain=2.2 if levels == 4 else 1.8)   # see gen_golden_r2.L6_GAIN
    m = VideoCompressor(spynet_levels=levels)
    m.load_state_dict(sd)
    m = m.to(dev).eval()
    m.impl = _impls()[-1][1]
    frames = synthetic_gop(2176, 3840, gop=2, gop_id=7)[:, 0]
    gold = _gold_from_oracle(sd, frames[1:2], frames[0:1], levels=levels)
    rep = _check_against(m, gold, dev, masked_free_run=True)
    _assert_full_size_latent_rates(rep, gold)
    print("4K L=%d parity report:" % levels, {k: float("%.3g" % v) for k, v in rep.items()})
    m.release()


@pytest.mark.parametrize("impl_name,impl", _impls())
def test_deeper_pyramid_matches_reference_golden_L6(dev, golden_pframe_L6_256, impl_name, impl):
    """SpyNet with 6 pyramid levels (configs[3] "deeper flow pyramid") against the reference class with ``self.L``
    patched to 6 and modelL5/modelL6 modules appended (oracle/gen_golden_r2.py::build_reference_model_levels)."""
    from fastvideocodec_b200 import VideoCompressor
    from fastvideocodec_b200.synthetic import init_state_dict
    g = golden_pframe_L6_256
    m = VideoCompressor(spynet_levels=6)
    m.load_state_dict(init_state_dict(0, spynet_levels=6, spynet_gain=1.8))
    m = m.to(dev).eval()
    m.impl = impl
    with torch.no_grad():
        out = m(g["cur"].to(dev), g["ref"].to(dev))
    for name in LATENTS:
        check_latent(name, m.get_intermediate(name).cpu(), g[name], g[PREQUANT[name]])
    for name in ("estmv", "mvfeature"):
        err = (m.get_intermediate(name).cpu() - g[name]).abs().max().item()
        assert err <= 5e-4 * max(1.0, g[name].abs().max().item()), (name, err)
    for i, n in enumerate(SCALAR_NAMES, start=1):
        assert abs(float(out[i]) - float(g[n])) <= 0.005 * abs(float(g[n])), n
    assert abs(_psnr(out[1]) - _psnr(g["mse"])) <= 0.02
    rec = m.decode_from_latents(g["ref"].to(dev), g["quant_mv"].to(dev), g["feat_hat"].to(dev))
    assert (rec.cpu() - g["clipped"]).abs().max().item() <= 1e-3
    m.release()


def test_config5_multiview_batch8(model, state_dict, dev):
    """BASELINE configs[4]: 8 camera views of 1280x720 (padded to 768) folded into the batch
    (train_multiview.py:232-233).  Views 0 and 7 of the B=8 run against the CPU oracle run on that view alone
    (views are independent samples: DVC has no cross-view op), and B=8 equals B=1 runs bit for bit."""
    from fastvideocodec_b200.synthetic import synthetic_gop
    model.impl = _impls()[-1][1]
    fr_host = synthetic_gop(768, 1280, gop=2, gop_id=11, batch=8)   # [2, 8, 3, H, W]
    fr = fr_host.to(dev)
    with torch.no_grad():
        both = model(fr[1], fr[0])
    lat = {n: model.get_intermediate(n).cpu() for n in LATENTS}
    rec8 = both[0].cpu()
    assert math.isfinite(float(both[7])) and float(both[0].min()) >= 0.0 and float(both[0].max()) <= 1.0
    bpps = {}
    for v in (0, 7):
        gold = _gold_from_oracle(state_dict, fr_host[1, v:v + 1], fr_host[0, v:v + 1])
        mask = torch.zeros((768, 1280), dtype=torch.bool)
        for n, scale in (("quant_mv", 16), ("feat_hat", 16), ("z_hat", 64)):
            a = lat[n][v:v + 1]
            check_latent(n, a, gold[n], gold[PREQUANT[n]], tie_rule=(n == "quant_mv"))
            if n == "feat_hat":
                _mask_feat_flips(mask, a != gold[n])
        assert mask.float().mean().item() <= 0.5
        assert (rec8[v:v + 1] - gold["clipped"]).abs().masked_fill(mask, 0.0).max().item() <= 1e-2
        # the same view as a B=1 run: bit-identical frame, and every gate of _check_against against the oracle
        _check_against(model, gold, dev, masked_free_run=True)
        with torch.no_grad():
            one = model(fr[1, v:v + 1], fr[0, v:v + 1])
        assert torch.equal(both[0][v:v + 1], one[0])
        bpps[v] = float(one[7])
    with torch.no_grad():
        mid = model(fr[1, 3:4], fr[0, 3:4])
    assert torch.equal(both[0][3:4], mid[0])
    model.release()


def _randomize_eb(module, gen):
    """Non-trivial EntropyBottleneck parameters (fresh init has zero factors and symmetric matrices)."""
    with torch.no_grad():
        for name, p in module.named_parameters():
            if "_matrix" in name or "_factor" in name:
                p.add_(torch.randn(p.shape, generator=gen) * 0.3)
            elif "quantiles" in name:
                p[:, 0, 1] = torch.randn(p.shape[0], generator=gen) * 0.4


def test_entropy_models_recprob_forward(dev):
    """entropy_models.RecProbModel.forward / get_estimate_bits (entropy_models.py:55-78), RPM_flag False, against
    the oracle's restatement of the CompressAI algorithm (parity UNPINNED: CompressAI is not vendored)."""
    from fastvideocodec_b200.entropy_models import RecProbModel
    g = torch.Generator().manual_seed(21)
    m = RecProbModel(128)
    _randomize_eb(m.entropy_bottleneck, g)
    x = torch.randn((2, 128, 17, 30), generator=g) * 4
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    want_xh, want_lik, want_prior = O.recprob_forward(sd, x)
    m = m.to(dev).eval()
    m.set_RPM(False)
    hidden = torch.zeros(1)
    with torch.no_grad():
        xh, lik, hid, prior = m(x.to(dev), hidden, training=False)
    assert torch.equal(xh.cpu(), want_xh) and torch.equal(prior.cpu(), want_prior)
    assert (lik.cpu() - want_lik).abs().max().item() <= 2e-6
    bits, want_bits = float(m.get_estimate_bits(lik)), float(O.estimate_bits_clamped(want_lik))
    assert abs(bits - want_bits) <= 1e-5 * want_bits
    # fused kernel: the same bits without materialising the reduction in PyTorch
    _, _, fused = m.entropy_bottleneck.forward_bits(x.to(dev))
    assert abs(float(fused) - want_bits) <= 1e-5 * want_bits


def test_entropy_models_meanscale_forward(dev):
    """entropy_models.MeanScaleHyperPriors.forward / get_estimate_bits (entropy_models.py:202-235): hyper convs on
    the tcgen05 engine, EntropyBottleneck + GaussianConditional kernels (parity UNPINNED, see above)."""
    from fastvideocodec_b200.entropy_models import MeanScaleHyperPriors
    g = torch.Generator().manual_seed(22)
    C = 64
    m = MeanScaleHyperPriors(C)
    _randomize_eb(m.entropy_bottleneck, g)
    with torch.no_grad():
        for name, p in m.named_parameters():
            if name.startswith("h_"):
                p.copy_(torch.randn(p.shape, generator=g) * (0.04 if p.dim() == 4 else 0.1))
    x = torch.randn((1, C, 32, 48), generator=g) * 3
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    want_xh, want_xl, want_zl, want_z, want_sigma, want_mu = O.meanscale_forward(sd, x, C)
    m = m.to(dev).eval()
    with torch.no_grad():
        xh, (xl, zl) = m(x.to(dev), training=False)
    assert (m.z.cpu() - want_z).abs().max().item() <= 5e-5 * max(1.0, want_z.abs().max().item())
    # Everything after the hyper-latent quantiser is compared outside the receptive field (4 convs 3x3 -> +-4 px)
    # of hyper-latents that sat on a rounding tie (fp32 evaluation order moves those across the tie).
    med = sd["entropy_bottleneck.quantiles"][:, 0, 1].view(1, C, 1, 1)
    flips = (torch.round(m.z.cpu() - med) != torch.round(want_z - med)).any(dim=1, keepdim=True).float()
    assert flips.sum().item() <= 2
    keep = torch.nn.functional.max_pool2d(flips, 9, 1, 4) == 0

    def err(a, b):
        return ((a.cpu() - b).abs() * keep).max().item()

    assert err(m.sigma, want_sigma) <= 1e-4 * want_sigma.abs().max().item()
    assert err(m.mu, want_mu) <= 1e-4 * max(1.0, want_mu.abs().max().item())
    assert err(zl, want_zl) <= 1e-5
    # x_hat = round(x - mu) + mu: identical up to mu's rounding noise except at ties of (x - mu)
    d = (xh.cpu() - want_xh).abs() * keep
    assert (d > 1e-3).float().mean().item() <= 1e-4 and d.max().item() <= 1.0 + 1e-3
    same = keep & (d <= 1e-3)
    assert ((xl.cpu() - want_xl).abs() * same).max().item() <= 1e-4
    if flips.sum().item() == 0 and (d > 1e-3).sum().item() == 0:
        bits, want_bits = float(m.get_estimate_bits((xl, zl))[0]), float(O.meanscale_bits(want_xl, want_zl)[0])
        assert abs(bits - want_bits) <= 0.005 * want_bits


@pytest.mark.parametrize("tag,name", [("tree", "LSVC-128"), ("chain", "LSVC-L-128"), ("onehop", "LSVC-O-128")])
def test_lsvc_forward_matches_reference_golden(dev, state_dict, tag, name):
    """SURVEY 8f N1 — LSVC batched / tree GOP forward (models.py:1344-1411) through fvc_lsvc_mv_forward /
    fvc_lsvc_mc_res_forward against the unmodified reference's outputs (tests/golden/lsvc_64.npz)."""
    from conftest import load_golden
    from fastvideocodec_b200.lsvc import LSVC
    gold = load_golden("lsvc_64.npz")
    m = LSVC(name)
    m.load_state_dict(state_dict)
    m = m.to(dev).eval()
    with torch.no_grad():
        out = m(gold["x"].to(dev))
    # warped / MC frames of the first tree layer depend on no quantiser of this GOP other than the MV latents
    for i, n in enumerate(["com", "mc", "warped"]):
        d = (out[i].cpu() - gold["%s_%s" % (tag, n)]).abs()
        # metric-level closed-loop gate (a tie flip in a parent frame moves its children)
        assert d.mean().item() <= 2e-3, (n, d.mean().item())
    for i, n in enumerate(["rec_loss", "warp_loss", "mc_loss", "bpp_res", "bpp"], start=3):
        a, b = float(out[i]), float(gold["%s_%s" % (tag, n)])
        assert abs(a - b) <= 0.005 * abs(b), (n, a, b)
    # GOP driver, LSVC branch (models.py:384-398)
    from fastvideocodec_b200 import parallel_compression
    with torch.no_grad():
        pc = parallel_compression(None, m, gold["x"].to(dev))
    x = gold["x"][1:]
    ref_psnr = [float(10 * torch.log10(1 / torch.mean((gold["%s_com" % tag][i] - x[i]) ** 2))) for i in range(4)]
    assert pc[0].shape == (5, 3, 64, 64) and all(abs(a - b) <= 0.02 for a, b in zip(pc[6], ref_psnr))
    assert abs(pc[3] - float(gold["%s_bpp" % tag])) <= 0.005 * float(gold["%s_bpp" % tag])
    m.release()


def test_lsvc_tree_equals_dvc_building_blocks(dev, state_dict):
    """Element level: LSVC with the linear graph on a 2-frame clip is one DVC P-frame whose flow is estimated
    against the ORIGINAL reference — identical to VideoCompressor.forward when reference == original (frame 1)."""
    from fastvideocodec_b200 import VideoCompressor
    from fastvideocodec_b200.lsvc import LSVC
    from fastvideocodec_b200.synthetic import synthetic_gop
    x = synthetic_gop(128, 192, gop=2, gop_id=6)[:, 0].to(dev)
    l = LSVC("LSVC-L-128")
    l.load_state_dict(state_dict)
    l = l.to(dev).eval()
    v = VideoCompressor()
    v.load_state_dict(state_dict)
    v = v.to(dev).eval()
    with torch.no_grad():
        lo = l(x)
        vo = v(x[1:2], x[0:1])
    assert torch.equal(lo[0], vo[0])                        # clipped reconstruction, bit for bit
    assert abs(float(lo[7]) - float(vo[7])) <= 1e-6 * float(vo[7])
    assert abs(float(lo[5]) - float(vo[3])) <= 1e-6 * float(vo[3]) and abs(float(lo[4]) - float(vo[2])) <= 1e-6 * float(vo[2])
    l.release()
    v.release()


@pytest.mark.parametrize("inverse", [False, True])
def test_compressai_gdn_layer(dev, inverse):
    """layers.GDN (CompressAI names) against the published formula in torch fp32 (tolerance: 2e-6 absolute + 2e-6 relative,
    i.e. fp32 rounding of the 22-bit hi/lo activation records and of the 64-term norm)."""
    from fastvideocodec_b200.layers import GDN
    torch.manual_seed(5)
    g = GDN(64, inverse=inverse)
    with torch.no_grad():
        g.beta.add_(0.3 * torch.rand(64))
        g.gamma.add_(0.05 * torch.rand(64, 64))
    x = torch.randn(2, 64, 24, 40)
    ped = 2.0 ** -36
    beta = torch.clamp(g.beta.detach(), min=(1e-6 + ped) ** 0.5) ** 2 - ped
    gamma = torch.clamp(g.gamma.detach(), min=ped ** 0.5) ** 2 - ped
    norm = torch.nn.functional.conv2d(x * x, gamma.view(64, 64, 1, 1), beta)
    want = x * (torch.sqrt(norm) if inverse else torch.rsqrt(norm))
    with torch.no_grad():
        got = g.to(dev)(x.to(dev)).cpu()
    assert ((got - want).abs() - 2e-6 * want.abs()).max().item() <= 2e-6


@pytest.mark.parametrize("case", [(64, 64, 3, 1, 0, 1, 33, 45), (128, 128, 3, 2, 1, 2, 9, 15), (8, 32, 7, 1, 0, 1, 24, 40),
                                  (64, 32, 7, 1, 0, 1, 16, 24)])
def test_conv2d_pair_mode_bit_identical_to_single_cta(dev, case, monkeypatch):
    """The CTA-pair engine (tcgen05 cta_group::2, M = 256) must give bit-identical results to the single-CTA
    engine: same accumulation groups and order, only the tile-to-SM mapping differs."""
    from fastvideocodec_b200 import ops
    cin, cout, k, stride, transposed, act, H, W = case
    torch.manual_seed(11)
    x = torch.randn(2, cin, H, W)
    w = torch.randn((cin, cout, k, k) if transposed else (cout, cin, k, k)) / (cin * k * k) ** 0.5
    b = torch.randn(cout)
    f = ops.conv_transpose2d if transposed else ops.conv2d
    monkeypatch.setenv("FVC_TC_PAIR", "1")
    y_pair = f(x.to(dev), w.to(dev), b.to(dev), stride, act).cpu()
    monkeypatch.setenv("FVC_TC_PAIR", "0")
    y_single = f(x.to(dev), w.to(dev), b.to(dev), stride, act).cpu()
    assert torch.equal(y_pair, y_single)


# ------------------------------------------------------------------------------------------------
# round 2: the reference's real SpyNet weights, HD closed loop against the reference, decoder-only, drop-in surface
# ------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def real_model(dev, real_state_dict):
    from fastvideocodec_b200 import VideoCompressor
    m = VideoCompressor()
    m.load_state_dict(real_state_dict)
    return m.to(dev).eval()


@pytest.mark.parametrize("impl_name,impl", _impls())
def test_real_spynet_weights_match_reference_golden(real_model, golden_pframe_real_128, dev, impl_name, impl):
    """The reference's own pretrained SpyNet weights (|w|max ~ 5: the dynamic range the fp16 hi/lo operand pairs must
    hold) through the CUDA path, against the unmodified reference run with them (pframe_real_128.npz)."""
    real_model.impl = impl
    _check_against(real_model, golden_pframe_real_128, dev)
    assert real_model.saturation_count() == 0


@pytest.mark.parametrize("size", [(256, 256), (1088, 1920)])
def test_real_spynet_weights_match_oracle(real_model, real_state_dict, dev, size):
    """Real SpyNet weights at 256x256 and at 1088x1920 (open loop) against the CPU oracle, which is pinned against
    the live reference with these weights (tests/test_oracle_golden.py)."""
    from fastvideocodec_b200.synthetic import synthetic_gop
    real_model.impl = _impls()[-1][1]
    frames = synthetic_gop(size[0], size[1], gop=2, gop_id=23)[:, 0]
    gold = _gold_from_oracle(real_state_dict, frames[1:2], frames[0:1])
    rep = _check_against(real_model, gold, dev, masked_free_run=size[0] >= 1088)
    if size[0] >= 1088:
        _assert_full_size_latent_rates(rep, gold)
    assert real_model.saturation_count() == 0
    print("real SpyNet weights %dx%d:" % size, {k: float("%.3g" % v) for k, v in rep.items()})
    real_model.release()


@pytest.mark.parametrize("gold_name", ["golden_pframe_64", "golden_pframe_128", "golden_pframe_real_128"])
def test_decoder_only_from_reference_latents(request, dev, state_dict, real_state_dict, gold_name):
    """SURVEY 7.2-1 caveat (i): decoder-only reconstruction.  The reference's quantised latents go into
    fvc_decode_from_latents (mvDecoder, motion compensation, resDecoder, clamp); the frame must match the
    reference's within 1e-2 EVERYWHERE, no mask (measured: ~1e-5)."""
    from fastvideocodec_b200 import VideoCompressor
    g = request.getfixturevalue(gold_name)
    m = VideoCompressor()
    m.load_state_dict(real_state_dict if "real" in gold_name else state_dict)
    m = m.to(dev).eval()
    for _, impl in _impls():
        m.impl = impl
        rec = m.decode_from_latents(g["ref"].to(dev), g["quant_mv"].to(dev), g["feat_hat"].to(dev))
        err = (rec.cpu() - g["clipped"]).abs().max().item()
        assert err <= 1e-3, (gold_name, impl, err)      # north star: 1e-2
        for name in ("mv_hat", "warpframe", "prediction", "recon_res"):
            e = (m.get_intermediate(name).cpu() - g[name]).abs().max().item()
            assert e <= 5e-4 * max(1.0, g[name].abs().max().item()), (name, e)
    m.release()


def test_decoder_only_hd(model, state_dict, dev):
    """Decoder-only at 1088x1920: oracle latents in, reconstruction within 1e-2 everywhere (no mask)."""
    from fastvideocodec_b200.synthetic import synthetic_gop
    frames = synthetic_gop(1088, 1920, gop=2, gop_id=5)[:, 0]
    gold = _gold_from_oracle(state_dict, frames[1:2], frames[0:1])
    model.impl = _impls()[-1][1]
    rec = model.decode_from_latents(frames[0:1].to(dev), gold["quant_mv"].to(dev), gold["feat_hat"].to(dev))
    err = (rec.cpu() - gold["clipped"]).abs().max().item()
    assert err <= 1e-3, err


def _sparse_tie_check(name, got_q, gold, pre_name, tie_rule=True, extra=0):
    """check_latent against hd_gop10.npz: int8 latents + the reference's pre-round values near ties (sparse)."""
    want = gold["f1_" + name].float()
    diff = (got_q != want)
    n_bad = int(diff.sum())
    assert n_bad <= max(2, int(1e-4 * want.numel())) + extra, (name, n_bad, extra)
    if n_bad == 0:
        return 0.0
    assert float((got_q - want)[diff].abs().max()) == 1.0
    if not tie_rule:
        return n_bad / want.numel()
    idx = diff.flatten().nonzero().flatten()
    tie_idx = gold["f1_%s_tie_idx" % pre_name].long()
    tie_val = gold["f1_%s_tie_val" % pre_name]
    pos = torch.searchsorted(tie_idx, idx)
    assert bool((pos < tie_idx.numel()).all()) and bool((tie_idx[pos] == idx).all()), \
        (name, "a flipped element is further than 2e-3 from a rounding tie in the reference")
    frac = tie_val[pos] - torch.floor(tie_val[pos])
    tol = TIE_TOL * max(1.0, float(gold["f1_%s_rms" % pre_name]))
    assert float((frac - 0.5).abs().max()) <= tol, (name, float((frac - 0.5).abs().max()), tol)
    return n_bad / want.numel()


def test_hd_gop10_closed_loop_matches_reference_golden(model, golden_hd_gop10, dev):
    """The GOP bench.py times (synthetic_gop(1088, 1920, gop=10, gop_id=0), init_state_dict(0)), closed loop, against
    the UNMODIFIED reference's own rows for it (tests/golden/hd_gop10.npz: 7 scalars + PSNR x 9 P-frames).
    Frame 1 (open loop) element level: latents vs the reference's (count, +-1, tie rule), reconstruction <= 1e-2
    outside the receptive field of flips.  Closed loop: GOP means bpp 0.5 % / PSNR 0.02 dB (north star), per frame
    1 % / 0.04 dB (a tie flip in frame t perturbs x_prev of frames t+1.. locally)."""
    from fastvideocodec_b200.synthetic import synthetic_gop
    g = golden_hd_gop10
    model.impl = _impls()[-1][1]
    frames = synthetic_gop(1088, 1920, gop=10, gop_id=int(g["gop_id"]))          # [10,1,3,H,W]
    with torch.no_grad():
        out = model(frames[1].to(dev), frames[0].to(dev))
    mask = torch.zeros((1088, 1920), dtype=torch.bool)
    fracs = {}
    n_mv = 0
    for name in LATENTS:                      # quant_mv first
        a = model.get_intermediate(name).cpu()
        fracs[name] = _sparse_tie_check(name, a, g, PREQUANT[name], tie_rule=(name == "quant_mv"),
                                        extra=0 if name == "quant_mv" else n_mv)
        if name == "quant_mv":
            n_mv = int((a != g["f1_quant_mv"].float()).sum())
        if name == "feat_hat":
            _mask_feat_flips(mask, a != g["f1_" + name].float())
    assert mask.float().mean().item() <= 0.5
    want = torch.from_numpy(g["f1_clipped_u16"].numpy().astype("float32")) / 65535.0
    err = (out[0].cpu() - want).abs().masked_fill(mask, 0.0).max().item()
    assert err <= 1e-2, err
    rows = g["rows"]
    for i in range(7):
        assert abs(float(out[1 + i]) - rows[0, i].item()) <= 0.005 * abs(rows[0, i].item()), (i, float(out[1 + i]))
    # teacher forced with the reference's own latents: our feature / z quantisers (tie rule) and the whole frame
    # against the reference's, no mask (<= 1e-3 + the u16 step of the fixture; north star 1e-2)
    forced = [g["f1_" + n].float().to(dev).contiguous() for n in LATENTS]
    model.force_latents(1, 1088, 1920, dev, *forced)
    try:
        with torch.no_grad():
            fout = model(frames[1].to(dev), frames[0].to(dev))
        for name in ("feat_hat", "z_hat"):
            q = torch.round(model.get_intermediate(PREQUANT[name]).cpu())
            fracs["forced_" + name] = _sparse_tie_check(name, q, g, PREQUANT[name])
        ferr = (fout[0].cpu() - want).abs().max().item()
        assert ferr <= 1e-3 + 1e-5, ferr
    finally:
        model.force_latents(1, 1088, 1920, dev)
    # closed loop through the host entry point (what bench.py's e2e leg calls)
    _, sc = model.gop_forward_host(frames.contiguous().pin_memory(), want_recon=False)
    got_bpp, want_bpp = float(sc[:, 6].double().mean()), rows[:, 6].mean().item()
    got_psnr = sum(_psnr(m) for m in sc[:, 0].tolist()) / 9
    want_psnr = rows[:, 7].mean().item()
    assert abs(got_bpp - want_bpp) <= 0.005 * want_bpp, (got_bpp, want_bpp)
    assert abs(got_psnr - want_psnr) <= 0.02, (got_psnr, want_psnr)
    for i in range(9):
        assert abs(float(sc[i, 6]) - rows[i, 6].item()) <= 0.01 * rows[i, 6].item(), i
        assert abs(_psnr(sc[i, 0]) - rows[i, 7].item()) <= 0.04, i
    print("HD GOP-10 closed loop: bpp %.5f vs %.5f, PSNR %.4f vs %.4f dB, frame-1 flips %s" %
          (got_bpp, want_bpp, got_psnr, want_psnr, fracs))


def test_saturation_is_reported(dev, state_dict):
    """Activations beyond the fp16 operand-pair range (|v| >= 65504) are clamped by the epilogue; the reference is
    fp32, so that must surface as an error: scalars turn NaN, the counter is non-zero, the GOP entry point raises."""
    from fastvideocodec_b200 import VideoCompressor, _lib
    from fastvideocodec_b200.synthetic import synthetic_gop
    sd = {k: v.clone() for k, v in state_dict.items()}
    sd["warpnet.feature_ext.weight"] = sd["warpnet.feature_ext.weight"] * 1e6
    m = VideoCompressor()
    m.load_state_dict(sd)
    m = m.to(dev).eval()
    m.impl = _impls()[-1][1]
    fr = synthetic_gop(64, 64, gop=2, gop_id=0)
    with torch.no_grad():
        out = m(fr[1].to(dev), fr[0].to(dev))
    assert all(math.isnan(float(v)) for v in out[1:])
    assert m.saturation_count() > 0
    with pytest.raises(_lib.FvcError):
        m.gop_forward_host(fr.contiguous())
    assert m.saturation_count(reset=True) > 0 and m.saturation_count() == 0
    m.release()


def test_gop_forward_host_rejects_bad_input(model, dev):
    with pytest.raises(TypeError):
        model.gop_forward_host(torch.zeros((2, 1, 3, 64, 64), device=dev))
    with pytest.raises(TypeError):
        model.gop_forward_host(torch.zeros((2, 1, 3, 64, 64), dtype=torch.float64))
    with pytest.raises(TypeError):
        model.gop_forward_host(torch.zeros((2, 1, 3, 64, 128))[..., ::2])
    with pytest.raises(ValueError):
        model.gop_forward_host(torch.zeros((1, 1, 3, 64, 64)))
    with pytest.raises(ValueError):
        model.gop_forward_host(torch.zeros((2, 1, 3, 60, 64)))


def test_context_cache_is_bounded(dev, state_dict):
    from fastvideocodec_b200 import VideoCompressor
    m = VideoCompressor()
    m.load_state_dict(state_dict)
    m = m.to(dev).eval()
    m.max_contexts = 2
    x = torch.rand((1, 3, 64, 192), device=dev)
    with torch.no_grad():
        first = m(x[..., :64], x[..., :64])[0].clone()
        m(x[..., :128], x[..., :128])
        m(x, x)
        assert len(m._ctxs) == 2
        again = m(x[..., :64], x[..., :64])[0]        # evicted context is rebuilt: same result
    assert torch.equal(first, again) and len(m._ctxs) == 2
    m.release()


def test_conv_op_handles_are_cached_and_follow_weight_updates(dev):
    """ops.conv2d keeps the packed weights / plan of a layer across calls (fvc_conv_op_*): the same call twice reuses
    one handle and returns identical bits; an in-place weight update (version bump) or a new input shape builds a new
    handle whose result follows the new weights; the cache is bounded; the one-shot C entry point agrees bit for bit."""
    import ctypes as C
    import torch.nn.functional as F
    from fastvideocodec_b200 import ops
    from fastvideocodec_b200._lib import check, lib, ptr, stream_ptr
    ops.conv_op_cache_clear()
    g = torch.Generator().manual_seed(77)
    x = torch.randn((1, 64, 16, 24), generator=g).to(dev)
    w = (torch.randn((64, 64, 3, 3), generator=g) / 24.0).to(dev)
    b = (torch.randn((64,), generator=g) * 0.1).to(dev)
    y0 = ops.conv2d(x, w, b, 1, ops.ACT_RELU)
    assert ops.conv_op_cache_size() == 1
    x2 = x * 0.5
    y1 = ops.conv2d(x2, w.detach(), b, 1, ops.ACT_RELU)           # detach(): same memory, same version -> same handle
    assert ops.conv_op_cache_size() == 1
    assert torch.equal(ops.conv2d(x, w, b, 1, ops.ACT_RELU), y0)
    want1 = torch.relu(F.conv2d(x2.cpu(), w.cpu(), b.cpu(), padding=1))
    assert (y1.cpu() - want1).abs().max().item() <= 1e-4 * max(1.0, want1.abs().max().item())
    # the one-shot entry point (create + run + destroy): same bits
    y_once = torch.empty_like(y0)
    check(lib().fvc_conv2d(ptr(x), ptr(w), ptr(b), ptr(y_once), 1, 64, 16, 24, 64, 3, 1, 0, ops.ACT_RELU, ops.IMPL_TC,
                           stream_ptr()), "fvc_conv2d")
    assert torch.equal(y_once, y0)
    # in-place update of the weights and of the bias: new handles, new results
    w.mul_(-1.0)
    y2 = ops.conv2d(x, w, b, 1, ops.ACT_RELU)
    assert ops.conv_op_cache_size() == 2
    want2 = torch.relu(F.conv2d(x.cpu(), w.cpu(), b.cpu(), padding=1))
    assert (y2.cpu() - want2).abs().max().item() <= 1e-4 * max(1.0, want2.abs().max().item())
    b.add_(1.0)
    y3 = ops.conv2d(x, w, b, 1, ops.ACT_RELU)
    want3 = torch.relu(F.conv2d(x.cpu(), w.cpu(), b.cpu(), padding=1))
    assert (y3.cpu() - want3).abs().max().item() <= 1e-4 * max(1.0, want3.abs().max().item())
    # another input shape: its own handle; the cache stays bounded
    ops.conv2d(x[..., :16].contiguous(), w, b, 1, ops.ACT_RELU)
    assert ops.conv_op_cache_size() == 4
    cap = ops._conv_ops.capacity
    try:
        ops._conv_ops.capacity = 2
        ops.conv2d(x[..., :8].contiguous(), w, b, 1, ops.ACT_RELU)
        assert ops.conv_op_cache_size() == 2
    finally:
        ops._conv_ops.capacity = cap
    # C ABI argument checks of the handle API
    h = C.c_void_p()
    assert lib().fvc_conv_op_create(C.byref(h), ptr(w), ptr(b), 1, 64, 16, 24, 64, 4, 1, 0, 0, ops.IMPL_TC, stream_ptr()) != 0
    assert lib().fvc_conv_op_run(None, ptr(x), ptr(y0), stream_ptr()) != 0
    ops.conv_op_cache_clear()
    assert ops.conv_op_cache_size() == 0


def test_subnet_module_forwards_match_oracle(model, state_dict, dev):
    """The drop-in module surface reference models.py classes call (sub-module forwards, motioncompensation,
    BitEstimator closures) against the oracle's restatement of each module, element-wise."""
    from fastvideocodec_b200.synthetic import synthetic_gop
    sd = state_dict
    fr = synthetic_gop(128, 192, gop=2, gop_id=13)[:, 0]
    cur, ref = fr[1:2], fr[0:1]
    g = torch.Generator().manual_seed(31)

    def close(a, b, tol=1e-4):
        err = (a.cpu() - b).abs().max().item()
        assert err <= tol * max(1.0, b.abs().max().item()), err

    with torch.no_grad():
        flow = O.me_spynet(sd, cur, ref)
        close(model.opticFlow(cur.to(dev), ref.to(dev)), flow)
        mvf = O.analysis_mv(sd, flow)
        close(model.mvEncoder(flow.to(dev)), mvf)
        q = torch.round(mvf)
        mvh = O.synthesis_mv(sd, q)
        close(model.mvDecoder(q.to(dev)), mvh)
        pred, warp = model.motioncompensation(ref.to(dev), mvh.to(dev))
        wo = O.flow_warp(ref, mvh)
        po = O.warp_net(sd, torch.cat((wo, ref), 1)) + wo
        close(warp, wo)
        close(pred, po)
        close(model.warpnet(torch.cat((wo, ref), 1).to(dev)), po - wo)
        res = cur - po
        feat = O.analysis(sd, res)
        close(model.resEncoder(res.to(dev)), feat)
        z = O.analysis_prior(sd, feat)
        close(model.respriorEncoder(feat.to(dev)), z)
        zq = torch.round(z)
        close(model.respriorDecoder(zq.to(dev)), O.synthesis_prior(sd, zq))
        fq = torch.round(feat)
        close(model.resDecoder(fq.to(dev)), O.synthesis(sd, fq))
        x = torch.randn((1, 64, 5, 7), generator=g) * 3
        qz, bits = model.bitEstimator_z.quant_bits(x.to(dev))
        wb, _ = O.factorized_bits(sd, "bitEstimator_z", torch.round(x))
        assert torch.equal(qz.cpu(), torch.round(x)) and abs(float(bits) - float(wb)) <= 1e-5 * float(wb)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs in one process")
def test_two_devices_in_one_process(state_dict):
    """cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is per device: a second GPU used from the same process must
    launch the >48 KB shared-memory kernels too, and give the same bits."""
    from fastvideocodec_b200 import VideoCompressor
    from fastvideocodec_b200.synthetic import synthetic_gop
    fr = synthetic_gop(64, 128, gop=2, gop_id=0)
    outs = []
    for d in (0, 1):
        dev = torch.device("cuda", d)
        m = VideoCompressor()
        m.load_state_dict(state_dict)
        m = m.to(dev).eval()
        with torch.no_grad():
            o = m(fr[1].to(dev), fr[0].to(dev))
        torch.cuda.synchronize(dev)
        outs.append((o[0].cpu(), float(o[7])))
        m.release()
    assert torch.equal(outs[0][0], outs[1][0]) and outs[0][1] == outs[1][1]


# ------------------------------------------------------------------------------------------------
# precision='fast' (one fp16 MMA per product): metric-level gates only
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", CONV_CASES)
def test_conv2d_fast_precision(dev, case):
    """Single-pass fp16 operands, fp32 accumulation: error bounded by the operand rounding (2^-11 relative per
    operand, random signs over the K = Cin*k*k products)."""
    import torch.nn.functional as F
    from fastvideocodec_b200 import ops
    cin, cout, k, stride, transposed, act, H, W = case
    g = torch.Generator().manual_seed(300 + cin * 7 + cout + k + stride)
    x = torch.randn((2, cin, H, W), generator=g)
    wshape = (cin, cout, k, k) if transposed else (cout, cin, k, k)
    w = torch.randn(wshape, generator=g) / math.sqrt(cin * k * k)
    b = torch.randn((cout,), generator=g) * 0.1
    if transposed:
        want = F.conv_transpose2d(x, w, b, stride=stride, padding=k // 2, output_padding=stride - 1)
        got = ops.conv_transpose2d(x.to(dev), w.to(dev), b.to(dev), stride, act, ops.IMPL_TC_FAST).cpu()
    else:
        want = F.conv2d(x, w, b, stride=stride, padding=k // 2)
        got = ops.conv2d(x.to(dev), w.to(dev), b.to(dev), stride, act, ops.IMPL_TC_FAST).cpu()
    want = {0: lambda t: t, 1: torch.relu, 2: lambda t: F.leaky_relu(t, 0.1), 3: torch.exp}[act](want)
    err = (got - want).abs().max().item()
    assert err <= 4e-3 * max(1.0, want.abs().max().item()), (case, err)
    assert err > 0.0 or cin * k * k < 64      # it really is the reduced-precision path


@pytest.fixture(scope="module")
def fast_model(dev, state_dict):
    from fastvideocodec_b200 import VideoCompressor
    m = VideoCompressor(precision="fast")
    m.load_state_dict(state_dict)
    return m.to(dev).eval()


def test_fast_precision_config1_gop_metric_parity(fast_model, state_dict, dev):
    """precision='fast' on BASELINE config 1 (256x256, GOP 10, closed loop) against the CPU oracle: the metric-level
    north-star gates (bpp 0.5 %, PSNR 0.02 dB on the GOP means).  Latent flip rates are reported, not gated."""
    from fastvideocodec_b200.synthetic import synthetic_gop
    frames = synthetic_gop(256, 256, gop=10, gop_id=0)[:, 0]
    rows, rec = O.gop_forward(state_dict, frames)
    _, sc = fast_model.gop_forward_host(frames.unsqueeze(1).contiguous())
    bpp = sum(r[0] for r in rows) / len(rows)
    psnr = sum(r[1] for r in rows) / len(rows)
    got_bpp = float(sc[:, 6].mean())
    got_psnr = sum(_psnr(m) for m in sc[:, 0].tolist()) / len(rows)
    print("fast 256^2 GOP: bpp %.5f vs %.5f (%.3g), PSNR %.4f vs %.4f dB" % (got_bpp, bpp, abs(got_bpp - bpp) / bpp,
                                                                        got_psnr, psnr))
    assert abs(got_bpp - bpp) <= 0.005 * bpp
    assert abs(got_psnr - psnr) <= 0.02


def test_fast_precision_hd_gop_metric_parity(fast_model, golden_hd_gop10, dev):
    """precision='fast' on the GOP bench.py times, closed loop, against the unmodified reference's rows:
    bpp within 0.5 %, PSNR within 0.02 dB on the GOP means; frame-1 latent flip rates printed."""
    from fastvideocodec_b200.synthetic import synthetic_gop
    g = golden_hd_gop10
    frames = synthetic_gop(1088, 1920, gop=10, gop_id=int(g["gop_id"]))
    with torch.no_grad():
        fast_model(frames[1].to(dev), frames[0].to(dev))
    flips = {n: (fast_model.get_intermediate(n).cpu() != g["f1_" + n].float()).float().mean().item() for n in LATENTS}
    _, sc = fast_model.gop_forward_host(frames.contiguous().pin_memory(), want_recon=False)
    rows = g["rows"]
    got_bpp, want_bpp = float(sc[:, 6].double().mean()), rows[:, 6].mean().item()
    got_psnr = sum(_psnr(m) for m in sc[:, 0].tolist()) / 9
    want_psnr = rows[:, 7].mean().item()
    print("fast HD GOP: bpp %.5f vs %.5f (%.3g), PSNR %.4f vs %.4f dB, frame-1 latent flip rates %s" %
          (got_bpp, want_bpp, abs(got_bpp - want_bpp) / want_bpp, got_psnr, want_psnr, flips))
    assert abs(got_bpp - want_bpp) <= 0.005 * want_bpp
    assert abs(got_psnr - want_psnr) <= 0.02
    fast_model.release()


def test_torch_library_ops_run_the_library(model, golden_pframe_64, dev):
    """torch.ops.fvc.* (torch_ops.py) are the same calls as the module / ops surface."""
    from fastvideocodec_b200 import ops
    g = golden_pframe_64
    model.impl = _impls()[-1][1]
    with torch.no_grad():
        out = model(g["cur"].to(dev), g["ref"].to(dev))
        ctx = model._last_ctx
        recon, sc = torch.ops.fvc.pframe_forward(g["cur"].to(dev), g["ref"].to(dev), ctx.handle)
    assert torch.equal(recon, out[0]) and torch.equal(sc, torch.stack(out[1:]))
    rec2 = torch.ops.fvc.decode_from_latents(g["ref"].to(dev), g["quant_mv"].to(dev), g["feat_hat"].to(dev), ctx.handle)
    assert (rec2.cpu() - g["clipped"]).abs().max().item() <= 1e-3
    img, flow = torch.rand((1, 3, 32, 48), device=dev), torch.randn((1, 2, 32, 48), device=dev)
    assert torch.equal(torch.ops.fvc.flow_warp(img, flow), ops.flow_warp(img, flow))
    w, b = torch.randn((16, 8, 3, 3), device=dev) * 0.1, torch.zeros(16, device=dev)
    x = torch.randn((1, 8, 16, 24), device=dev)
    assert torch.equal(torch.ops.fvc.conv2d(x, w, b, 1, False, 1, 1), ops.conv2d(x, w, b, 1, 1))


@pytest.mark.parametrize("case", CONV_CASES + [(64, 64, 3, 1, 0, 1, 40, 72), (128, 128, 3, 2, 1, 2, 20, 36),
                                               (128, 128, 3, 1, 0, 2, 34, 50), (6, 64, 3, 1, 0, 1, 30, 44)])
@pytest.mark.parametrize("layout", [1, 2])
def test_tma_store_epilogue_bit_identical_to_lane_stores(dev, case, layout, monkeypatch):
    """The TMA-store epilogue (ACT tile staged in shared memory, cp.async.bulk.tensor stores; plain, parity-planar and
    stride-2-phase outputs) must write exactly the bytes of the per-lane st.global epilogue: same arithmetic, only
    the way the records reach memory differs.  Outputs go through ACT records (FVC_CONV2D_VIA_ACT), as in the frame
    pipeline; image sizes include partial tiles at both edges."""
    from fastvideocodec_b200 import ops
    cin, cout, k, stride, transposed, act, H, W = case
    if cout < 8:
        pytest.skip("2-3 output channels: fp32 outputs only in the pipeline")
    torch.manual_seed(17)
    x = torch.randn(2, cin, H, W)
    w = torch.randn((cin, cout, k, k) if transposed else (cout, cin, k, k)) / (cin * k * k) ** 0.5
    b = torch.randn(cout)
    f = ops.conv_transpose2d if transposed else ops.conv2d
    monkeypatch.setenv("FVC_CONV2D_VIA_ACT", str(layout))
    monkeypatch.setenv("FVC_TC_TMAST", "2")
    y_tma = f(x.to(dev), w.to(dev), b.to(dev), stride, act).cpu()
    monkeypatch.setenv("FVC_TC_TMAST", "0")
    y_lane = f(x.to(dev), w.to(dev), b.to(dev), stride, act).cpu()
    monkeypatch.delenv("FVC_CONV2D_VIA_ACT")
    y_f32 = f(x.to(dev), w.to(dev), b.to(dev), stride, act).cpu()
    assert torch.equal(y_tma, y_lane), (y_tma - y_lane).abs().max().item()
    assert (y_lane - y_f32).abs().max().item() <= 2e-6 * max(1.0, y_f32.abs().max().item())


def test_fused_gdn_bit_identical_to_separate_norm_convolution(dev, state_dict, monkeypatch):
    """(I)GDN fused into the producing convolution's epilogue (second MMA on the squared tile) against the round-1
    form (raw + squared ACT tensors, separate 1x1 "norm" convolution launch): same tiles, same MMA order, same
    rounding of x -> bit-identical frames, latents and scalars; six launches fewer per frame."""
    from fastvideocodec_b200 import VideoCompressor
    from fastvideocodec_b200.synthetic import synthetic_gop
    fr = synthetic_gop(192, 320, gop=2, gop_id=8).to(dev)       # partial tiles at every resolution of the residual codec
    res = {}
    for fused in ("1", "0"):
        monkeypatch.setenv("FVC_GDN_FUSED", fused)
        m = VideoCompressor()
        m.load_state_dict(state_dict)
        m = m.to(dev).eval()
        m.impl = _impls()[-1][1]
        with torch.no_grad():
            out = m(fr[1], fr[0])
        res[fused] = (out, m.get_intermediate("feature"), m.get_intermediate("recon_res"), m.launch_count())
        m.release()
    a, b = res["1"], res["0"]
    assert torch.equal(a[0][0], b[0][0]) and all(float(x) == float(y) for x, y in zip(a[0][1:], b[0][1:]))
    assert torch.equal(a[1], b[1]) and torch.equal(a[2], b[2])
    print("launches fused %d vs separate %d" % (a[3], b[3]))


def test_fused_tails_match_separate_convolutions(dev, state_dict, monkeypatch):
    """mvDecoder.deconv7->deconv8 and Warp_net conv5.conv2->conv6 fused (the wide tensor stays in the SM: second MMA
    against the tail's weights regrouped per (tap, channel), then the 9-tap sum) against the separate 3x3 launches.
    Same products from the same 22-bit rounded y; only the fp32 summation order over the 9 taps differs."""
    from fastvideocodec_b200 import VideoCompressor
    from fastvideocodec_b200.synthetic import synthetic_gop
    fr = synthetic_gop(192, 320, gop=2, gop_id=8).to(dev)
    res = {}
    for fused in ("1", "0"):
        monkeypatch.setenv("FVC_TAIL_FUSED", fused)
        m = VideoCompressor()
        m.load_state_dict(state_dict)
        m = m.to(dev).eval()
        m.impl = _impls()[-1][1]
        with torch.no_grad():
            out = m(fr[1], fr[0])
        res[fused] = (out, m.get_intermediate("mv_hat"), m.get_intermediate("warpnet_res"), m.get_intermediate("quant_mv"),
                      m.launch_count())
        m.release()
    a, b = res["1"], res["0"]
    assert torch.equal(a[3], b[3])                                   # upstream of both tails: identical
    assert (a[1] - b[1]).abs().max().item() <= 2e-6 * max(1.0, b[1].abs().max().item())
    assert (a[2] - b[2]).abs().max().item() <= 2e-6 * max(1.0, b[2].abs().max().item())
    assert abs(float(a[0][7]) - float(b[0][7])) <= 1e-4 * float(b[0][7])
    print("launches fused %d vs separate %d" % (a[4], b[4]))


def test_three_partial_buffers_bit_identical_to_two(dev, state_dict, monkeypatch):
    """The fused GDN / tail kernels rotate THREE partial-accumulator buffers in TMEM (FVC_TC_NAB3, default on): buffering
    depth only, the chains and their order are unchanged -> every output bit-identical to the two-buffer ring, also over
    partial tiles and several tiles per CTA (the ring index is g mod 3, its phase bit flips every third chain)."""
    from fastvideocodec_b200 import VideoCompressor
    from fastvideocodec_b200.synthetic import synthetic_gop
    res = {}
    for (H, W, gid) in ((192, 320, 8), (448, 704, 9)):
        fr = synthetic_gop(H, W, gop=2, gop_id=gid).to(dev)
        for nab3 in ("1", "0"):
            monkeypatch.setenv("FVC_TC_NAB3", nab3)
            m = VideoCompressor()
            m.load_state_dict(state_dict)
            m = m.to(dev).eval()
            m.impl = _impls()[-1][1]
            with torch.no_grad():
                out = m(fr[1], fr[0])
            res[nab3] = (out, [m.get_intermediate(n) for n in ("mv_hat", "warpnet_res", "feature", "recon_res")])
            m.release()
        a, b = res["1"], res["0"]
        assert torch.equal(a[0][0], b[0][0]) and all(float(x) == float(y) for x, y in zip(a[0][1:], b[0][1:]))
        assert all(torch.equal(x, y) for x, y in zip(a[1], b[1]))


def test_lean_kernels_bit_identical_to_generic(dev, state_dict, monkeypatch):
    """Most layers run a LEAN instantiation of the convolution kernel (k_conv_tc<.., 4>: the epilogue branches they never
    take are compiled out, FVC_TC_LEAN, default on).  Same arithmetic in the same order: every tensor of the frame is
    bit-identical to the generic kernels', in both precision modes."""
    from fastvideocodec_b200 import VideoCompressor
    from fastvideocodec_b200.synthetic import synthetic_gop
    fr = synthetic_gop(320, 448, gop=2, gop_id=10).to(dev)
    names = ("estmv", "mvfeature", "mv_hat", "prediction", "feature", "z", "sigma", "recon_res")
    for precision in ("exact", "fast"):
        res = {}
        for lean in ("1", "0"):
            monkeypatch.setenv("FVC_TC_LEAN", lean)
            m = VideoCompressor(precision=precision)
            m.load_state_dict(state_dict)
            m = m.to(dev).eval()
            with torch.no_grad():
                out = m(fr[1], fr[0])
            res[lean] = (out, [m.get_intermediate(n) for n in names])
            m.release()
        a, b = res["1"], res["0"]
        assert torch.equal(a[0][0], b[0][0]) and all(float(x) == float(y) for x, y in zip(a[0][1:], b[0][1:])), precision
        assert all(torch.equal(x, y) for x, y in zip(a[1], b[1])), precision


def test_lsvc_forward_matches_oracle_larger_frames(dev, state_dict):
    """LSVC tree GOP (models.py:1344-1411) at 192x320 with 6 P-frames (three tree layers: batches of 6 / 2 / 4) against
    the oracle's restatement (pinned to the unmodified reference at 64x64): bpp / losses within 0.5 %, frames of the
    FIRST tree layer (they depend on no quantiser of this GOP but the mv latents) element-wise, all frames in the mean."""
    from fastvideocodec_b200.lsvc import LSVC, graph_from_batch, refidx_from_graph
    from fastvideocodec_b200.synthetic import synthetic_gop
    x = synthetic_gop(192, 320, gop=7, gop_id=14)[:, 0]
    g, layers, parents = graph_from_batch(6)
    want = O.lsvc_forward(state_dict, x, layers, parents, refidx_from_graph(g, 6))
    m = LSVC("LSVC-128")
    m.load_state_dict(state_dict)
    m = m.to(dev).eval()
    with torch.no_grad():
        out = m(x.to(dev))
    for i, n in enumerate(["rec_loss", "warp_loss", "mc_loss", "bpp_res", "bpp"], start=3):
        a, b = float(out[i]), float(want[i])
        assert abs(a - b) <= 0.005 * abs(b), (n, a, b)
    first = [t - 1 for t in layers[0] if t <= 6]
    for i, n in ((1, "mc"), (2, "warped")):
        d = (out[i].cpu() - want[i]).abs()
        assert d[first].max().item() <= 5e-3, (n, d[first].max().item())      # a tie-flipped mv latent moves these by ~2e-3
        assert d.mean().item() <= 2e-3, (n, d.mean().item())
    assert (out[0].cpu() - want[0]).abs().mean().item() <= 2e-3
    m.release()


@pytest.mark.parametrize("shape", [(1, 64, 64), (3, 5, 7), (2, 1088, 1920), (1, 1, 1), (2, 17, 6)])
def test_to_tensor_u8_bit_identical_to_torchvision_rule(dev, shape):
    """fvc_u8hwc_to_f32chw = transforms.ToTensor() of the reference's ingest (dataset.py:75): HWC -> CHW, x / 255 in fp32
    (IEEE division), bit for bit; aligned fast path, odd sizes (scalar tail) and unaligned views."""
    from fastvideocodec_b200 import ops
    n, H, W = shape
    g = torch.Generator().manual_seed(n * 1000 + H + W)
    u8 = torch.randint(0, 256, (n, H, W, 3), generator=g, dtype=torch.uint8)
    want = u8.permute(0, 3, 1, 2).float().div(255)            # torchvision.transforms.functional.to_tensor
    got = ops.to_tensor_u8(u8.to(dev))
    assert got.shape == want.shape and torch.equal(got.cpu(), want)
    if H * W >= 8:                                            # unaligned source (offset by one pixel = 3 bytes)
        flat = torch.cat([torch.zeros(3, dtype=torch.uint8), u8.reshape(-1)]).to(dev)
        got2 = ops.to_tensor_u8(flat[3:].view(n, H, W, 3))
        assert torch.equal(got2.cpu(), want)
    with pytest.raises(TypeError):
        ops.to_tensor_u8(u8)                                  # CPU tensor: no CPU path


def test_gop_forward_host_uint8_frames_equal_float_frames(model, dev):
    """fvc_gop_forward_host_u8: the GOP call fed with uint8 HWC frames gives exactly what the float32 call gives for
    ToTensor() of the same frames (scalars and reconstructions bit for bit)."""
    from fastvideocodec_b200.synthetic import synthetic_gop
    fr = synthetic_gop(128, 192, gop=4, gop_id=5)                               # [G,1,3,H,W] float
    u8 = (fr * 255.0).round().clamp(0, 255).to(torch.uint8).permute(0, 1, 3, 4, 2).contiguous()
    as_float = u8.permute(0, 1, 4, 2, 3).float().div(255).contiguous()
    rec_f, sc_f = model.gop_forward_host(as_float.pin_memory())
    rec_u, sc_u = model.gop_forward_host(u8.pin_memory())
    assert torch.equal(sc_f, sc_u) and torch.equal(rec_f, rec_u)
    assert bool(torch.isfinite(sc_u).all()) and rec_u.shape == (3, 1, 3, 128, 192)
    with pytest.raises(TypeError):
        model.gop_forward_host(u8.permute(0, 1, 4, 2, 3))                       # uint8 but channel-planar / strided


def test_rpm_prior_network_matches_reference(dev):
    """entropy_models.RPM / ConvLSTM (reference entropy_models.py:328-378) on the tcgen05 engine — 128-channel 3x3
    convolutions, the 256 -> 512 gate convolution and conv8 128 -> 256 as 128-channel blocks — against two recurrent
    steps of the reference's own classes (tests/golden/rpm_128.npz, oracle/gen_golden_r2.py rpm)."""
    from conftest import load_golden
    from fastvideocodec_b200 import ops
    from fastvideocodec_b200.entropy_models import RPM, RecProbModel
    from fastvideocodec_b200.synthetic import init_rpm_state_dict
    gold = load_golden("rpm_128.npz")
    rpm = RPM(128)
    sd = init_rpm_state_dict(128, 7)
    assert list(rpm.state_dict().keys()) == list(sd.keys())      # the generator asserts the same against the reference class
    rpm.load_state_dict(sd, strict=True)
    rpm = rpm.to(dev).eval()
    ops.conv_op_cache_clear()

    def close(a, b, what):
        err = (a.cpu() - b).abs().max().item()
        assert err <= 1e-4 * max(1.0, b.abs().max().item()), (what, err)

    with torch.no_grad():
        s0, m0, h1 = rpm(gold["x0"].to(dev), gold["h0"])                 # hidden may arrive on the CPU (reference: .to(x.device))
        n_handles = ops.conv_op_cache_size()
        s1, m1, h2 = rpm(gold["x1"].to(dev), h1)
    assert ops.conv_op_cache_size() == n_handles == 7 + 2 + 8            # conv1-7, conv8 (2 blocks), lstm (4 x 2 blocks): reused
    for got, name in ((s0, "sigma0"), (m0, "mu0"), (h1, "h1"), (s1, "sigma1"), (m1, "mu1"), (h2, "h2")):
        close(got, gold[name], name)
    # inside RecProbModel: the conditional-Gaussian branch driven by the network (entropy_models.py:58-63)
    m = RecProbModel(128)
    m.RPM.load_state_dict(sd)
    m = m.to(dev).eval()
    m.set_RPM(True)
    x = (gold["x1"] + 0.3).to(dev)
    with torch.no_grad():
        xh, lik, hid, prior = m(x, gold["h0"], training=False, prior_latent=gold["x0"].to(dev))
    close(m.sigma, torch.exp(torch.clamp(gold["sigma0"], min=-7.0)) / 10, "RecProbModel.sigma")
    close(hid, gold["h1"], "RecProbModel.hidden")
    assert torch.equal(prior, torch.round(x)) and bool(((lik > 0) & (lik <= 1)).all())
    ops.conv_op_cache_clear()


@pytest.mark.parametrize("shape", [(1, 128, 192), (2, 64, 64), (1, 1088, 1920)])
def test_iframe_codec_matches_oracle_and_round_trips(dev, state_dict, shape):
    """SURVEY 8f N4 — fvc_iframe_forward: a frame coded by the residual branch alone (net.py:86-116 with a zero
    prediction), against the oracle's composition of the same reference modules (O.iframe_forward):
      free run      : pre-round latents, quantised latents (count / +-1 / tie rule), rates
      teacher forced: with the oracle's latents, the reconstruction within 1e-3 everywhere
      real coding   : iframe_compress -> iframe_decompress reproduces the encoder's reconstruction bit for bit."""
    from fastvideocodec_b200 import VideoCompressor
    from fastvideocodec_b200.synthetic import synthetic_gop
    B, H, W = shape
    x = synthetic_gop(H, W, gop=1 + B, gop_id=17)[1:1 + B, 0].contiguous()
    want, cap = O.iframe_forward(state_dict, x, capture=True)
    m = VideoCompressor()
    m.load_state_dict(state_dict)
    m = m.to(dev).eval()
    xd = x.to(dev)
    with torch.no_grad():
        recon, mse, bpp_f, bpp_z, bpp = m.iframe_forward(xd)
    for name in ("feature", "z", "sigma"):
        got = m.get_intermediate(name).cpu()
        tol = 5e-4 if name != "sigma" else 2e-3        # sigma = exp(.) of a network of the (possibly flipped) z_hat
        if name != "sigma" or torch.equal(m.get_intermediate("z_hat").cpu(), cap["z_hat"]):
            err = (got - cap[name]).abs().max().item()
            assert err <= tol * max(1.0, cap[name].abs().max().item()), (name, err)
    nf = check_latent("feat_hat", m.get_intermediate("feat_hat").cpu(), cap["feat_hat"], cap["feature"])
    nz = check_latent("z_hat", m.get_intermediate("z_hat").cpu(), cap["z_hat"], cap["z"])
    if nf == 0 and nz == 0:
        assert (recon.cpu() - want[0]).abs().max().item() <= 1e-3
        for got, ref in ((mse, want[1]), (bpp_f, want[2]), (bpp_z, want[3]), (bpp, want[4])):
            assert abs(float(got) - float(ref)) <= 2e-4 * abs(float(ref)), (float(got), float(ref))
    else:
        assert abs(float(bpp) - float(want[4])) <= 5e-3 * float(want[4])
    # teacher forced: the oracle's quantised latents in, reconstruction everywhere
    m.force_latents(B, H, W, dev, z_hat=cap["z_hat"].to(dev), feat_hat=cap["feat_hat"].to(dev))
    with torch.no_grad():
        recon_f = m.iframe_forward(xd)[0]
    m.force_latents(B, H, W, dev)
    assert (recon_f.cpu() - want[0]).abs().max().item() <= 1e-3
    assert (m.get_intermediate("recon_res").cpu() - cap["recon_res"]).abs().max().item() <= 5e-4 * max(1.0, cap["recon_res"].abs().max().item())
    # real entropy coding and the decoder
    with torch.no_grad():
        streams, rec_enc, sc = m.iframe_compress(xd)
        assert set(streams) == {"feature", "z"} and torch.equal(rec_enc, recon)
        rec_dec = m.iframe_decompress(streams, (B, 3, H, W))
    assert torch.equal(rec_dec, rec_enc)
    real_bpp = 8.0 * (len(streams["feature"]) + len(streams["z"])) / (B * H * W)
    assert abs(float(sc[6]) - real_bpp) <= 1e-6 * real_bpp and float(sc[5]) == 0.0
    assert float(bpp) <= real_bpp <= 1.01 * float(bpp) + 8.0 * 160 / (B * H * W)
    # the I_compression-shaped hook
    with torch.no_grad():
        y, b, p = m.i_codec(xd)
    assert torch.equal(y, recon) and abs(float(p) - 10 * math.log10(1 / float(mse))) <= 1e-3
    # a P-frame after the intra frame on the same context is unaffected by the intra call
    if B == 1 and H <= 256:
        fr = synthetic_gop(H, W, gop=2, gop_id=3)[:, 0]
        with torch.no_grad():
            a = m(fr[1:2].to(dev), fr[0:1].to(dev))
            m.iframe_forward(xd)
            b2 = m(fr[1:2].to(dev), fr[0:1].to(dev))
        assert all(torch.equal(u, v) for u, v in zip(a, b2))
    with pytest.raises(TypeError):
        m.iframe_forward(x)
    m.release()


def test_parallel_compression_with_device_i_codec(model, dev):
    """models.py:258-262 / 412-429: ``parallel_compression(..., compressI=True)`` with the device intra codec in the place
    of the bpgenc / bpgdec shell-out: frame 0 is replaced by its intra reconstruction, its bpp / PSNR lead the lists, and
    the P-frames are predicted from it."""
    from fastvideocodec_b200 import parallel_compression
    from fastvideocodec_b200.synthetic import synthetic_gop
    model.r = 1024
    data = synthetic_gop(64, 128, gop=3, gop_id=2)[:, 0].to(dev)
    with torch.no_grad():
        rec0, bpp0, psnr0 = model.i_codec(data[0:1])
        out = parallel_compression(None, model, data.clone(), compressI=True, i_codec=model.i_codec)
        p1 = model(data[1:2], rec0)
    assert len(out) == 11 and len(out[6]) == 3                     # I + 2 P PSNRs
    assert abs(out[6][0] - float(psnr0)) <= 1e-4
    assert torch.equal(out[0][0:1], p1[0])                          # first P-frame predicted from the intra recon
    want_bpp = (float(bpp0) + float(p1[7])) / 2                     # running mean over the first two entries is inside out[3]
    assert out[3] > 0 and math.isfinite(out[3]) and want_bpp > 0
