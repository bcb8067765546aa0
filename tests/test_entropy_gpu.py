"""GPU tests of the real entropy coding path (reference DVC/net.py:123-138, 155-168, 183-195) through the
C ABI: integer CDF tables against the oracle's restatement of the reference + torchac conversion, the rANS coder
byte-exact against the oracle's coder, encode -> decode round trips, and the whole codec (compress -> decompress)."""
import math

import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import dvc_oracle as O
from oracle import entropy_oracle as E

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def gold():
    return load_golden("entropy_tables.npz")


def _params(sd, prefix, dev):
    out = []
    for i in (1, 2, 3, 4):
        for p in ("h", "b", "a"):
            if not (i == 4 and p == "a"):
                out.append(sd[f"{prefix}.f{i}.{p}"].to(dev))
    return out


def test_cdf_tables_match_reference_model(dev, gold):
    """The 2*mxrange-entry integer CDFs the coder uses == the reference's float tables (from the unmodified reference,
    entropy_tables.npz) pushed through torchac's published int conversion: within 1 unit of 2^-16 everywhere (GPU
    tanh / sigmoid / expm1 vs CPU), identical for >= 99 % of the entries."""
    from fastvideocodec_b200 import ops
    mx = int(gold["mxrange"])
    sd = {k[3:]: v for k, v in gold.items() if k.startswith("sd.")}
    for name in ("z", "mv"):
        got = ops.cdf_table_factorized(_params(sd, "bitEstimator_" + name, dev), mx).cpu().numpy().astype(np.int64)
        want = E.strictly_increasing(E.torchac_int_cdf(gold["cdf_" + name])).astype(np.int64)
        assert got.shape == want.shape
        assert np.abs(got - want).max() <= 1, (name, np.abs(got - want).max())
        assert (got == want).mean() >= 0.99
        assert (np.diff(got, axis=-1) >= 1).all() and (got[:, -1] == 65536).all()
    got = ops.cdf_table_laplace(gold["lap_sigma"].to(dev), mx).cpu().numpy().astype(np.int64)
    want = E.torchac_int_cdf(gold["cdf_lap"]).astype(np.int64)
    assert np.abs(got - want).max() <= 1 and (got == want).mean() >= 0.99
    assert (np.diff(got, axis=-1) >= 1).all()


@pytest.mark.parametrize("n_pix,lane", [(1, 8192), (5, 16), (37, 64), (300, 8192), (1031, 512)])
def test_rans_encoder_byte_exact_and_round_trip_factorized(dev, gold, n_pix, lane):
    """Encoder bytes == the oracle's rANS coder fed with the same (start, freq) pairs; decode(encode(x)) == round(x).
    Ragged sizes: one symbol, lanes of unequal length, exactly full lanes."""
    from fastvideocodec_b200 import ops
    mx = int(gold["mxrange"])
    sd = {k[3:]: v for k, v in gold.items() if k.startswith("sd.")}
    table = ops.cdf_table_factorized(_params(sd, "bitEstimator_mv", dev), mx)
    g = torch.Generator().manual_seed(n_pix)
    x = torch.randn((n_pix, 128), generator=g) * 4
    x[0, :4] = torch.tensor([0.5, 1.5, -149.6, 148.4])                     # ties (half to even) and the range limits
    stream = ops.entropy_encode_factorized(x.to(dev), table, mx, lane)
    q = torch.round(x)
    sym = (q.flatten().numpy().astype(np.int64) + mx)
    chan = np.arange(sym.size) % 128
    start, freq = E.intervals_from_table(table.cpu().numpy().astype(np.uint32), sym, chan)
    assert stream == E.rans_encode(start, freq, lane)
    back = ops.entropy_decode_factorized(stream, x.shape, table, mx, lane)
    assert torch.equal(back.cpu(), q)
    assert len(stream) <= 16 + 4 + 2 * ((sym.size + lane - 1) // lane) + (1.003 * E.ideal_bits(start, freq) + 48 * ((sym.size + lane - 1) // lane)) / 8


def test_rans_round_trip_laplace(dev):
    from fastvideocodec_b200 import ops
    g = torch.Generator().manual_seed(3)
    x = torch.randn((68 * 30, 96), generator=g) * 3
    sigma = torch.exp(torch.randn(x.shape, generator=g) * 2)
    sigma[0, :4] = torch.tensor([0.0, 1e-7, 1e12, 1e-5])
    x[0, :4] = 0.0
    stream = ops.entropy_encode_laplace(x.to(dev), sigma.to(dev), 150, 4096)
    back = ops.entropy_decode_laplace(stream, sigma.to(dev), 150, 4096)
    assert torch.equal(back.cpu(), torch.round(x))
    # code length against the ideal code length of the reference's model for these symbols
    tab = E.torchac_int_cdf(E.reference_cdf_laplace(sigma.flatten(), 150))
    start, freq = E.intervals_from_table(tab, torch.round(x).flatten().numpy().astype(np.int64) + 150)
    nl = (x.numel() + 4095) // 4096
    assert abs(8 * len(stream) - E.ideal_bits(start, np.maximum(freq, 1))) <= 0.004 * 8 * len(stream) + 64 * nl + 256


def test_uncodable_symbol_is_an_error(dev, gold):
    """A latent outside [-mxrange, mxrange-2]: the reference's torchac call raises (check_input_bounds=True); so do we."""
    from fastvideocodec_b200 import _lib, ops
    sd = {k[3:]: v for k, v in gold.items() if k.startswith("sd.")}
    table = ops.cdf_table_factorized(_params(sd, "bitEstimator_z", dev), 8)
    x = torch.zeros((4, 64))
    x[2, 5] = 7.0        # q + mxrange = 15 > 2*mxrange - 2
    with pytest.raises(_lib.FvcError):
        ops.entropy_encode_factorized(x.to(dev), table, 8, 64)


def _psnr(mse):
    return 10.0 * math.log10(1.0 / float(mse))


@pytest.mark.parametrize("size", [(256, 256), (1088, 1920)])
def test_codec_round_trip_and_real_bits(dev, state_dict, size):
    """compress -> three byte streams -> decompress reproduces the encoder's reconstruction BIT FOR BIT (the decoder
    re-derives sigma from the decoded z_hat and decodes feature under it); the real bpp (8 x bytes, net.py:136) sits
    within 1 % above the estimate and the calrealbits forward returns exactly that."""
    from fastvideocodec_b200 import VideoCompressor
    from fastvideocodec_b200.synthetic import synthetic_gop
    H, W = size
    m = VideoCompressor()
    m.load_state_dict(state_dict)
    m = m.to(dev).eval()
    fr = synthetic_gop(H, W, gop=2, gop_id=4).to(dev)
    with torch.no_grad():
        est = m(fr[1], fr[0])
        lat = {n: m.get_intermediate(n).clone() for n in ("quant_mv", "z_hat", "feat_hat")}
        streams, recon, sc = m.compress(fr[1], fr[0])
        assert torch.equal(recon, est[0])
        back = m.decompress(streams, fr[0])
        assert torch.equal(back, recon), (back - recon).abs().max().item()
        for n in lat:                                     # the decoder's latents are the encoder's
            assert torch.equal(m.get_intermediate(n), lat[n]), n
        m.calrealbits = True
        real = m(fr[1], fr[0])
        m.calrealbits = False
    npx = H * W
    nbytes = {k: len(v) for k, v in streams.items()}
    assert abs(float(real[4]) - 8 * nbytes["feature"] / npx) <= 1e-6 * float(real[4])
    assert abs(float(real[5]) - 8 * nbytes["z"] / npx) <= 1e-6 * float(real[5])
    assert abs(float(real[6]) - 8 * nbytes["mv"] / npx) <= 1e-6 * float(real[6])
    assert abs(float(real[7]) - float(sc[6])) <= 1e-6 * float(real[7])
    assert float(real[1]) == float(est[1])                # distortion unchanged
    gap = (float(real[7]) - float(est[7])) / float(est[7])
    print("real vs estimated bpp at %dx%d: %.5f vs %.5f (%+.3f %%), bytes %s" % (H, W, float(real[7]), float(est[7]),
                                                                              100 * gap, nbytes))
    assert -0.002 <= gap <= 0.01, gap
    # a corrupted container is rejected, not decoded into garbage silently
    bad = dict(streams)
    bad["z"] = b"XXXX" + streams["z"][4:]
    from fastvideocodec_b200 import _lib
    with pytest.raises(_lib.FvcError):
        m.decompress(bad, fr[0])
    m.release()


# ------------------------------------------------------------------------------------------------
# the indexed-table coder and the CompressAI-side compress / decompress of entropy_models.py
# (parity UNPINNED against CompressAI itself; checked against the oracle's restatement, oracle/compressai_oracle.py)
# ------------------------------------------------------------------------------------------------
def _scale_table(n=16, hi=64.0):
    return np.exp(np.linspace(np.log(0.11), np.log(hi), n))


@pytest.mark.parametrize("n,lane", [(1, 8192), (63, 16), (3000, 64), (20000, 8192)])
def test_indexed_coder_byte_exact_and_round_trip(dev, n, lane):
    from fastvideocodec_b200 import ops
    from oracle import compressai_oracle as CA
    st = _scale_table()
    cdf, ln, off = CA.gaussian_tables(st)
    rng = np.random.default_rng(100 + n)
    idx = rng.integers(0, len(st), n)
    sym = np.rint(rng.standard_normal(n) * st[idx]).astype(np.int64)
    esc = rng.random(n) < 0.03                            # escapes: both signs, 1..8 bypass digits
    sym[esc] = (rng.integers(1, 1 << 30, esc.sum()) * rng.choice([-1, 1], esc.sum())) >> rng.integers(0, 28, esc.sum())
    if n >= 4:
        sym[:4] = [off[idx[0]] - 1, off[idx[1]] + ln[idx[1]] - 2, off[idx[2]] + ln[idx[2]] - 3, off[idx[3]]]
    t = lambda a: torch.from_numpy(np.asarray(a)).to(dev)
    got = ops.entropy_encode_indexed(t(sym), t(idx), t(cdf), t(ln), t(off), lane)
    want = CA.encode_indexed(sym, idx, cdf, ln, off, lane)
    assert got == want
    back = ops.entropy_decode_indexed(got, t(idx), t(cdf), t(ln), t(off), lane)
    assert back.dtype == torch.int32 and np.array_equal(back.cpu().numpy(), sym)
    # a one-symbol table (freq = 2^16) codes for free and round-trips
    cdf1 = np.asarray([[0, 65536, 0], [0, 40000, 65536]], dtype=np.int32)
    ln1, off1 = np.asarray([2, 3], dtype=np.int32), np.asarray([0, 0], dtype=np.int32)
    s1 = np.zeros(n, dtype=np.int64)
    i1 = (np.arange(n) % 2).astype(np.int64)
    g1 = ops.entropy_encode_indexed(t(s1), t(i1), t(cdf1), t(ln1), t(off1), lane)
    assert g1 == CA.encode_indexed(s1, i1, cdf1, ln1, off1, lane)
    assert np.array_equal(ops.entropy_decode_indexed(g1, t(i1), t(cdf1), t(ln1), t(off1), lane).cpu().numpy(), s1)


def test_indexed_coder_errors(dev):
    from fastvideocodec_b200 import ops
    from fastvideocodec_b200._lib import FvcError
    from oracle import compressai_oracle as CA
    cdf, ln, off = CA.gaussian_tables(_scale_table(4, 8.0))
    t = lambda a: torch.from_numpy(np.asarray(a)).to(dev)
    sym, idx = np.zeros(10, dtype=np.int64), np.zeros(10, dtype=np.int64)
    bad = idx.copy()
    bad[3] = 4
    with pytest.raises(FvcError):                                  # index outside the table set
        ops.entropy_encode_indexed(t(sym), t(bad), t(cdf), t(ln), t(off))
    broken = cdf.copy()
    broken[3, 5] = broken[3, 4]                                    # empty interval in the widest table
    s2 = sym.copy()
    s2[0] = off[3] + 4
    with pytest.raises(FvcError):
        ops.entropy_encode_indexed(t(s2), t(idx + 3), t(broken), t(ln), t(off))
    good = ops.entropy_encode_indexed(t(sym), t(idx), t(cdf), t(ln), t(off))
    with pytest.raises(FvcError):                                  # truncated container
        ops.entropy_decode_indexed(good[:12], t(idx), t(cdf), t(ln), t(off))
    with pytest.raises(FvcError):                                  # wrong element count
        ops.entropy_decode_indexed(good, t(idx[:5]), t(cdf), t(ln), t(off))
    with pytest.raises(ValueError):
        ops.entropy_encode_indexed(t(sym), t(idx[:5]), t(cdf), t(ln), t(off))
    # zero symbols: the empty string, both ways
    empty = torch.zeros(0, dtype=torch.int32, device=dev)
    assert ops.entropy_encode_indexed(empty, empty, t(cdf), t(ln), t(off)) == b""
    assert ops.entropy_decode_indexed(b"", empty, t(cdf), t(ln), t(off)).numel() == 0
    with pytest.raises(FvcError):
        ops.entropy_decode_indexed(good, empty, t(cdf), t(ln), t(off))


def _trained_like(eb, g):
    C = eb.channels
    with torch.no_grad():
        for name, p in eb.named_parameters():
            if "_matrix" in name or "_factor" in name:
                p.add_(torch.randn(p.shape, generator=g) * 0.3)
        med = torch.randn(C, generator=g) * 0.7
        eb.quantiles[:, 0, 0] = med - 3.0 - torch.rand(C, generator=g) * 9
        eb.quantiles[:, 0, 1] = med
        eb.quantiles[:, 0, 2] = med + 3.0 + torch.rand(C, generator=g) * 14


def test_recprob_update_compress_decompress(dev):
    """RecProbModel.update / compress / decompress / get_actual_bits (entropy_models.py:43-48, 70-72, 80-94), factorized
    branch and conditional-Gaussian branch (RPM outputs supplied by a stub network)."""
    from fastvideocodec_b200.entropy_models import RecProbModel
    from oracle import compressai_oracle as CA
    g = torch.Generator().manual_seed(31)
    C = 32

    def rpm(prior_latent, hidden):
        return prior_latent * 0.05 - 1.0, prior_latent * 0.9, hidden + 1

    m = RecProbModel(C, rpm=rpm)
    _trained_like(m.entropy_bottleneck, g)
    with pytest.raises(ValueError):
        m.entropy_bottleneck.compress(torch.zeros((1, C, 2, 2), device=dev))     # update() not run
    m = m.to(dev).eval()
    assert m.update(force=True) is True and m.update() is False
    x = (torch.randn((2, C, 9, 14), generator=g) * 3).to(dev)
    x[0, 0, 0, :4] = torch.tensor([500.0, -700.0, 40.0, -40.0])                  # escapes
    m.set_RPM(False)
    hidden = torch.zeros(1)
    with torch.no_grad():
        xh, lik, _, prior = m(x, hidden, training=False)
        strings = m.compress(x)
        assert isinstance(strings, list) and len(strings) == 2 and all(isinstance(s, bytes) for s in strings)
        back = m.decompress(strings, x.shape[-2:])
    assert torch.equal(back, xh)
    # byte-exact against the oracle coder on the tables the module built (NCHW order, one string per batch element)
    eb = m.entropy_bottleneck
    cdf, ln, off = eb._quantized_cdf.cpu().numpy(), eb._cdf_length.cpu().numpy(), eb._offset.cpu().numpy()
    med = eb.quantiles[:, 0, 1].detach().view(1, C, 1, 1)
    sym = torch.round(x - med).long().cpu().numpy()
    idx = np.broadcast_to(np.arange(C).reshape(1, C, 1, 1), sym.shape)
    for b in range(2):
        assert strings[b] == CA.encode_indexed(sym[b].reshape(-1), idx[b].reshape(-1), cdf, ln, off, eb.lane_len)
    bits = float(m.get_actual_bits(strings))
    assert bits == 8 * sum(len(s) for s in strings)
    ideal = sum(CA.ideal_bits(sym[b].reshape(-1), idx[b].reshape(-1), cdf, ln, off) for b in range(2))
    assert ideal <= bits <= ideal + 2 * (34 + 8 * 24)
    # the table model is the likelihood model: actual bits track the (unclamped) estimate outside the escapes
    est = float(-torch.log2(lik).sum())
    assert abs(bits - est) <= 0.05 * est + 2000
    # conditional-Gaussian branch
    m.set_RPM(True)
    with torch.no_grad():
        xh, lik, hid, _ = m(x, hidden, training=False, prior_latent=prior)
        strings = m.compress(x)
        back = m.decompress(strings, x.shape[-2:])
        assert torch.equal(back, xh)
        x2, s2, h2, p2 = m.compress_slow(x, hidden, prior)
        assert s2 == strings and torch.equal(x2, xh) and torch.equal(p2, torch.round(xh))
        x3, h3, p3 = m.decompress_slow(s2, x.shape[-2:], hidden, prior)
        assert torch.equal(x3, xh) and torch.equal(p3, p2) and m.enc_t > 0 and m.dec_t > 0
    gc = m.gaussian_conditional
    idx = CA.build_indexes(m.sigma.cpu(), gc.scale_table.cpu()).numpy()
    sym = torch.round(x - m.mu).long().cpu().numpy()
    cdf, ln, off = gc._quantized_cdf.cpu().numpy(), gc._cdf_length.cpu().numpy(), gc._offset.cpu().numpy()
    assert strings[1] == CA.encode_indexed(sym[1].reshape(-1), idx[1].reshape(-1), cdf, ln, off, gc.lane_len)
    # state_dict round trip keeps the tables (CompressAI layout)
    from fastvideocodec_b200.entropy_models import RecProbModel as R2
    m2 = R2(C, rpm=rpm)
    m2.load_state_dict(m.state_dict())
    assert torch.equal(m2.entropy_bottleneck._quantized_cdf, eb._quantized_cdf.cpu())
    assert m2.update() is False                                   # tables came with the checkpoint


@pytest.mark.parametrize("trick", [True, False])
def test_meanscale_compress_decompress(dev, trick):
    """MeanScaleHyperPriors.update / compress / decompress / compress_slow / decompress_slow / get_actual_bits
    (entropy_models.py:194-197, 221-226, 237-324)."""
    from fastvideocodec_b200.entropy_models import MeanScaleHyperPriors
    g = torch.Generator().manual_seed(41)
    C = 64
    m = MeanScaleHyperPriors(C, entropy_trick=trick)
    _trained_like(m.entropy_bottleneck, g)
    with torch.no_grad():
        for name, p in m.named_parameters():
            if name.startswith("h_"):
                p.copy_(torch.randn(p.shape, generator=g) * (0.04 if p.dim() == 4 else 0.1))
    m = m.to(dev).eval()
    assert m.update(force=True)
    x = (torch.randn((2, C, 16, 24), generator=g) * 3).to(dev)
    with torch.no_grad():
        xh, (xl, zl) = m(x, training=False)
        strings = m.compress(x)
        assert len(strings) == 2 and len(strings[0]) == 2 and len(strings[1]) == 2
        assert torch.equal(m.decompress(strings, x.shape[-2:]), xh)
        bits = m.get_actual_bits(strings)
        assert bits.shape == (2,) and bits.tolist() == [8.0 * (len(strings[0][b]) + len(strings[1][b])) for b in range(2)]
        # code length against the ideal -log2 sum of the intervals actually coded (x under the Gaussian tables picked by
        # build_indexes, escapes included; the random hyper-networks here make the model a poor fit, which is fine)
        from oracle import compressai_oracle as CA
        gc = m.gaussian_conditional
        cdf, ln, off = gc._quantized_cdf.cpu().numpy(), gc._cdf_length.cpu().numpy(), gc._offset.cpu().numpy()
        idx = CA.build_indexes(m.sigma.cpu(), gc.scale_table.cpu()).numpy()
        sym = torch.round(x - m.mu).long().cpu().numpy()
        for b in range(2):
            ideal = CA.ideal_bits(sym[b].reshape(-1), idx[b].reshape(-1), cdf, ln, off)
            nl = -(-sym[b].size // gc.lane_len)
            assert ideal <= 8 * len(strings[0][b]) <= ideal + 34 * nl + 8 * (20 + 2 * nl) + 16 * nl
            assert strings[0][b] == CA.encode_indexed(sym[b].reshape(-1), idx[b].reshape(-1), cdf, ln, off, gc.lane_len)
        # the slow path recomputes everything from x on the encoder and from the strings alone on the decoder
        x_hat, s, z_size = m.compress_slow(x, decode=True)
        assert len(s[0]) == (1 if trick else 2) and tuple(z_size) == ((2, 16, 24) if trick else (16, 24))
        back = m.decompress_slow(s, z_size)
        assert back.shape == x.shape and torch.equal(back, x_hat)
        assert (x_hat - xh).abs().max().item() <= 1e-3              # same model as the fast path
        assert m.enc_t > 0 and m.dec_t > 0
