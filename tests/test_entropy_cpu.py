"""CPU tests of the entropy-coding oracle (oracle/entropy_oracle.py): the reference's float CDF tables (pinned against
tables produced by the unmodified reference, tests/golden/entropy_tables.npz), the restated torchac integer conversion,
and the rANS lane format."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import entropy_oracle as E


@pytest.fixture(scope="module")
def gold():
    return load_golden("entropy_tables.npz")


def _sd(gold):
    return {k[3:]: v for k, v in gold.items() if k.startswith("sd.")}


def test_float_tables_match_reference(gold):
    mx = int(gold["mxrange"])
    sd = _sd(gold)
    for name in ("z", "mv"):
        got = E.reference_cdf_factorized(sd, "bitEstimator_" + name, mx)
        assert got.shape == gold["cdf_" + name].shape
        assert (got - gold["cdf_" + name]).abs().max().item() <= 1e-6
    got = E.reference_cdf_laplace(gold["lap_sigma"], mx)
    assert (got - gold["cdf_lap"]).abs().max().item() <= 1e-6


def test_torchac_int_conversion_properties(gold):
    mx = int(gold["mxrange"])
    for key in ("cdf_z", "cdf_mv", "cdf_lap"):
        t = E.torchac_int_cdf(gold[key])
        assert t.shape[-1] == 2 * mx and t.dtype == np.uint32
        assert (t[..., -1] == 65536).all() and (t[..., :-1] <= 65535).all()
        inc = E.strictly_increasing(t)
        assert (np.diff(inc.astype(np.int64), axis=-1) >= 1).all()       # every symbol codable
        assert (inc != t).mean() <= 1e-3                                  # the fix-up is the rare exception
    # known answers: F = 0, 0.5, 1 at table index 0, 150, 298 (Lp = 300: scale 65237)
    f = torch.zeros(1, 300)
    f[0, 150], f[0, 298] = 0.5, 1.0
    t = E.torchac_int_cdf(f)
    assert t[0, 0] == 0 and t[0, 150] == round(0.5 * 65237) + 150 and t[0, 298] == 65237 + 298 and t[0, 299] == 65536


def test_rans_round_trip_and_length(gold):
    mx = int(gold["mxrange"])
    table = E.strictly_increasing(E.torchac_int_cdf(gold["cdf_z"]))       # [64, 300]
    rng = np.random.default_rng(5)
    for n, L in ((1, 8), (7, 8), (8, 8), (9, 8), (1000, 64), (5000, 8192)):
        chan = np.arange(n) % 64
        sym = np.clip(np.rint(rng.normal(0, 3, n)).astype(np.int64) + mx, 0, 2 * mx - 2)
        start, freq = E.intervals_from_table(table, sym, chan)
        stream = E.rans_encode(start, freq, L)

        def lookup(k, slot):
            T = table[k % 64]
            s = int(np.searchsorted(T, slot, side="right") - 1)
            return s, int(T[s]), int(T[s + 1] - T[s])

        assert (E.rans_decode(stream, n, L, lookup) == sym).all()
        nlanes = (n + L - 1) // L
        payload_bits = 8 * (len(stream) - 16 - ((nlanes * 2 + 3) & ~3))
        ideal = E.ideal_bits(start, freq)
        # rANS with a 32-bit state, 16-bit words and 16-bit probabilities: 32 bits of final state per lane, up to one
        # word of slack, and <= 0.3 % coding redundancy (state precision 2^16 / probability resolution 2^16)
        assert ideal - 1e-6 <= payload_bits <= 1.003 * ideal + 48 * nlanes, (n, L, payload_bits, ideal)


# ------------------------------------------------------------------------------------------------
# CompressAI-side tables and the indexed coder (oracle/compressai_oracle.py; parity UNPINNED, CompressAI
# is not available: the published algorithm restated twice, once as the oracle and once as the product's host logic)
# ------------------------------------------------------------------------------------------------
def _trained_like_eb(C, seed):
    from fastvideocodec_b200.entropy_models import EntropyBottleneck
    g = torch.Generator().manual_seed(seed)
    eb = EntropyBottleneck(C)
    with torch.no_grad():
        for name, p in eb.named_parameters():
            if "_matrix" in name or "_factor" in name:
                p.add_(torch.randn(p.shape, generator=g) * 0.3)
        med = torch.randn(C, generator=g) * 0.7
        lo = 2.0 + torch.rand(C, generator=g) * 9
        hi = 2.0 + torch.rand(C, generator=g) * 14
        eb.quantiles[:, 0, 0], eb.quantiles[:, 0, 1], eb.quantiles[:, 0, 2] = med - lo, med, med + hi
    return eb


def _eb_parts(eb):
    m = [getattr(eb, "_matrix%d" % i).detach() for i in range(5)]
    b = [getattr(eb, "_bias%d" % i).detach() for i in range(5)]
    f = [getattr(eb, "_factor%d" % i).detach() for i in range(4)]
    return m, b, f


def test_pmf_to_quantized_cdf_known_answers_and_properties():
    from fastvideocodec_b200.entropy_models import pmf_to_quantized_cdf
    from oracle import compressai_oracle as CA
    # hand-computed (ops.cpp: round(p * 2^16), renormalise by the total, cumulate, last = 2^16, repair empty bins)
    assert pmf_to_quantized_cdf([0.5, 0.25, 0.25]).tolist() == [0, 32768, 49152, 65536]
    assert pmf_to_quantized_cdf([0.5, 0.0, 0.5]).tolist() == [0, 32767, 32768, 65536]      # bin 1 steals from bin 0
    assert pmf_to_quantized_cdf([0.0, 0.75, 0.25]).tolist() == [0, 1, 49153, 65536]        # bin 0 steals from bin 2 (narrowest): cdf[1..2] += 1
    assert pmf_to_quantized_cdf([1.0]).tolist() == [0, 65536]
    rng = np.random.default_rng(3)
    for trial in range(60):
        n = int(rng.integers(1, 400))
        p = rng.random(n).astype(np.float32) ** int(rng.integers(1, 12))
        p[rng.random(n) < 0.3] = 0
        if trial % 3 == 0:
            p[int(rng.integers(0, n))] = 1.0
        if p.sum() == 0:
            p[0] = 1
        p = p / p.sum()
        got, want = pmf_to_quantized_cdf(p), CA.pmf_to_quantized_cdf(p)
        assert np.array_equal(got, want)
        assert got[0] == 0 and got[-1] == 65536 and np.all(np.diff(got) > 0)
    with pytest.raises(ValueError):
        pmf_to_quantized_cdf([0.0, 0.0])
    with pytest.raises(ValueError):
        pmf_to_quantized_cdf([0.5, float("nan")])


def test_update_tables_match_oracle():
    from fastvideocodec_b200.entropy_models import GaussianConditional, get_scale_table
    from oracle import compressai_oracle as CA
    eb = _trained_like_eb(24, 5)
    assert eb.update() is True and eb.update() is False and eb.update(force=True) is True
    cdf, ln, off = CA.eb_tables(*_eb_parts(eb), eb.quantiles)
    assert np.array_equal(eb._quantized_cdf.numpy(), cdf) and np.array_equal(eb._cdf_length.numpy(), ln)
    assert np.array_equal(eb._offset.numpy(), off)
    q = eb.quantiles.detach()
    # support of channel c: [median - ceil(median - q_lo), median + ceil(q_hi - median)], + the escape bin, + 1 for the cdf
    want_len = torch.ceil(q[:, 0, 1] - q[:, 0, 0]) + torch.ceil(q[:, 0, 2] - q[:, 0, 1]) + 1 + 2
    assert torch.equal(eb._cdf_length.float(), want_len)
    for c in range(24):
        row = eb._quantized_cdf[c, :eb._cdf_length[c]].numpy()
        assert row[0] == 0 and row[-1] == 65536 and np.all(np.diff(row) > 0)
        assert np.all(eb._quantized_cdf[c, eb._cdf_length[c]:].numpy() == 0)
    # the quantised pmf is the model's pmf: compare with the likelihood forward of the oracle at the integer offsets
    # (on the channel with the longest support: for shorter ones CompressAI takes the upper tail at the end of the
    # LONGEST support, so their bins are renormalised upwards by the missing mass; restated as published)
    c = int(torch.argmax(eb._cdf_length))
    k = torch.arange(int(eb._cdf_length[c]) - 2).float() + float(eb._offset[c]) + q[c, 0, 1]
    x = torch.zeros((1, 24, 1, len(k))) + q[:, 0, 1].view(1, 24, 1, 1)
    x[0, c, 0] = k
    from oracle import dvc_oracle as O
    _, lik = O.eb_forward(*_eb_parts(eb), q[:, 0, 1], x)
    pm = np.diff(eb._quantized_cdf[c, :eb._cdf_length[c] - 1].numpy()) / 65536.0
    assert np.abs(pm - lik[0, c, 0].numpy()).max() <= 3.0 / 65536

    gc = GaussianConditional(None)
    assert gc.update_scale_table(get_scale_table()) is True and gc.update_scale_table(get_scale_table()) is False
    cdf, ln, off = CA.gaussian_tables(get_scale_table().numpy())
    assert np.array_equal(gc._quantized_cdf.numpy(), cdf) and np.array_equal(gc._cdf_length.numpy(), ln)
    assert np.array_equal(gc._offset.numpy(), off)
    assert gc._quantized_cdf.shape == (64, 3133)          # ceil(256 * 6.1094) = 1565 -> 2 * 1565 + 1 + 2
    # symmetric tables: pmf(k) == pmf(-k)
    row = np.diff(gc._quantized_cdf[20, :gc._cdf_length[20] - 1].numpy())
    assert np.abs(row - row[::-1]).max() <= 2
    g = torch.Generator().manual_seed(2)
    scales = torch.exp(torch.randn((2, 3, 5, 7), generator=g) * 2.5)
    scales.view(-1)[:64] = get_scale_table()              # exact ties with the table
    scales.view(-1)[64:67] = torch.tensor([0.0, 0.05, 1e9])
    assert torch.equal(gc.build_indexes(scales), CA.build_indexes(scales, get_scale_table()))
    with pytest.raises(ValueError):
        GaussianConditional(None).update_scale_table([1.0, 0.5])


def test_indexed_coder_oracle_round_trip_with_escapes():
    from oracle import compressai_oracle as CA
    cdf, ln, off = CA.gaussian_tables(np.exp(np.linspace(np.log(0.11), np.log(64), 16)))
    rng = np.random.default_rng(11)
    n = 3000
    idx = rng.integers(0, 16, n)
    sym = np.rint(rng.standard_normal(n) * np.exp(np.linspace(np.log(0.11), np.log(64), 16))[idx]).astype(np.int64)
    esc = rng.random(n) < 0.02                            # far outside the tables: bypass digits, both signs, up to 2^30
    sym[esc] = (rng.integers(1, 1 << 30, esc.sum()) * rng.choice([-1, 1], esc.sum())) >> rng.integers(0, 28, esc.sum())
    sym[:4] = [off[idx[0]] - 1, off[idx[1]] + ln[idx[1]] - 2, off[idx[2]] + ln[idx[2]] - 3, off[idx[3]]]   # both edges
    for lane in (64, 8192):
        s = CA.encode_indexed(sym, idx, cdf, ln, off, lane)
        assert np.array_equal(CA.decode_indexed(s, idx, cdf, ln, off, lane), sym)
        ideal = CA.ideal_bits(sym, idx, cdf, ln, off)
        nl = -(-n // lane)
        assert ideal <= 8 * len(s) <= ideal + 34 * nl + 8 * (16 + 2 * nl + 4) + 16 * nl
    # escape layout: count digit(s) then the raw value, least significant digit first
    iv = CA.element_intervals(int(off[3]) - 3, 3, cdf, ln, off)          # v = -3 -> raw = 5 -> one digit
    assert iv[0][0] == cdf[3][ln[3] - 2] and iv[1:] == [(1 << 12, 1 << 12), (5 << 12, 1 << 12)]
    iv = CA.element_intervals(int(off[3]) + int(ln[3]) - 2, 3, cdf, ln, off)   # v = max -> raw = 0 -> no digits
    assert iv[1:] == [(0, 1 << 12)]
