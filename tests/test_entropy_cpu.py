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
