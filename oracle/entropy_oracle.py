"""CPU restatement of the reference's REAL entropy-coding branch (``calrealbits``) — TEST INFRASTRUCTURE ONLY.

Follows DVC/net.py:123-138 (feature, Laplace(0, sigma)), 155-168 (z, bitEstimator_z) and 183-195 (mv,
bitEstimator_mv): per element a float CDF table with ``2*mxrange`` entries ``cdf[i] = F(i - mxrange - 0.5)``, symbols
``x + mxrange``, ``torchac.encode_float_cdf`` and ``real_bits = len(byte_stream) * 8``.

``torchac`` is an un-vendored, un-pinned dependency of the reference (``import torchac``, net.py:15; not installable
here): its published float->int16 CDF conversion is restated in ``torchac_int_cdf`` — PARITY UNPINNED at that boundary,
the float tables in front of it are pinned against the reference's own BitEstimator / torch.distributions calls
(tests/test_entropy_cpu.py).  torchac's coder is an arithmetic coder; the product codes the same integer model with
rANS lanes (fastvideocodec_b200/csrc/fvc_entropy.cu), whose byte format is restated here bit for bit
(``rans_encode`` / ``rans_decode``) so that the CUDA coder can be checked byte-exactly.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from . import dvc_oracle as O

RANS_L = 1 << 16
MAGIC = 0x31525646  # "FVR1"


# ------------------------------------------------------------------------------------------------
# the reference's float CDF tables
# ------------------------------------------------------------------------------------------------
def reference_cdf_factorized(sd, prefix, mxrange=150):
    """net.py:158-159 / 186-187: ``cdfs.append(self.bitEstimator(i - 0.5))`` for i in range(-mxrange, mxrange).
    Returns [C, 2*mxrange] (the table does not depend on the position)."""
    C = sd[prefix + ".f1.h"].numel()
    i = torch.arange(-mxrange, mxrange, dtype=torch.float32).view(1, 1, 1, -1).repeat(1, C, 1, 1)
    return O.bit_estimator_cdf(sd, prefix, i - 0.5)[0, :, 0, :]


def reference_cdf_laplace(sigma, mxrange=150):
    """net.py:127-128, 141-143: Laplace(0, clamp(sigma, 1e-5, 1e10)).cdf(i - 0.5); returns sigma.shape + [2*mxrange]."""
    sg = sigma.clamp(1e-5, 1e10).unsqueeze(-1)
    i = torch.arange(-mxrange, mxrange, dtype=torch.float32)
    return O.laplace_cdf((i - 0.5).expand(sg.shape[:-1] + (2 * mxrange,)), sg)


def torchac_int_cdf(cdf_float):
    """torchac._convert_to_int_and_normalize (needs_normalization=True, PRECISION=16), restated: round(cdf * (2^16 -
    (Lp - 1))) + arange(Lp) in 16-bit arithmetic; and the coder's ``c_high = 2^16 for the last symbol`` rule folded in
    as table[..., Lp-1] = 2^16.  Returns uint32 [..., Lp]: symbol s occupies [table[s], table[s+1])."""
    Lp = cdf_float.shape[-1]
    q = torch.round(cdf_float.float() * float(65536 - (Lp - 1))).to(torch.int64)
    q = (q + torch.arange(Lp, dtype=torch.int64)) & 0xFFFF
    q[..., Lp - 1] = 65536
    return q.numpy().astype(np.uint32)


def strictly_increasing(table):
    """The fix-up the product applies to the per-channel tables (float evaluation of saturated tails may wobble)."""
    t = table.astype(np.int64).copy()
    Lp = t.shape[-1]
    for i in range(1, Lp - 1):
        bad = t[..., i] <= t[..., i - 1]
        t[..., i] = np.where(bad, t[..., i - 1] + 1, t[..., i])
    return t.astype(np.uint32)


def ideal_bits(starts, freqs):
    """sum(-log2(freq / 2^16)): the code length both an arithmetic coder (torchac) and rANS approach."""
    return float(np.sum(16.0 - np.log2(freqs.astype(np.float64))))


# ------------------------------------------------------------------------------------------------
# rANS lanes, byte-exact restatement of fvc_entropy.cu
# ------------------------------------------------------------------------------------------------
def rans_encode(starts, freqs, lane_len):
    """starts, freqs: integer arrays [n] (freq >= 1, start + freq <= 2^16).  Returns the FVR1 container (bytes)."""
    n = len(starts)
    nlanes = (n + lane_len - 1) // lane_len
    lanes = []
    for l in range(nlanes):
        a, b = l * lane_len, min(n, (l + 1) * lane_len)
        x = RANS_L
        words = []
        for k in range(b - 1, a - 1, -1):
            s, f = int(starts[k]), int(freqs[k])
            if x >= (f << 16):
                words.append(x & 0xFFFF)
                x >>= 16
            x = ((x // f) << 16) + (x % f) + s
        words.append(x & 0xFFFF)
        words.append(x >> 16)
        lanes.append(np.asarray(words[::-1], dtype=np.uint16))
    hdr = np.asarray([MAGIC, n, lane_len, nlanes], dtype=np.uint32).tobytes()
    lw = np.zeros(((nlanes + 1) // 2) * 2, dtype=np.uint16)
    lw[:nlanes] = [len(w) for w in lanes]
    return hdr + lw.tobytes() + b"".join(w.tobytes() for w in lanes)


def rans_decode(stream, n, lane_len, lookup):
    """lookup(k, slot) -> (symbol, start, freq) for element k.  Returns the int symbols [n]."""
    h = np.frombuffer(stream[:16], dtype=np.uint32)
    assert h[0] == MAGIC and h[1] == n and h[2] == lane_len
    nlanes = int(h[3])
    lw = np.frombuffer(stream[16:16 + 2 * nlanes], dtype=np.uint16).astype(np.int64)
    base = 16 + ((nlanes * 2 + 3) & ~3)
    words = np.frombuffer(stream[base:], dtype=np.uint16)
    out = np.zeros(n, dtype=np.int64)
    off = 0
    for l in range(nlanes):
        w = words[off:off + lw[l]]
        off += int(lw[l])
        x = (int(w[0]) << 16) | int(w[1])
        p = 2
        for k in range(l * lane_len, min(n, (l + 1) * lane_len)):
            sym, s, f = lookup(k, x & 0xFFFF)
            out[k] = sym
            x = f * (x >> 16) + (x & 0xFFFF) - s
            if x < RANS_L:
                x = (x << 16) | int(w[p])
                p += 1
    return out


def intervals_from_table(table, sym, chan=None):
    """(start, freq) of symbols `sym` under per-channel tables [C, Lp] (chan = channel of each symbol) or per-element
    tables [n, Lp] (chan None)."""
    idx = np.arange(len(sym)) if chan is None else chan
    start = table[idx, sym].astype(np.int64)
    end = table[idx, sym + 1].astype(np.int64)
    return start, end - start
