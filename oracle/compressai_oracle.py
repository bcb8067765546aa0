"""CPU restatement of the CompressAI range-coding side the reference's ``entropy_models.py`` calls — TEST INFRASTRUCTURE ONLY.

The reference's ``RecProbModel`` / ``MeanScaleHyperPriors`` (entropy_models.py:26-324) inherit ``update()``,
``compress()`` and ``decompress()`` from CompressAI (``from compressai.entropy_models import EntropyModel,
GaussianConditional, EntropyBottleneck``, entropy_models.py:9; pip dependency, docker/Dockerfile:46, not vendored, no
version pinned).  Call sites on the reference side: ``update`` 43-48 / 194-197, ``compress`` 80-86 / 237-242,
``decompress`` 88-94 / 244-247, ``get_actual_bits`` 70-72 / 221-226.

PARITY UNPINNED: CompressAI is not available in this build environment; what follows restates its published algorithm
(compressai/entropy_models/entropy_models.py and cpp_exts/ops/ops.cpp, cpp_exts/rans/rans_interface.cpp, v1.1-1.2):
  * ``pmf_to_quantized_cdf``      — 16-bit CDF from a float pmf, zero-width bins repaired by stealing from the cheapest
  * ``eb_tables``                 — EntropyBottleneck.update(): support from the learned quantiles, pmf from the learned CDF
  * ``gaussian_tables``           — GaussianConditional.update(): one zero-mean table per entry of the scale table
  * ``build_indexes``             — GaussianConditional.build_indexes
  * ``element_intervals``         — BufferedRansEncoder.encode_with_indexes: table symbol + 4-bit bypass digits for escapes
CompressAI's own byte format (one rans64 stream per image) is NOT reproduced: the product codes the same symbol /
interval sequence with its rANS lanes (fvc_entropy.cu), restated byte-exactly in ``encode_indexed`` / ``decode_indexed``.
"""
from __future__ import annotations

import numpy as np
import torch

from . import dvc_oracle as O
from .entropy_oracle import MAGIC, RANS_L

PRECISION = 16
BYPASS_PRECISION = 4
MAX_BYPASS = (1 << BYPASS_PRECISION) - 1


def pmf_to_quantized_cdf(pmf, precision=PRECISION):
    """ops.cpp pmf_to_quantized_cdf: returns int64 [len(pmf) + 1], cdf[0] = 0, cdf[-1] = 2^precision, strictly increasing."""
    pmf = np.asarray(pmf, dtype=np.float32)
    assert np.all(np.isfinite(pmf)) and np.all(pmf >= 0)
    cdf = np.zeros(len(pmf) + 1, dtype=np.int64)
    # std::round on float: half away from zero (values are non-negative)
    cdf[1:] = np.floor(pmf.astype(np.float32) * np.float32(1 << precision) + np.float32(0.5)).astype(np.int64)
    total = int(cdf.sum())
    assert total > 0
    cdf = ((1 << precision) * cdf) // total
    cdf = np.cumsum(cdf)
    cdf[-1] = 1 << precision
    n = len(cdf)
    for i in range(n - 1):
        if cdf[i] == cdf[i + 1]:
            freqs = cdf[1:] - cdf[:-1]
            cand = np.where(freqs > 1)[0]
            assert len(cand) > 0
            best = int(cand[np.argmin(freqs[cand])])          # first minimum, as the C++ strict '<' scan keeps
            if best < i:
                cdf[best + 1:i + 1] -= 1
            else:
                assert best > i
                cdf[i + 1:best + 1] += 1
    assert cdf[0] == 0 and cdf[-1] == (1 << precision) and np.all(cdf[1:] > cdf[:-1])
    return cdf


def _pmf_to_cdf(pmf, tail_mass, pmf_length, max_length):
    """EntropyModel._pmf_to_cdf: int32 [ntab, max_length + 2], row i = quantised cdf of pmf[i, :len_i] + tail_mass[i]."""
    out = np.zeros((len(pmf_length), max_length + 2), dtype=np.int32)
    for i in range(len(pmf_length)):
        prob = np.concatenate([pmf[i, :pmf_length[i]], tail_mass[i]])
        c = pmf_to_quantized_cdf(prob)
        out[i, :len(c)] = c
    return out


def eb_tables(matrices, biases, factors, quantiles):
    """EntropyBottleneck.update(): (quantized_cdf, cdf_length, offset) as int32 numpy arrays.
    matrices/biases/factors: the ``_matrix{i}`` / ``_bias{i}`` / ``_factor{i}`` parameters, quantiles [C, 1, 3]."""
    q = quantiles.detach().float()
    medians = q[:, 0, 1]
    minima = torch.clamp(torch.ceil(medians - q[:, 0, 0]).int(), min=0)
    maxima = torch.clamp(torch.ceil(q[:, 0, 2] - medians).int(), min=0)
    offset = -minima
    pmf_start = medians - minima
    pmf_length = maxima + minima + 1
    max_length = int(pmf_length.max())
    samples = torch.arange(max_length)[None, :] + pmf_start[:, None, None]
    lower = O.eb_logits_cumulative(matrices, biases, factors, samples - 0.5)
    upper = O.eb_logits_cumulative(matrices, biases, factors, samples + 0.5)
    sign = -torch.sign(lower + upper)
    pmf = torch.abs(torch.sigmoid(sign * upper) - torch.sigmoid(sign * lower))[:, 0, :]
    tail_mass = torch.sigmoid(lower[:, 0, :1]) + torch.sigmoid(-upper[:, 0, -1:])
    cdf = _pmf_to_cdf(pmf.numpy(), tail_mass.numpy(), pmf_length.numpy(), max_length)
    return cdf, (pmf_length + 2).numpy().astype(np.int32), offset.numpy().astype(np.int32)


def gaussian_tables(scale_table, tail_mass=1e-9):
    """GaussianConditional.update() for a scale table [S]: (quantized_cdf, cdf_length, offset)."""
    import scipy.stats
    st = torch.as_tensor(scale_table, dtype=torch.float32)
    multiplier = -scipy.stats.norm.ppf(tail_mass / 2)
    pmf_center = torch.ceil(st * multiplier).int()
    pmf_length = 2 * pmf_center + 1
    max_length = int(pmf_length.max())
    samples = torch.abs(torch.arange(max_length).int() - pmf_center[:, None]).float()
    scale = st.unsqueeze(1).float()
    const = -(2 ** -0.5)
    upper = 0.5 * torch.erfc(const * ((0.5 - samples) / scale))
    lower = 0.5 * torch.erfc(const * ((-0.5 - samples) / scale))
    pmf = upper - lower
    tail = 2 * lower[:, :1]
    cdf = _pmf_to_cdf(pmf.numpy(), tail.numpy(), pmf_length.numpy(), max_length)
    return cdf, (pmf_length + 2).numpy().astype(np.int32), (-pmf_center).numpy().astype(np.int32)


def build_indexes(scales, scale_table, scale_bound=0.11):
    """GaussianConditional.build_indexes: number of table entries below each (lower-bounded) scale."""
    s = torch.clamp(scales, min=scale_bound)
    idx = torch.full(s.shape, len(scale_table) - 1, dtype=torch.int32)
    for t in scale_table[:-1]:
        idx -= (s <= t).int()
    return idx


def element_intervals(symbol, index, cdf, cdf_length, offset):
    """The (start, freq) intervals one element contributes, in DECODING order (rans_interface.cpp encode_with_indexes)."""
    T = cdf[index]
    maxv = int(cdf_length[index]) - 2
    v = int(symbol) - int(offset[index])
    raw = 0
    if v < 0:
        raw = -2 * v - 1
        v = maxv
    elif v >= maxv:
        raw = 2 * (v - maxv)
        v = maxv
    out = [(int(T[v]), int(T[v + 1]) - int(T[v]))]
    if v == maxv:
        nb = 0
        while (raw >> (nb * BYPASS_PRECISION)) != 0:
            nb += 1
        val = nb
        while val >= MAX_BYPASS:
            out.append((MAX_BYPASS << 12, 1 << 12))
            val -= MAX_BYPASS
        out.append((val << 12, 1 << 12))
        for j in range(nb):
            out.append((((raw >> (j * BYPASS_PRECISION)) & MAX_BYPASS) << 12, 1 << 12))
    return out


def ideal_bits(symbols, indexes, cdf, cdf_length, offset):
    tot = 0.0
    for s, i in zip(symbols, indexes):
        for _, f in element_intervals(s, i, cdf, cdf_length, offset):
            tot += 16.0 - np.log2(float(f))
    return tot


def encode_indexed(symbols, indexes, cdf, cdf_length, offset, lane_len):
    """FVR1 container of the indexed-table coder, byte-exact restatement of k_rans_encode_indexed / k_rans_pack."""
    n = len(symbols)
    nlanes = (n + lane_len - 1) // lane_len
    lanes = []
    for l in range(nlanes):
        a, b = l * lane_len, min(n, (l + 1) * lane_len)
        x = RANS_L
        words = []
        for k in range(b - 1, a - 1, -1):
            for s, f in reversed(element_intervals(symbols[k], indexes[k], cdf, cdf_length, offset)):
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


def decode_indexed(stream, indexes, cdf, cdf_length, offset, lane_len):
    n = len(indexes)
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
        st = {"x": (int(w[0]) << 16) | int(w[1]), "p": 2}

        def advance(s, f):
            x = f * (st["x"] >> 16) + (st["x"] & 0xFFFF) - s
            if x < RANS_L:
                x = (x << 16) | int(w[st["p"]])
                st["p"] += 1
            st["x"] = x

        def digit():
            d = (st["x"] & 0xFFFF) >> 12
            advance(d << 12, 1 << 12)
            return d

        for k in range(l * lane_len, min(n, (l + 1) * lane_len)):
            T = cdf[indexes[k]]
            maxv = int(cdf_length[indexes[k]]) - 2
            slot = st["x"] & 0xFFFF
            v = int(np.searchsorted(T[:maxv + 2], slot, side="right")) - 1
            v = min(v, maxv)
            advance(int(T[v]), int(T[v + 1]) - int(T[v]))
            if v == maxv:
                d = digit()
                nb = d
                while d == MAX_BYPASS:
                    d = digit()
                    nb += d
                raw = 0
                for j in range(nb):
                    raw |= digit() << (j * BYPASS_PRECISION)
                v = raw >> 1
                v = -v - 1 if raw & 1 else v + maxv
            out[k] = v + int(offset[indexes[k]])
    return out
