"""Mirror of the likelihood / bit-estimation forward of the reference ``entropy_models.py``.

The reference builds these classes on CompressAI (``EntropyBottleneck``, ``GaussianConditional``,
``CompressionModel``), an un-vendored, un-pinned dependency (docker/Dockerfile:46).  This module keeps the
reference's class names, constructor arguments, ``forward`` / ``get_estimate_bits`` signatures and the
CompressAI parameter and buffer names (``_matrix{0-4}``, ``_bias{0-4}``, ``_factor{0-3}``, ``quantiles``, ``target``
[3], ``_offset`` / ``_quantized_cdf`` / ``_cdf_length``, ``likelihood_lower_bound.bound``, ``lower_bound_scale.bound``,
``scale_table``, ``scale_bound``; as published for CompressAI 1.1-1.2, restated from memory) so that checkpoints load, and evaluates the likelihoods with the library's fused CUDA kernels
(``fvc_eb_forward`` / ``fvc_gaussian_forward``) and the hyper-prior convolutions with the tcgen05 engine.

In scope (eval): ``RecProbModel.forward`` without the recurrent prior (entropy_models.py:55-68, ``RPM_flag=False``)
and with externally supplied RPM outputs, ``MeanScaleHyperPriors.forward`` (202-219), both ``get_estimate_bits``
(74-78, 228-235), and the range-coding side: ``update()`` (43-48, 194-197: CompressAI's quantised CDF
tables ``_quantized_cdf`` / ``_cdf_length`` / ``_offset`` from the learned densities), ``compress`` / ``decompress``
(80-94, 237-247), ``compress_slow`` / ``decompress_slow`` (97-147, 250-324) and ``get_actual_bits`` (70-72, 221-226).
The coder is the library's rANS-lane kernel over the same table / index / escape model CompressAI hands its range coder
(``fvc_entropy_encode_indexed``); the byte strings are this library's "FVR1" containers, not CompressAI's rans64 format
(sizes agree to within the per-lane state overhead; only ``len()`` and the round trip are used by the reference).
``RPM`` / ``ConvLSTM`` (328-378, the recurrent prior network; pinned against the reference's own classes,
tests/golden/rpm_128.npz) run their 3x3 convolutions on the tcgen05 engine.
Out of scope: training-mode noise, ``aux_loss`` / ``loss`` (training of the quantiles).

PARITY UNPINNED at this boundary: CompressAI is not available to check against; the algorithm follows its
published ``EntropyBottleneck._logits_cumulative/_likelihood`` and ``GaussianConditional._likelihood``
(oracle/dvc_oracle.py: eb_forward / gaussian_forward).
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn as nn

from . import ops

SCALES_MIN = 0.11
SCALES_MAX = 256
SCALES_LEVELS = 64


def get_scale_table(min=SCALES_MIN, max=SCALES_MAX, levels=SCALES_LEVELS):
    """reference entropy_models.py:22-23."""
    return torch.exp(torch.linspace(math.log(min), math.log(max), levels))


class _LowerBound(nn.Module):
    """State-dict shape of CompressAI's ``LowerBound`` (one buffer ``bound``); the bound itself is applied inside the
    CUDA kernels."""

    def __init__(self, bound):
        super().__init__()
        self.register_buffer("bound", torch.Tensor([float(bound)]))


def pmf_to_quantized_cdf(pmf, precision=16):
    """CompressAI ``pmf_to_quantized_cdf`` (cpp_exts/ops/ops.cpp): 16-bit CDF of a float pmf whose last entry is the
    tail mass; every bin keeps a non-zero width (a zero-width bin takes one count from the narrowest bin wider than 1)."""
    p = np.asarray(pmf, dtype=np.float32)
    if p.ndim != 1 or p.size == 0 or not np.all(np.isfinite(p)) or np.any(p < 0):
        raise ValueError("pmf must be a non-empty vector of finite, non-negative values")
    freq = np.floor(p * np.float32(1 << precision) + np.float32(0.5)).astype(np.int64)
    total = int(freq.sum())
    if total == 0:
        raise ValueError("pmf sums to zero")
    cdf = np.concatenate([[0], np.cumsum(((1 << precision) * freq) // total)])
    cdf[-1] = 1 << precision
    for i in np.flatnonzero(cdf[1:] == cdf[:-1]):     # ascending; earlier repairs may already have widened bin i
        if cdf[i] != cdf[i + 1]:
            continue
        width = cdf[1:] - cdf[:-1]
        wide = np.flatnonzero(width > 1)
        if wide.size == 0:
            raise ValueError("more symbols than 2^%d counts" % precision)
        steal = int(wide[np.argmin(width[wide])])
        if steal < i:
            cdf[steal + 1:i + 1] -= 1
        else:
            cdf[i + 1:steal + 1] += 1
    return cdf


class _EntropyModelBuffers(nn.Module):
    """The buffers every CompressAI ``EntropyModel`` registers (``_offset``, ``_quantized_cdf``, ``_cdf_length``: the
    range-coder tables written by ``update()``, empty until then) and ``likelihood_lower_bound.bound``, so that a
    CompressAI-layout ``state_dict`` loads with ``strict=True``.  Table buffers take the checkpoint's shape on load,
    as CompressAI's ``update_registered_buffers`` does."""

    _RESIZABLE = ("_offset", "_quantized_cdf", "_cdf_length", "scale_table")

    def _register_entropy_buffers(self, likelihood_bound):
        self.likelihood_lower_bound = _LowerBound(likelihood_bound)
        self.register_buffer("_offset", torch.IntTensor())
        self.register_buffer("_quantized_cdf", torch.IntTensor())
        self.register_buffer("_cdf_length", torch.IntTensor())

    lane_len = 8192          # symbols per rANS lane of compress() / decompress()

    def _pmf_to_cdf(self, pmf, tail_mass, pmf_length, max_length):
        """CompressAI ``EntropyModel._pmf_to_cdf``: int32 [ntab, max_length + 2], one quantised CDF per row."""
        pmf, tail_mass, pmf_length = pmf.detach().cpu().numpy(), tail_mass.detach().cpu().numpy(), pmf_length.cpu().numpy()
        cdf = np.zeros((len(pmf_length), max_length + 2), dtype=np.int32)
        for i, n in enumerate(pmf_length):
            row = pmf_to_quantized_cdf(np.concatenate([pmf[i, :n], tail_mass[i]]))
            cdf[i, :len(row)] = row
        return torch.from_numpy(cdf)

    def _check_tables(self):
        if self._offset.numel() == 0 or self._quantized_cdf.dim() != 2 or self._cdf_length.numel() != self._quantized_cdf.shape[0]:
            raise ValueError("Uninitialized CDFs. Run update() first")

    def compress(self, inputs, indexes, means=None):
        """CompressAI ``EntropyModel.compress``: one byte string per batch element; symbols = round(inputs - means),
        coded in the element's flattened order under ``_quantized_cdf[indexes]``."""
        self._check_tables()
        if inputs.dim() < 2 or tuple(inputs.shape) != tuple(indexes.shape):
            raise ValueError("inputs and indexes must have the same shape [B, C, ...]")
        sym = inputs.detach()
        if means is not None:
            sym = sym - means
        sym = torch.round(sym).to(torch.int32)
        return [ops.entropy_encode_indexed(sym[i].reshape(-1), indexes[i].reshape(-1), self._quantized_cdf, self._cdf_length,
                                           self._offset, self.lane_len) for i in range(sym.shape[0])]

    def decompress(self, strings, indexes, dtype=torch.float, means=None):
        """CompressAI ``EntropyModel.decompress``: symbols back from the strings, plus ``means``."""
        self._check_tables()
        if not isinstance(strings, (tuple, list)) or len(strings) != indexes.shape[0]:
            raise ValueError("one string per batch element of indexes is required")
        if means is not None and (means.shape[:2] != indexes.shape[:2]):
            raise ValueError("means must match indexes in batch and channel size")
        out = torch.stack([ops.entropy_decode_indexed(s, indexes[i].reshape(-1), self._quantized_cdf, self._cdf_length,
                                                      self._offset, self.lane_len).reshape(indexes[i].shape)
                           for i, s in enumerate(strings)]).to(dtype)
        return out + means if means is not None else out

    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        for name in self._RESIZABLE:
            key = prefix + name
            if key in state_dict and hasattr(self, name) and getattr(self, name).shape != state_dict[key].shape:
                buf = getattr(self, name)
                setattr(self, name, torch.empty(state_dict[key].shape, dtype=state_dict[key].dtype, device=buf.device))
        return super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)


class EntropyBottleneck(_EntropyModelBuffers):
    """CompressAI-compatible factorized prior (filters (3,3,3,3), init_scale 10, tail_mass 1e-9).

    Eval forward: ``x_hat = round(x - median) + median``; likelihood from the learned cumulative,
    lower-bounded at 1e-9.  Returns ``(x_hat, likelihood)`` like CompressAI.
    """

    def __init__(self, channels, tail_mass=1e-9, init_scale=10.0, filters=(3, 3, 3, 3), likelihood_bound=1e-9):
        super().__init__()
        if tuple(filters) != (3, 3, 3, 3):
            raise ValueError("the CUDA kernel is specialised for filters (3,3,3,3)")
        self.channels = int(channels)
        self.filters = tuple(int(f) for f in filters)
        self.init_scale = float(init_scale)
        self.tail_mass = float(tail_mass)
        self.likelihood_bound = float(likelihood_bound)
        f = (1,) + self.filters + (1,)
        scale = self.init_scale ** (1 / (len(self.filters) + 1))
        for i in range(len(self.filters) + 1):
            init = np.log(np.expm1(1 / scale / f[i + 1]))
            self.register_parameter("_matrix%d" % i, nn.Parameter(torch.full((channels, f[i + 1], f[i]), float(init))))
            self.register_parameter("_bias%d" % i,
                                    nn.Parameter(torch.empty(channels, f[i + 1], 1).uniform_(-0.5, 0.5)))
            if i < len(self.filters):
                self.register_parameter("_factor%d" % i, nn.Parameter(torch.zeros(channels, f[i + 1], 1)))
        init_q = torch.tensor([-self.init_scale, 0.0, self.init_scale])
        self.quantiles = nn.Parameter(init_q.repeat(channels, 1, 1))
        target = math.log(2 / self.tail_mass - 1)
        self.register_buffer("target", torch.Tensor([-target, 0, target]))      # CompressAI: shape [3]
        self._register_entropy_buffers(likelihood_bound)

    def _medians(self):
        return self.quantiles[:, 0, 1].detach()

    def _packed(self):
        n = len(self.filters)
        return ops.pack_eb_params([getattr(self, "_matrix%d" % i).detach() for i in range(n + 1)],
                                  [getattr(self, "_bias%d" % i).detach() for i in range(n + 1)],
                                  [getattr(self, "_factor%d" % i).detach() for i in range(n)])

    def forward(self, x, training=None):
        if training if training is not None else self.training:
            raise NotImplementedError("training-mode (additive noise) likelihoods are outside the hot path")
        xh, lik, _ = ops.eb_forward(x, self._packed().to(x.device), self._medians().to(x.device))
        return xh, lik

    def forward_bits(self, x):
        """(x_hat, likelihood, sum clamp(-log2(lik + 1e-5), 0, 50)) in one fused kernel."""
        return ops.eb_forward(x, self._packed().to(x.device), self._medians().to(x.device))

    def _logits_cumulative(self, inputs):
        """CompressAI ``EntropyBottleneck._logits_cumulative`` on [C, 1, N] sample points (table construction only)."""
        logits = inputs
        n = len(self.filters)
        for i in range(n + 1):
            logits = torch.matmul(torch.nn.functional.softplus(getattr(self, "_matrix%d" % i).detach()), logits)
            logits = logits + getattr(self, "_bias%d" % i).detach()
            if i < n:
                logits = logits + torch.tanh(getattr(self, "_factor%d" % i).detach()) * torch.tanh(logits)
        return logits

    def update(self, force=False):
        """CompressAI ``EntropyBottleneck.update``: per channel, the support [median - ceil(median - q_lo),
        median + ceil(q_hi - median)] from the learned quantiles, its pmf from the learned CDF, the two tails as the
        escape bin, quantised to 16 bits.  Returns False when tables exist and ``force`` is not set."""
        if self._offset.numel() > 0 and not force:
            return False
        q = self.quantiles.detach()
        medians = q[:, 0, 1]
        minima = torch.clamp(torch.ceil(medians - q[:, 0, 0]).int(), min=0)
        maxima = torch.clamp(torch.ceil(q[:, 0, 2] - medians).int(), min=0)
        pmf_start = medians - minima
        pmf_length = maxima + minima + 1
        max_length = int(pmf_length.max().item())
        samples = torch.arange(max_length, device=q.device)[None, :] + pmf_start[:, None, None]
        lower = self._logits_cumulative(samples - 0.5)
        upper = self._logits_cumulative(samples + 0.5)
        sign = -torch.sign(lower + upper)
        pmf = torch.abs(torch.sigmoid(sign * upper) - torch.sigmoid(sign * lower))[:, 0, :]
        tail_mass = torch.sigmoid(lower[:, 0, :1]) + torch.sigmoid(-upper[:, 0, -1:])
        dev = self._offset.device
        self._quantized_cdf = self._pmf_to_cdf(pmf, tail_mass, pmf_length, max_length).to(dev)
        self._offset = (-minima).to(torch.int32).to(dev)
        self._cdf_length = (pmf_length + 2).to(torch.int32).to(dev)
        return True

    def _build_indexes(self, size):
        """channel index of every element of a [B, C, ...] tensor."""
        view = [1, size[1]] + [1] * (len(size) - 2)
        return torch.arange(size[1], dtype=torch.int32, device=self._offset.device).view(view).expand(*size)

    def _expanded_medians(self, size):
        return self._medians().view([1, size[1]] + [1] * (len(size) - 2)).expand(*size)

    def compress(self, x):
        """CompressAI ``EntropyBottleneck.compress``: x [B, C, ...] -> one string per batch element."""
        idx = self._build_indexes(tuple(x.shape)).to(x.device)
        return super().compress(x, idx, self._expanded_medians(tuple(x.shape)).to(x.device))

    def decompress(self, strings, size):
        """CompressAI ``EntropyBottleneck.decompress``: ``size`` = the spatial dims; returns [len(strings), C, *size]."""
        self._check_tables()
        full = (len(strings), self._quantized_cdf.size(0)) + tuple(size)
        idx = self._build_indexes(full)
        return super().decompress(strings, idx, torch.float, self._expanded_medians(full).to(idx.device))


class GaussianConditional(_EntropyModelBuffers):
    """CompressAI-compatible conditional Gaussian: scale lower bound 0.11, likelihood bound 1e-9."""

    def __init__(self, scale_table=None, scale_bound=0.11, tail_mass=1e-9, likelihood_bound=1e-9):
        super().__init__()
        if scale_bound != 0.11 or likelihood_bound != 1e-9:
            raise ValueError("the CUDA kernel is specialised for scale_bound 0.11 and likelihood_bound 1e-9")
        self.tail_mass = float(tail_mass)
        self._register_entropy_buffers(likelihood_bound)
        self.lower_bound_scale = _LowerBound(scale_bound)
        self.register_buffer("scale_table", torch.as_tensor(scale_table, dtype=torch.float32)
                             if scale_table is not None else torch.Tensor())
        self.register_buffer("scale_bound", torch.Tensor([float(scale_bound)]))

    def update_scale_table(self, scale_table, force=False):
        """CompressAI ``GaussianConditional.update_scale_table``: (re)build the tables for a scale table [S]."""
        if self._offset.numel() > 0 and not force:
            return False
        st = torch.as_tensor(scale_table, dtype=torch.float32)
        if st.dim() != 1 or st.numel() < 1 or bool((st <= 0).any()) or bool((st[1:] <= st[:-1]).any()):
            raise ValueError("scale_table must be a sorted vector of positive scales")
        self.scale_table = st.to(self.scale_table.device)
        self.update()
        return True

    def update(self):
        """CompressAI ``GaussianConditional.update``: per scale s a zero-mean table over [-c, c], c = ceil(s * m) with
        m the normal quantile of tail_mass / 2; pmf = Phi((0.5 - |k|) / s) - Phi((-0.5 - |k|) / s), escape bin 2 * lower tail."""
        import scipy.stats
        multiplier = -scipy.stats.norm.ppf(self.tail_mass / 2)
        st = self.scale_table.float()
        pmf_center = torch.ceil(st * multiplier).int()
        pmf_length = 2 * pmf_center + 1
        max_length = int(pmf_length.max().item())
        samples = torch.abs(torch.arange(max_length, device=st.device).int() - pmf_center[:, None]).float()
        scale = st.unsqueeze(1)
        const = -(2 ** -0.5)
        upper = 0.5 * torch.erfc(const * ((0.5 - samples) / scale))
        lower = 0.5 * torch.erfc(const * ((-0.5 - samples) / scale))
        dev = self._offset.device
        self._quantized_cdf = self._pmf_to_cdf(upper - lower, 2 * lower[:, :1], pmf_length, max_length).to(dev)
        self._offset = (-pmf_center).to(torch.int32).to(dev)
        self._cdf_length = (pmf_length + 2).to(torch.int32).to(dev)

    def build_indexes(self, scales):
        """CompressAI ``GaussianConditional.build_indexes``: table index = number of table scales below the (lower-bounded)
        scale."""
        if self.scale_table.numel() == 0:
            raise ValueError("empty scale table: run update_scale_table() first")
        s = torch.clamp(scales.detach(), min=float(self.scale_bound))
        table = self.scale_table.to(s.device)
        # sum over the table of (s <= t), for all but the last entry == searchsorted from the left on the sorted table
        below = torch.searchsorted(table[:-1].contiguous(), s.contiguous(), right=False)
        return below.to(torch.int32)

    def forward(self, x, scales, means=None, training=None):
        if training if training is not None else self.training:
            raise NotImplementedError("training-mode (additive noise) likelihoods are outside the hot path")
        xh, lik, _ = ops.gaussian_forward(x, scales, means)
        return xh, lik


def _estimate_bits_clamped(likelihoods):
    """reference entropy_models.py:74-78."""
    return torch.sum(torch.clamp(-1.0 * torch.log(likelihoods + 1e-5) / math.log(2.0), 0, 50))


class _SplitConv(nn.Conv2d):
    """nn.Conv2d(cin, cout, 3, 1, 1) evaluated on the tcgen05 engine, which takes at most 128 input and 128 output
    channels per launch: wider layers (the 2C -> 4C gate convolution of the ConvLSTM, C -> 2C of RPM.conv8) run as
    blocks of 128 output channels, each the fp32 sum over blocks of 128 input channels.  The weight blocks are cut once
    per weight version, so the op-level handle cache (ops.conv2d) keeps their packed form."""

    def __init__(self, cin, cout):
        super().__init__(cin, cout, kernel_size=3, stride=1, padding=1)
        self._blocks = None

    def _weight_blocks(self):
        key = (self.weight.data_ptr(), self.weight._version, self.bias.data_ptr(), self.bias._version)
        if self._blocks is None or self._blocks[0] != key:
            w, b = self.weight.detach(), self.bias.detach()
            outs = []
            for o in range(0, w.shape[0], 128):
                ins = [w[o:o + 128, i:i + 128].contiguous() for i in range(0, w.shape[1], 128)]
                outs.append((ins, b[o:o + 128].contiguous()))
            self._blocks = (key, outs)
        return self._blocks[1]

    def forward(self, x, act=ops.ACT_NONE):
        outs = []
        for ins, b in self._weight_blocks():
            if len(ins) == 1:
                outs.append(ops.conv2d(x, ins[0], b, 1, act))
                continue
            y = None
            for k, w in enumerate(ins):
                part = ops.conv2d(x[:, 128 * k:128 * (k + 1)].contiguous(), w, b if k == 0 else None, 1, ops.ACT_NONE)
                y = part if y is None else y + part
            outs.append(torch.relu(y) if act == ops.ACT_RELU else y)
            if act not in (ops.ACT_NONE, ops.ACT_RELU):
                raise NotImplementedError("only ReLU is applied after a split-input convolution")
        return outs[0] if len(outs) == 1 else torch.cat(outs, dim=1)


class ConvLSTM(nn.Module):
    """reference entropy_models.py:359-378: one 3x3 convolution of cat(x, h) to the four gates (j, i, f, o)."""

    def __init__(self, channels=128, forget_bias=1.0, activation=torch.relu):
        super().__init__()
        self.conv = _SplitConv(2 * channels, 4 * channels)
        self._forget_bias = forget_bias
        self._activation = activation
        self._channels = channels

    def forward(self, x, state):
        c, h = torch.split(state, self._channels, dim=1)
        y = self.conv(torch.cat((x, h), dim=1).contiguous())
        j, i, f, o = torch.split(y, self._channels, dim=1)
        f = torch.sigmoid(f + self._forget_bias)
        i = torch.sigmoid(i)
        c = c * f + i * self._activation(j)
        o = torch.sigmoid(o)
        h = o * self._activation(c)
        return h, torch.cat((c, h), dim=1)


class RPM(nn.Module):
    """reference entropy_models.py:328-357: the recurrent prior network of RecProbModel — four 3x3 conv + ReLU, the
    ConvLSTM, three 3x3 conv + ReLU, conv C -> 2C + ReLU split into (sigma, mu).  Same parameter names as the reference
    (``conv1..conv8``, ``lstm.conv``): its checkpoints load."""

    def __init__(self, channels=128, act=torch.tanh):
        super().__init__()
        for i in range(1, 8):
            setattr(self, "conv%d" % i, _SplitConv(channels, channels))
        self.conv8 = _SplitConv(channels, 2 * channels)
        self.channels = channels
        self.lstm = ConvLSTM(channels)

    def forward(self, x, hidden):
        x = x.contiguous()
        for i in range(1, 5):
            x = getattr(self, "conv%d" % i)(x, ops.ACT_RELU)
        x, hidden = self.lstm(x, hidden.to(x.device))
        for i in range(5, 8):
            x = getattr(self, "conv%d" % i)(x.contiguous(), ops.ACT_RELU)
        sigma_mu = self.conv8(x, ops.ACT_RELU)
        sigma, mu = torch.split(sigma_mu, self.channels, dim=1)
        return sigma, mu, hidden


class RecProbModel(nn.Module):
    """reference entropy_models.py:26-148 (forward / get_estimate_bits only).

    ``RPM_flag=False``: factorized ``entropy_bottleneck``.  ``RPM_flag=True``: the recurrent prior network ``RPM``
    (entropy_models.py:328-378) supplies sigma / mu of the conditional-Gaussian branch (58-63); ``rpm=`` replaces it by
    any callable ``rpm(prior_latent, rpm_hidden) -> (sigma, mu, rpm_hidden)``.
    """

    def __init__(self, channels, rpm=None):
        super().__init__()
        self.channels = int(channels)
        self.entropy_bottleneck = EntropyBottleneck(channels)
        self.gaussian_conditional = GaussianConditional(None)
        self.sigma = self.mu = self.prior_latent = None
        self.RPM = rpm if rpm is not None else RPM(channels)
        self.RPM_flag = False

    def set_RPM(self, RPM_flag):
        self.RPM_flag = RPM_flag

    def forward(self, x, rpm_hidden, training=None, prior_latent=None):
        if self.RPM_flag:
            assert prior_latent is not None, 'prior latent is none!'
            self.sigma, self.mu, rpm_hidden = self.RPM(prior_latent, rpm_hidden.to(x.device))
            self.sigma = torch.maximum(self.sigma, torch.FloatTensor([-7.0]).to(x.device))
            self.sigma = torch.exp(self.sigma) / 10
            x_hat, likelihood = self.gaussian_conditional(x, self.sigma, means=self.mu, training=training)
        else:
            x_hat, likelihood = self.entropy_bottleneck(x, training=training)
        prior_latent = torch.round(x).detach()
        return x_hat, likelihood, rpm_hidden.detach(), prior_latent

    def get_estimate_bits(self, likelihoods):
        return _estimate_bits_clamped(likelihoods)

    def update(self, scale_table=None, force=False):
        """reference entropy_models.py:43-48."""
        if scale_table is None:
            scale_table = get_scale_table()
        updated = self.gaussian_conditional.update_scale_table(scale_table, force=force)
        updated |= self.entropy_bottleneck.update(force=force)
        return updated

    def get_actual_bits(self, string):
        """reference entropy_models.py:70-72."""
        return torch.FloatTensor([len(b''.join(string)) * 8]).squeeze(0)

    def _rpm_params(self, prior_latent, rpm_hidden):
        sigma, mu, rpm_hidden = self.RPM(prior_latent, rpm_hidden.to(prior_latent.device))
        sigma = torch.maximum(sigma, torch.FloatTensor([-7.0]).to(sigma.device))
        return torch.exp(sigma) / 10, mu, rpm_hidden

    def compress(self, x):
        """reference entropy_models.py:80-86 (uses sigma / mu of the last forward when RPM_flag is set)."""
        if self.RPM_flag:
            indexes = self.gaussian_conditional.build_indexes(self.sigma)
            return self.gaussian_conditional.compress(x, indexes, means=self.mu)
        return self.entropy_bottleneck.compress(x)

    def decompress(self, string, shape):
        """reference entropy_models.py:88-94."""
        if self.RPM_flag:
            indexes = self.gaussian_conditional.build_indexes(self.sigma)
            return self.gaussian_conditional.decompress(string, indexes, means=self.mu)
        return self.entropy_bottleneck.decompress(string, shape)

    def compress_slow(self, x, rpm_hidden, prior_latent):
        """reference entropy_models.py:97-124: (x_hat, string, rpm_hidden, prior_latent); times in eNet_t / eAC_t / enc_t."""
        import time
        self.eAC_t = self.eNet_t = 0
        if self.RPM_flag:
            assert prior_latent is not None, 'prior latent is none!'
            t_0 = time.perf_counter()
            sigma, mu, rpm_hidden = self._rpm_params(prior_latent, rpm_hidden)
            self.eNet_t += time.perf_counter() - t_0
            t_0 = time.perf_counter()
            indexes = self.gaussian_conditional.build_indexes(sigma)
            string = self.gaussian_conditional.compress(x, indexes, means=mu)
            x_hat, _ = self.gaussian_conditional(x, sigma.contiguous(), means=mu.contiguous(), training=self.training)
            self.eAC_t += time.perf_counter() - t_0
        else:
            t_0 = time.perf_counter()
            string = self.entropy_bottleneck.compress(x)
            x_hat, _ = self.entropy_bottleneck(x, training=self.training)
            self.eAC_t += time.perf_counter() - t_0
        prior_latent = torch.round(x_hat).detach()
        self.enc_t = self.eNet_t + self.eAC_t
        return x_hat, string, rpm_hidden.detach(), prior_latent

    def decompress_slow(self, string, shape, rpm_hidden, prior_latent):
        """reference entropy_models.py:126-147: (x_hat, rpm_hidden, prior_latent); times in dnet_t / dAC_t / dec_t."""
        import time
        self.dAC_t = self.dnet_t = 0
        if self.RPM_flag:
            assert prior_latent is not None, 'prior latent is none!'
            t_0 = time.perf_counter()
            sigma, mu, rpm_hidden = self._rpm_params(prior_latent, rpm_hidden)
            self.dnet_t += time.perf_counter() - t_0
            t_0 = time.perf_counter()
            indexes = self.gaussian_conditional.build_indexes(sigma)
            x_hat = self.gaussian_conditional.decompress(string, indexes, means=mu)
            self.dAC_t += time.perf_counter() - t_0
        else:
            t_0 = time.perf_counter()
            x_hat = self.entropy_bottleneck.decompress(string, shape)
            self.dAC_t += time.perf_counter() - t_0
        prior_latent = torch.round(x_hat).detach()
        self.dec_t = self.dnet_t + self.dAC_t
        return x_hat, rpm_hidden.detach(), prior_latent


class _HyperConv(nn.Sequential):
    """Two 3x3 convolutions as in entropy_models.py:165-188; LeakyReLU() default slope 0.01."""

    def __init__(self, channels, out_channels, last_act):
        super().__init__(nn.Conv2d(channels, channels, kernel_size=3, stride=1, padding=1),
                         nn.LeakyReLU(inplace=True),
                         nn.Conv2d(channels, out_channels, kernel_size=3, stride=1, padding=1),
                         *([nn.LeakyReLU(inplace=True)] if last_act else []))
        self.last_act = last_act

    def forward(self, x):
        # tcgen05 engine (op-level entry point: weights are packed per call; fine for this 1/16-resolution path)
        x = ops.conv2d(x, self[0].weight.detach(), self[0].bias.detach(), 1, ops.ACT_LRELU001)
        return ops.conv2d(x, self[2].weight.detach(), self[2].bias.detach(), 1,
                          ops.ACT_LRELU001 if self.last_act else ops.ACT_NONE)


class MeanScaleHyperPriors(nn.Module):
    """reference entropy_models.py:150-324 (forward / get_estimate_bits only)."""

    def __init__(self, channels, entropy_trick=True):
        super().__init__()
        self.channels = int(channels)
        self.entropy_bottleneck = EntropyBottleneck(channels)
        self.gaussian_conditional = GaussianConditional(None)
        self.sigma = self.mu = self.z_string = None
        self.h_a1 = _HyperConv(channels, channels, True)
        self.h_a2 = _HyperConv(channels, channels, False)
        self.h_s1 = _HyperConv(channels, channels, True)
        self.h_s2 = _HyperConv(channels, channels * 2, False)
        self.scale_table = get_scale_table()
        self.entropy_trick = entropy_trick

    def forward(self, x, training=None):
        z = self.h_a1(x)
        z = self.h_a2(z)
        z_hat, z_likelihood = self.entropy_bottleneck(z, training=training)
        self.z = z
        g = self.h_s1(z_hat)
        gaussian_params = self.h_s2(g)
        self.sigma, self.mu = torch.split(gaussian_params, self.channels, dim=1)
        self.sigma = torch.maximum(self.sigma, torch.FloatTensor([-7.0]).to(x.device))
        self.sigma = torch.exp(self.sigma)
        x_hat, x_likelihood = self.gaussian_conditional(x, self.sigma.contiguous(), means=self.mu.contiguous(),
                                                        training=training)
        return x_hat, (x_likelihood, z_likelihood)

    def update(self, scale_table=None, force=False):
        """reference entropy_models.py:194-197."""
        updated = self.gaussian_conditional.update_scale_table(self.scale_table, force=force)
        updated |= self.entropy_bottleneck.update(force=force)
        return updated

    def get_actual_bits(self, string):
        """reference entropy_models.py:221-226."""
        (x_string, z_string) = string
        x_act = torch.FloatTensor([len(s) * 8 for s in x_string])
        z_act = torch.FloatTensor([len(s) * 8 for s in z_string])
        return x_act + z_act

    def _gaussian_params(self, z_hat):
        g = self.h_s1(z_hat)
        sigma, mu = torch.split(self.h_s2(g), self.channels, dim=1)
        sigma = torch.maximum(sigma, torch.FloatTensor([-7.0]).to(sigma.device))
        return torch.exp(sigma).contiguous(), mu.contiguous()

    def compress(self, x):
        """reference entropy_models.py:237-242: the fast path, with z / sigma / mu kept by the last forward."""
        z_string = self.entropy_bottleneck.compress(self.z)
        indexes = self.gaussian_conditional.build_indexes(self.sigma)
        x_string = self.gaussian_conditional.compress(x, indexes, means=self.mu)
        return (x_string, z_string)

    def decompress(self, string, shape):
        """reference entropy_models.py:244-247."""
        indexes = self.gaussian_conditional.build_indexes(self.sigma)
        return self.gaussian_conditional.decompress(string[0], indexes, means=self.mu)

    def compress_slow(self, x, decode=False):
        """reference entropy_models.py:250-294: (x_hat or None, (x_string, z_string), z_size).  With ``entropy_trick`` the
        batch is folded into one string ([B,C,H,W] -> [1,C,B,H,W])."""
        import time
        self.eAC_t = self.eNet_t = 0
        t_0 = time.perf_counter()
        z = self.h_a2(self.h_a1(x))
        self.eNet_t += time.perf_counter() - t_0
        t_0 = time.perf_counter()
        z_hat, _ = self.entropy_bottleneck(z, training=self.training)
        self.eAC_t += time.perf_counter() - t_0
        t_0 = time.perf_counter()
        sigma, mu = self._gaussian_params(z_hat)
        self.eNet_t += time.perf_counter() - t_0
        t_0 = time.perf_counter()
        x_hat = self.gaussian_conditional(x, sigma, means=mu, training=self.training)[0] if decode else None
        if self.entropy_trick:
            z = z.permute(1, 0, 2, 3).unsqueeze(0).contiguous()
            z_size = z.size()[-3:]
        else:
            z_size = z.size()[-2:]
        z_string = self.entropy_bottleneck.compress(z)
        indexes = self.gaussian_conditional.build_indexes(sigma)
        if self.entropy_trick:
            x = x.permute(1, 0, 2, 3).unsqueeze(0).contiguous()
            indexes = indexes.permute(1, 0, 2, 3).unsqueeze(0).contiguous()
            mu = mu.permute(1, 0, 2, 3).unsqueeze(0).contiguous()
        x_string = self.gaussian_conditional.compress(x, indexes, means=mu)
        self.eAC_t += time.perf_counter() - t_0
        self.enc_t = self.eNet_t + self.eAC_t
        return x_hat, (x_string, z_string), z_size

    def decompress_slow(self, string, shape):
        """reference entropy_models.py:296-324: ``shape`` = the z_size compress_slow returned."""
        import time
        self.dAC_t = self.dnet_t = 0
        t_0 = time.perf_counter()
        z_hat = self.entropy_bottleneck.decompress(string[1], shape)
        if self.entropy_trick:
            z_hat = z_hat.squeeze(0).permute(1, 0, 2, 3).contiguous()
        self.dAC_t += time.perf_counter() - t_0
        t_0 = time.perf_counter()
        sigma, mu = self._gaussian_params(z_hat)
        self.dnet_t += time.perf_counter() - t_0
        t_0 = time.perf_counter()
        indexes = self.gaussian_conditional.build_indexes(sigma)
        if self.entropy_trick:
            indexes = indexes.permute(1, 0, 2, 3).unsqueeze(0).contiguous()
            mu = mu.permute(1, 0, 2, 3).unsqueeze(0).contiguous()
        x_hat = self.gaussian_conditional.decompress(string[0], indexes, means=mu)
        if self.entropy_trick:
            x_hat = x_hat.squeeze(0).permute(1, 0, 2, 3).contiguous()
        self.dAC_t += time.perf_counter() - t_0
        self.dec_t = self.dnet_t + self.dAC_t
        return x_hat

    def get_estimate_bits(self, likelihoods):
        """reference entropy_models.py:228-235 (plain log2 sum per batch element)."""
        (x_likelihood, z_likelihood) = likelihoods
        log2 = torch.log(torch.FloatTensor([2])).squeeze(0).to(x_likelihood.device)
        bs = x_likelihood.size(0)
        x_est = torch.sum(torch.log(x_likelihood.view(bs, -1)), dim=-1) / (-log2)
        z_est = torch.sum(torch.log(z_likelihood.view(bs, -1)), dim=-1) / (-log2)
        return x_est + z_est
