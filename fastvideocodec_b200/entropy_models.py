"""Mirror of the likelihood / bit-estimation forward of the reference ``entropy_models.py``.

The reference builds these classes on CompressAI (``EntropyBottleneck``, ``GaussianConditional``,
``CompressionModel``), an un-vendored, un-pinned dependency (docker/Dockerfile:46).  This module keeps the
reference's class names, constructor arguments, ``forward`` / ``get_estimate_bits`` signatures and the
CompressAI parameter and buffer names (``_matrix{0-4}``, ``_bias{0-4}``, ``_factor{0-3}``, ``quantiles``, ``target``
[3], ``_offset`` / ``_quantized_cdf`` / ``_cdf_length``, ``likelihood_lower_bound.bound``, ``lower_bound_scale.bound``,
``scale_table``, ``scale_bound``; as published for CompressAI 1.1-1.2, restated from memory) so that checkpoints load, and evaluates the likelihoods with the library's fused CUDA kernels
(``fvc_eb_forward`` / ``fvc_gaussian_forward``) and the hyper-prior convolutions with the tcgen05 engine.

In scope (eval forward only): ``RecProbModel.forward`` without the recurrent prior
(entropy_models.py:55-68, ``RPM_flag=False``) and with externally supplied RPM outputs,
``MeanScaleHyperPriors.forward`` (202-219), both ``get_estimate_bits`` (74-78, 228-235).
Out of scope: ``compress`` / ``decompress*`` range coding (torchac), ``update()`` CDF tables,
the ``RPM`` / ``ConvLSTM`` recurrent prior networks, training-mode noise.

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
        self.scale_table = torch.as_tensor(scale_table, dtype=torch.float32)
        return True

    def forward(self, x, scales, means=None, training=None):
        if training if training is not None else self.training:
            raise NotImplementedError("training-mode (additive noise) likelihoods are outside the hot path")
        xh, lik, _ = ops.gaussian_forward(x, scales, means)
        return xh, lik


def _estimate_bits_clamped(likelihoods):
    """reference entropy_models.py:74-78."""
    return torch.sum(torch.clamp(-1.0 * torch.log(likelihoods + 1e-5) / math.log(2.0), 0, 50))


class RecProbModel(nn.Module):
    """reference entropy_models.py:26-148 (forward / get_estimate_bits only).

    ``RPM_flag=False``: factorized ``entropy_bottleneck``.  ``RPM_flag=True``: the recurrent prior network
    (RPM + ConvLSTM) is out of scope; pass a module as ``rpm`` (called as ``rpm(prior_latent, rpm_hidden)`` ->
    ``sigma, mu, rpm_hidden``) to use the conditional-Gaussian branch (entropy_models.py:58-63).
    """

    def __init__(self, channels, rpm=None):
        super().__init__()
        self.channels = int(channels)
        self.entropy_bottleneck = EntropyBottleneck(channels)
        self.gaussian_conditional = GaussianConditional(None)
        self.sigma = self.mu = self.prior_latent = None
        self.RPM = rpm
        self.RPM_flag = False

    def set_RPM(self, RPM_flag):
        self.RPM_flag = RPM_flag

    def forward(self, x, rpm_hidden, training=None, prior_latent=None):
        if self.RPM_flag:
            assert prior_latent is not None, 'prior latent is none!'
            if self.RPM is None:
                raise NotImplementedError("the RPM/ConvLSTM prior network is outside the hot path; pass rpm=...")
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

    def get_estimate_bits(self, likelihoods):
        """reference entropy_models.py:228-235 (plain log2 sum per batch element)."""
        (x_likelihood, z_likelihood) = likelihoods
        log2 = torch.log(torch.FloatTensor([2])).squeeze(0).to(x_likelihood.device)
        bs = x_likelihood.size(0)
        x_est = torch.sum(torch.log(x_likelihood.view(bs, -1)), dim=-1) / (-log2)
        z_est = torch.sum(torch.log(z_likelihood.view(bs, -1)), dim=-1) / (-log2)
        return x_est + z_est
