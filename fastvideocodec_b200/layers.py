"""CompressAI-compatible layer mirrors on the CUDA kernels.

The reference builds its non-DVC codecs on ``compressai.layers.GDN`` (models.py:23, 529-538).  CompressAI is an
un-vendored dependency (docker/Dockerfile:46), so this module restates its published GDN with the library's
parameter / buffer names - checkpoints of the reference's models load - and evaluates it with ``fvc_gdn``
(the same kernel that serves DVC/subnet/GDN.py:63-93: the two GDN implementations share the
non-negative re-parametrisation ``max(p, bound)^2 - pedestal`` with pedestal 2^-36, beta_min 1e-6).

PARITY UNPINNED against CompressAI itself (not installed here); pinned indirectly through the DVC GDN, whose
outputs are golden vectors of the unmodified reference (tests/golden/ops.npz).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops


class _LowerBound(nn.Module):
    def __init__(self, bound):
        super().__init__()
        self.register_buffer("bound", torch.tensor([float(bound)]))


class _NonNegativeParametrizer(nn.Module):
    """compressai.ops.parametrizers.NonNegativeParametrizer (buffers only; the math runs in the kernel)."""

    def __init__(self, minimum=0.0, reparam_offset=2 ** -18):
        super().__init__()
        pedestal = float(reparam_offset) ** 2
        self.register_buffer("pedestal", torch.tensor([pedestal]))
        self.lower_bound = _LowerBound((float(minimum) + pedestal) ** 0.5)

    def init(self, x):
        return torch.sqrt(torch.max(x + self.pedestal, self.pedestal))


class GDN(nn.Module):
    """compressai.layers.GDN(in_channels, inverse=False, beta_min=1e-6, gamma_init=0.1): eval forward.

    state_dict keys: ``beta``, ``gamma``, ``beta_reparam.pedestal``, ``beta_reparam.lower_bound.bound``,
    ``gamma_reparam.pedestal``, ``gamma_reparam.lower_bound.bound`` (as in CompressAI).
    """

    def __init__(self, in_channels, inverse=False, beta_min=1e-6, gamma_init=0.1):
        super().__init__()
        if abs(float(beta_min) - 1e-6) > 1e-12:
            raise ValueError("the CUDA kernel is specialised for beta_min = 1e-6")
        if in_channels % 8 or not (8 <= in_channels <= 64):
            raise ValueError("fvc_gdn supports 8..64 channels in multiples of 8")
        self.inverse = bool(inverse)
        self.beta_reparam = _NonNegativeParametrizer(minimum=beta_min)
        self.beta = nn.Parameter(self.beta_reparam.init(torch.ones(in_channels)))
        self.gamma_reparam = _NonNegativeParametrizer()
        self.gamma = nn.Parameter(self.gamma_reparam.init(float(gamma_init) * torch.eye(in_channels)))

    def forward(self, x):
        if torch.is_grad_enabled() and (x.requires_grad or self.beta.requires_grad and self.training):
            raise NotImplementedError("training / autograd is outside the B200 inference hot path")
        return ops.gdn(x, self.beta.detach(), self.gamma.detach(), inverse=self.inverse)
