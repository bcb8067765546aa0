"""Parameter-holding mirrors of the reference DVC sub-networks (reference DVC/subnet/*.py).

Each class keeps the reference's attribute names, so ``state_dict()`` keys and shapes are identical
to the reference and reference checkpoints load unchanged.  ``forward`` of every
module runs through libfvc_b200 (ops.py); nothing here computes with PyTorch kernels.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from . import ops
from .synthetic import OUT_CHANNEL_M, OUT_CHANNEL_MV, OUT_CHANNEL_N

out_channel_N = OUT_CHANNEL_N
out_channel_M = OUT_CHANNEL_M
out_channel_mv = OUT_CHANNEL_MV


class Conv(nn.Module):
    """nn.Conv2d(cin, cout, k, stride, padding=k//2) parameter holder."""

    def __init__(self, cin, cout, k, stride=1, act=ops.ACT_NONE):
        super().__init__()
        self.weight = nn.Parameter(torch.zeros(cout, cin, k, k))
        self.bias = nn.Parameter(torch.zeros(cout))
        self.stride, self.act = stride, act

    def forward(self, x, act=None):
        return ops.conv2d(x, self.weight.detach(), self.bias.detach(), self.stride,
                          self.act if act is None else act)


class Deconv(nn.Module):
    """nn.ConvTranspose2d(cin, cout, k, stride, padding=k//2, output_padding=stride-1) holder."""

    def __init__(self, cin, cout, k, stride=2, act=ops.ACT_NONE):
        super().__init__()
        self.weight = nn.Parameter(torch.zeros(cin, cout, k, k))
        self.bias = nn.Parameter(torch.zeros(cout))
        self.stride, self.act = stride, act

    def forward(self, x, act=None):
        return ops.conv_transpose2d(x, self.weight.detach(), self.bias.detach(), self.stride,
                                    self.act if act is None else act)


class GDN(nn.Module):
    """reference DVC/subnet/GDN.py:26-93 (forward only)."""

    def __init__(self, ch, inverse=False):
        super().__init__()
        self.inverse = inverse
        ped = (2.0 ** -18) ** 2
        self.beta = nn.Parameter(torch.sqrt(torch.ones(ch) + ped))
        self.gamma = nn.Parameter(torch.sqrt(0.1 * torch.eye(ch) + ped))

    def forward(self, x):
        return ops.gdn(x, self.beta.detach(), self.gamma.detach(), self.inverse)


class MEBasic(nn.Module):
    """reference endecoder.py:142-169."""

    def __init__(self):
        super().__init__()
        self.conv1 = Conv(8, 32, 7, 1, ops.ACT_RELU)
        self.conv2 = Conv(32, 64, 7, 1, ops.ACT_RELU)
        self.conv3 = Conv(64, 32, 7, 1, ops.ACT_RELU)
        self.conv4 = Conv(32, 16, 7, 1, ops.ACT_RELU)
        self.conv5 = Conv(16, 2, 7, 1)

    def forward(self, x):
        return self.conv5(self.conv4(self.conv3(self.conv2(self.conv1(x)))))


class ME_Spynet(nn.Module):
    """reference endecoder.py:312-356; ``L`` is a constructor argument here."""

    def __init__(self, L=4):
        super().__init__()
        self.L = L
        self.moduleBasic = nn.ModuleList([MEBasic() for _ in range(L)])

    def forward(self, im1, im2):
        im1l, im2l = [im1], [im2]
        for _ in range(self.L - 1):
            im1l.append(ops.avg_pool2(im1l[-1]))
            im2l.append(ops.avg_pool2(im2l[-1]))
        B, _, h, w = im2l[-1].shape
        flow = torch.zeros((B, 2, h // 2, w // 2), device=im1.device, dtype=torch.float32)
        for lvl in range(self.L):
            up = ops.upsample2x_bilinear(flow, False, 2.0)
            a, b = im1l[self.L - 1 - lvl], ops.flow_warp(im2l[self.L - 1 - lvl], up)
            flow = up + self.moduleBasic[lvl](torch.cat([a, b, up], 1))
        return flow


class Analysis_mv_net(nn.Module):
    """reference analysis_mv.py:8-66 (useAttn=False branch)."""

    def __init__(self):
        super().__init__()
        c = out_channel_mv
        for i in range(1, 9):
            setattr(self, f"conv{i}", Conv(2 if i == 1 else c, c, 3, 2 if i % 2 else 1,
                                           ops.ACT_LRELU01 if i < 8 else ops.ACT_NONE))

    def forward(self, x):
        for i in range(1, 9):
            x = getattr(self, f"conv{i}")(x)
        return x


class Synthesis_mv_net(nn.Module):
    """reference synthesis_mv.py:9-79 (useAttn=False branch)."""

    def __init__(self):
        super().__init__()
        c = out_channel_mv
        for i in range(1, 9):
            act = ops.ACT_LRELU01 if i < 8 else ops.ACT_NONE
            if i % 2:
                setattr(self, f"deconv{i}", Deconv(c, c, 3, 2, act))
            else:
                setattr(self, f"deconv{i}", Conv(c, 2 if i == 8 else c, 3, 1, act))

    def forward(self, x):
        for i in range(1, 9):
            x = getattr(self, f"deconv{i}")(x)
        return x


class ResBlock(nn.Module):
    """reference endecoder.py:228-260 (pre-activation; relu on tensors is a torch elementwise op
    only in this module-level convenience path — the fused path lives in fvc_pframe_forward)."""

    def __init__(self, ch=64, k=3):
        super().__init__()
        self.conv1 = Conv(ch, ch, k, 1, ops.ACT_RELU)
        self.conv2 = Conv(ch, ch, k, 1)

    def forward(self, x):
        return x + self.conv2(self.conv1(torch.relu(x)))


class Warp_net(nn.Module):
    """reference endecoder.py:262-296."""

    def __init__(self):
        super().__init__()
        ch = 64
        self.feature_ext = Conv(6, ch, 3, 1, ops.ACT_RELU)
        for i in range(6):
            setattr(self, f"conv{i}", ResBlock(ch, 3))
        self.conv6 = Conv(ch, 3, 3, 1)

    def forward(self, x):
        f = self.feature_ext(x)
        c0 = self.conv0(f)
        c1 = self.conv1(ops.avg_pool2(c0))
        c2 = self.conv2(ops.avg_pool2(c1))
        c3 = self.conv3(c2)
        c4 = self.conv4(c1 + ops.upsample2x_bilinear(c3, True))
        c5 = self.conv5(c0 + ops.upsample2x_bilinear(c4, True))
        return self.conv6(c5)


class Analysis_net(nn.Module):
    """reference analysis.py:10-60."""

    def __init__(self):
        super().__init__()
        N, M = out_channel_N, out_channel_M
        self.conv1, self.gdn1 = Conv(3, N, 5, 2), GDN(N)
        self.conv2, self.gdn2 = Conv(N, N, 5, 2), GDN(N)
        self.conv3, self.gdn3 = Conv(N, N, 5, 2), GDN(N)
        self.conv4 = Conv(N, M, 5, 2)

    def forward(self, x):
        x = self.gdn1(self.conv1(x))
        x = self.gdn2(self.conv2(x))
        x = self.gdn3(self.conv3(x))
        return self.conv4(x)


class Synthesis_net(nn.Module):
    """reference synthesis.py:8-58."""

    def __init__(self):
        super().__init__()
        N, M = out_channel_N, out_channel_M
        self.deconv1, self.igdn1 = Deconv(M, N, 5, 2), GDN(N, inverse=True)
        self.deconv2, self.igdn2 = Deconv(N, N, 5, 2), GDN(N, inverse=True)
        self.deconv3, self.igdn3 = Deconv(N, N, 5, 2), GDN(N, inverse=True)
        self.deconv4 = Deconv(N, 3, 5, 2)

    def forward(self, x):
        x = self.igdn1(self.deconv1(x))
        x = self.igdn2(self.deconv2(x))
        x = self.igdn3(self.deconv3(x))
        return self.deconv4(x)


class Analysis_prior_net(nn.Module):
    """reference analysis_prior.py:10-56."""

    def __init__(self):
        super().__init__()
        N, M = out_channel_N, out_channel_M
        self.conv1 = Conv(M, N, 3, 1, ops.ACT_RELU)
        self.conv2 = Conv(N, N, 5, 2, ops.ACT_RELU)
        self.conv3 = Conv(N, N, 5, 2)

    def forward(self, x):
        return self.conv3(self.conv2(self.conv1(torch.abs(x))))


class Synthesis_prior_net(nn.Module):
    """reference synthesis_prior.py:11-58."""

    def __init__(self):
        super().__init__()
        N, M = out_channel_N, out_channel_M
        self.deconv1 = Deconv(N, N, 5, 2, ops.ACT_RELU)
        self.deconv2 = Deconv(N, N, 5, 2, ops.ACT_RELU)
        self.deconv3 = Deconv(N, M, 3, 1, ops.ACT_EXP)

    def forward(self, x):
        return self.deconv3(self.deconv2(self.deconv1(x)))


class Bitparm(nn.Module):
    """reference bitEstimator.py:6-25 (parameters only; evaluation is fused in the CUDA kernel)."""

    def __init__(self, channel, final=False):
        super().__init__()
        self.final = final
        self.h = nn.Parameter(torch.zeros(1, channel, 1, 1))
        self.b = nn.Parameter(torch.zeros(1, channel, 1, 1))
        if not final:
            self.a = nn.Parameter(torch.zeros(1, channel, 1, 1))
        else:
            self.a = None


class BitEstimator(nn.Module):
    """reference bitEstimator.py:27-42."""

    def __init__(self, channel):
        super().__init__()
        self.f1, self.f2, self.f3 = Bitparm(channel), Bitparm(channel), Bitparm(channel)
        self.f4 = Bitparm(channel, True)

    def param_list(self):
        out = []
        for f in (self.f1, self.f2, self.f3):
            out += [f.h.detach(), f.b.detach(), f.a.detach()]
        out += [self.f4.h.detach(), self.f4.b.detach()]
        return out

    def quant_bits(self, x):
        """(round(x), total bits) — the closures iclr18_estrate_bits_z/mv of net.py:153-205."""
        return ops.quant_bits_factorized(x, self.param_list())


def flow_warp(im, flow):
    """reference endecoder.py:116-119."""
    return ops.flow_warp(im, flow)
