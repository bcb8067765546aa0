"""Deterministic random-init weights and synthetic GOPs for the DVC P-frame path.

There is no network in the build/bench environment, so neither the trained
``DVC/snapshot/*.model`` checkpoints nor the SpyNet ``.npy`` weights the
reference loads (reference DVC/subnet/endecoder.py:122-139) are available on
the GPU box.  ``init_state_dict`` reproduces the reference constructors'
initialisers (same distributions, gains and constants; cited per block) under a
seed, with the reference's exact ``state_dict`` key/shape layout (SURVEY.md
section 8b), so the tensors load into the reference ``VideoCompressor`` and into
ours alike.  SpyNet gets a scaled default-conv init that yields sub-pixel to
few-pixel flows on the synthetic frames.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import torch

OUT_CHANNEL_N = 64    # reference DVC/subnet/basics.py:23
OUT_CHANNEL_M = 96    # reference DVC/subnet/basics.py:24
OUT_CHANNEL_MV = 128  # reference DVC/subnet/basics.py:27


def _xavier_normal(g, shape, gain, fan_in, fan_out):
    std = gain * math.sqrt(2.0 / float(fan_in + fan_out))
    return torch.randn(shape, generator=g, dtype=torch.float32) * std


def _xavier_uniform(g, shape, fan_in, fan_out):
    a = math.sqrt(3.0) * math.sqrt(2.0 / float(fan_in + fan_out))
    return (torch.rand(shape, generator=g, dtype=torch.float32) * 2.0 - 1.0) * a


def _conv(sd, g, name, cout, cin, k, gain, bias=0.01, transposed=False, uniform=False):
    # torch fan computation: dim1*rf is fan_in, dim0*rf is fan_out, for both layouts
    shape = (cin, cout, k, k) if transposed else (cout, cin, k, k)
    fan_in = shape[1] * k * k
    fan_out = shape[0] * k * k
    if uniform:
        w = _xavier_uniform(g, shape, fan_in, fan_out)
    else:
        w = _xavier_normal(g, shape, gain, fan_in, fan_out)
    sd[name + ".weight"] = w
    sd[name + ".bias"] = torch.full((cout,), float(bias), dtype=torch.float32)


def _gdn(sd, name, ch):
    # reference DVC/subnet/GDN.py:45-61
    ped = (2.0 ** -18) ** 2
    sd[name + ".beta"] = torch.sqrt(torch.ones(ch) + ped)
    sd[name + ".gamma"] = torch.sqrt(0.1 * torch.eye(ch) + ped)


def _bitest(sd, g, name, ch):
    # reference DVC/subnet/bitEstimator.py:10-18 : normal(0, 0.01), f4 has no `a`
    for i in (1, 2, 3, 4):
        for p in ("h", "b", "a"):
            if i == 4 and p == "a":
                continue
            sd[f"{name}.f{i}.{p}"] = torch.randn((1, ch, 1, 1), generator=g) * 0.01


def init_state_dict(seed: int = 0, spynet_levels: int = 4, spynet_gain: float = 2.2):
    """Random-init weights with the reference VideoCompressor state_dict layout."""
    g = torch.Generator().manual_seed(int(seed))
    sd = OrderedDict()
    # --- opticFlow: ME_Spynet / MEBasic (endecoder.py:142-169, 312-320) -----------
    chans = [(8, 32), (32, 64), (64, 32), (32, 16), (16, 2)]
    for lvl in range(spynet_levels):
        for i, (cin, cout) in enumerate(chans):
            bound = 1.0 / math.sqrt(cin * 49)
            w = (torch.rand((cout, cin, 7, 7), generator=g) * 2 - 1) * bound * spynet_gain
            b = (torch.rand((cout,), generator=g) * 2 - 1) * bound
            sd[f"opticFlow.moduleBasic.{lvl}.conv{i + 1}.weight"] = w
            sd[f"opticFlow.moduleBasic.{lvl}.conv{i + 1}.bias"] = b
    mv, N, M = OUT_CHANNEL_MV, OUT_CHANNEL_N, OUT_CHANNEL_M
    # --- mvEncoder (analysis_mv.py:14-44) -----------------------------------------
    _conv(sd, g, "mvEncoder.conv1", mv, 2, 3, math.sqrt(2 * (2 + mv) / 4))
    for i in range(2, 9):
        _conv(sd, g, f"mvEncoder.conv{i}", mv, mv, 3, math.sqrt(2))
    # --- mvDecoder (synthesis_mv.py:15-45) ----------------------------------------
    for i in (1, 3, 5, 7):
        _conv(sd, g, f"mvDecoder.deconv{i}", mv, mv, 3, math.sqrt(2), transposed=True)
    for i in (2, 4, 6):
        _conv(sd, g, f"mvDecoder.deconv{i}", mv, mv, 3, math.sqrt(2))
    _conv(sd, g, "mvDecoder.deconv8", 2, mv, 3, math.sqrt(2 * (mv + 2) / (mv + mv)))
    # --- warpnet (endecoder.py:262-283), xavier_uniform, bias 0 ---------------------
    _conv(sd, g, "warpnet.feature_ext", 64, 6, 3, 1.0, bias=0.0, uniform=True)
    for i in range(6):
        _conv(sd, g, f"warpnet.conv{i}.conv1", 64, 64, 3, 1.0, bias=0.0, uniform=True)
        _conv(sd, g, f"warpnet.conv{i}.conv2", 64, 64, 3, 1.0, bias=0.0, uniform=True)
    _conv(sd, g, "warpnet.conv6", 3, 64, 3, 1.0, bias=0.0, uniform=True)
    # --- resEncoder (analysis.py:16-30) ---------------------------------------------
    _conv(sd, g, "resEncoder.conv1", N, 3, 5, math.sqrt(2 * (3 + N) / 6))
    _gdn(sd, "resEncoder.gdn1", N)
    _conv(sd, g, "resEncoder.conv2", N, N, 5, math.sqrt(2))
    _gdn(sd, "resEncoder.gdn2", N)
    _conv(sd, g, "resEncoder.conv3", N, N, 5, math.sqrt(2))
    _gdn(sd, "resEncoder.gdn3", N)
    _conv(sd, g, "resEncoder.conv4", M, N, 5, math.sqrt(2 * (M + N) / (N + N)))
    # --- resDecoder (synthesis.py:14-28) ----------------------------------------------
    _conv(sd, g, "resDecoder.deconv1", N, M, 5, math.sqrt(2 * (N + M) / (M + M)), transposed=True)
    _gdn(sd, "resDecoder.igdn1", N)
    _conv(sd, g, "resDecoder.deconv2", N, N, 5, math.sqrt(2), transposed=True)
    _gdn(sd, "resDecoder.igdn2", N)
    _conv(sd, g, "resDecoder.deconv3", N, N, 5, math.sqrt(2), transposed=True)
    _gdn(sd, "resDecoder.igdn3", N)
    _conv(sd, g, "resDecoder.deconv4", 3, N, 5, math.sqrt(2 * (N + 3) / (N + N)), transposed=True)
    # --- respriorEncoder (analysis_prior.py:16-26) ---------------------------------
    _conv(sd, g, "respriorEncoder.conv1", N, M, 3, math.sqrt(2 * (M + N) / (M + M)))
    _conv(sd, g, "respriorEncoder.conv2", N, N, 5, math.sqrt(2))
    _conv(sd, g, "respriorEncoder.conv3", N, N, 5, math.sqrt(2))
    # --- respriorDecoder (synthesis_prior.py:17-27) -----------------------------------
    _conv(sd, g, "respriorDecoder.deconv1", N, N, 5, math.sqrt(2), transposed=True)
    _conv(sd, g, "respriorDecoder.deconv2", N, N, 5, math.sqrt(2), transposed=True)
    _conv(sd, g, "respriorDecoder.deconv3", M, N, 3, math.sqrt(2 * (N + M) / (N + N)), transposed=True)
    # --- bit estimators (net.py:48-49) ------------------------------------------------
    _bitest(sd, g, "bitEstimator_z", N)
    _bitest(sd, g, "bitEstimator_mv", mv)
    return sd


def synthetic_gop(height: int, width: int, gop: int = 10, gop_id: int = 0, batch: int = 1):
    """Smooth content with a known 2-3 px/frame translation.

    Returns fp32 ``[gop, batch*3 -> (batch,3)]`` i.e. a tensor ``[gop, batch, 3, H, W]`` in [0,1].
    """
    g = torch.Generator().manual_seed(1234 + int(gop_id))
    base = torch.rand((batch, 3, max(height // 8, 2), max(width // 8, 2)), generator=g)
    f0 = torch.nn.functional.interpolate(base, (height, width), mode="bicubic", align_corners=False).clamp(0, 1)
    frames = []
    for t in range(gop):
        f = torch.roll(f0, shifts=(2 * t, 3 * t), dims=(2, 3))
        f = f + 0.01 * torch.randn(f.shape, generator=g)
        frames.append(f.clamp(0, 1))
    return torch.stack(frames, 0).contiguous()


def init_rpm_state_dict(channels: int = 128, seed: int = 0):
    """Seeded weights with the state_dict layout of the reference's RPM prior network (entropy_models.py:328-378):
    ``conv1..conv7`` C->C, ``conv8`` C->2C, ``lstm.conv`` 2C->4C, all 3x3; PyTorch's default conv init range."""
    g = torch.Generator().manual_seed(int(seed))
    sd = OrderedDict()
    C = int(channels)
    shapes = [("conv%d" % i, C, C) for i in range(1, 8)] + [("conv8", 2 * C, C), ("lstm.conv", 4 * C, 2 * C)]
    for name, cout, cin in shapes:
        bound = 1.0 / math.sqrt(cin * 9)
        sd[name + ".weight"] = (torch.rand((cout, cin, 3, 3), generator=g) * 2 - 1) * bound * 1.7
        sd[name + ".bias"] = (torch.rand((cout,), generator=g) * 2 - 1) * bound
    return sd
