"""CPU fp32 restatement of the reference DVC P-frame forward.  TEST INFRASTRUCTURE ONLY.

Every function cites the reference lines it follows (paths relative to the
reference checkout).  The non-convolution ops (pooling, the two bilinear
up-samplers, the grid_sample-based backward warp, GDN, the factorized and
Laplace bit estimators, the losses) are restated as explicit index arithmetic,
not by calling the torch op the reference calls, so that the CUDA kernels are
checked against formulas and the formulas are pinned against the reference by
``tests/test_oracle_golden.py``.  Convolutions use ``F.conv2d`` /
``F.conv_transpose2d`` on CPU fp32 (the definition of the op).

Parity: PINNED against outputs of the unmodified reference (see
``oracle/gen_golden.py`` and ``tests/golden/``).
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------------------
# memory-bound geometry ops
# --------------------------------------------------------------------------------------
def avg_pool2(x):
    """F.avg_pool2d(k=2, s=2) — DVC/subnet/endecoder.py:344-346, 273/275 (AvgPool2d)."""
    a = x[..., 0::2, 0::2]
    b = x[..., 0::2, 1::2]
    c = x[..., 1::2, 0::2]
    d = x[..., 1::2, 1::2]
    return (((a + b) + c) + d) / 4.0


def _lin_idx(n_out, n_in, align_corners, dtype):
    o = torch.arange(n_out, dtype=dtype)
    if align_corners:
        src = o * (float(n_in - 1) / float(n_out - 1)) if n_out > 1 else o * 0
    else:
        src = (o + 0.5) * (float(n_in) / float(n_out)) - 0.5
        src = src.clamp(min=0)
    i0 = src.floor().long().clamp(max=n_in - 1)
    i1 = (i0 + 1).clamp(max=n_in - 1)
    w1 = src - i0.to(dtype)
    return i0, i1, w1


def upsample2x_bilinear(x, align_corners=False):
    """F.interpolate(size=2x, mode='bilinear') — endecoder.py:173-179 (align_corners False,
    SpyNet flow) and 180-184 (align_corners True, Warp_net skip connections)."""
    B, C, H, W = x.shape
    y0, y1, wy = _lin_idx(2 * H, H, align_corners, x.dtype)
    x0, x1, wx = _lin_idx(2 * W, W, align_corners, x.dtype)
    wy = wy.view(1, 1, -1, 1)
    wx = wx.view(1, 1, 1, -1)
    r0 = x[:, :, y0, :]
    r1 = x[:, :, y1, :]
    top = r0[:, :, :, x0] * (1 - wx) + r0[:, :, :, x1] * wx
    bot = r1[:, :, :, x0] * (1 - wx) + r1[:, :, :, x1] * wx
    return top * (1 - wy) + bot * wy


def flow_warp(img, flow):
    """torch_warp / flow_warp — endecoder.py:52-67, 116-119.

    grid = linspace(-1,1,W)[j] + flow_x/((W-1)/2) fed to F.grid_sample(bilinear, border) with the
    default align_corners=False: source x = ((g+1)*W-1)/2, clamped to [0,W-1]; 4 taps, taps whose
    index falls outside the image contribute 0 (they carry weight 0 after the clamp).
    """
    B, C, H, W = img.shape
    dt = img.dtype
    gx = torch.linspace(-1.0, 1.0, W, dtype=dt).view(1, 1, W) + flow[:, 0] / ((W - 1.0) / 2.0)
    gy = torch.linspace(-1.0, 1.0, H, dtype=dt).view(1, H, 1) + flow[:, 1] / ((H - 1.0) / 2.0)
    sx = ((gx + 1.0) * W - 1.0) / 2.0
    sy = ((gy + 1.0) * H - 1.0) / 2.0
    sx = sx.clamp(0.0, W - 1.0)
    sy = sy.clamp(0.0, H - 1.0)
    x0 = sx.floor()
    y0 = sy.floor()
    tx = sx - x0
    ty = sy - y0
    x0 = x0.long()
    y0 = y0.long()
    x1 = x0 + 1
    y1 = y0 + 1
    out = torch.zeros_like(img)
    flat = img.reshape(B, C, H * W)
    for (yy, xx, ww) in ((y0, x0, (1 - tx) * (1 - ty)), (y0, x1, tx * (1 - ty)),
                         (y1, x0, (1 - tx) * ty), (y1, x1, tx * ty)):
        valid = ((xx >= 0) & (xx < W) & (yy >= 0) & (yy < H)).to(dt)
        idx = (yy.clamp(0, H - 1) * W + xx.clamp(0, W - 1)).view(B, 1, H * W).expand(B, C, H * W)
        out = out + (flat.gather(2, idx).view(B, C, H, W) * (ww * valid).unsqueeze(1))
    return out


# --------------------------------------------------------------------------------------
# sub-networks
# --------------------------------------------------------------------------------------
def _conv(sd, name, x, stride=1):
    w = sd[name + ".weight"]
    return F.conv2d(x, w, sd[name + ".bias"], stride=stride, padding=w.shape[-1] // 2)


def _deconv(sd, name, x, stride=2):
    w = sd[name + ".weight"]
    k = w.shape[-1]
    return F.conv_transpose2d(x, w, sd[name + ".bias"], stride=stride, padding=k // 2,
                              output_padding=stride - 1)


def me_basic(sd, prefix, x):
    """MEBasic.forward — endecoder.py:162-169."""
    for i in (1, 2, 3, 4):
        x = F.relu(_conv(sd, f"{prefix}.conv{i}", x))
    return _conv(sd, f"{prefix}.conv5", x)


def me_spynet(sd, im1, im2, levels=4):
    """ME_Spynet.forward — endecoder.py:337-356 (im1 = current frame, im2 = reference)."""
    im1l, im2l = [im1], [im2]
    for _ in range(levels - 1):
        im1l.append(avg_pool2(im1l[-1]))
        im2l.append(avg_pool2(im2l[-1]))
    B = im1.shape[0]
    hc, wc = im2l[-1].shape[2] // 2, im2l[-1].shape[3] // 2
    flow = torch.zeros((B, 2, hc, wc), dtype=im1.dtype)
    for lvl in range(levels):
        up = upsample2x_bilinear(flow, align_corners=False) * 2.0
        a = im1l[levels - 1 - lvl]
        b = flow_warp(im2l[levels - 1 - lvl], up)
        flow = up + me_basic(sd, f"opticFlow.moduleBasic.{lvl}", torch.cat([a, b, up], 1))
    return flow


def analysis_mv(sd, x):
    """Analysis_mv_net.forward — analysis_mv.py:58-66."""
    strides = (2, 1, 2, 1, 2, 1, 2, 1)
    for i, s in enumerate(strides, start=1):
        x = _conv(sd, f"mvEncoder.conv{i}", x, stride=s)
        if i < 8:
            x = F.leaky_relu(x, 0.1)
    return x


def synthesis_mv(sd, x):
    """Synthesis_mv_net.forward — synthesis_mv.py:71-79."""
    for i in range(1, 9):
        if i % 2 == 1:
            x = _deconv(sd, f"mvDecoder.deconv{i}", x)
        else:
            x = _conv(sd, f"mvDecoder.deconv{i}", x)
        if i < 8:
            x = F.leaky_relu(x, 0.1)
    return x


def res_block(sd, prefix, x):
    """ResBlock.forward — endecoder.py:251-260 (pre-activation)."""
    y = _conv(sd, prefix + ".conv1", F.relu(x))
    y = _conv(sd, prefix + ".conv2", F.relu(y))
    return x + y


def warp_net(sd, x):
    """Warp_net.forward — endecoder.py:282-296."""
    f = F.relu(_conv(sd, "warpnet.feature_ext", x))
    c0 = res_block(sd, "warpnet.conv0", f)
    c1 = res_block(sd, "warpnet.conv1", avg_pool2(c0))
    c2 = res_block(sd, "warpnet.conv2", avg_pool2(c1))
    c3 = res_block(sd, "warpnet.conv3", c2)
    c3u = c1 + upsample2x_bilinear(c3, align_corners=True)
    c4 = res_block(sd, "warpnet.conv4", c3u)
    c4u = c0 + upsample2x_bilinear(c4, align_corners=True)
    c5 = res_block(sd, "warpnet.conv5", c4u)
    return _conv(sd, "warpnet.conv6", c5)


def gdn(sd, name, x, inverse=False):
    """GDN.forward — GDN.py:63-93 (build 45-61)."""
    ped = (2.0 ** -18) ** 2
    beta_bound = (1e-6 + ped) ** 0.5
    gamma_bound = 2.0 ** -18
    beta = torch.clamp(sd[name + ".beta"], min=beta_bound) ** 2 - ped
    gamma = torch.clamp(sd[name + ".gamma"], min=gamma_bound) ** 2 - ped
    C = x.shape[1]
    norm = torch.sqrt(F.conv2d(x * x, gamma.view(C, C, 1, 1), beta))
    return x * norm if inverse else x / norm


def analysis(sd, x):
    """Analysis_net.forward — analysis.py:44-48."""
    x = gdn(sd, "resEncoder.gdn1", _conv(sd, "resEncoder.conv1", x, 2))
    x = gdn(sd, "resEncoder.gdn2", _conv(sd, "resEncoder.conv2", x, 2))
    x = gdn(sd, "resEncoder.gdn3", _conv(sd, "resEncoder.conv3", x, 2))
    return _conv(sd, "resEncoder.conv4", x, 2)


def synthesis(sd, x):
    """Synthesis_net.forward — synthesis.py:54-58."""
    x = gdn(sd, "resDecoder.igdn1", _deconv(sd, "resDecoder.deconv1", x), inverse=True)
    x = gdn(sd, "resDecoder.igdn2", _deconv(sd, "resDecoder.deconv2", x), inverse=True)
    x = gdn(sd, "resDecoder.igdn3", _deconv(sd, "resDecoder.deconv3", x), inverse=True)
    return _deconv(sd, "resDecoder.deconv4", x)


def analysis_prior(sd, x):
    """Analysis_prior_net.forward — analysis_prior.py:40-56."""
    x = torch.abs(x)
    x = F.relu(_conv(sd, "respriorEncoder.conv1", x, 1))
    x = F.relu(_conv(sd, "respriorEncoder.conv2", x, 2))
    return _conv(sd, "respriorEncoder.conv3", x, 2)


def synthesis_prior(sd, x):
    """Synthesis_prior_net.forward — synthesis_prior.py:42-58."""
    x = F.relu(_deconv(sd, "respriorDecoder.deconv1", x))
    x = F.relu(_deconv(sd, "respriorDecoder.deconv2", x))
    x = _deconv(sd, "respriorDecoder.deconv3", x, stride=1)
    return torch.exp(x)


# --------------------------------------------------------------------------------------
# entropy-model bit estimation
# --------------------------------------------------------------------------------------
def _softplus(x):
    """F.softplus(beta=1, threshold=20)."""
    return torch.where(x > 20.0, x, torch.log1p(torch.exp(x)))


def bit_estimator_cdf(sd, prefix, x):
    """BitEstimator.forward / Bitparm.forward — bitEstimator.py:20-42.  x is [B,C,H,W]."""
    for i in (1, 2, 3):
        h, b, a = (sd[f"{prefix}.f{i}.{p}"] for p in ("h", "b", "a"))
        x = x * _softplus(h) + b
        x = x + torch.tanh(x) * torch.tanh(a)
    h, b = sd[f"{prefix}.f4.h"], sd[f"{prefix}.f4.b"]
    return torch.sigmoid(x * _softplus(h) + b)


def clamp_log2_bits(prob):
    """sum(clamp(-log(p + 1e-5)/log 2, 0, 50)) — net.py:145, 170, 198; entropy_models.py:74-78."""
    return torch.sum(torch.clamp(-1.0 * torch.log(prob + 1e-5) / math.log(2.0), 0, 50))


def factorized_bits(sd, prefix, q):
    """iclr18_estrate_bits_z / _mv — net.py:153-178, 181-205 (calrealbits False)."""
    prob = bit_estimator_cdf(sd, prefix, q + 0.5) - bit_estimator_cdf(sd, prefix, q - 0.5)
    return clamp_log2_bits(prob), prob


def laplace_cdf(v, sigma):
    """torch.distributions.Laplace(0, sigma).cdf(v) = 0.5 - 0.5*sign(v)*expm1(-|v|/sigma)."""
    return 0.5 - 0.5 * torch.sign(v) * torch.expm1(-torch.abs(v) / sigma)


def laplace_bits(q, sigma):
    """feature_probs_based_sigma — net.py:121-151 (calrealbits False)."""
    sigma = sigma.clamp(1e-5, 1e10)
    prob = laplace_cdf(q + 0.5, sigma) - laplace_cdf(q - 0.5, sigma)
    return clamp_log2_bits(prob), prob


# CompressAI-facing likelihoods used by entropy_models.py (RecProbModel 55-68,
# MeanScaleHyperPriors 202-219).  CompressAI is an un-vendored, un-pinned dependency
# (docker/Dockerfile:46, era 1.1-1.2); these follow its published algorithm:
# EntropyBottleneck._logits_cumulative/_likelihood, GaussianConditional._likelihood.
# PARITY UNPINNED at this boundary (no CompressAI available to check against).
def eb_logits_cumulative(matrices, biases, factors, x):
    """x: [C,1,N].  matrices[i]: [C,f_{i+1},f_i], biases[i]: [C,f_{i+1},1], factors[i]: [C,f_{i+1},1]."""
    logits = x
    for i in range(len(matrices)):
        logits = torch.matmul(F.softplus(matrices[i]), logits) + biases[i]
        if i < len(factors):
            logits = logits + torch.tanh(factors[i]) * torch.tanh(logits)
    return logits


def eb_forward(matrices, biases, factors, medians, x, likelihood_bound=1e-9):
    """EntropyBottleneck.forward in eval mode: x_hat = round(x - median) + median; likelihood."""
    B, C, H, W = x.shape
    med = medians.view(1, C, 1, 1)
    xh = torch.round(x - med) + med
    v = xh.permute(1, 0, 2, 3).reshape(C, 1, -1)
    lower = eb_logits_cumulative(matrices, biases, factors, v - 0.5)
    upper = eb_logits_cumulative(matrices, biases, factors, v + 0.5)
    sign = -torch.sign(lower + upper)
    lik = torch.abs(torch.sigmoid(sign * upper) - torch.sigmoid(sign * lower))
    lik = lik.clamp(min=likelihood_bound)
    return xh, lik.reshape(C, B, H, W).permute(1, 0, 2, 3)


def gaussian_forward(x, scales, means, scale_bound=0.11, likelihood_bound=1e-9):
    """GaussianConditional.forward in eval mode: quantize round(x-mu)+mu; Phi-difference likelihood."""
    xh = torch.round(x - means) + means
    s = scales.clamp(min=scale_bound)
    v = torch.abs(xh - means)
    const = -(2 ** -0.5)
    upper = 0.5 * torch.erfc(const * ((0.5 - v) / s))
    lower = 0.5 * torch.erfc(const * ((-0.5 - v) / s))
    return xh, (upper - lower).clamp(min=likelihood_bound)


def estimate_bits_clamped(likelihoods):
    """RecProbModel.get_estimate_bits — entropy_models.py:74-78."""
    return torch.sum(torch.clamp(-1.0 * torch.log(likelihoods + 1e-5) / math.log(2.0), 0, 50))


def _eb_params(sd, prefix):
    m = [sd[prefix + "_matrix%d" % i] for i in range(5)]
    b = [sd[prefix + "_bias%d" % i] for i in range(5)]
    f = [sd[prefix + "_factor%d" % i] for i in range(4)]
    return m, b, f, sd[prefix + "quantiles"][:, 0, 1]


def recprob_forward(sd, x):
    """RecProbModel.forward with RPM_flag False — entropy_models.py:55-68: (x_hat, likelihood, prior_latent)."""
    m, b, f, med = _eb_params(sd, "entropy_bottleneck.")
    xh, lik = eb_forward(m, b, f, med, x)
    return xh, lik, torch.round(x)


def meanscale_forward(sd, x, channels):
    """MeanScaleHyperPriors.forward — entropy_models.py:202-219 (nn.LeakyReLU() default slope 0.01)."""
    def hconv(prefix, t, last_act):
        t = F.leaky_relu(F.conv2d(t, sd[prefix + ".0.weight"], sd[prefix + ".0.bias"], padding=1), 0.01)
        t = F.conv2d(t, sd[prefix + ".2.weight"], sd[prefix + ".2.bias"], padding=1)
        return F.leaky_relu(t, 0.01) if last_act else t
    z = hconv("h_a2", hconv("h_a1", x, True), False)
    m, b, f, med = _eb_params(sd, "entropy_bottleneck.")
    z_hat, z_lik = eb_forward(m, b, f, med, z)
    gp = hconv("h_s2", hconv("h_s1", z_hat, True), False)
    sigma, mu = torch.split(gp, channels, dim=1)
    sigma = torch.exp(torch.maximum(sigma, torch.tensor(-7.0)))
    x_hat, x_lik = gaussian_forward(x, sigma, mu)
    return x_hat, x_lik, z_lik, z, sigma, mu


def meanscale_bits(x_lik, z_lik):
    """MeanScaleHyperPriors.get_estimate_bits — entropy_models.py:228-235."""
    bs = x_lik.size(0)
    return (torch.sum(torch.log(x_lik.view(bs, -1)), -1) + torch.sum(torch.log(z_lik.view(bs, -1)), -1)) / (-math.log(2.0))


# --------------------------------------------------------------------------------------
# whole P-frame
# --------------------------------------------------------------------------------------
def pframe_forward(sd, cur, ref, levels=4, capture=False):
    """VideoCompressor.forward in eval mode — net.py:70-220.

    Returns the reference 8-tuple; with ``capture`` also a dict of intermediates.
    """
    estmv = me_spynet(sd, cur, ref, levels)
    mvfeature = analysis_mv(sd, estmv)
    quant_mv = torch.round(mvfeature)
    mv_hat = synthesis_mv(sd, quant_mv)
    warpframe = flow_warp(ref, mv_hat)
    prediction = warp_net(sd, torch.cat((warpframe, ref), 1)) + warpframe
    residual = cur - prediction
    feature = analysis(sd, residual)
    z = analysis_prior(sd, feature)
    z_hat = torch.round(z)
    sigma = synthesis_prior(sd, z_hat)
    feat_hat = torch.round(feature)
    recon_res = synthesis(sd, feat_hat)
    recon = prediction + recon_res
    clipped = recon.clamp(0.0, 1.0)
    mse = torch.mean((recon - cur).pow(2))
    warploss = torch.mean((warpframe - cur).pow(2))
    interloss = torch.mean((prediction - cur).pow(2))
    bits_feature, _ = laplace_bits(feat_hat, sigma)
    bits_z, _ = factorized_bits(sd, "bitEstimator_z", z_hat)
    bits_mv, _ = factorized_bits(sd, "bitEstimator_mv", quant_mv)
    B, _, H, W = cur.shape
    den = B * H * W
    bpp_feature, bpp_z, bpp_mv = bits_feature / den, bits_z / den, bits_mv / den
    bpp = bpp_feature + bpp_z + bpp_mv
    out = (clipped, mse, warploss, interloss, bpp_feature, bpp_z, bpp_mv, bpp)
    if capture:
        inter = dict(estmv=estmv, mvfeature=mvfeature, quant_mv=quant_mv, mv_hat=mv_hat,
                     warpframe=warpframe, prediction=prediction, feature=feature, z=z, z_hat=z_hat,
                     sigma=sigma, feat_hat=feat_hat, recon_res=recon_res, recon=recon)
        return out, inter
    return out


def iframe_forward(sd, x, capture=False):
    """Intra frame through the residual branch of VideoCompressor.forward (net.py:86-116) with a zero prediction: the
    composition behind fvc_iframe_forward (the reference itself codes I-frames with bpgenc / bpgdec,
    models.py:412-429, so this pins the library's composition of the reference's own modules, not a reference path).
    Returns (clipped, mse, bpp_feature, bpp_z, bpp)."""
    feature = analysis(sd, x)
    z = analysis_prior(sd, feature)
    z_hat = torch.round(z)
    sigma = synthesis_prior(sd, z_hat)
    feat_hat = torch.round(feature)
    recon = synthesis(sd, feat_hat)
    clipped = recon.clamp(0.0, 1.0)
    mse = torch.mean((recon - x).pow(2))
    bits_feature, _ = laplace_bits(feat_hat, sigma)
    bits_z, _ = factorized_bits(sd, "bitEstimator_z", z_hat)
    B, _, H, W = x.shape
    den = B * H * W
    out = (clipped, mse, bits_feature / den, bits_z / den, (bits_feature + bits_z) / den)
    if capture:
        return out, dict(feature=feature, z=z, z_hat=z_hat, sigma=sigma, feat_hat=feat_hat, recon_res=recon)
    return out


def lsvc_forward(sd, x, layers, parents, ref_index, levels=4):
    """LSVC.forward in eval mode (128-channel, non-attention variants) — models.py:1344-1411.  ``layers``,
    ``parents``, ``ref_index`` are the reference's GOP graph (models.py:683-728, 923-949)."""
    inp = x[1:]
    bs, c, h, w = inp.shape
    estmv = me_spynet(sd, inp, x[torch.as_tensor(ref_index)], levels)      # flow against ORIGINAL references
    mvfeature = analysis_mv(sd, estmv)
    quant_mv = torch.round(mvfeature)
    mv_up = synthesis_mv(sd, quant_mv)
    bits_mv, _ = factorized_bits(sd, "bitEstimator_mv", quant_mv)
    com = [None] * bs
    mc = [None] * bs
    wp = [None] * bs
    bits_res = 0.0
    for layer in layers:
        tars = [t for t in layer if t <= bs]
        if not tars:
            continue
        ref = torch.cat([x[:1] if parents[t] == 0 else com[parents[t] - 1] for t in tars], 0)
        diff = torch.cat([mv_up[t - 1:t] for t in tars], 0)
        target = torch.cat([inp[t - 1:t] for t in tars], 0)
        warpframe = flow_warp(ref, diff)
        pred = warp_net(sd, torch.cat((warpframe, ref), 1)) + warpframe
        feature = analysis(sd, target - pred)
        z_hat = torch.round(analysis_prior(sd, feature))
        sigma = synthesis_prior(sd, z_hat)
        feat_hat = torch.round(feature)
        res_hat = synthesis(sd, feat_hat)
        bf, _ = laplace_bits(feat_hat, sigma)
        bz, _ = factorized_bits(sd, "bitEstimator_z", z_hat)
        bits_res = bits_res + bf + bz
        cf = torch.clip(res_hat + pred, min=0, max=1)
        for i, t in enumerate(tars):
            com[t - 1], mc[t - 1], wp[t - 1] = cf[i:i + 1], pred[i:i + 1], warpframe[i:i + 1]
    com, mc, wp = torch.cat(com, 0), torch.cat(mc, 0), torch.cat(wp, 0)
    rec_loss = torch.mean((com - inp).pow(2))
    warp_loss = torch.mean((wp - inp).pow(2))
    mc_loss = torch.mean((mc - inp).pow(2))
    bpp_res = bits_res / (bs * h * w)
    bpp = bpp_res + bits_mv / (bs * h * w)
    return com, mc, wp, rec_loss, warp_loss, mc_loss, bpp_res, bpp


def gop_forward(sd, frames, levels=4):
    """parallel_compression, 'DVC-pretrained' branch — models.py:368-383, 400-410.

    ``frames`` is [G,3,H,W] with frame 0 the (already decoded) I-frame.  Returns the per-frame
    list of (bpp, psnr, mse) and the reconstructed frames [G-1,3,H,W].
    """
    x_prev = frames[0:1]
    rows, recs = [], []
    for i in range(1, frames.shape[0]):
        out = pframe_forward(sd, frames[i:i + 1], x_prev, levels)
        x_prev = out[0].detach()
        mse = out[1]
        psnr = 10.0 * torch.log(1 / mse) / math.log(10.0)
        rows.append((float(out[7]), float(psnr), float(mse)))
        recs.append(x_prev)
    return rows, torch.cat(recs, 0)
