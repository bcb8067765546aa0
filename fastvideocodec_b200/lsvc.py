"""LSVC batched / tree GOP forward — mirrors reference models.py:1157-1411 (class LSVC) and the
GOP-graph helpers (models.py:683-728 ``generate_graph``, 923-949 ``graph_from_batch`` / ``refidx_from_graph``).

LSVC uses the same sub-networks as DVC; what differs is the schedule: optical flow and the MV autoencoder run on
ALL P-frames of the GOP at once against their ORIGINAL reference frames, then motion compensation and the
residual codec run layer by layer of a reference tree against the RECONSTRUCTED parents.  Here both phases are
single calls into the C ABI (``fvc_lsvc_mv_forward`` / ``fvc_lsvc_mc_res_forward``) on contexts of the matching
batch size.  Supported: the non-attention variants with 128 MV channels (``'LSVC-128'``, ``'LSVC-L-128'``,
``'LSVC-O-128'``, ``-D``), whose parameters are exactly the ``VideoCompressor`` ones.  The ``-A`` / ``-S``
attention variants and the 96-channel default are outside this build.
"""
from __future__ import annotations

import time

import torch

from ._lib import check, lib, ptr, stream_ptr
from .net import VideoCompressor


def generate_graph(graph_type='default'):
    """reference models.py:683-728: (children map, layers, parents) of the GOP reference tree."""
    if graph_type == 'default':
        g = {k: [k + 1] for k in range(30)}
        layers = [[i + 1] for i in range(30)]
        parents = {i + 1: i for i in range(30)}
    elif graph_type == 'onehop':
        g = {0: [i + 1 for i in range(14)]}
        layers = [[i + 1 for i in range(14)]]
        parents = {i + 1: 0 for i in range(14)}
    elif graph_type == '2layers':
        g = {0: [1, 2]}
        layers = [[1, 2]]
        parents = {1: 0, 2: 0}
    elif graph_type == '3layers':
        g = {0: [1, 4], 1: [2, 3], 4: [5, 6]}
        layers = [[1, 4], [2, 3, 5, 6]]
        parents = {1: 0, 4: 0, 2: 1, 3: 1, 5: 4, 6: 4}
    elif graph_type == '4layers':
        g = {0: [1, 8], 1: [2, 5], 8: [9, 12], 2: [3, 4], 5: [6, 7], 9: [10, 11], 12: [13, 14]}
        layers = [[1, 8], [2, 5, 9, 12], [3, 4, 6, 7, 10, 11, 13, 14]]
        parents = {1: 0, 8: 0, 2: 1, 5: 1, 9: 8, 12: 8, 3: 2, 4: 2, 6: 5, 7: 5, 10: 9, 11: 9, 13: 12, 14: 12}
    elif graph_type == '5layers':
        g = {0: [1, 16], 1: [2, 9], 16: [17, 24], 2: [3, 6], 9: [10, 13], 17: [18, 21], 24: [25, 28],
             3: [4, 5], 6: [7, 8], 10: [11, 12], 13: [14, 15], 18: [19, 20], 21: [22, 23], 25: [26, 27], 28: [29, 30]}
        layers = [[1, 16], [2, 9, 17, 24], [3, 6, 10, 13, 18, 21, 25, 28],
                  [4, 5, 7, 8, 11, 12, 14, 15, 19, 20, 22, 23, 26, 27, 29, 30]]
        parents = {1: 0, 16: 0, 2: 1, 9: 1, 17: 16, 24: 16, 3: 2, 6: 2, 10: 9, 13: 9, 18: 17, 21: 17, 25: 24, 28: 24,
                   4: 3, 5: 3, 7: 6, 8: 6, 11: 10, 12: 10, 14: 13, 15: 13, 19: 18, 20: 18, 22: 21, 23: 21, 26: 25,
                   27: 25, 29: 28, 30: 28}
    else:
        raise ValueError('Undefined graph type: %s' % graph_type)
    return g, layers, parents


def graph_from_batch(bs, isLinear=False, isOnehop=False):
    """reference models.py:923-941."""
    if isLinear:
        return generate_graph('default')
    if isOnehop:
        return generate_graph('onehop')
    if bs <= 2:
        return generate_graph('2layers')
    if bs <= 6:
        return generate_graph('3layers')
    if bs <= 14:
        return generate_graph('4layers')
    if bs <= 30:
        return generate_graph('5layers')
    raise ValueError('Batch size not supported yet: %d' % bs)


def refidx_from_graph(g, bs):
    """reference models.py:943-949: index (into x) of the reference frame of every P-frame."""
    ref_index = [-1 for _ in range(bs)]
    for start in g:
        if start > bs:
            continue
        for k in g[start]:
            if k > bs:
                continue
            ref_index[k - 1] = start
    return ref_index


class LSVC(VideoCompressor):
    """Drop-in for reference ``models.LSVC`` (eval forward): ``forward(x)`` with ``x = [I-frame, P_1..P_bs]``
    returns ``(com_frames, MC_frames, warped_frames, rec_loss, warp_loss, mc_loss, bpp_res, bpp)``."""

    def __init__(self, name, loss_type='P', compression_level=3, use_split=True):
        if '-A' in name or '-S' in name:
            raise NotImplementedError("attention variants of LSVC (-A / -S) are outside this build")
        if '-128' not in name:
            raise NotImplementedError("only the 128-channel MV variants ('-128') share the DVC parameter set")
        super().__init__()
        self.name = name
        self.useAttn = False
        self.loss_type = loss_type
        self.channels = 128
        self.compression_level = compression_level
        self.use_split = use_split          # the reference's 2-GPU model split: a no-op here (one GPU per process)
        # models.py:68-78 (init_training_params)
        self.r_img, self.r_bpp, self.r_aux = 1, 1, 1
        self.stage = 'REC'
        psnr_list = [256, 512, 1024, 2048, 4096, 8192, 16384, 16384 * 2, 16384 * 4]
        msssim_list = [8, 16, 32, 64]
        I_lvl_list = [37, 32, 27, 22, 17, 12, 7, 2, 1]
        self.r = psnr_list[compression_level] if loss_type == 'P' else msssim_list[compression_level]
        self.I_level = I_lvl_list[compression_level]
        self.encoding_time = self.decoding_time = 0.0

    def forward(self, x):
        if self.training:
            raise NotImplementedError("training-mode forward is outside the B200 inference hot path")
        if not (x.is_cuda and x.dtype == torch.float32 and x.dim() == 4 and x.shape[1] == 3 and x.shape[0] >= 2):
            raise TypeError("x must be a CUDA float32 [1+bs,3,H,W] tensor (I-frame first)")
        x = x.contiguous()
        input_image = x[1:]
        bs, c, h, w = input_image.size()
        if h % 64 or w % 64:
            raise ValueError("H and W must be multiples of 64 (got %dx%d)" % (h, w))
        g, layers, parents = graph_from_batch(bs, isLinear=('-L' in self.name), isOnehop=('-O' in self.name))
        ref_index = refidx_from_graph(g, bs)
        t0 = time.perf_counter()
        dev = x.device
        with torch.cuda.device(dev):
            # ---- phase A: flow + MV codec for all frames at once (models.py:1350-1351) --------------------
            ctx = self._context(bs, h, w, dev)
            ref_orig = x[torch.as_tensor(ref_index, device=dev)].contiguous()
            mv_hat = torch.empty((bs, 2, h, w), device=dev, dtype=torch.float32)
            bits_mv = torch.empty((), device=dev, dtype=torch.float32)
            check(lib().fvc_lsvc_mv_forward(ctx.handle, ptr(input_image), ptr(ref_orig), ptr(mv_hat), ptr(bits_mv),
                                            stream_ptr()), "fvc_lsvc_mv_forward")
            # ---- phase B: tree compensation, layer by layer (models.py:1353-1391) ------------------------
            com = torch.empty_like(input_image)
            mc = torch.empty_like(input_image)
            warped = torch.empty_like(input_image)
            sums = torch.zeros(5, device=dev, dtype=torch.float32)
            for layer in layers:
                tars = [t for t in layer if t <= bs]
                if not tars:
                    continue
                n = len(tars)
                ref = torch.cat([x[:1] if parents[t] == 0 else com[parents[t] - 1:parents[t]] for t in tars], 0)
                idx = torch.as_tensor([t - 1 for t in tars], device=dev)
                diff = mv_hat[idx].contiguous()
                target = input_image[idx].contiguous()
                lctx = self._context(n, h, w, dev)
                o_com, o_mc, o_warp = (torch.empty((n, 3, h, w), device=dev) for _ in range(3))
                o_sums = torch.empty(5, device=dev, dtype=torch.float32)
                check(lib().fvc_lsvc_mc_res_forward(lctx.handle, ptr(target), ptr(ref.contiguous()), ptr(diff),
                                                    ptr(o_com), ptr(o_mc), ptr(o_warp), ptr(o_sums), stream_ptr()),
                      "fvc_lsvc_mc_res_forward")
                com[idx], mc[idx], warped[idx] = o_com, o_mc, o_warp
                sums += o_sums
        self._last_ctx = ctx
        self.encoding_time = self.decoding_time = time.perf_counter() - t0
        cnt = float(bs * c * h * w)
        rec_loss, warp_loss, mc_loss = sums[0] / cnt, sums[1] / cnt, sums[2] / cnt
        bpp_res = (sums[3] + sums[4]) / (bs * h * w)
        bpp_mv = bits_mv / (bs * h * w)
        bpp = (bpp_res + bpp_mv) * self.r_bpp
        return com, mc, warped, rec_loss, warp_loss, mc_loss, bpp_res, bpp
