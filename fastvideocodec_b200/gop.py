"""GOP driver and codec registry for the DVC path — mirrors reference models.py.

* ``get_codec_model('DVC-pretrained', ...)`` / ``get_DVC_pretrained`` — models.py:32-36, 1432-1445
* ``parallel_compression`` ('DVC-pretrained' branch) — models.py:233-410 (branch 368-383)
* ``PSNR`` — models.py:460-473, ``AverageMeter`` — models.py:1414-1430
* GOP sharding across ranks + the one statistics all-reduce: GOPs are independent,
  frames inside a GOP are sequential, so rank r codes GOPs g = r (mod world) with no collective on
  the data path.
"""
from __future__ import annotations

import math
import os

import torch
import torch.nn as nn


class AverageMeter(object):
    """reference models.py:1414-1430."""

    def __init__(self):
        self.reset()

    def reset(self):
        self.val = 0
        self.avg = 0
        self.sum = 0
        self.count = 0

    def update(self, val, n=1):
        self.val = val
        self.sum += val * n
        self.count += n
        self.avg = self.sum / self.count


def PSNR(Y1_raw, Y1_com, use_list=False):
    """reference models.py:460-473."""
    Y1_com = Y1_com.to(Y1_raw.device)
    log10 = torch.log(torch.FloatTensor([10])).squeeze(0).to(Y1_raw.device)
    if not use_list:
        train_mse = torch.mean(torch.pow(Y1_raw - Y1_com, 2))
        return 10.0 * torch.log(1 / train_mse) / log10
    quality = []
    for i in range(Y1_raw.size()[0]):
        train_mse = torch.mean(torch.pow(Y1_raw[i:i + 1] - Y1_com[i:i + 1].unsqueeze(0), 2))
        quality.append(10.0 * torch.log(1 / train_mse) / log10)
    return quality


def get_DVC_pretrained(level, snapshot_dir="DVC/snapshot", device="cuda"):
    """reference models.py:1432-1445.  Snapshots are loaded when present (they are not shipped)."""
    from .net import VideoCompressor, load_model
    model = VideoCompressor()
    model.name = 'DVC-pretrained'
    model.compression_level = level
    model.loss_type = 'P'
    ratio_list = [256, 512, 1024, 2048, 2048 * 2, 2048 * 4, 2048 * 8]
    I_lvl_list = [37, 32, 27, 22, 17, 12, 7]
    model.I_level = I_lvl_list[level]
    model.r = ratio_list[level]
    path = os.path.join(snapshot_dir, f'{ratio_list[level]}.model')
    if level < 4 and os.path.exists(path):
        load_model(model, path)
    return model.to(device).eval()


def get_codec_model(name, loss_type='P', compression_level=2, noMeasure=True, use_split=True, num_views=0,
                    resilience=0, use_attn=True, load_with_copy=False):
    """reference models.py:32-66 — 'DVC-pretrained' (the hot path) and the LSVC names that share its
    sub-networks are built."""
    if name in ['DVC-pretrained']:
        return get_DVC_pretrained(compression_level)
    if 'LSVC' in name:
        from .lsvc import LSVC
        return LSVC(name, loss_type=loss_type, compression_level=compression_level, use_split=use_split)
    raise ValueError("codec %r is outside the B200 hot path ('DVC-pretrained' and 'LSVC*-128' are built)" % (name,))


def parallel_compression(args, model, data, compressI=False, level=0, batch_idx=0, i_codec=None):
    """reference models.py:233-410, 'DVC-pretrained' branch.

    ``data``: [G,3,H,W] on the GPU.  The reference shells out to bpgenc/bpgdec for frame 0
    (I_compression, models.py:412-429: external binaries, out of scope); here frame 0 is used as the
    decoded I-frame unless ``i_codec(frame) -> (x_hat, bpp, psnr)`` is supplied.
    Returns the reference 11-tuple.
    """
    img_loss_list, bpp_list, psnr_list, aux_loss_list, aux2_loss_list = [], [], [], [], []
    if data.dim() != 4:
        raise ValueError("the DVC-pretrained branch takes a [G,3,H,W] GOP")
    if i_codec is not None:
        x_hat0, bpp0, psnr0 = i_codec(data[0:1])
        data[0:1] = x_hat0
        if compressI:
            bpp_list += [bpp0.to(data.device)]
            psnr_list += [psnr0.to(data.device)]
    if 'LSVC' in getattr(model, 'name', ''):
        return _parallel_compression_lsvc(model, data, bpp_list, psnr_list)
    log10 = torch.log(torch.FloatTensor([10])).squeeze(0).to(data.device)
    B = data.size(0)
    x_prev = data[0:1]
    x_hat_list = []
    for i in range(1, B):
        x_prev, mseloss, warploss, interloss, bpp_feature, bpp_z, bpp_mv, bpp = model(data[i:i + 1], x_prev)
        x_prev = x_prev.detach()
        img_loss_list += [model.r * mseloss.to(data.device)]
        aux_loss_list += [10.0 * torch.log(1 / warploss) / log10]
        bpp_list += [bpp.to(data.device)]
        psnr_list += [10.0 * torch.log(1 / mseloss) / log10]
        aux2_loss_list += [10.0 * torch.log(1 / interloss) / log10]
        x_hat_list.append(x_prev)
    x_hat = torch.cat(x_hat_list, dim=0)
    loss = 0
    be_loss = torch.stack(bpp_list, dim=0).mean(dim=0).cpu().data.item()
    be_res_loss = 0
    img_loss = 0
    psnr = torch.stack(psnr_list, dim=0).mean(dim=0).cpu().data.item()
    aux_loss = torch.stack(aux_loss_list, dim=0).mean(dim=0).cpu().data.item() if aux_loss_list else 0
    aux2_loss = torch.stack(aux2_loss_list, dim=0).mean(dim=0).cpu().data.item() if aux2_loss_list else 0
    return (x_hat, loss, img_loss, be_loss, be_res_loss, psnr, torch.stack(psnr_list, dim=0).tolist(), aux_loss,
            aux2_loss, 0, 0)


def _parallel_compression_lsvc(model, data, bpp_list, psnr_list):
    """reference models.py:384-398: one batched / tree forward for the whole GOP."""
    B = data.size(0)
    x_hat, x_mc, x_wp, rec_loss, warp_loss, mc_loss, bpp_res, bpp = model(data.detach())
    img_loss_list = [rec_loss * model.r]
    all_loss_list = [(rec_loss * model.r + bpp).to(data.device)]
    psnr_list += PSNR(data[1:], x_hat, use_list=True)
    aux2_loss_list = PSNR(data[1:], x_mc, use_list=True)
    aux_loss_list = PSNR(data[1:], x_wp, use_list=True)
    x_hat = torch.cat([data[0:1], x_hat], dim=0)
    bpp_list += [bpp.to(data.device) for _ in range(B - 1)]
    loss = torch.stack(all_loss_list, dim=0).sum(dim=0)
    be_loss = torch.stack(bpp_list, dim=0).mean(dim=0).cpu().data.item()
    img_loss = torch.stack(img_loss_list, dim=0).mean(dim=0).cpu().data.item()
    psnr = torch.stack(psnr_list, dim=0).mean(dim=0).cpu().data.item()
    aux_loss = torch.stack(aux_loss_list, dim=0).mean(dim=0).cpu().data.item()
    aux2_loss = torch.stack(aux2_loss_list, dim=0).mean(dim=0).cpu().data.item()
    return (x_hat, loss, img_loss, be_loss, 0, psnr, torch.stack(psnr_list, dim=0).tolist(), aux_loss, aux2_loss, 0, 0)


# ----------------------------------------------------------------------------------------------
# multi-GPU: shard by GOP, reduce statistics once
# ----------------------------------------------------------------------------------------------
def shard_gops(n_gops, rank, world):
    """GOP ids owned by ``rank``: round-robin g = rank (mod world)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    return list(range(rank, n_gops, world))


def stats_vector(scalars):
    """Per-rank sufficient statistics from per-frame rows [n,7] (mse, warp, inter, bpp_f, bpp_z, bpp_mv, bpp):
    [sum bpp, sum psnr, sum mse, n_frames] in fp64."""
    s = torch.as_tensor(scalars, dtype=torch.float64).reshape(-1, 7)
    if s.numel() == 0:
        return torch.zeros(4, dtype=torch.float64)
    psnr = 10.0 * torch.log10(1.0 / s[:, 0])
    return torch.stack([s[:, 6].sum(), psnr.sum(), s[:, 0].sum(), torch.tensor(float(s.shape[0]), dtype=torch.float64)])


def reduce_stats(vec, group=None):
    """The only collective of the path: all_reduce(SUM) of the 4-element statistics vector.

    NCCL needs a CUDA tensor; gloo (CPU tests) takes the CPU tensor as is.
    """
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return vec
    backend = dist.get_backend(group)
    t = vec.clone()
    if backend == "nccl":
        t = t.cuda()
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t.cpu()


def summarize(vec):
    n = max(float(vec[3]), 1.0)
    return {"bpp": float(vec[0]) / n, "psnr": float(vec[1]) / n, "mse": float(vec[2]) / n, "frames": int(vec[3])}
