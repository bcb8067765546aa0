"""``torch.library`` registration of the C-ABI entry points (north_star: "PyTorch custom ops reached through a thin
C-ABI extension"; SURVEY 8b last clause).

The ops live in the ``fvc::`` namespace, so ``torch.no_grad``, stream ordering, ``torch.ops.fvc.*`` look-ups and fake
(meta) shape propagation compose with the rest of a PyTorch program.  Each op is a thin shim over the same ctypes
binding the rest of the package uses: CUDA fp32 tensors in, CUDA fp32 tensors out, the work happens in libfvc_b200.so on
torch's current stream.  There is no autograd formula (inference path; the reference's training branch is out of scope).

    fvc::pframe_forward(cur, ref, ctx)            -> (recon [B,3,H,W], scalars [7])      net.py:70-220
    fvc::decode_from_latents(ref, qmv, fhat, ctx) -> recon                               net.py:77-80, 101-105
    fvc::flow_warp(im, flow)                      -> warped                              endecoder.py:52-67, 116-119
    fvc::conv2d(x, w, b, stride, transposed, act, impl) -> y                             nn.Conv2d / nn.ConvTranspose2d + act
    fvc::quant_bits_factorized(x, params[11])     -> (q, bits)                           net.py:153-205
    fvc::quant_bits_laplace(x, sigma)             -> (q, bits)                           net.py:121-151
"""
from __future__ import annotations

from typing import List, Tuple

import torch

from . import ops
from ._lib import check, lib, ptr, stream_ptr

_lib_def = torch.library.Library("fvc", "DEF")
_lib_def.define("pframe_forward(Tensor cur, Tensor ref, int ctx) -> (Tensor, Tensor)")
_lib_def.define("decode_from_latents(Tensor ref, Tensor quant_mv, Tensor feat_hat, int ctx) -> Tensor")
_lib_def.define("flow_warp(Tensor im, Tensor flow) -> Tensor")
_lib_def.define("conv2d(Tensor x, Tensor weight, Tensor bias, int stride, bool transposed, int act, int impl) -> Tensor")
_lib_def.define("quant_bits_factorized(Tensor x, Tensor[] params) -> (Tensor, Tensor)")
_lib_def.define("quant_bits_laplace(Tensor x, Tensor sigma) -> (Tensor, Tensor)")


def _pframe_forward(cur: torch.Tensor, ref: torch.Tensor, ctx: int) -> Tuple[torch.Tensor, torch.Tensor]:
    cur, ref = cur.contiguous(), ref.contiguous()
    recon = torch.empty_like(cur)
    scalars = torch.empty(7, device=cur.device, dtype=torch.float32)
    with torch.cuda.device(cur.device):
        check(lib().fvc_pframe_forward(ctx, ptr(cur), ptr(ref), ptr(recon), ptr(scalars), stream_ptr()),
              "fvc_pframe_forward")
    return recon, scalars


def _decode_from_latents(ref: torch.Tensor, quant_mv: torch.Tensor, feat_hat: torch.Tensor, ctx: int) -> torch.Tensor:
    ref, quant_mv, feat_hat = ref.contiguous(), quant_mv.contiguous(), feat_hat.contiguous()
    recon = torch.empty_like(ref)
    with torch.cuda.device(ref.device):
        check(lib().fvc_decode_from_latents(ctx, ptr(ref), ptr(quant_mv), ptr(feat_hat), ptr(recon), stream_ptr()),
              "fvc_decode_from_latents")
    return recon


def _conv2d(x, weight, bias, stride: int, transposed: bool, act: int, impl: int):
    return (ops.conv_transpose2d if transposed else ops.conv2d)(x, weight, bias, stride, act, impl)


def _quant_bits_factorized(x: torch.Tensor, params: List[torch.Tensor]):
    return ops.quant_bits_factorized(x, list(params))


_lib_impl = torch.library.Library("fvc", "IMPL", "CUDA")
_lib_impl.impl("pframe_forward", _pframe_forward)
_lib_impl.impl("decode_from_latents", _decode_from_latents)
_lib_impl.impl("flow_warp", ops.flow_warp)
_lib_impl.impl("conv2d", _conv2d)
_lib_impl.impl("quant_bits_factorized", _quant_bits_factorized)
_lib_impl.impl("quant_bits_laplace", ops.quant_bits_laplace)


# fake (meta) kernels: shapes only, so the ops can be traced / shape-propagated
def _scalar_like(x):
    return x.new_empty(())


@torch.library.register_fake("fvc::pframe_forward")
def _(cur, ref, ctx):
    return torch.empty_like(cur), cur.new_empty((7,))


@torch.library.register_fake("fvc::decode_from_latents")
def _(ref, quant_mv, feat_hat, ctx):
    return torch.empty_like(ref)


@torch.library.register_fake("fvc::flow_warp")
def _(im, flow):
    return torch.empty_like(im)


@torch.library.register_fake("fvc::conv2d")
def _(x, weight, bias, stride, transposed, act, impl):
    B, _, H, W = x.shape
    if transposed:
        return x.new_empty((B, weight.shape[1], H * stride, W * stride))
    return x.new_empty((B, weight.shape[0], H // stride, W // stride))


@torch.library.register_fake("fvc::quant_bits_factorized")
def _(x, params):
    return torch.empty_like(x), _scalar_like(x)


@torch.library.register_fake("fvc::quant_bits_laplace")
def _(x, sigma):
    return torch.empty_like(x), _scalar_like(x)
