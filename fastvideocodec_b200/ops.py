"""Op-level Python surface over the C ABI: one function per reference op on the hot path.

Tensors are torch CUDA fp32 tensors used purely as device buffers (PyTorch is plumbing here); each
function mirrors the reference op it replaces (same argument meaning, same result layout).
"""
from __future__ import annotations

import collections
import ctypes as C
import os

import torch

from . import _lib
from ._lib import (ACT_EXP, ACT_LRELU001, ACT_LRELU01, ACT_NONE, ACT_RELU, IMPL_SIMT, IMPL_TC, IMPL_TC_FAST, check, lib, ptr,
                   stream_ptr)

__all__ = ["avg_pool2", "upsample2x_bilinear", "to_tensor_u8", "flow_warp", "conv2d", "conv_transpose2d", "conv_op_cache_clear", "gdn",
           "quant_bits_factorized", "quant_bits_laplace", "recon_losses", "eb_forward", "gaussian_forward",
           "pack_eb_params", "cdf_table_factorized", "cdf_table_laplace", "entropy_encode_factorized",
           "entropy_encode_laplace", "entropy_decode_factorized", "entropy_decode_laplace", "entropy_encode_indexed",
           "entropy_decode_indexed", "ACT_NONE", "ACT_RELU", "ACT_LRELU01", "ACT_LRELU001", "ACT_EXP", "IMPL_SIMT", "IMPL_TC", "IMPL_TC_FAST"]


def _cuda_f32(t, name):
    if not (torch.is_tensor(t) and t.is_cuda and t.dtype == torch.float32):
        raise TypeError("%s must be a CUDA float32 tensor (libfvc_b200 has no CPU path)" % name)
    return t.contiguous()


def avg_pool2(x):
    """F.avg_pool2d(x, 2, 2) — reference DVC/subnet/endecoder.py:344-346."""
    x = _cuda_f32(x, "x")
    B, Cc, H, W = x.shape
    y = torch.empty((B, Cc, H // 2, W // 2), device=x.device, dtype=torch.float32)
    check(lib().fvc_avg_pool2(ptr(x), ptr(y), B * Cc, H, W, stream_ptr()), "fvc_avg_pool2")
    return y


def upsample2x_bilinear(x, align_corners=False, scale=1.0):
    """F.interpolate(x, 2x, 'bilinear', align_corners) * scale — endecoder.py:173-184, 353."""
    x = _cuda_f32(x, "x")
    B, Cc, H, W = x.shape
    y = torch.empty((B, Cc, 2 * H, 2 * W), device=x.device, dtype=torch.float32)
    check(lib().fvc_upsample2x_bilinear(ptr(x), ptr(y), B * Cc, H, W, int(bool(align_corners)), float(scale),
                                        stream_ptr()), "fvc_upsample2x_bilinear")
    return y


def to_tensor_u8(frames_u8):
    """transforms.ToTensor() on the device (dataset.py:75): uint8 [..., H, W, 3] -> float32 [..., 3, H, W] = x / 255."""
    if not (torch.is_tensor(frames_u8) and frames_u8.is_cuda and frames_u8.dtype == torch.uint8 and frames_u8.dim() >= 3
            and frames_u8.shape[-1] == 3):
        raise TypeError("frames_u8 must be a CUDA uint8 tensor [..., H, W, 3] (libfvc_b200 has no CPU path)")
    src = frames_u8.contiguous()
    H, W = src.shape[-3], src.shape[-2]
    n = src.numel() // max(1, H * W * 3)
    out = torch.empty(tuple(src.shape[:-3]) + (3, H, W), device=src.device, dtype=torch.float32)
    with torch.cuda.device(src.device):
        for i0 in range(0, n, 65535):
            k = min(65535, n - i0)
            check(lib().fvc_u8hwc_to_f32chw(C.c_void_p(src.data_ptr() + i0 * H * W * 3),
                                            C.c_void_p(out.data_ptr() + i0 * H * W * 12), k, H, W, stream_ptr()),
                  "fvc_u8hwc_to_f32chw")
    return out


def flow_warp(im, flow):
    """flow_warp(im, flow) — endecoder.py:116-119 / torch_warp 52-67."""
    im, flow = _cuda_f32(im, "im"), _cuda_f32(flow, "flow")
    B, Cc, H, W = im.shape
    if tuple(flow.shape) != (B, 2, H, W):
        raise ValueError("flow must be [B,2,H,W]")
    out = torch.empty_like(im)
    check(lib().fvc_flow_warp(ptr(im), ptr(flow), ptr(out), B, Cc, H, W, stream_ptr()), "fvc_flow_warp")
    return out


class _ConvOpCache:
    """Handles of fvc_conv_op_* (packed weights + engine plan + staging tensors) kept across calls of conv2d /
    conv_transpose2d, so a layer applied once per forward (entropy_models.py:160-190) packs its weights once.

    An entry is keyed by the identity of the weight / bias memory and its version counter (in-place updates such as an
    optimizer step bump it and the next call builds a new handle); the entry holds a reference to both tensors, so their
    memory cannot be freed and handed to other tensors while the key is live."""

    def __init__(self, capacity):
        self.capacity = capacity
        self.entries = collections.OrderedDict()

    def _drop(self, entry):
        torch.cuda.synchronize(entry["device"])        # the handle's buffers may still be in use by queued launches
        lib().fvc_conv_op_destroy(entry["handle"])

    def clear(self):
        while self.entries:
            self._drop(self.entries.popitem()[1])

    def get(self, key, build):
        e = self.entries.get(key)
        if e is not None:
            self.entries.move_to_end(key)
            return e
        e = build()
        self.entries[key] = e
        while len(self.entries) > self.capacity:
            self._drop(self.entries.popitem(last=False)[1])
        return e


_conv_ops = _ConvOpCache(int(os.environ.get("FVC_CONV_OP_CACHE", "32")))


def conv_op_cache_clear():
    """Destroy every cached convolution handle (ops.conv2d / conv_transpose2d / torch.ops.fvc.conv2d)."""
    _conv_ops.clear()


def conv_op_cache_size():
    return len(_conv_ops.entries)


def _conv(x, weight, bias, stride, transposed, act, impl):
    # a non-contiguous weight / bias is copied per call: its memory identity means nothing, so it is never cached
    cacheable = torch.is_tensor(weight) and weight.is_contiguous() and (bias is None or (torch.is_tensor(bias) and bias.is_contiguous()))
    x, weight = _cuda_f32(x, "x"), _cuda_f32(weight, "weight")
    B, Cin, H, W = x.shape
    k = weight.shape[-1]
    if transposed:
        if weight.shape[0] != Cin:
            raise ValueError("ConvTranspose2d weight must be [Cin,Cout,k,k]")
        Cout = weight.shape[1]
        Ho, Wo = H * stride, W * stride
    else:
        if weight.shape[1] != Cin:
            raise ValueError("Conv2d weight must be [Cout,Cin,k,k]")
        Cout = weight.shape[0]
        Ho, Wo = H // stride, W // stride
    if bias is not None:
        bias = _cuda_f32(bias, "bias")
    y = torch.empty((B, Cout, Ho, Wo), device=x.device, dtype=torch.float32)
    if _conv_ops.capacity <= 0 or not cacheable:       # FVC_CONV_OP_CACHE=0: one-shot call, nothing kept
        b = bias if bias is not None else torch.zeros(Cout, device=x.device, dtype=torch.float32)
        check(lib().fvc_conv2d(ptr(x), ptr(weight), ptr(b), ptr(y), B, Cin, H, W, Cout, k, stride, int(transposed),
                               int(act), int(impl), stream_ptr()), "fvc_conv2d")
        return y
    try:
        versions = (weight._version, None if bias is None else bias._version)
    except RuntimeError:                               # inference tensors carry no version counter: do not cache
        versions = None
    if versions is None:
        b = bias if bias is not None else torch.zeros(Cout, device=x.device, dtype=torch.float32)
        check(lib().fvc_conv2d(ptr(x), ptr(weight), ptr(b), ptr(y), B, Cin, H, W, Cout, k, stride, int(transposed),
                               int(act), int(impl), stream_ptr()), "fvc_conv2d")
        return y
    # the engine's planner reads FVC_* environment switches when a plan is built: they are part of the key
    key = (x.device.index, weight.data_ptr(), None if bias is None else bias.data_ptr(), versions,
           B, Cin, H, W, Cout, k, stride, bool(transposed), int(act), int(impl),
           tuple(sorted(kv for kv in os.environ.items() if kv[0].startswith("FVC_"))))

    def build():
        b = bias if bias is not None else torch.zeros(Cout, device=x.device, dtype=torch.float32)
        h = C.c_void_p()
        check(lib().fvc_conv_op_create(C.byref(h), ptr(weight), ptr(b), B, Cin, H, W, Cout, k, stride, int(transposed),
                                       int(act), int(impl), stream_ptr()), "fvc_conv_op_create")
        return {"handle": h, "device": x.device, "weight": weight, "bias": bias}

    with torch.cuda.device(x.device):
        e = _conv_ops.get(key, build)
        check(lib().fvc_conv_op_run(e["handle"], ptr(x), ptr(y), stream_ptr()), "fvc_conv_op_run")
    return y


def conv2d(x, weight, bias=None, stride=1, act=ACT_NONE, impl=IMPL_TC):
    """nn.Conv2d(.., k, stride, padding=k//2) + activation (all convs of the DVC sub-networks)."""
    return _conv(x, weight, bias, stride, False, act, impl)


def conv_transpose2d(x, weight, bias=None, stride=2, act=ACT_NONE, impl=IMPL_TC):
    """nn.ConvTranspose2d(.., k, stride, padding=k//2, output_padding=stride-1) + activation."""
    return _conv(x, weight, bias, stride, True, act, impl)


def gdn(x, beta, gamma, inverse=False):
    """GDN.forward with RAW beta/gamma parameters — reference DVC/subnet/GDN.py:63-93."""
    x, beta, gamma = _cuda_f32(x, "x"), _cuda_f32(beta, "beta"), _cuda_f32(gamma, "gamma")
    B, Cc, H, W = x.shape
    y = torch.empty_like(x)
    check(lib().fvc_gdn(ptr(x), ptr(beta), ptr(gamma), ptr(y), B, Cc, H, W, int(bool(inverse)), stream_ptr()),
          "fvc_gdn")
    return y


def quant_bits_factorized(x, params):
    """round(x) and total bits under a BitEstimator — net.py:153-178 / bitEstimator.py:20-42.

    ``params``: the 11 tensors f1.h f1.b f1.a f2.h f2.b f2.a f3.h f3.b f3.a f4.h f4.b (any shape
    holding C values).  Returns (q, bits) with bits a 0-dim tensor.
    """
    x = _cuda_f32(x, "x")
    B, Cc, H, W = x.shape
    ps = [_cuda_f32(p, "param").reshape(-1) for p in params]
    if len(ps) != 11 or any(p.numel() != Cc for p in ps):
        raise ValueError("need 11 parameter vectors of length C")
    arr = (C.c_void_p * 11)(*[p.data_ptr() for p in ps])
    q = torch.empty_like(x)
    bits = torch.empty((), device=x.device, dtype=torch.float32)
    check(lib().fvc_quant_bits_factorized(ptr(x), arr, ptr(q), ptr(bits), B, Cc, H, W, stream_ptr()),
          "fvc_quant_bits_factorized")
    return q, bits


def quant_bits_laplace(x, sigma):
    """round(x) and total bits under Laplace(0, clamp(sigma)) — net.py:121-151."""
    x, sigma = _cuda_f32(x, "x"), _cuda_f32(sigma, "sigma")
    if x.shape != sigma.shape:
        raise ValueError("x and sigma must have the same shape")
    q = torch.empty_like(x)
    bits = torch.empty((), device=x.device, dtype=torch.float32)
    check(lib().fvc_quant_bits_laplace(ptr(x), ptr(sigma), ptr(q), ptr(bits), x.numel(), stream_ptr()),
          "fvc_quant_bits_laplace")
    return q, bits


def recon_losses(cur, pred, warp, res):
    """clamp(pred+res, 0, 1) and the three distortion means — net.py:103-116."""
    cur, pred, warp, res = (_cuda_f32(t, n) for t, n in ((cur, "cur"), (pred, "pred"), (warp, "warp"), (res, "res")))
    clipped = torch.empty_like(cur)
    means = torch.empty(3, device=cur.device, dtype=torch.float32)
    check(lib().fvc_recon_losses(ptr(cur), ptr(pred), ptr(warp), ptr(res), ptr(clipped), ptr(means), cur.numel(),
                                 stream_ptr()), "fvc_recon_losses")
    return clipped, means


def pack_eb_params(matrices, biases, factors):
    """Packs CompressAI EntropyBottleneck parameters (filters (3,3,3,3)) to the [C,58] layout."""
    Cc = matrices[0].shape[0]
    cols = []
    for i in range(5):
        cols.append(matrices[i].reshape(Cc, -1))
        cols.append(biases[i].reshape(Cc, -1))
        if i < 4:
            cols.append(factors[i].reshape(Cc, -1))
    packed = torch.cat(cols, 1).contiguous().float()
    if packed.shape[1] != 58:
        raise ValueError("expected filters (3,3,3,3): got %d values per channel" % packed.shape[1])
    return packed


def eb_forward(x, packed, medians):
    """EntropyBottleneck eval forward (x_hat, likelihood, bits) — entropy_models.py:66, 74-78."""
    x, packed, medians = _cuda_f32(x, "x"), _cuda_f32(packed, "packed"), _cuda_f32(medians, "medians")
    B, Cc, H, W = x.shape
    xh, lik = torch.empty_like(x), torch.empty_like(x)
    bits = torch.empty((), device=x.device, dtype=torch.float32)
    check(lib().fvc_eb_forward(ptr(x), ptr(packed), ptr(medians.reshape(-1)), ptr(xh), ptr(lik), ptr(bits), B, Cc, H,
                               W, stream_ptr()), "fvc_eb_forward")
    return xh, lik, bits


def gaussian_forward(x, scales, means=None):
    """GaussianConditional eval forward (x_hat, likelihood, bits) — entropy_models.py:63, 218."""
    x, scales = _cuda_f32(x, "x"), _cuda_f32(scales, "scales")
    mp = ptr(_cuda_f32(means, "means")) if means is not None else C.c_void_p(0)
    xh, lik = torch.empty_like(x), torch.empty_like(x)
    bits = torch.empty((), device=x.device, dtype=torch.float32)
    check(lib().fvc_gaussian_forward(ptr(x), ptr(scales), mp, ptr(xh), ptr(lik), ptr(bits), x.numel(), stream_ptr()),
          "fvc_gaussian_forward")
    return xh, lik, bits


# ------------------------------------------------------------------------------------------------
# real entropy coding (calrealbits branch of net.py:123-138, 155-168, 183-195; csrc/fvc_entropy.cu)
# ------------------------------------------------------------------------------------------------
def cdf_table_factorized(params, mxrange=150):
    """Integer CDF tables [C, 2*mxrange] (int64 view of the library's uint32) of a BitEstimator: what the reference
    hands to torchac, ``cdf[i] = bitEstimator(i - mxrange - 0.5)`` converted to 16 bits (net.py:158-160)."""
    ps = [_cuda_f32(p, "param").reshape(-1) for p in params]
    Cc = ps[0].numel()
    if len(ps) != 11 or any(p.numel() != Cc for p in ps):
        raise ValueError("need 11 parameter vectors of length C")
    arr = (C.c_void_p * 11)(*[p.data_ptr() for p in ps])
    out = torch.empty((Cc, 2 * mxrange), device=ps[0].device, dtype=torch.int32)
    check(lib().fvc_cdf_table_factorized(arr, Cc, int(mxrange), ptr(out), stream_ptr()), "fvc_cdf_table_factorized")
    return out


def cdf_table_laplace(sigma, mxrange=150):
    """Per-element integer CDF tables sigma.shape + [2*mxrange] of Laplace(0, clamp(sigma)) (net.py:127-128, 141-143).
    Test use: the coder itself never materialises them."""
    sigma = _cuda_f32(sigma, "sigma")
    out = torch.empty(tuple(sigma.shape) + (2 * mxrange,), device=sigma.device, dtype=torch.int32)
    check(lib().fvc_cdf_table_laplace(ptr(sigma), sigma.numel(), int(mxrange), ptr(out), stream_ptr()),
          "fvc_cdf_table_laplace")
    return out


def _entropy_finish(buf, nbytes, err, what):
    e = err.cpu().tolist()
    if e[0] or e[1] or e[2]:
        raise _lib.FvcError("%s: %d symbols outside [-mxrange, mxrange-2], %d empty intervals, %d unreadable lanes"
                            % (what, e[0], e[1], e[2]))
    return bytes(buf[:int(nbytes.item())].cpu().numpy().tobytes())


def entropy_encode_factorized(x_nhwc, table, mxrange=150, lane_len=8192):
    """rANS-codes round(x) (x: [..., C] channels-last, fp32) under per-channel tables [C, 2*mxrange]; returns bytes."""
    x = _cuda_f32(x_nhwc, "x")
    Cc, n = x.shape[-1], x.numel()
    cap = lib().fvc_entropy_stream_capacity(n, lane_len)
    buf = torch.empty(cap, device=x.device, dtype=torch.uint8)
    nbytes = torch.zeros(1, device=x.device, dtype=torch.int32)
    err = torch.zeros(3, device=x.device, dtype=torch.int32)
    check(lib().fvc_entropy_encode_factorized(ptr(x), n, Cc, ptr(table.contiguous()), int(mxrange), int(lane_len), ptr(buf),
                                              cap, ptr(nbytes), ptr(err), stream_ptr()), "fvc_entropy_encode_factorized")
    return _entropy_finish(buf, nbytes, err, "entropy_encode_factorized")


def entropy_encode_laplace(x, sigma, mxrange=150, lane_len=8192):
    x, sigma = _cuda_f32(x, "x"), _cuda_f32(sigma, "sigma")
    n = x.numel()
    cap = lib().fvc_entropy_stream_capacity(n, lane_len)
    buf = torch.empty(cap, device=x.device, dtype=torch.uint8)
    nbytes = torch.zeros(1, device=x.device, dtype=torch.int32)
    err = torch.zeros(3, device=x.device, dtype=torch.int32)
    check(lib().fvc_entropy_encode_laplace(ptr(x), ptr(sigma), n, int(mxrange), int(lane_len), ptr(buf), cap, ptr(nbytes),
                                           ptr(err), stream_ptr()), "fvc_entropy_encode_laplace")
    return _entropy_finish(buf, nbytes, err, "entropy_encode_laplace")


def _stream_tensor(stream, device):
    import numpy as np
    pad = (-len(stream)) % 4
    return torch.from_numpy(np.frombuffer(stream + b"\0" * pad, dtype=np.uint8).copy()).to(device)


def entropy_decode_factorized(stream, shape_nhwc, table, mxrange=150, lane_len=8192):
    """Inverse of entropy_encode_factorized: returns the integer-valued fp32 tensor of shape ``shape_nhwc``."""
    dev = table.device
    st = _stream_tensor(stream, dev)
    q = torch.empty(tuple(shape_nhwc), device=dev, dtype=torch.float32)
    err = torch.zeros(3, device=dev, dtype=torch.int32)
    check(lib().fvc_entropy_decode_factorized(ptr(st), len(stream), q.numel(), shape_nhwc[-1], ptr(table.contiguous()),
                                              int(mxrange), int(lane_len), ptr(q), ptr(err), stream_ptr()),
          "fvc_entropy_decode_factorized")
    if int(err[2]):
        raise _lib.FvcError("entropy_decode_factorized: %d lanes could not be opened" % int(err[2]))
    return q


def entropy_decode_laplace(stream, sigma, mxrange=150, lane_len=8192):
    sigma = _cuda_f32(sigma, "sigma")
    st = _stream_tensor(stream, sigma.device)
    q = torch.empty_like(sigma)
    err = torch.zeros(3, device=sigma.device, dtype=torch.int32)
    check(lib().fvc_entropy_decode_laplace(ptr(st), len(stream), q.numel(), ptr(sigma), int(mxrange), int(lane_len), ptr(q),
                                           ptr(err), stream_ptr()), "fvc_entropy_decode_laplace")
    if int(err[2]):
        raise _lib.FvcError("entropy_decode_laplace: %d lanes could not be opened" % int(err[2]))
    return q


def _cuda_i32(t, name, device=None):
    if not torch.is_tensor(t):
        raise TypeError("%s must be a tensor" % name)
    if device is not None and t.device != device:
        t = t.to(device)
    if not t.is_cuda:
        raise TypeError("%s must live on a CUDA device (libfvc_b200 has no CPU path)" % name)
    return t.to(torch.int32).contiguous()


def entropy_encode_indexed(symbols, indexes, quantized_cdf, cdf_length, offset, lane_len=8192):
    """rANS-codes integer ``symbols`` under per-element tables ``quantized_cdf[indexes]`` with CompressAI's escape rule
    for out-of-range values (fvc_entropy_encode_indexed; the coder behind EntropyModel.compress).  Returns bytes."""
    sym = _cuda_i32(symbols, "symbols")
    dev = sym.device
    idx = _cuda_i32(indexes, "indexes", dev)
    cdf, ln, off = _cuda_i32(quantized_cdf, "quantized_cdf", dev), _cuda_i32(cdf_length, "cdf_length", dev), _cuda_i32(offset, "offset", dev)
    if idx.numel() != sym.numel() or cdf.dim() != 2 or ln.numel() != cdf.shape[0] or off.numel() != cdf.shape[0]:
        raise ValueError("indexes must match symbols; quantized_cdf [ntab, stride] with one cdf_length / offset per table")
    n = sym.numel()
    if n == 0:                                          # nothing to code: the empty string round-trips
        return b""
    cap = lib().fvc_entropy_stream_capacity_indexed(n, lane_len)
    buf = torch.empty(cap, device=dev, dtype=torch.uint8)
    nbytes = torch.zeros(1, device=dev, dtype=torch.int32)
    err = torch.zeros(3, device=dev, dtype=torch.int32)
    with torch.cuda.device(dev):
        check(lib().fvc_entropy_encode_indexed(ptr(sym), ptr(idx), n, ptr(cdf), cdf.shape[0], cdf.shape[1], ptr(ln), ptr(off),
                                               int(lane_len), ptr(buf), cap, ptr(nbytes), ptr(err), stream_ptr()),
              "fvc_entropy_encode_indexed")
    e = err.cpu().tolist()
    if e[0] or e[1]:
        raise _lib.FvcError("entropy_encode_indexed: %d indexes outside the table set, %d empty intervals (run update())"
                            % (e[0], e[1]))
    return bytes(buf[:int(nbytes.item())].cpu().numpy().tobytes())


def entropy_decode_indexed(stream, indexes, quantized_cdf, cdf_length, offset, lane_len=8192):
    """Inverse of entropy_encode_indexed: int32 symbols with the shape of ``indexes``."""
    idx = _cuda_i32(indexes, "indexes")
    dev = idx.device
    cdf, ln, off = _cuda_i32(quantized_cdf, "quantized_cdf", dev), _cuda_i32(cdf_length, "cdf_length", dev), _cuda_i32(offset, "offset", dev)
    if idx.numel() == 0:
        if len(stream):
            raise _lib.FvcError("entropy_decode_indexed: a non-empty stream for zero symbols")
        return torch.empty(idx.shape, device=dev, dtype=torch.int32)
    st = _stream_tensor(stream, dev)
    out = torch.empty(idx.shape, device=dev, dtype=torch.int32)
    err = torch.zeros(3, device=dev, dtype=torch.int32)
    with torch.cuda.device(dev):
        check(lib().fvc_entropy_decode_indexed(ptr(st), len(stream), idx.numel(), ptr(idx), ptr(cdf), cdf.shape[0],
                                               cdf.shape[1], ptr(ln), ptr(off), int(lane_len), ptr(out), ptr(err),
                                               stream_ptr()), "fvc_entropy_decode_indexed")
    e = err.cpu().tolist()
    if e[0] or e[2]:
        raise _lib.FvcError("entropy_decode_indexed: %d indexes outside the table set, %d lanes could not be opened"
                            % (e[0], e[2]))
    return out
