"""ctypes binding of libfvc_b200.so (C ABI declared in include/fvc_b200.h).

The library is the product: if it is missing, or there is no CUDA device when a compute entry
point is called, we raise — there is no CPU or PyTorch fallback on this path.
"""
from __future__ import annotations

import ctypes as C
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfvc_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "fvc_b200.h")

ACT_NONE, ACT_RELU, ACT_LRELU01, ACT_EXP, ACT_LRELU001 = 0, 1, 2, 3, 4
IMPL_SIMT, IMPL_TC, IMPL_TC_FAST = 0, 1, 2

_lib = None

_f = C.c_void_p   # device float*
_i = C.c_int
_l = C.c_int64
_s = C.c_void_p   # cudaStream_t

_SIGNATURES = {
    "fvc_version": (C.c_int, []),
    "fvc_last_error": (C.c_char_p, []),
    "fvc_avg_pool2": (_i, [_f, _f, _i, _i, _i, _s]),
    "fvc_upsample2x_bilinear": (_i, [_f, _f, _i, _i, _i, _i, C.c_float, _s]),
    "fvc_flow_warp": (_i, [_f, _f, _f, _i, _i, _i, _i, _s]),
    "fvc_conv2d": (_i, [_f, _f, _f, _f, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _s]),
    "fvc_conv_op_create": (_i, [C.POINTER(C.c_void_p), _f, _f, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _s]),
    "fvc_conv_op_run": (_i, [C.c_void_p, _f, _f, _s]),
    "fvc_conv_op_destroy": (None, [C.c_void_p]),
    "fvc_gdn": (_i, [_f, _f, _f, _f, _i, _i, _i, _i, _i, _s]),
    "fvc_quant_bits_factorized": (_i, [_f, C.POINTER(C.c_void_p), _f, _f, _i, _i, _i, _i, _s]),
    "fvc_quant_bits_laplace": (_i, [_f, _f, _f, _f, _l, _s]),
    "fvc_recon_losses": (_i, [_f, _f, _f, _f, _f, _f, _l, _s]),
    "fvc_eb_forward": (_i, [_f, _f, _f, _f, _f, _f, _i, _i, _i, _i, _s]),
    "fvc_gaussian_forward": (_i, [_f, _f, _f, _f, _f, _f, _l, _s]),
    "fvc_ctx_create": (C.c_void_p, [_i, _i, _i, _i, _i]),
    "fvc_ctx_destroy": (None, [C.c_void_p]),
    "fvc_ctx_set_param": (_i, [C.c_void_p, C.c_char_p, _f, _l, _s]),
    "fvc_ctx_missing_params": (_i, [C.c_void_p]),
    "fvc_pframe_forward": (_i, [C.c_void_p, _f, _f, _f, _f, _s]),
    "fvc_ctx_get_tensor": (_l, [C.c_void_p, C.c_char_p, _f, _l, _s]),
    "fvc_gop_forward_host": (_i, [C.c_void_p, C.c_void_p, _i, C.c_void_p, C.c_void_p, _s]),
    "fvc_gop_forward_host_u8": (_i, [C.c_void_p, C.c_void_p, _i, C.c_void_p, C.c_void_p, _s]),
    "fvc_u8hwc_to_f32chw": (_i, [C.c_void_p, _f, _i, _i, _i, _s]),
    "fvc_iframe_forward": (_i, [C.c_void_p, _f, _f, _f, _s]),
    "fvc_iframe_decode_bitstreams": (_i, [C.c_void_p, C.c_void_p, _l, C.c_void_p, _l, _f, _s]),
    "fvc_lsvc_mv_forward": (_i, [C.c_void_p, _f, _f, _f, _f, _s]),
    "fvc_lsvc_mc_res_forward": (_i, [C.c_void_p, _f, _f, _f, _f, _f, _f, _f, _s]),
    "fvc_decode_from_latents": (_i, [C.c_void_p, _f, _f, _f, _f, _s]),
    "fvc_ctx_force_latents": (_i, [C.c_void_p, _f, _f, _f]),
    "fvc_ctx_saturation_count": (_l, [C.c_void_p, _i, _s]),
    "fvc_ctx_set_realbits": (_i, [C.c_void_p, _i, _i]),
    "fvc_ctx_get_bitstream": (_l, [C.c_void_p, _i, C.c_void_p, _l, _s]),
    "fvc_decode_bitstreams": (_i, [C.c_void_p, _f, C.c_void_p, _l, C.c_void_p, _l, C.c_void_p, _l, _f, _s]),
    "fvc_cdf_table_factorized": (_i, [C.POINTER(C.c_void_p), _i, _i, C.c_void_p, _s]),
    "fvc_cdf_table_laplace": (_i, [_f, _l, _i, C.c_void_p, _s]),
    "fvc_entropy_encode_factorized": (_i, [_f, _l, _i, C.c_void_p, _i, _i, C.c_void_p, _l, C.c_void_p, C.c_void_p, _s]),
    "fvc_entropy_encode_laplace": (_i, [_f, _f, _l, _i, _i, C.c_void_p, _l, C.c_void_p, C.c_void_p, _s]),
    "fvc_entropy_decode_factorized": (_i, [C.c_void_p, _l, _l, _i, C.c_void_p, _i, _i, _f, C.c_void_p, _s]),
    "fvc_entropy_decode_laplace": (_i, [C.c_void_p, _l, _l, _f, _i, _i, _f, C.c_void_p, _s]),
    "fvc_entropy_stream_capacity": (_l, [_l, _i]),
    "fvc_entropy_encode_indexed": (_i, [C.c_void_p, C.c_void_p, _l, C.c_void_p, _i, _i, C.c_void_p, C.c_void_p, _i,
                                        C.c_void_p, _l, C.c_void_p, C.c_void_p, _s]),
    "fvc_entropy_decode_indexed": (_i, [C.c_void_p, _l, _l, C.c_void_p, C.c_void_p, _i, _i, C.c_void_p, C.c_void_p, _i,
                                        C.c_void_p, C.c_void_p, _s]),
    "fvc_entropy_stream_capacity_indexed": (_l, [_l, _i]),
    "fvc_ctx_launch_count": (_l, [C.c_void_p]),
    "fvc_ctx_last_conv_seconds": (C.c_double, [C.c_void_p]),
    "fvc_ctx_profile_text": (C.c_char_p, [C.c_void_p]),
}


def declared_symbols():
    """Every function include/fvc_b200.h declares (parsed from the header)."""
    with open(HEADER_PATH) as f:
        text = f.read()
    return sorted(set(re.findall(r"FVC_API\s+[\w\s\*]+?\b(fvc_\w+)\s*\(", text)))


def lib():
    """Loads (once) and returns the shared library; raises if it is not built."""
    global _lib
    if _lib is None:
        path = os.environ.get("FVC_LIB_PATH", LIB_PATH)   # override: A/B of two builds on one box (tools/ab_lib.py)
        if not os.path.exists(path):
            raise RuntimeError(
                "libfvc_b200.so is not built (%s). Run `python -m fastvideocodec_b200.build`; "
                "there is no fallback path." % path)
        handle = C.CDLL(path)
        for name, (res, args) in _SIGNATURES.items():
            if path != LIB_PATH and not hasattr(handle, name):
                continue                                    # an older build under test lacks newer entry points
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


class FvcError(RuntimeError):
    pass


def check(rc, what=""):
    if rc is not None and rc < 0:
        msg = lib().fvc_last_error()
        raise FvcError("%s failed (%d): %s" % (what or "libfvc_b200 call", rc, (msg or b"").decode()))
    return rc


def stream_ptr():
    """Raw cudaStream_t of torch's current stream (so calls compose with torch ordering)."""
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    return C.c_void_p(t.data_ptr())
