// C ABI: op-level entry points (one per reference op; see include/fvc_b200.h).
#include <vector>
#include <cstring>
#include <algorithm>

#include "fvc_kernels.cuh"

using namespace fvc;

namespace {

struct TmpPool {  // stream-ordered temporaries, released on scope exit
    cudaStream_t s;
    std::vector<void*> ptrs;
    explicit TmpPool(cudaStream_t st) : s(st) {}
    template <typename T>
    int get(T** p, size_t bytes) {
        void* q = nullptr;
        FVC_CUDA(cudaMallocAsync(&q, bytes ? bytes : 16, s));
        ptrs.push_back(q);
        *p = reinterpret_cast<T*>(q);
        return 0;
    }
    ~TmpPool() {
        for (void* p : ptrs) cudaFreeAsync(p, s);
    }
};

int pad_c(int c) { return c <= 32 ? 32 : (c <= 64 ? 64 : 128); }

int have_device() {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        set_error("no CUDA device (%s); libfvc_b200 has no CPU fallback", cudaGetErrorString(e));
        return FVC_ERR_CUDA;
    }
    return 0;
}
#define NEED_DEVICE() do { int _r = have_device(); if (_r) return _r; } while (0)

}  // namespace

extern "C" {

int fvc_avg_pool2(const float* x, float* y, int planes, int H, int W, void* stream) {
    NEED_DEVICE();
    FVC_ARG(x && y && planes >= 0 && H >= 0 && W >= 0 && (H % 2 == 0) && (W % 2 == 0));
    return launch_avg_pool2_planar(x, y, planes, H, W, (cudaStream_t)stream);
}

int fvc_upsample2x_bilinear(const float* x, float* y, int planes, int H, int W, int align_corners, float scale,
                            void* stream) {
    NEED_DEVICE();
    FVC_ARG(x && y && planes >= 0 && H >= 0 && W >= 0);
    return launch_upsample2x_planar(x, y, planes, H, W, align_corners, scale, (cudaStream_t)stream);
}

int fvc_u8hwc_to_f32chw(const uint8_t* src, float* dst, int n, int H, int W, void* stream) {
    NEED_DEVICE();
    FVC_ARG(src && dst && n >= 0 && H >= 0 && W >= 0);
    return launch_u8hwc_to_f32chw(src, dst, n, H, W, (cudaStream_t)stream);
}

int fvc_flow_warp(const float* img, const float* flow, float* out, int B, int C, int H, int W, void* stream) {
    NEED_DEVICE();
    FVC_ARG(img && flow && out && B >= 0 && C >= 0 && H >= 2 && W >= 2);
    return launch_flow_warp_nchw(img, flow, out, B, C, H, W, (cudaStream_t)stream);
}

// One convolution with its packed weights, its engine plan and its staging tensors, kept across calls.  The plan binds
// the input / output buffers (TMA descriptors), so the handle owns them and is specific to one input shape.
struct fvc_conv_op {
    ConvLayer L;
    int B = 0, Cin = 0, H = 0, W = 0, Cout = 0, Ho = 0, Wo = 0, impl = 0, device = 0;
    ActT in{}, via{};
    float* out_nhwc = nullptr;
    float* bias = nullptr;
    TcPlan* plan = nullptr;
    SimtWeights sw{};
    Epilogue ep{};
};

void fvc_conv_op_destroy(fvc_conv_op* op) {
    if (!op) return;
    if (op->plan) tc_plan_destroy(op->plan);
    cudaFree(op->sw.w);
    cudaFree(op->in.p);
    cudaFree(op->via.p);
    cudaFree(op->out_nhwc);
    cudaFree(op->bias);
    delete op;
}

int fvc_conv_op_create(fvc_conv_op** out, const float* w, const float* bias, int B, int Cin, int H, int W, int Cout, int k,
                       int stride, int transposed, int act, int impl, void* stream) {
    NEED_DEVICE();
    FVC_ARG(out && w && bias);
    FVC_ARG(B >= 1 && Cin >= 1 && Cin <= 128 && Cout >= 1 && Cout <= 128 && (k == 1 || k == 3 || k == 5 || k == 7));
    FVC_ARG(H >= 1 && W >= 1 && (stride == 1 || stride == 2));
    FVC_ARG(impl == FVC_IMPL_SIMT || impl == FVC_IMPL_TC || impl == FVC_IMPL_TC_FAST);
    FVC_ARG(stride == 1 || (H % 2 == 0 && W % 2 == 0) || transposed);
    cudaStream_t s = (cudaStream_t)stream;
    fvc_conv_op* op = new fvc_conv_op();
    struct Guard {   // releases the half-built handle on every error return
        fvc_conv_op* op;
        ~Guard() { if (op) fvc_conv_op_destroy(op); }
    } guard{op};
    FVC_CUDA(cudaGetDevice(&op->device));
    make_conv_layer(op->L, Cin, Cout, k, stride, transposed);
    op->B = B; op->Cin = Cin; op->H = H; op->W = W; op->Cout = Cout; op->impl = impl;
    op->Ho = transposed ? H * stride : H / stride;
    op->Wo = transposed ? W * stride : W / stride;
    op->in.B = B; op->in.H = H; op->in.W = W;
    op->in.Cp = (impl != FVC_IMPL_SIMT && Cin <= 8) ? 8 : pad_c(Cin);   // narrow records for the input layers
    op->in.parity = (!transposed && stride == 2) ? 1 : 0;
    FVC_CUDA(cudaMalloc(&op->in.p, act_bytes(B, H, W, op->in.Cp)));
    FVC_CUDA(cudaMemsetAsync(op->in.p, 0, act_bytes(B, H, W, op->in.Cp), s));   // padding channels stay zero
    FVC_CUDA(cudaMalloc(&op->out_nhwc, (size_t)B * op->Ho * op->Wo * Cout * 4));
    FVC_CUDA(cudaMalloc(&op->bias, (size_t)Cout * 4));
    FVC_CUDA(cudaMemcpyAsync(op->bias, bias, (size_t)Cout * 4, cudaMemcpyDeviceToDevice, s));
    Epilogue& ep = op->ep;
    memset(&ep, 0, sizeof(ep));
    ep.bias = op->bias;
    ep.acc_scale = 1.f;
    ep.act = act;
    ep.out_f32 = op->out_nhwc;
    // test hook: FVC_CONV2D_VIA_ACT=1 (2: parity-planar) writes the result as an ACT record tensor, the way the layers
    // of the frame pipeline hand their outputs on (exercises the ACT / TMA-store epilogues), and converts it back
    {
        const char* va = getenv("FVC_CONV2D_VIA_ACT");
        const int mode = va ? atoi(va) : 0;
        if (mode && impl != FVC_IMPL_SIMT) {
            ActT& via = op->via;
            via.B = B; via.H = op->Ho; via.W = op->Wo; via.Cp = pad_c(Cout);
            via.parity = (mode == 2 && op->Ho % 2 == 0 && op->Wo % 2 == 0 && !(transposed && stride == 2)) ? 1 : 0;
            FVC_CUDA(cudaMalloc(&via.p, act_bytes(B, op->Ho, op->Wo, via.Cp)));
            FVC_CUDA(cudaMemsetAsync(via.p, 0, act_bytes(B, op->Ho, op->Wo, via.Cp), s));
            ep.out_act = via;
            ep.out_f32 = nullptr;
        }
    }
    int rc = 0;
    if (impl != FVC_IMPL_SIMT) {
        if (!tc_supported(op->L, op->in.Cp)) {
            set_error("fvc_conv2d: shape not supported by the tcgen05 engine");
            return FVC_ERR_ARG;
        }
        rc = tc_plan_create(op->L, w, op->in, op->Ho, op->Wo, ep, &op->plan, s, impl == FVC_IMPL_TC_FAST);
    } else {
        rc = simt_pack_weights(op->L, w, op->in.Cp, std::max(pad_c(Cout), 32), &op->sw, s);
    }
    if (rc) return rc;
    guard.op = nullptr;
    *out = op;
    return 0;
}

int fvc_conv_op_run(fvc_conv_op* op, const float* x, float* y, void* stream) {
    NEED_DEVICE();
    FVC_ARG(op && x && y);
    int dev = -1;
    FVC_CUDA(cudaGetDevice(&dev));
    if (dev != op->device) {
        set_error("fvc_conv_op_run: the op was created on device %d, the current device is %d", op->device, dev);
        return FVC_ERR_STATE;
    }
    cudaStream_t s = (cudaStream_t)stream;
    int rc = launch_nchw_to_act(x, op->in, op->Cin, 0, s);
    if (rc) return rc;
    if (op->plan) rc = tc_plan_launch(op->plan, s);
    else rc = launch_conv_simt(op->L, op->sw, op->in, op->Ho, op->Wo, op->ep, s);
    if (rc) return rc;
    if (op->via.p) return launch_act_to_nchw(op->via, op->Cout, y, s);
    return launch_nhwc_to_nchw(op->out_nhwc, y, op->B, op->Cout, op->Ho, op->Wo, s);
}

int fvc_conv2d(const float* x, const float* w, const float* bias, float* y, int B, int Cin, int H, int W, int Cout,
               int k, int stride, int transposed, int act, int impl, void* stream) {
    FVC_ARG(x && w && bias && y);
    fvc_conv_op* op = nullptr;
    int rc = fvc_conv_op_create(&op, w, bias, B, Cin, H, W, Cout, k, stride, transposed, act, impl, stream);
    if (rc) return rc;
    rc = fvc_conv_op_run(op, x, y, stream);
    // the handle's buffers are plain allocations: wait for the work before releasing them
    if (cudaStreamSynchronize((cudaStream_t)stream) != cudaSuccess && rc == 0)
        rc = cuda_fail(cudaGetLastError(), "sync", __FILE__, __LINE__);
    fvc_conv_op_destroy(op);
    return rc;
}

int fvc_gdn(const float* x, const float* beta, const float* gamma, float* y, int B, int C, int H, int W,
            int inverse, void* stream) {
    NEED_DEVICE();
    FVC_ARG(x && beta && gamma && y && C >= 8 && C <= 64 && C % 8 == 0);
    cudaStream_t s = (cudaStream_t)stream;
    TmpPool tmp(s);
    ActT in, out;
    in.B = B; in.H = H; in.W = W; in.Cp = pad_c(C); in.parity = 0;
    out = in;
    float *be = nullptr, *ge = nullptr;
    if (tmp.get(&in.p, act_bytes(B, H, W, in.Cp)) || tmp.get(&out.p, act_bytes(B, H, W, in.Cp)) ||
        tmp.get(&be, C * 4) || tmp.get(&ge, (size_t)C * C * 4))
        return FVC_ERR_CUDA;
    int rc = launch_nchw_to_act(x, in, C, 0, s);
    if (!rc) rc = launch_gdn_reparam(beta, gamma, be, ge, C, s);
    if (!rc) rc = launch_gdn_act(in, C, be, ge, inverse, out, s);
    if (!rc) rc = launch_act_to_nchw(out, C, y, s);
    return rc;
}

int fvc_quant_bits_factorized(const float* x, const float* const* params, float* q_out, float* bits_out, int B,
                              int C, int H, int W, void* stream) {
    NEED_DEVICE();
    FVC_ARG(params && bits_out && B >= 0 && C >= 1 && H >= 0 && W >= 0);
    cudaStream_t s = (cudaStream_t)stream;
    if ((int64_t)B * C * H * W == 0) {
        FVC_CUDA(cudaMemsetAsync(bits_out, 0, 4, s));
        return 0;
    }
    FVC_ARG(x != nullptr);
    TmpPool tmp(s);
    float* partials = nullptr;
    if (tmp.get(&partials, (size_t)bits_max_blocks() * 4)) return FVC_ERR_CUDA;
    FactorizedParams prm;
    for (int i = 0; i < 11; ++i) {
        FVC_ARG(params[i] != nullptr);
        prm.p[i] = params[i];
    }
    ActT none;
    memset(&none, 0, sizeof(none));
    int nb = 0;
    if ((int64_t)B * C * H * W == 0) {
        FVC_CUDA(cudaMemsetAsync(bits_out, 0, 4, s));
        return 0;
    }
    int rc = launch_quant_bits_factorized(x, 0, B, C, H * W, prm, q_out, none, partials, &nb, s);
    if (rc) return rc;
    return launch_reduce_partials(partials, nb, 1, 1.0, bits_out, s);
}

int fvc_quant_bits_laplace(const float* x, const float* sigma, float* q_out, float* bits_out, int64_t n,
                           void* stream) {
    NEED_DEVICE();
    FVC_ARG(bits_out && n >= 0);
    cudaStream_t s = (cudaStream_t)stream;
    if (n == 0) {
        FVC_CUDA(cudaMemsetAsync(bits_out, 0, 4, s));
        return 0;
    }
    FVC_ARG(x && sigma);
    TmpPool tmp(s);
    float* partials = nullptr;
    if (tmp.get(&partials, (size_t)bits_max_blocks() * 4)) return FVC_ERR_CUDA;
    ActT none;
    memset(&none, 0, sizeof(none));
    int nb = 0;
    int rc = launch_quant_bits_laplace(x, sigma, n, 1, q_out, none, partials, &nb, s);
    if (rc) return rc;
    return launch_reduce_partials(partials, nb, 1, 1.0, bits_out, s);
}

int fvc_recon_losses(const float* cur, const float* pred, const float* warp, const float* res, float* clipped_out,
                     float* means_out, int64_t n, void* stream) {
    NEED_DEVICE();
    FVC_ARG(cur && pred && warp && res && clipped_out && means_out && n > 0 && n % 3 == 0);
    cudaStream_t s = (cudaStream_t)stream;
    TmpPool tmp(s);
    float* partials = nullptr;
    if (tmp.get(&partials, (size_t)148 * 8 * 3 * 4)) return FVC_ERR_CUDA;
    int nb = 0;
    int rc = launch_recon_losses(cur, pred, warp, res, 0, 1, (int)(n / 3), clipped_out, partials, &nb, s);
    if (rc) return rc;
    return launch_reduce_partials(partials, nb, 3, 1.0 / (double)n, means_out, s);
}

int fvc_eb_forward(const float* x, const float* packed_params, const float* medians, float* xhat_out, float* lik_out,
                   float* bits_out, int B, int C, int H, int W, void* stream) {
    NEED_DEVICE();
    FVC_ARG(x && packed_params && medians && bits_out);
    cudaStream_t s = (cudaStream_t)stream;
    if ((int64_t)B * C * H * W == 0) {
        FVC_CUDA(cudaMemsetAsync(bits_out, 0, 4, s));
        return 0;
    }
    TmpPool tmp(s);
    float* partials = nullptr;
    if (tmp.get(&partials, (size_t)bits_max_blocks() * 4)) return FVC_ERR_CUDA;
    int nb = 0;
    int rc = launch_eb_forward(x, packed_params, medians, xhat_out, lik_out, partials, &nb, B, C, H * W, s);
    if (rc) return rc;
    return launch_reduce_partials(partials, nb, 1, 1.0, bits_out, s);
}

int fvc_gaussian_forward(const float* x, const float* scales, const float* means, float* xhat_out, float* lik_out,
                         float* bits_out, int64_t n, void* stream) {
    NEED_DEVICE();
    FVC_ARG(x && scales && bits_out && n >= 0);
    cudaStream_t s = (cudaStream_t)stream;
    if (n == 0) {
        FVC_CUDA(cudaMemsetAsync(bits_out, 0, 4, s));
        return 0;
    }
    TmpPool tmp(s);
    float* partials = nullptr;
    if (tmp.get(&partials, (size_t)bits_max_blocks() * 4)) return FVC_ERR_CUDA;
    int nb = 0;
    int rc = launch_gaussian_forward(x, scales, means, xhat_out, lik_out, partials, &nb, n, s);
    if (rc) return rc;
    return launch_reduce_partials(partials, nb, 1, 1.0, bits_out, s);
}

/* ---- entropy coding, op level (net.py:123-138, 155-168, 183-195; fvc_entropy.cu) -------------------------------- */
int fvc_cdf_table_factorized(const float* const* params, int C, int mxrange, uint32_t* table_out, void* stream) {
    NEED_DEVICE();
    FVC_ARG(params && table_out && C >= 1);
    FactorizedParams prm;
    for (int i = 0; i < 11; ++i) {
        FVC_ARG(params[i] != nullptr);
        prm.p[i] = params[i];
    }
    return launch_cdf_table_factorized(prm, C, mxrange, table_out, (cudaStream_t)stream);
}

int fvc_cdf_table_laplace(const float* sigma, int64_t n, int mxrange, uint32_t* table_out, void* stream) {
    NEED_DEVICE();
    FVC_ARG(sigma && table_out && n >= 0);
    return launch_cdf_table_laplace(sigma, n, mxrange, table_out, (cudaStream_t)stream);
}

static int entropy_encode_common(bool laplace, const float* x, const float* sigma, const uint32_t* table, int64_t n, int C,
                                 int R, int L, uint8_t* out, int64_t capacity, uint32_t* nbytes_out, uint32_t* err_out,
                                 cudaStream_t s) {
    FVC_ARG(x && out && nbytes_out && err_out && n >= 1 && L >= 1 && L <= 65533);
    FVC_ARG((int64_t)entropy_stream_capacity(n, L) <= capacity);
    TmpPool tmp(s);
    uint32_t *packed = nullptr, *lane_words = nullptr;
    uint16_t* words = nullptr;
    if (tmp.get(&packed, (size_t)n * 8) || tmp.get(&words, entropy_words_capacity(n, L) * 2) ||
        tmp.get(&lane_words, (size_t)cdiv64(n, L) * 4))
        return FVC_ERR_CUDA;
    FVC_CUDA(cudaMemsetAsync(err_out, 0, 12, s));
    int rc = laplace ? launch_sym_laplace(x, sigma, n, R, packed, err_out, s)
                     : launch_sym_factorized(x, n, C, R, table, packed, err_out, s);
    if (rc) return rc;
    return launch_rans_encode(packed, n, L, words, lane_words, out, nbytes_out, s);
}

int fvc_entropy_encode_factorized(const float* x, int64_t n, int C, const uint32_t* table, int mxrange, int lane_len,
                                  void* stream_out, int64_t capacity, uint32_t* nbytes_out, uint32_t* err_out,
                                  void* stream) {
    NEED_DEVICE();
    FVC_ARG(table && C >= 1);
    return entropy_encode_common(false, x, nullptr, table, n, C, mxrange, lane_len, (uint8_t*)stream_out, capacity,
                                 nbytes_out, err_out, (cudaStream_t)stream);
}

int fvc_entropy_encode_laplace(const float* x, const float* sigma, int64_t n, int mxrange, int lane_len,
                               void* stream_out, int64_t capacity, uint32_t* nbytes_out, uint32_t* err_out,
                               void* stream) {
    NEED_DEVICE();
    FVC_ARG(sigma != nullptr);
    return entropy_encode_common(true, x, sigma, nullptr, n, 1, mxrange, lane_len, (uint8_t*)stream_out, capacity,
                                 nbytes_out, err_out, (cudaStream_t)stream);
}

int fvc_entropy_decode_factorized(const void* stream_in, int64_t nbytes, int64_t n, int C, const uint32_t* table,
                                  int mxrange, int lane_len, float* q_out, uint32_t* err_out, void* stream) {
    NEED_DEVICE();
    FVC_ARG(stream_in && table && q_out && err_out && n >= 1 && C >= 1);
    FVC_CUDA(cudaMemsetAsync(err_out, 0, 12, (cudaStream_t)stream));
    return launch_rans_decode_factorized((const uint8_t*)stream_in, nbytes, n, lane_len, C, mxrange, table, q_out,
                                         err_out, (cudaStream_t)stream);
}

int fvc_entropy_decode_laplace(const void* stream_in, int64_t nbytes, int64_t n, const float* sigma, int mxrange,
                               int lane_len, float* q_out, uint32_t* err_out, void* stream) {
    NEED_DEVICE();
    FVC_ARG(stream_in && sigma && q_out && err_out && n >= 1);
    FVC_CUDA(cudaMemsetAsync(err_out, 0, 12, (cudaStream_t)stream));
    return launch_rans_decode_laplace((const uint8_t*)stream_in, nbytes, n, lane_len, mxrange, sigma, q_out, err_out,
                                      (cudaStream_t)stream);
}

int fvc_entropy_encode_indexed(const int32_t* symbols, const int32_t* indexes, int64_t n, const int32_t* cdf, int ntab,
                               int cdf_stride, const int32_t* cdf_length, const int32_t* offset, int lane_len,
                               void* stream_out, int64_t capacity, uint32_t* nbytes_out, uint32_t* err_out,
                               void* stream) {
    NEED_DEVICE();
    FVC_ARG(symbols && indexes && cdf && cdf_length && offset && stream_out && nbytes_out && err_out);
    FVC_ARG(n >= 1 && ntab >= 1 && cdf_stride >= 2 && lane_len >= 1 && entropy_indexed_slot_words(lane_len) <= 65535);
    FVC_ARG((int64_t)entropy_stream_capacity_indexed(n, lane_len) <= capacity);
    cudaStream_t s = (cudaStream_t)stream;
    TmpPool tmp(s);
    uint32_t* lane_words = nullptr;
    uint16_t* words = nullptr;
    const int64_t nlanes = cdiv64(n, lane_len);
    if (tmp.get(&words, (size_t)nlanes * entropy_indexed_slot_words(lane_len) * 2) || tmp.get(&lane_words, (size_t)nlanes * 4))
        return FVC_ERR_CUDA;
    FVC_CUDA(cudaMemsetAsync(err_out, 0, 12, s));
    return launch_rans_encode_indexed(symbols, indexes, n, lane_len, cdf, ntab, cdf_stride, cdf_length, offset, words,
                                      lane_words, (uint8_t*)stream_out, nbytes_out, err_out, s);
}

int fvc_entropy_decode_indexed(const void* stream_in, int64_t nbytes, int64_t n, const int32_t* indexes,
                               const int32_t* cdf, int ntab, int cdf_stride, const int32_t* cdf_length,
                               const int32_t* offset, int lane_len, int32_t* symbols_out, uint32_t* err_out,
                               void* stream) {
    NEED_DEVICE();
    FVC_ARG(stream_in && indexes && cdf && cdf_length && offset && symbols_out && err_out);
    FVC_ARG(n >= 1 && ntab >= 1 && cdf_stride >= 2 && lane_len >= 1);
    FVC_CUDA(cudaMemsetAsync(err_out, 0, 12, (cudaStream_t)stream));
    return launch_rans_decode_indexed((const uint8_t*)stream_in, nbytes, n, lane_len, indexes, cdf, ntab, cdf_stride,
                                      cdf_length, offset, symbols_out, err_out, (cudaStream_t)stream);
}

int64_t fvc_entropy_stream_capacity_indexed(int64_t n, int lane_len) {
    if (n < 0 || lane_len < 1) return FVC_ERR_ARG;
    return (int64_t)entropy_stream_capacity_indexed(n, lane_len);
}

int64_t fvc_entropy_stream_capacity(int64_t n, int lane_len) {
    if (n < 0 || lane_len < 1) return FVC_ERR_ARG;
    return (int64_t)entropy_stream_capacity(n, lane_len);
}

}  // extern "C"
