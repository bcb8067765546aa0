// Shared device/host helpers for libfvc_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <algorithm>
#include <utility>

#include "../../include/fvc_b200.h"

namespace fvc {

// ----------------------------------------------------------------------------------------------
// error plumbing
// ----------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define FVC_CUDA(expr)                                                        \
    do {                                                                      \
        cudaError_t _e = (expr);                                              \
        if (_e != cudaSuccess) return fvc::cuda_fail(_e, #expr, __FILE__, __LINE__); \
    } while (0)
#define FVC_CHECK_LAUNCH() FVC_CUDA(cudaGetLastError())
#define FVC_ARG(cond)                                                         \
    do {                                                                      \
        if (!(cond)) {                                                        \
            fvc::set_error("bad argument: %s (%s:%d)", #cond, __FILE__, __LINE__); \
            return FVC_ERR_ARG;                                               \
        }                                                                     \
    } while (0)

extern thread_local int64_t g_launch_count;  // kernels launched by this library on this thread

// ----------------------------------------------------------------------------------------------
// Activation record format ("ACT"): NHWC, per pixel [hi: Cp x 16 bit][lo: Cp x 16 bit]; value = hi + lo,
// hi = rn16(v), lo = rn16(v - hi).  Both halves feed tcgen05 kind::f16 MMAs directly.
//   FVC_SPLIT_FP16=1 (default): IEEE half pairs, 22 significant bits (abs floor 3e-8, |v| clamped to
//                               65504) - the precision the parity gates need (measured);
//   FVC_SPLIT_FP16=0          : bfloat16 pairs, 16 significant bits, fp32 range.
// parity layout (inputs of stride-2 convs): 4 planes [(y&1)*2+(x&1)][H/2][W/2][record].
// ----------------------------------------------------------------------------------------------
#ifndef FVC_SPLIT_FP16
#define FVC_SPLIT_FP16 1
#endif
typedef uint16_t e16;  // one 16-bit storage element of a record (half or bfloat16 bits)

__device__ __forceinline__ e16 f2e(float v) {
#if FVC_SPLIT_FP16
    return __half_as_ushort(__float2half_rn(v));
#else
    return __bfloat16_as_ushort(__float2bfloat16_rn(v));
#endif
}
__device__ __forceinline__ float e2f(e16 h) {
#if FVC_SPLIT_FP16
    return __half2float(__ushort_as_half(h));
#else
    return __uint_as_float((uint32_t)h << 16);
#endif
}
// two packed elements (low half = first) -> floats
__device__ __forceinline__ void e2f2(uint32_t u, float& a, float& b) {
#if FVC_SPLIT_FP16
    float2 f = __half22float2(*reinterpret_cast<__half2*>(&u));
    a = f.x; b = f.y;
#else
    a = __uint_as_float(u << 16);
    b = __uint_as_float(u & 0xffff0000u);
#endif
}

struct ActT {
    e16* p;
    int B, H, W, Cp;
    int parity;
};

__host__ __device__ inline size_t act_bytes(int B, int H, int W, int Cp) {
    return (size_t)B * H * W * Cp * 2 * sizeof(e16);
}

__device__ __forceinline__ size_t act_pixel_offset(const ActT& t, int b, int y, int x) {
    size_t pix;
    if (t.parity) {
        int plane = ((y & 1) << 1) | (x & 1);
        pix = (((size_t)(b * 4 + plane) * (t.H >> 1) + (y >> 1)) * (t.W >> 1) + (x >> 1));
    } else {
        pix = ((size_t)b * t.H + y) * t.W + x;
    }
    return pix * (size_t)(2 * t.Cp);
}

__device__ __forceinline__ void split16(float v, e16& hi, e16& lo) {
#if FVC_SPLIT_FP16
    v = fminf(fmaxf(v, -65504.f), 65504.f);  // keep hi finite (NaN passes through)
#endif
    hi = f2e(v);
    lo = f2e(v - e2f(hi));
}

__device__ __forceinline__ float act_load(const ActT& t, size_t pixoff, int c) {
    return e2f(t.p[pixoff + c]) + e2f(t.p[pixoff + t.Cp + c]);
}

__device__ __forceinline__ void act_store(const ActT& t, size_t pixoff, int c, float v) {
    e16 hi, lo;
    split16(v, hi, lo);
    t.p[pixoff + c] = hi;
    t.p[pixoff + t.Cp + c] = lo;
}

__device__ __forceinline__ uint32_t pack16x2(e16 a, e16 b) { return (uint32_t)a | ((uint32_t)b << 16); }

// warp + block sum (blockDim.x multiple of 32, <= 1024); result valid in thread 0
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float block_sum(float v, float* smem32) {
    v = warp_sum(v);
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) smem32[w] = v;
    __syncthreads();
    int nw = (blockDim.x + 31) >> 5;
    v = (threadIdx.x < nw) ? smem32[threadIdx.x] : 0.f;
    if (w == 0) v = warp_sum(v);
    return v;
}

// ----------------------------------------------------------------------------------------------
// Programmatic dependent launch (PDL): every kernel of the frame pipeline starts with pdl_sync() and is launched with
// the programmatic-stream-serialization attribute, so its CTAs may become resident and run their prologue (barrier
// init, TMEM allocation, descriptor prefetch, bias staging) while the previous kernel drains; griddepcontrol.wait
// then blocks until the previous grid has completed and its writes are visible.  FVC_PDL=0 launches normally.
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_sync() {
    asm volatile("griddepcontrol.launch_dependents;");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

inline int cdiv(int a, int b) { return (a + b - 1) / b; }
inline int64_t cdiv64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ----------------------------------------------------------------------------------------------
// Convolution description shared by both engines.  A layer is 1 sub-convolution (conv, or
// stride-1 transposed conv) or 4 (stride-2 transposed conv, one per output parity phase); every
// sub-convolution is a gather  out(q*os + ph) = sum_taps in(q*st + d_tap) * W_tap.
// ----------------------------------------------------------------------------------------------
#define FVC_MAX_TAPS 49

struct SubConv {
    int ntaps;
    int py, px;                 // output phase (transposed stride 2), else 0
    int8_t dy[FVC_MAX_TAPS];    // input offset of the tap on the (strided) q grid
    int8_t dx[FVC_MAX_TAPS];
    int8_t r[FVC_MAX_TAPS];     // kernel coordinates of the tap in the reference weight
    int8_t s[FVC_MAX_TAPS];
};

struct ConvLayer {
    int Cin, Cout, k, stride, transposed;
    int nsub;
    SubConv sub[4];
    int st;   // input step on the q grid (2 for a stride-2 conv, else 1)
    int os;   // output step (2 for a stride-2 transposed conv, else 1)
};

void make_conv_layer(ConvLayer& L, int Cin, int Cout, int k, int stride, int transposed);

// Epilogue description (device-visible POD)
struct Epilogue {
    const float* bias;        // [Cout]
    float acc_scale;          // accumulators are multiplied by this before the bias (1/weight scale)
    int act;                  // FVC_ACT_*
    ActT res_act;             // optional residual (p == nullptr: none), same geometry as the output
    const float* res_f32;     // optional fp32 NHWC residual [B,Hout,Wout,Cout]
    ActT out_act;             // optional ACT output (p == nullptr: none)
    ActT out_act_relu;        // optional second ACT output holding relu(y)
    float* out_f32;           // optional fp32 NHWC output [B,Hout,Wout,Cout]
    // (I)GDN fused into the producing convolution's epilogue (tcgen05 engine; GDN.py:63-93): y = x / sqrt(norm) or
    // x * sqrt(norm), norm_i = beta_i + sum_j gamma_ij x_j^2 as a second 128 x 64 x 64 MMA on the squared tile
    const float* gdn_beta;    // [64] beta_eff; nullptr: no fused GDN
    const e16* gdn_gamma;     // packed gamma_eff stream of the 1x1 "norm" convolution: tiles [hi][lo][hi], 64 rows x 128 B
    int gdn_inverse;
    float gdn_scale;          // accumulator scale of the norm MMA: 1 / (sq_scale * weight scale of gamma)
    // Fused 3x3 "tail" convolution with 2-3 output channels (mvDecoder.deconv8 after deconv7, Warp_net conv6 after
    // conv5.conv2; tcgen05 engine): this layer's output y never reaches memory; a second MMA multiplies the staged y
    // tile by the tail's weights regrouped as a 1x1 convolution with 9 * Cout_tail (padded to 32) outputs, one per
    // (tap, channel); the per-pixel partial products P go out as fp32 [B,H,W,tap_cq] and k_tapsum adds the 9 shifted
    // taps.  nullptr: not fused.
    const e16* tap_w;         // packed stream of the regrouped weights: 32-row tiles, per 64-channel segment [hi][lo], then [hi]
    float* tap_out;           // P
    float tap_scale;          // 1 / weight scale of tap_w
    int tap_cq;               // floats per pixel of P (multiple of 4, <= 32)
    // tcgen05 engine only:
    ActT out_act_sq;          // optional ACT output holding sq_scale * y^2 (input of the GDN 1x1 convolution)
    float sq_scale;
    int res_mode;             // how res_act enters: 0: y += r;  1: y = r / sqrt(y) (GDN);  2: y = r * sqrt(y) (IGDN)
    unsigned int* sat_count;  // optional: incremented once per epilogue tile that stored an ACT value with |hi| >= 65504
                              // (the fp16 pair range; such values are clamped by cvt.satfinite and the result is wrong)
};

}  // namespace fvc
