// fp32 CUDA-core convolution for layers with 2-3 output channels and few input channels (SpyNet conv5:
// 16->2, 7x7): on the tensor cores these need N = 16 MMAs, which are bound by the
// A-operand shared-memory bandwidth at ~20 % of the pipe (tools/mma_rate.cu) and still cost three MMAs per
// product for the hi/lo split.  With so few outputs the FLOPs are small; exact fp32 FMAs on the CUDA cores are
// 3-5x faster here (measured) and need no operand split.
//
// Block = 128 threads = 32 (x) by 4; output tile 32 x 16; a thread owns 4 vertically adjacent pixels of one
// column, so neighbouring threads read neighbouring float4s (conflict-free) and each input value loaded from
// shared memory is used for up to 4 x CO x 4 FMAs.  Input channels are processed in chunks of 16.
#include <atomic>

#include "fvc_kernels.cuh"
#include "fvc_epilogue.cuh"

namespace fvc {

struct FewParams {
    ActT in;              // stride-1, non-parity ACT records
    const float* w;       // packed [chunk][tap][c4][co] float4 (4 consecutive input channels)
    const float* bias;    // [CO]
    const float* res_f32; // optional NHWC [B,H,W,CO]
    float* out_f32;       // NHWC [B,H,W,CO]
    ActT out_act;         // optional ACT output (channels >= CO are left untouched: buffer is zero-initialised)
    int Cin, H, W, B;
};

__global__ void k_few_pack(const float* __restrict__ w, float* __restrict__ out, int Cin, int Cout, int K) {
    // out[(((chunk*K*K + tap)*4 + c4)*Cout + co)*4 + ch] = w[co][chunk*16 + c4*4 + ch][r][s]
    const int nchunk = (Cin + 15) / 16;
    const size_t n = (size_t)nchunk * K * K * 4 * Cout * 4;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int ch = (int)(i & 3);
    size_t j = i >> 2;
    int co = (int)(j % Cout); j /= Cout;
    int c4 = (int)(j & 3); j >>= 2;
    int tap = (int)(j % (K * K));
    int chunk = (int)(j / (K * K));
    int ci = chunk * 16 + c4 * 4 + ch;
    out[i] = ci < Cin ? w[((size_t)co * Cin + ci) * K * K + tap] : 0.f;
}

// Packed fp32 FMA (Blackwell FFMA2): d.{x,y} += a.{x,y} * b.{x,y}, two IEEE fp32 FMAs per instruction.  The operands are
// adjacent register pairs of the float4s already loaded from shared memory, so no moves are needed.
__device__ __forceinline__ void ffma2(float2& d, float ax, float ay, float bx, float by) {
    asm("{\n\t.reg .b64 a, b, c;\n\t"
        "mov.b64 a, {%2, %3};\n\t"
        "mov.b64 b, {%4, %5};\n\t"
        "mov.b64 c, {%0, %1};\n\t"
        "fma.rn.f32x2 c, a, b, c;\n\t"
        "mov.b64 {%0, %1}, c;\n\t}"
        : "+f"(d.x), "+f"(d.y)
        : "f"(ax), "f"(ay), "f"(bx), "f"(by));
}

template <int K, int CO>
__global__ void __launch_bounds__(128) k_conv_few(const __grid_constant__ FewParams P) {
    pdl_sync();
    constexpr int TW = 32, TH = 16, PW = TW + K - 1, PH = TH + K - 1, R = K / 2;
    extern __shared__ float4 smf[];
    float4* patch = smf;                       // [PH][4][PW]
    float4* wq = smf + PH * 4 * PW;            // [K*K][4][CO]
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH, b = blockIdx.z;
    // two partial sums per (pixel, output): even and odd input channels of each float4 (packed FFMA2); added at the end
    float2 acc[4][CO];
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int c = 0; c < CO; ++c) acc[p][c] = make_float2(0.f, 0.f);
    const int nchunk = P.Cin / 16;
    const int Cp = P.in.Cp;
    for (int chunk = 0; chunk < nchunk; ++chunk) {
        __syncthreads();
        // ---- fill: 16 channels of the halo'ed patch as fp32 (hi + lo), zero outside the image ----------
        for (int i = threadIdx.x; i < PH * PW; i += 128) {
            const int py = i / PW, px = i - py * PW;
            const int y = y0 + py - R, x = x0 + px - R;
            float v[16];
            if (y >= 0 && y < P.H && x >= 0 && x < P.W) {
                const e16* rec = P.in.p + (((size_t)b * P.H + y) * P.W + x) * (size_t)(2 * Cp) + chunk * 16;
                uint32_t hh[8], ll[8];   // 16 channels = one 32-byte sector each for the hi and the lo halves
                ld_global_nc_v8(rec, hh);
                ld_global_nc_v8(rec + Cp, ll);
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    float a0, a1, b0, b1;
                    e2f2(hh[q], a0, a1);
                    e2f2(ll[q], b0, b1);
                    v[2 * q] = a0 + b0;
                    v[2 * q + 1] = a1 + b1;
                }
            } else {
#pragma unroll
                for (int q = 0; q < 16; ++q) v[q] = 0.f;
            }
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4)
                patch[(py * 4 + c4) * PW + px] = make_float4(v[4 * c4], v[4 * c4 + 1], v[4 * c4 + 2], v[4 * c4 + 3]);
        }
        {
            const float4* src = reinterpret_cast<const float4*>(P.w) + (size_t)chunk * K * K * 4 * CO;
            for (int i = threadIdx.x; i < K * K * 4 * CO; i += 128) wq[i] = __ldg(src + i);
        }
        __syncthreads();
        // ---- compute ---------------------------------------------------------------------------------
#pragma unroll 1
        for (int s = 0; s < K; ++s) {
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) {
                float4 v[K + 3];
#pragma unroll
                for (int r = 0; r < K + 3; ++r) v[r] = patch[((ty * 4 + r) * 4 + c4) * PW + tx + s];
#pragma unroll
                for (int r = 0; r < K; ++r) {
#pragma unroll
                    for (int c = 0; c < CO; ++c) {
                        const float4 w = wq[((r * K + s) * 4 + c4) * CO + c];
#pragma unroll
                        for (int p = 0; p < 4; ++p) {
                            ffma2(acc[p][c], v[p + r].x, v[p + r].y, w.x, w.y);
                            ffma2(acc[p][c], v[p + r].z, v[p + r].w, w.z, w.w);
                        }
                    }
                }
            }
        }
    }
    // ---- epilogue ------------------------------------------------------------------------------------
    const int x = x0 + tx;
    if (x >= P.W) return;
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const int y = y0 + ty * 4 + p;
        if (y >= P.H) continue;
        const size_t pix = ((size_t)b * P.H + y) * P.W + x;
        float o[CO];
#pragma unroll
        for (int c = 0; c < CO; ++c) {
            o[c] = (acc[p][c].x + acc[p][c].y) + P.bias[c];
            if (P.res_f32) o[c] += P.res_f32[pix * CO + c];
        }
        if (CO == 2) *reinterpret_cast<float2*>(P.out_f32 + pix * 2) = make_float2(o[0], o[1]);
        else {
#pragma unroll
            for (int c = 0; c < CO; ++c) P.out_f32[pix * CO + c] = o[c];
        }
        if (P.out_act.p) {
            e16* rec = P.out_act.p + act_pixel_offset(P.out_act, b, y, x);
#pragma unroll
            for (int c = 0; c < CO; ++c) {
                e16 hi, lo;
                split16(o[c], hi, lo);
                rec[c] = hi;
                rec[P.out_act.Cp + c] = lo;
            }
        }
    }
}

bool few_supported(const ConvLayer& L, int CinP) {
    // measured: wins for few input channels and large kernels (SpyNet conv5: 0.92 -> 0.29 ms at 1080p);
    // loses for 3x3 layers with 64-128 input channels (per-chunk patch refills dominate), which stay on
    // the tensor-core engine
    return !L.transposed && L.stride == 1 && L.k == 7 && (L.Cout == 2 || L.Cout == 3) && L.Cin == 16 &&
           CinP >= L.Cin;
}

int few_pack_weights(const ConvLayer& L, const float* w_ref, float** out, cudaStream_t s) {
    const int nchunk = L.Cin / 16;
    const size_t n = (size_t)nchunk * L.k * L.k * 4 * L.Cout * 4;
    if (!*out) FVC_CUDA(cudaMalloc(out, n * sizeof(float)));
    k_few_pack<<<(unsigned)cdiv64((int64_t)n, 256), 256, 0, s>>>(w_ref, *out, L.Cin, L.Cout, L.k);
    g_launch_count++;
    FVC_CHECK_LAUNCH();
    return 0;
}

template <int K, int CO>
static int few_launch(const FewParams& P, cudaStream_t s) {
    constexpr int PW = 32 + K - 1, PH = 16 + K - 1;
    const size_t smem = ((size_t)PH * 4 * PW + (size_t)K * K * 4 * CO) * sizeof(float4);
    static std::atomic<unsigned long long> attr{0};   // per-device attribute: one bit per device ordinal
    int dev = 0;
    FVC_CUDA(cudaGetDevice(&dev));
    const unsigned long long bit = 1ull << (dev & 63);
    if (!(attr.load(std::memory_order_acquire) & bit)) {
        FVC_CUDA(cudaFuncSetAttribute(k_conv_few<K, CO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr.fetch_or(bit, std::memory_order_release);
    }
    dim3 grid(cdiv(P.W, 32), cdiv(P.H, 16), P.B);
    FVC_CUDA(launch_pdl(k_conv_few<K, CO>, grid, 128, smem, s, P));
    g_launch_count++;
    FVC_CHECK_LAUNCH();
    return 0;
}

int launch_conv_few(const ConvLayer& L, const float* w_packed, const float* bias, ActT in, int Hout, int Wout,
                    const Epilogue& ep, cudaStream_t s) {
    FVC_ARG(few_supported(L, in.Cp) && !in.parity && in.H == Hout && in.W == Wout);
    FVC_ARG(ep.act == FVC_ACT_NONE && !ep.res_act.p && !ep.out_act_relu.p && !ep.out_act_sq.p && ep.out_f32);
    FewParams P;
    P.in = in; P.w = w_packed; P.bias = bias; P.res_f32 = ep.res_f32; P.out_f32 = ep.out_f32; P.out_act = ep.out_act;
    P.Cin = L.Cin; P.H = Hout; P.W = Wout; P.B = in.B;
    if (L.k == 7 && L.Cout == 2) return few_launch<7, 2>(P, s);
    if (L.k == 3 && L.Cout == 2) return few_launch<3, 2>(P, s);
    if (L.k == 3 && L.Cout == 3) return few_launch<3, 3>(P, s);
    if (L.k == 7 && L.Cout == 3) return few_launch<7, 3>(P, s);
    set_error("launch_conv_few: unsupported shape");
    return FVC_ERR_ARG;
}

}  // namespace fvc
