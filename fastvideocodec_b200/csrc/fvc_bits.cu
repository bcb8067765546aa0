// Fused quantise + likelihood + clamp-log2 + block reduction kernels (sm_100a).
// One pass over the latent: read x (and sigma), write round(x) (as the decoder's ACT input and/or
// fp32), and reduce the per-element bit cost with warp shuffles to one partial per block; partials
// are summed in a fixed order afterwards (deterministic).  Accurate libm-grade intrinsics are used
// on purpose: p = F(q+.5)-F(q-.5) cancels, and the bpp parity gate is on the sum.
#include "fvc_kernels.cuh"
#include "fvc_bits.cuh"

namespace fvc {

static const int kBitsThreads = 256;
static const int kBitsMaxBlocks = 148 * 8;
int bits_max_blocks() { return kBitsMaxBlocks; }

__global__ void __launch_bounds__(kBitsThreads)
k_quant_bits_factorized(const float* __restrict__ x, int nhwc, int B, int C, int HW, FactorizedParams prm,
                        float* __restrict__ q_f32, ActT q_act, float* __restrict__ partials) {
    pdl_sync();
    extern __shared__ float smf[];  // ChanParams[C]
    ChanParams* cp = reinterpret_cast<ChanParams*>(smf);
    __shared__ float red[32];
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        ChanParams p;
        p.sp1 = softplusf(prm.p[0][c]); p.b1 = prm.p[1][c]; p.ta1 = tanhf(prm.p[2][c]);
        p.sp2 = softplusf(prm.p[3][c]); p.b2 = prm.p[4][c]; p.ta2 = tanhf(prm.p[5][c]);
        p.sp3 = softplusf(prm.p[6][c]); p.b3 = prm.p[7][c]; p.ta3 = tanhf(prm.p[8][c]);
        p.sp4 = softplusf(prm.p[9][c]); p.b4 = prm.p[10][c];
        cp[c] = p;
    }
    __syncthreads();
    int64_t n = (int64_t)B * C * HW;
    float acc = 0.f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int c;
        int64_t pix;  // b*HW + p
        if (nhwc) {
            c = (int)(i % C);
            pix = i / C;
        } else {
            int p = (int)(i % HW);
            c = (int)((i / HW) % C);
            pix = (i / ((int64_t)HW * C)) * HW + p;
        }
        float q = rintf(x[i]);  // torch.round: half to even (net.py:76, 91)
        if (q_f32) q_f32[i] = q;
        if (q_act.p) {
            // geometry of q_act equals the latent grid; pix indexes [B,H,W] row-major
            size_t off = (size_t)pix * (size_t)(2 * q_act.Cp);
            e16 hi, lo;
            split16(q, hi, lo);
            q_act.p[off + c] = hi;
            q_act.p[off + q_act.Cp + c] = lo;
        }
        const ChanParams& p = cp[c];
        float prob = factorized_cdf(q + 0.5f, p) - factorized_cdf(q - 0.5f, p);
        acc += bits_of(prob);
    }
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) partials[blockIdx.x] = acc;
}

// zero the padding channels of an ACT record tensor (C..Cp-1), both halves
__global__ void k_zero_pad_channels(ActT t, int C) {
    pdl_sync();
    int pad = t.Cp - C;
    int64_t n = (int64_t)t.B * t.H * t.W * pad;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int c = C + (int)(i % pad);
    size_t off = (size_t)(i / pad) * (size_t)(2 * t.Cp);
    t.p[off + c] = 0;
    t.p[off + t.Cp + c] = 0;
}
static int zero_pad(ActT t, int C, cudaStream_t s) {
    if (!t.p || t.Cp == C) return 0;
    int64_t n = (int64_t)t.B * t.H * t.W * (t.Cp - C);
    FVC_CUDA(launch_pdl(k_zero_pad_channels, (unsigned)cdiv64(n, 256), 256, 0, s, t, C));
    g_launch_count++;
    FVC_CHECK_LAUNCH();
    return 0;
}

int launch_quant_bits_factorized(const float* x, int nhwc, int B, int C, int HW, FactorizedParams prm, float* q_f32,
                                 ActT q_act, float* partials, int* nblocks, cudaStream_t s) {
    FVC_ARG(C <= 1024);
    FVC_ARG(!q_act.p || (q_act.parity == 0 && q_act.Cp >= C));
    int64_t n = (int64_t)B * C * HW;
    int blocks = (int)std::min<int64_t>(std::max<int64_t>(cdiv64(n, kBitsThreads * 4), 1), kBitsMaxBlocks);
    size_t smem = (size_t)C * sizeof(ChanParams);
    FVC_CUDA(launch_pdl(k_quant_bits_factorized, blocks, kBitsThreads, smem, s, x, nhwc, B, C, HW, prm, q_f32, q_act, partials));
    g_launch_count++;
    FVC_CHECK_LAUNCH();
    *nblocks = blocks;
    return zero_pad(q_act, C, s);
}


// x, sigma: fp32 NHWC [npix, C]; net.py:121-151
__global__ void __launch_bounds__(kBitsThreads)
k_quant_bits_laplace(const float* __restrict__ x, const float* __restrict__ sigma, int64_t n, int C,
                     float* __restrict__ q_f32, ActT q_act, float* __restrict__ partials) {
    pdl_sync();
    __shared__ float red[32];
    float acc = 0.f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float q = rintf(x[i]);  // net.py:100
        float sg = fminf(fmaxf(sigma[i], 1e-5f), 1e10f);
        if (q_f32) q_f32[i] = q;
        if (q_act.p) {
            int c = (int)(i % C);
            size_t off = (size_t)(i / C) * (size_t)(2 * q_act.Cp);
            e16 hi, lo;
            split16(q, hi, lo);
            q_act.p[off + c] = hi;
            q_act.p[off + q_act.Cp + c] = lo;
        }
        float prob = laplace_cdf(q + 0.5f, sg) - laplace_cdf(q - 0.5f, sg);
        acc += bits_of(prob);
    }
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) partials[blockIdx.x] = acc;
}
int launch_quant_bits_laplace(const float* x, const float* sigma, int64_t n, int C, float* q_f32, ActT q_act,
                              float* partials, int* nblocks, cudaStream_t s) {
    FVC_ARG(!q_act.p || (q_act.parity == 0 && q_act.Cp >= C));
    int blocks = (int)std::min<int64_t>(std::max<int64_t>(cdiv64(n, kBitsThreads * 4), 1), kBitsMaxBlocks);
    FVC_CUDA(launch_pdl(k_quant_bits_laplace, blocks, kBitsThreads, 0, s, x, sigma, n, C, q_f32, q_act, partials));
    g_launch_count++;
    FVC_CHECK_LAUNCH();
    *nblocks = blocks;
    return zero_pad(q_act, C, s);
}

// ----------------------------------------------------------------------------------------------
// CompressAI-compatible likelihoods (entropy_models.py:55-68, 202-219).  PARITY UNPINNED.
// EntropyBottleneck with filters (3,3,3,3): per channel 58 floats packed as
//   M0[3x1] b0[3] f0[3] | M1[3x3] b1[3] f1[3] | M2[3x3] b2[3] f2[3] | M3[3x3] b3[3] f3[3] | M4[1x3] b4[1]
// (softplus of M and tanh of f are applied here, as CompressAI does at run time).
// ----------------------------------------------------------------------------------------------
#define FVC_EB_STRIDE 58
__device__ __forceinline__ float eb_logits(float v, const float* __restrict__ p) {
    float l[3], t[3];
    // layer 0: 1 -> 3
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        float z = softplusf(p[j]) * v + p[3 + j];
        l[j] = z + tanhf(p[6 + j]) * tanhf(z);
    }
    p += 9;
    // layers 1..3: 3 -> 3
#pragma unroll
    for (int k = 0; k < 3; ++k) {
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            float z = softplusf(p[j * 3 + 0]) * l[0] + softplusf(p[j * 3 + 1]) * l[1] +
                      softplusf(p[j * 3 + 2]) * l[2] + p[9 + j];
            t[j] = z + tanhf(p[12 + j]) * tanhf(z);
        }
        l[0] = t[0]; l[1] = t[1]; l[2] = t[2];
        p += 15;
    }
    // layer 4: 3 -> 1, no factor
    return softplusf(p[0]) * l[0] + softplusf(p[1]) * l[1] + softplusf(p[2]) * l[2] + p[3];
}

__global__ void __launch_bounds__(kBitsThreads)
k_eb_forward(const float* __restrict__ x, const float* __restrict__ packed, const float* __restrict__ medians,
             float* __restrict__ xhat, float* __restrict__ lik, float* __restrict__ partials, int B, int C, int HW) {
    __shared__ float red[32];
    int64_t n = (int64_t)B * C * HW;
    float acc = 0.f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int c = (int)((i / HW) % C);
        float med = medians[c];
        float v = rintf(x[i] - med) + med;
        const float* p = packed + (size_t)c * FVC_EB_STRIDE;
        float lower = eb_logits(v - 0.5f, p);
        float upper = eb_logits(v + 0.5f, p);
        float sm = lower + upper;
        float sgn = (sm > 0.f) ? -1.f : ((sm < 0.f) ? 1.f : 0.f);
        float l = fabsf(sigmoidf(sgn * upper) - sigmoidf(sgn * lower));
        l = fmaxf(l, 1e-9f);
        if (xhat) xhat[i] = v;
        if (lik) lik[i] = l;
        acc += bits_of(l);
    }
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) partials[blockIdx.x] = acc;
}
int launch_eb_forward(const float* x, const float* packed, const float* medians, float* xhat, float* lik,
                      float* partials, int* nblocks, int B, int C, int HW, cudaStream_t s) {
    int64_t n = (int64_t)B * C * HW;
    int blocks = (int)std::min<int64_t>(std::max<int64_t>(cdiv64(n, kBitsThreads * 4), 1), kBitsMaxBlocks);
    k_eb_forward<<<blocks, kBitsThreads, 0, s>>>(x, packed, medians, xhat, lik, partials, B, C, HW);
    g_launch_count++;
    FVC_CHECK_LAUNCH();
    *nblocks = blocks;
    return 0;
}

__global__ void __launch_bounds__(kBitsThreads)
k_gaussian_forward(const float* __restrict__ x, const float* __restrict__ scales, const float* __restrict__ means,
                   float* __restrict__ xhat, float* __restrict__ lik, float* __restrict__ partials, int64_t n) {
    __shared__ float red[32];
    float acc = 0.f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float mu = means ? means[i] : 0.f;
        float v = rintf(x[i] - mu) + mu;
        float s = fmaxf(scales[i], 0.11f);
        float a = fabsf(v - mu);
        const float cst = -0.70710678118654752440f;
        float upper = 0.5f * erfcf(cst * ((0.5f - a) / s));
        float lower = 0.5f * erfcf(cst * ((-0.5f - a) / s));
        float l = fmaxf(upper - lower, 1e-9f);
        if (xhat) xhat[i] = v;
        if (lik) lik[i] = l;
        acc += bits_of(l);
    }
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) partials[blockIdx.x] = acc;
}
int launch_gaussian_forward(const float* x, const float* scales, const float* means, float* xhat, float* lik,
                            float* partials, int* nblocks, int64_t n, cudaStream_t s) {
    int blocks = (int)std::min<int64_t>(std::max<int64_t>(cdiv64(n, kBitsThreads * 4), 1), kBitsMaxBlocks);
    k_gaussian_forward<<<blocks, kBitsThreads, 0, s>>>(x, scales, means, xhat, lik, partials, n);
    g_launch_count++;
    FVC_CHECK_LAUNCH();
    *nblocks = blocks;
    return 0;
}

}  // namespace fvc
