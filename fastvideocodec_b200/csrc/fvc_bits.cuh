// Device functions of the entropy models shared by the bit-estimation kernels (fvc_bits.cu) and the entropy coder
// (fvc_entropy.cu): BitEstimator CDF (bitEstimator.py:20-42) and Laplace CDF (torch.distributions.Laplace).
#pragma once
#include "fvc_kernels.cuh"

namespace fvc {

__device__ __forceinline__ float softplusf(float x) {  // F.softplus(beta=1, threshold=20)
    return x > 20.f ? x : log1pf(expf(x));
}
__device__ __forceinline__ float sigmoidf(float x) { return 1.f / (1.f + expf(-x)); }
__device__ __forceinline__ float bits_of(float p) {  // clamp(-log(p+1e-5)/log(2), 0, 50)
    float b = -1.0f * logf(p + 1e-5f) / 0.6931471805599453f;
    return fminf(fmaxf(b, 0.f), 50.f);
}

// BitEstimator CDF (bitEstimator.py:20-42) with per-channel constants sp = softplus(h), ta = tanh(a)
struct ChanParams {
    float sp1, b1, ta1, sp2, b2, ta2, sp3, b3, ta3, sp4, b4;
};
__device__ __forceinline__ float factorized_cdf(float x, const ChanParams& c) {
    x = x * c.sp1 + c.b1;
    x = x + tanhf(x) * c.ta1;
    x = x * c.sp2 + c.b2;
    x = x + tanhf(x) * c.ta2;
    x = x * c.sp3 + c.b3;
    x = x + tanhf(x) * c.ta3;
    return sigmoidf(x * c.sp4 + c.b4);
}

__device__ __forceinline__ ChanParams make_chan_params(const FactorizedParams& prm, int c) {
    ChanParams p;
    p.sp1 = softplusf(prm.p[0][c]); p.b1 = prm.p[1][c]; p.ta1 = tanhf(prm.p[2][c]);
    p.sp2 = softplusf(prm.p[3][c]); p.b2 = prm.p[4][c]; p.ta2 = tanhf(prm.p[5][c]);
    p.sp3 = softplusf(prm.p[6][c]); p.b3 = prm.p[7][c]; p.ta3 = tanhf(prm.p[8][c]);
    p.sp4 = softplusf(prm.p[9][c]); p.b4 = prm.p[10][c];
    return p;
}

// Laplace(0, sigma).cdf(v) = 0.5 - 0.5*sign(v)*expm1(-|v|/sigma)   (torch.distributions.Laplace)
__device__ __forceinline__ float laplace_cdf(float v, float sigma) {
    float sgn = (v > 0.f) ? 1.f : ((v < 0.f) ? -1.f : 0.f);
    return 0.5f - 0.5f * sgn * expm1f(-fabsf(v) / sigma);
}

}  // namespace fvc
