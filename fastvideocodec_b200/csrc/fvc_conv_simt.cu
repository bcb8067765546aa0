// fp32 CUDA-core implicit-GEMM convolution over ACT records (sm_100a).
// Role: bring-up path and on-device checker for the tcgen05 engine (same inputs, same epilogue);
// it reconstructs activations as hi+lo and multiplies by the original fp32 weights.
#include "fvc_kernels.cuh"
#include "fvc_epilogue.cuh"

namespace fvc {

void make_conv_layer(ConvLayer& L, int Cin, int Cout, int k, int stride, int transposed) {
    L.Cin = Cin; L.Cout = Cout; L.k = k; L.stride = stride; L.transposed = transposed;
    int p = k / 2;
    if (!transposed) {
        L.nsub = 1; L.st = stride; L.os = 1;
        SubConv& S = L.sub[0];
        S.ntaps = 0; S.py = S.px = 0;
        for (int r = 0; r < k; ++r)
            for (int s = 0; s < k; ++s) {
                int t = S.ntaps++;
                S.dy[t] = (int8_t)(r - p); S.dx[t] = (int8_t)(s - p); S.r[t] = (int8_t)r; S.s[t] = (int8_t)s;
            }
    } else if (stride == 1) {
        // out(o) = sum_r in(o + p - r) * Wt[r]
        L.nsub = 1; L.st = 1; L.os = 1;
        SubConv& S = L.sub[0];
        S.ntaps = 0; S.py = S.px = 0;
        for (int r = 0; r < k; ++r)
            for (int s = 0; s < k; ++s) {
                int t = S.ntaps++;
                S.dy[t] = (int8_t)(p - r); S.dx[t] = (int8_t)(p - s); S.r[t] = (int8_t)r; S.s[t] = (int8_t)s;
            }
    } else {
        // stride-2 transposed conv, padding k//2, output_padding 1: out(2q+ph) = sum over taps r with
        // (ph + p - r) even of in(q + (ph+p-r)/2) * Wt[r]
        L.nsub = 4; L.st = 1; L.os = 2;
        for (int py = 0; py < 2; ++py)
            for (int px = 0; px < 2; ++px) {
                SubConv& S = L.sub[py * 2 + px];
                S.ntaps = 0; S.py = py; S.px = px;
                for (int r = 0; r < k; ++r) {
                    if ((py + p - r) & 1) continue;
                    for (int s = 0; s < k; ++s) {
                        if ((px + p - s) & 1) continue;
                        int t = S.ntaps++;
                        S.dy[t] = (int8_t)((py + p - r) / 2); S.dx[t] = (int8_t)((px + p - s) / 2);
                        S.r[t] = (int8_t)r; S.s[t] = (int8_t)s;
                    }
                }
            }
    }
}

// ----------------------------------------------------------------------------------------------
// weight packing: reference layout -> [sub][tap][CinP][CoutS] fp32 (zero padded)
// ----------------------------------------------------------------------------------------------
struct PackTaps {
    int ntaps[4];
    int8_t r[4][FVC_MAX_TAPS], s[4][FVC_MAX_TAPS];
};
__global__ void k_simt_pack(const float* __restrict__ w, float* __restrict__ out, PackTaps pt, int nsub, int Cin,
                            int Cout, int k, int transposed, int CinP, int CoutS, size_t sub_stride) {
    size_t n = (size_t)nsub * sub_stride;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int sub = (int)(i / sub_stride);
    size_t j = i % sub_stride;
    int co = (int)(j % CoutS);
    int ci = (int)((j / CoutS) % CinP);
    int t = (int)(j / ((size_t)CoutS * CinP));
    float v = 0.f;
    if (t < pt.ntaps[sub] && ci < Cin && co < Cout) {
        int r = pt.r[sub][t], s = pt.s[sub][t];
        v = transposed ? w[(((size_t)ci * Cout + co) * k + r) * k + s] : w[(((size_t)co * Cin + ci) * k + r) * k + s];
    }
    out[i] = v;
}

int simt_pack_weights(const ConvLayer& L, const float* w_ref, int CinP, int CoutS, SimtWeights* out,
                      cudaStream_t s) {
    int maxt = 0;
    PackTaps pt;
    for (int i = 0; i < L.nsub; ++i) {
        pt.ntaps[i] = L.sub[i].ntaps;
        maxt = std::max(maxt, L.sub[i].ntaps);
        for (int t = 0; t < L.sub[i].ntaps; ++t) {
            pt.r[i][t] = L.sub[i].r[t];
            pt.s[i][t] = L.sub[i].s[t];
        }
    }
    out->CinP = CinP;
    out->CoutS = CoutS;
    out->sub_stride = (size_t)maxt * CinP * CoutS;
    size_t n = out->sub_stride * L.nsub;
    if (!out->w) FVC_CUDA(cudaMalloc(&out->w, n * sizeof(float)));
    k_simt_pack<<<(unsigned)cdiv64((int64_t)n, 256), 256, 0, s>>>(w_ref, out->w, pt, L.nsub, L.Cin, L.Cout, L.k,
                                                                  L.transposed, CinP, CoutS, out->sub_stride);
    g_launch_count++;
    FVC_CHECK_LAUNCH();
    return 0;
}

// ----------------------------------------------------------------------------------------------
// the kernel: block 32x8 threads, q-tile 32x16, 2 pixels x 16 output channels per thread
// ----------------------------------------------------------------------------------------------
struct SimtParams {
    ActT in;
    const float* w;     // this launch's weights, [sub][tap][CinP][CoutS]
    size_t sub_stride;
    int CinP, CoutS, Cout;
    int nsub, ncog;
    int Hq, Wq, Hout, Wout, os;
    SubConv sub[4];
    Epilogue ep;
};

#define SIMT_CI 8
#define SIMT_CO 16
#define SIMT_TW 32
#define SIMT_TH 16

template <int ST>
__global__ void __launch_bounds__(256) k_conv_simt(const __grid_constant__ SimtParams P) {
    extern __shared__ float sm[];
    const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * 32 + tx;
    int z = blockIdx.z;
    const int cog = z % P.ncog; z /= P.ncog;
    const int sub = z % P.nsub;
    const int b = z / P.nsub;
    const SubConv& S = P.sub[sub];
    const int co0 = cog * SIMT_CO;
    const int qx0 = blockIdx.x * SIMT_TW, qy0 = blockIdx.y * SIMT_TH;

    int dymin = 127, dymax = -127, dxmin = 127, dxmax = -127;
    for (int t = 0; t < S.ntaps; ++t) {
        dymin = min(dymin, (int)S.dy[t]); dymax = max(dymax, (int)S.dy[t]);
        dxmin = min(dxmin, (int)S.dx[t]); dxmax = max(dxmax, (int)S.dx[t]);
    }
    const int PH = (SIMT_TH - 1) * ST + (dymax - dymin) + 1;
    const int PW = (SIMT_TW - 1) * ST + (dxmax - dxmin) + 1;
    float* patch = sm;                            // [SIMT_CI][PH][PW]
    float* wts = sm + SIMT_CI * PH * PW;          // [ntaps][SIMT_CI][SIMT_CO]

    float acc0[SIMT_CO], acc1[SIMT_CO];
#pragma unroll
    for (int j = 0; j < SIMT_CO; ++j) acc0[j] = acc1[j] = 0.f;

    const bool active_cog = co0 < P.Cout;
    if (active_cog) {
        const float* wsub = P.w + (size_t)sub * P.sub_stride;
        for (int c0 = 0; c0 < P.CinP; c0 += SIMT_CI) {
            // stage the input patch: one (pixel, 8-channel group) per thread iteration
            for (int i = tid; i < PH * PW; i += 256) {
                int py = i / PW, px = i - py * PW;
                int iy = qy0 * ST + dymin + py, ix = qx0 * ST + dxmin + px;
                float v[8];
                if (iy >= 0 && iy < P.in.H && ix >= 0 && ix < P.in.W) {
                    ep_load8(P.in.p + act_pixel_offset(P.in, b, iy, ix), P.in.Cp, c0, v);
                } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[j] = 0.f;
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) patch[(j * PH + py) * PW + px] = v[j];
            }
            for (int i = tid; i < S.ntaps * SIMT_CI * SIMT_CO; i += 256) {
                int j = i % SIMT_CO;
                int c = (i / SIMT_CO) % SIMT_CI;
                int t = i / (SIMT_CO * SIMT_CI);
                wts[i] = wsub[((size_t)t * P.CinP + c0 + c) * P.CoutS + co0 + j];
            }
            __syncthreads();
            for (int t = 0; t < S.ntaps; ++t) {
                const int oy = S.dy[t] - dymin, ox = S.dx[t] - dxmin;
                const float* p0 = patch + (ty * ST + oy) * PW + tx * ST + ox;
                const float* p1 = p0 + 8 * ST * PW;
                const float4* wt = reinterpret_cast<const float4*>(wts + t * SIMT_CI * SIMT_CO);
#pragma unroll
                for (int c = 0; c < SIMT_CI; ++c) {
                    float a0 = p0[c * PH * PW], a1 = p1[c * PH * PW];
#pragma unroll
                    for (int j4 = 0; j4 < SIMT_CO / 4; ++j4) {
                        float4 w4 = wt[c * (SIMT_CO / 4) + j4];
                        acc0[4 * j4 + 0] = fmaf(a0, w4.x, acc0[4 * j4 + 0]);
                        acc0[4 * j4 + 1] = fmaf(a0, w4.y, acc0[4 * j4 + 1]);
                        acc0[4 * j4 + 2] = fmaf(a0, w4.z, acc0[4 * j4 + 2]);
                        acc0[4 * j4 + 3] = fmaf(a0, w4.w, acc0[4 * j4 + 3]);
                        acc1[4 * j4 + 0] = fmaf(a1, w4.x, acc1[4 * j4 + 0]);
                        acc1[4 * j4 + 1] = fmaf(a1, w4.y, acc1[4 * j4 + 1]);
                        acc1[4 * j4 + 2] = fmaf(a1, w4.z, acc1[4 * j4 + 2]);
                        acc1[4 * j4 + 3] = fmaf(a1, w4.w, acc1[4 * j4 + 3]);
                    }
                }
            }
            __syncthreads();
        }
    }
    const int qx = qx0 + tx;
    if (qx < P.Wq) {
        int qy = qy0 + ty;
        if (qy < P.Hq)
            epilogue_apply<SIMT_CO>(P.ep, P.Cout, P.Hout, P.Wout, b, qy * P.os + S.py, qx * P.os + S.px, co0, acc0);
        qy += 8;
        if (qy < P.Hq)
            epilogue_apply<SIMT_CO>(P.ep, P.Cout, P.Hout, P.Wout, b, qy * P.os + S.py, qx * P.os + S.px, co0, acc1);
    }
}

int launch_conv_simt(const ConvLayer& L, const SimtWeights& W, ActT in, int Hout, int Wout, const Epilogue& ep,
                     cudaStream_t s) {
    FVC_ARG(W.w != nullptr && W.CinP == in.Cp);
    FVC_ARG(ep.gdn_beta == nullptr);  // fused GDN exists only in the tcgen05 engine
    SimtParams P;
    P.in = in;
    P.w = W.w;
    P.sub_stride = W.sub_stride;
    P.CinP = W.CinP; P.CoutS = W.CoutS; P.Cout = L.Cout;
    P.nsub = L.nsub;
    int chans = L.Cout;
    if (ep.out_act.p) chans = std::max(chans, ep.out_act.Cp);
    if (ep.out_act_relu.p) chans = std::max(chans, ep.out_act_relu.Cp);
    P.ncog = cdiv(chans, SIMT_CO);
    FVC_ARG(cdiv(L.Cout, SIMT_CO) * SIMT_CO <= W.CoutS);
    P.os = L.os;
    P.Hout = Hout; P.Wout = Wout;
    P.Hq = Hout / L.os; P.Wq = Wout / L.os;
    for (int i = 0; i < L.nsub; ++i) P.sub[i] = L.sub[i];
    P.ep = ep;
    // shared memory: worst case over sub-convolutions
    size_t smem = 0;
    for (int i = 0; i < L.nsub; ++i) {
        const SubConv& S = L.sub[i];
        int dymin = 127, dymax = -127, dxmin = 127, dxmax = -127;
        for (int t = 0; t < S.ntaps; ++t) {
            dymin = std::min(dymin, (int)S.dy[t]); dymax = std::max(dymax, (int)S.dy[t]);
            dxmin = std::min(dxmin, (int)S.dx[t]); dxmax = std::max(dxmax, (int)S.dx[t]);
        }
        int PH = (SIMT_TH - 1) * L.st + (dymax - dymin) + 1;
        int PW = (SIMT_TW - 1) * L.st + (dxmax - dxmin) + 1;
        size_t b = ((size_t)SIMT_CI * PH * PW + (size_t)S.ntaps * SIMT_CI * SIMT_CO) * sizeof(float);
        smem = std::max(smem, b);
    }
    dim3 grid(cdiv(P.Wq, SIMT_TW), cdiv(P.Hq, SIMT_TH), in.B * L.nsub * P.ncog);
    dim3 block(32, 8);
    if (L.st == 1) {
        FVC_CUDA(cudaFuncSetAttribute(k_conv_simt<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_conv_simt<1><<<grid, block, smem, s>>>(P);
    } else {
        FVC_CUDA(cudaFuncSetAttribute(k_conv_simt<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_conv_simt<2><<<grid, block, smem, s>>>(P);
    }
    g_launch_count++;
    FVC_CHECK_LAUNCH();
    return 0;
}

}  // namespace fvc
