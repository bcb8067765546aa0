// Memory-bound geometry / elementwise kernels of the DVC P-frame path (sm_100a).
// All are one-pass, coalesced, HBM-bound; op semantics follow SURVEY.md appendix A.
#include "fvc_kernels.cuh"

namespace fvc {

#define LAUNCH_1D(n, threads) dim3((unsigned)cdiv64((int64_t)(n), (threads))), dim3(threads)

// ----------------------------------------------------------------------------------------------
// bilinear helpers
// ----------------------------------------------------------------------------------------------
// F.interpolate 2x source index (ATen area_pixel_compute_source_index), in -> 2*in
__device__ __forceinline__ void up2_index(int dst, int in, int align_corners, int& i0, int& i1, float& l1) {
    float src;
    if (align_corners) {
        float scale = (2 * in > 1) ? (float)(in - 1) / (float)(2 * in - 1) : 0.f;
        src = scale * (float)dst;
    } else {
        src = 0.5f * ((float)dst + 0.5f) - 0.5f;
        if (src < 0.f) src = 0.f;
    }
    i0 = (int)src;
    if (i0 > in - 1) i0 = in - 1;
    i1 = i0 + ((i0 < in - 1) ? 1 : 0);
    l1 = src - (float)i0;
}

// torch.linspace(-1, 1, n)[i] in fp32 (symmetric evaluation as ATen does)
__device__ __forceinline__ float linspace_m1_1(int i, int n) {
    float step = 2.0f / (float)(n - 1);
    return (i < n / 2) ? (-1.0f + step * (float)i) : (1.0f - step * (float)(n - 1 - i));
}

// grid_sample(bilinear, border, align_corners=False) source coordinates and taps for the
// reference's grid  g = linspace(-1,1,n)[i] + flow / ((n-1)/2)   (endecoder.py:52-67)
struct WarpTaps {
    int x0, y0;
    float wnw, wne, wsw, wse;
    bool vx1, vy1;  // whether the +1 taps are inside the image
};
__device__ __forceinline__ WarpTaps warp_taps(int x, int y, float fx, float fy, int W, int H) {
    float gx = linspace_m1_1(x, W) + fx / (((float)W - 1.0f) / 2.0f);
    float gy = linspace_m1_1(y, H) + fy / (((float)H - 1.0f) / 2.0f);
    float ix = ((gx + 1.f) * (float)W - 1.f) / 2.f;
    float iy = ((gy + 1.f) * (float)H - 1.f) / 2.f;
    ix = fminf((float)(W - 1), fmaxf(ix, 0.f));
    iy = fminf((float)(H - 1), fmaxf(iy, 0.f));
    float fx0 = floorf(ix), fy0 = floorf(iy);
    WarpTaps t;
    t.x0 = (int)fx0;
    t.y0 = (int)fy0;
    float x1 = fx0 + 1.f, y1 = fy0 + 1.f;
    t.wnw = (x1 - ix) * (y1 - iy);
    t.wne = (ix - fx0) * (y1 - iy);
    t.wsw = (x1 - ix) * (iy - fy0);
    t.wse = (ix - fx0) * (iy - fy0);
    t.vx1 = (t.x0 + 1) < W;
    t.vy1 = (t.y0 + 1) < H;
    return t;
}
__device__ __forceinline__ float warp_sample(const float* __restrict__ plane, int W, const WarpTaps& t) {
    const float* r0 = plane + (size_t)t.y0 * W + t.x0;
    float v = r0[0] * t.wnw;
    if (t.vx1) v += r0[1] * t.wne;
    if (t.vy1) {
        v += r0[W] * t.wsw;
        if (t.vx1) v += r0[W + 1] * t.wse;
    }
    return v;
}

// ----------------------------------------------------------------------------------------------
// transforms.ToTensor() of the reference's frame ingest (dataset.py:75; models.py:425) on the device:
// uint8 HWC [n,H,W,3] -> fp32 CHW [n,3,H,W], x / 255 (IEEE division: bit-identical to torchvision's .div(255)).
// A thread converts 4 pixels = 12 bytes (three 32-bit loads) and writes one float4 per plane; the tail is scalar.
// ----------------------------------------------------------------------------------------------
__global__ void k_u8hwc_to_f32chw(const uint8_t* __restrict__ src, float* __restrict__ dst, int n, int64_t hw) {
    pdl_sync();
    const int64_t quads = hw >> 2;
    const int img = blockIdx.y;
    const uint8_t* s = src + (size_t)img * hw * 3;
    float* d = dst + (size_t)img * hw * 3;
    const bool aligned = ((reinterpret_cast<uintptr_t>(s) & 3) == 0) && ((reinterpret_cast<uintptr_t>(d) & 15) == 0) && ((hw & 3) == 0);
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < quads; q += (int64_t)gridDim.x * blockDim.x) {
        uint8_t b[12];
        if (aligned) {
            const uint32_t* w = reinterpret_cast<const uint32_t*>(s + q * 12);
            const uint32_t w0 = __ldg(w), w1 = __ldg(w + 1), w2 = __ldg(w + 2);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                b[k] = (uint8_t)(w0 >> (8 * k));
                b[4 + k] = (uint8_t)(w1 >> (8 * k));
                b[8 + k] = (uint8_t)(w2 >> (8 * k));
            }
        } else {
#pragma unroll
            for (int k = 0; k < 12; ++k) b[k] = s[q * 12 + k];
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float4 v = make_float4(b[c] / 255.0f, b[3 + c] / 255.0f, b[6 + c] / 255.0f, b[9 + c] / 255.0f);
            if (aligned) *reinterpret_cast<float4*>(d + (size_t)c * hw + q * 4) = v;
            else { float* o = d + (size_t)c * hw + q * 4; o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w; }
        }
    }
    if (blockIdx.x == 0 && threadIdx.x < (hw & 3)) {          // up to 3 trailing pixels
        const int64_t px = quads * 4 + threadIdx.x;
#pragma unroll
        for (int c = 0; c < 3; ++c) d[(size_t)c * hw + px] = s[px * 3 + c] / 255.0f;
    }
}
int launch_u8hwc_to_f32chw(const uint8_t* src, float* dst, int n, int H, int W, cudaStream_t s) {
    FVC_ARG(n >= 0 && n <= 65535 && H >= 0 && W >= 0);
    const int64_t hw = (int64_t)H * W;
    if (n == 0 || hw == 0) return 0;
    dim3 grid((unsigned)std::min<int64_t>(std::max<int64_t>(cdiv64(hw >> 2, 256), 1), 148 * 8), (unsigned)n);
    FVC_CUDA(launch_pdl(k_u8hwc_to_f32chw, grid, 256, 0, s, src, dst, n, hw));
    g_launch_count++;
    FVC_CHECK_LAUNCH();
    return 0;
}

// ----------------------------------------------------------------------------------------------
// avg_pool2d(2,2) on planar fp32
// ----------------------------------------------------------------------------------------------
__global__ void k_avg_pool2_planar(const float* __restrict__ x, float* __restrict__ y, int planes, int H, int W) {
    pdl_sync();
    int Wo = W >> 1, Ho = H >> 1;
    int64_t n = (int64_t)planes * Ho * Wo;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int xo = (int)(i % Wo);
    int yo = (int)((i / Wo) % Ho);
    int p = (int)(i / ((int64_t)Wo * Ho));
    const float* r = x + ((size_t)p * H + 2 * yo) * W + 2 * xo;
    float2 a = *reinterpret_cast<const float2*>(r);
    float2 b = *reinterpret_cast<const float2*>(r + W);
    y[i] = (((a.x + a.y) + b.x) + b.y) / 4.0f;
}
int launch_avg_pool2_planar(const float* x, float* y, int planes, int H, int W, cudaStream_t s) {
    int64_t n = (int64_t)planes * (H / 2) * (W / 2);
    if (n == 0) return 0;
    FVC_CUDA(launch_pdl(k_avg_pool2_planar, LAUNCH_1D(n, 256), 0, s, x, y, planes, H, W));
    g_launch_count++;
    FVC_CHECK_LAUNCH();
    return 0;
}

// ----------------------------------------------------------------------------------------------
// 2x bilinear up-sampling on planar fp32 (op-level entry)
// ----------------------------------------------------------------------------------------------
__global__ void k_upsample2x_planar(const float* __restrict__ x, float* __restrict__ y, int planes, int H, int W,
                                    int ac, float scale) {
    int Wo = 2 * W, Ho = 2 * H;
    int64_t n = (int64_t)planes * Ho * Wo;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int xo = (int)(i % Wo);
    int yo = (int)((i / Wo) % Ho);
    int p = (int)(i / ((int64_t)Wo * Ho));
    int x0, x1, y0, y1;
    float lx, ly;
    up2_index(xo, W, ac, x0, x1, lx);
    up2_index(yo, H, ac, y0, y1, ly);
    const float* pl = x + (size_t)p * H * W;
    float v = (1.f - ly) * ((1.f - lx) * pl[y0 * W + x0] + lx * pl[y0 * W + x1]) +
              ly * ((1.f - lx) * pl[y1 * W + x0] + lx * pl[y1 * W + x1]);
    y[i] = v * scale;
}
int launch_upsample2x_planar(const float* x, float* y, int planes, int H, int W, int ac, float scale,
                             cudaStream_t s) {
    int64_t n = (int64_t)planes * H * W * 4;
    if (n == 0) return 0;
    k_upsample2x_planar<<<LAUNCH_1D(n, 256), 0, s>>>(x, y, planes, H, W, ac, scale);
    g_launch_count++;
    FVC_CHECK_LAUNCH();
    return 0;
}

// ----------------------------------------------------------------------------------------------
// flow_warp on NCHW fp32 (op-level entry)
// ----------------------------------------------------------------------------------------------
__global__ void k_flow_warp_nchw(const float* __restrict__ img, const float* __restrict__ flow,
                                 float* __restrict__ out, int B, int C, int H, int W) {
    int64_t n = (int64_t)B * H * W;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int x = (int)(i % W);
    int y = (int)((i / W) % H);
    int b = (int)(i / ((int64_t)W * H));
    size_t hw = (size_t)H * W;
    float fx = flow[((size_t)b * 2 + 0) * hw + (size_t)y * W + x];
    float fy = flow[((size_t)b * 2 + 1) * hw + (size_t)y * W + x];
    WarpTaps t = warp_taps(x, y, fx, fy, W, H);
    for (int c = 0; c < C; ++c)
        out[((size_t)b * C + c) * hw + (size_t)y * W + x] = warp_sample(img + ((size_t)b * C + c) * hw, W, t);
}
int launch_flow_warp_nchw(const float* img, const float* flow, float* out, int B, int C, int H, int W,
                          cudaStream_t s) {
    int64_t n = (int64_t)B * H * W;
    if (n == 0) return 0;
    k_flow_warp_nchw<<<LAUNCH_1D(n, 256), 0, s>>>(img, flow, out, B, C, H, W);
    g_launch_count++;
    FVC_CHECK_LAUNCH();
    return 0;
}

// ----------------------------------------------------------------------------------------------
// record writers: one thread writes a whole Cp=32 record whose first `nv` channels are given
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void store_record32(e16* rec, const float* v, int nv) {
    // rec: 64 x 16 bit = 128 B: [hi 32][lo 32]
    uint32_t hi[16], lo[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        e16 h0 = 0, l0 = 0, h1 = 0, l1 = 0;
        if (2 * j < nv) split16(v[2 * j], h0, l0);
        if (2 * j + 1 < nv) split16(v[2 * j + 1], h1, l1);
        hi[j] = pack16x2(h0, h1);
        lo[j] = pack16x2(l0, l1);
    }
    uint4* d = reinterpret_cast<uint4*>(rec);
#pragma unroll
    for (int j = 0; j < 4; ++j) d[j] = make_uint4(hi[4 * j], hi[4 * j + 1], hi[4 * j + 2], hi[4 * j + 3]);
#pragma unroll
    for (int j = 0; j < 4; ++j) d[4 + j] = make_uint4(lo[4 * j], lo[4 * j + 1], lo[4 * j + 2], lo[4 * j + 3]);
}

// narrow record (Cp = 8): [hi 8][lo 8] = 32 B
__device__ __forceinline__ void store_record8(e16* rec, const float* v, int nv) {
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        e16 h0 = 0, l0 = 0, h1 = 0, l1 = 0;
        if (2 * j < nv) split16(v[2 * j], h0, l0);
        if (2 * j + 1 < nv) split16(v[2 * j + 1], h1, l1);
        hi[j] = pack16x2(h0, h1);
        lo[j] = pack16x2(l0, l1);
    }
    uint4* d = reinterpret_cast<uint4*>(rec);
    d[0] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    d[1] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}
__device__ __forceinline__ void store_record(const ActT& t, e16* rec, const float* v, int nv) {
    if (t.Cp == 8) store_record8(rec, v, nv);
    else store_record32(rec, v, nv);
}

// ----------------------------------------------------------------------------------------------
// SpyNet level preparation (endecoder.py:352-354): up = 2*up2(flow_prev); X = [im1, warp(im2,up), up]
// ----------------------------------------------------------------------------------------------
__global__ void k_spynet_prep(const float* __restrict__ im1, const float* __restrict__ im2,
                              const float* __restrict__ flow_prev, ActT X, float* __restrict__ flow_up) {
    pdl_sync();
    const int H = X.H, W = X.W;
    // 2-D grid: x from blockIdx.x, (b, y) from blockIdx.y (no 64-bit divisions per thread)
    const int x = (int)(blockIdx.x * blockDim.x + threadIdx.x);
    const int y = (int)blockIdx.y % H, b = (int)blockIdx.y / H;
    if (x >= W) return;
    const int64_t i = ((int64_t)b * H + y) * W + x;
    float ux = 0.f, uy = 0.f;
    if (flow_prev) {
        int h = H >> 1, w = W >> 1;
        int x0, x1, y0, y1;
        float lx, ly;
        up2_index(x, w, 0, x0, x1, lx);
        up2_index(y, h, 0, y0, y1, ly);
        const float2* fp = reinterpret_cast<const float2*>(flow_prev) + (size_t)b * h * w;
        float2 v00 = fp[y0 * w + x0], v01 = fp[y0 * w + x1], v10 = fp[y1 * w + x0], v11 = fp[y1 * w + x1];
        ux = ((1.f - ly) * ((1.f - lx) * v00.x + lx * v01.x) + ly * ((1.f - lx) * v10.x + lx * v11.x)) * 2.0f;
        uy = ((1.f - ly) * ((1.f - lx) * v00.y + lx * v01.y) + ly * ((1.f - lx) * v10.y + lx * v11.y)) * 2.0f;
    }
    WarpTaps t = warp_taps(x, y, ux, uy, W, H);
    size_t hw = (size_t)H * W;
    size_t po = (size_t)y * W + x;
    float v[8];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        v[c] = im1[((size_t)b * 3 + c) * hw + po];
        v[3 + c] = warp_sample(im2 + ((size_t)b * 3 + c) * hw, W, t);
    }
    v[6] = ux;
    v[7] = uy;
    store_record(X, X.p + act_pixel_offset(X, b, y, x), v, 8);
    reinterpret_cast<float2*>(flow_up)[i] = make_float2(ux, uy);
}
int launch_spynet_prep(const float* im1, const float* im2, const float* flow_prev, ActT X, float* flow_up,
                       cudaStream_t s) {
    FVC_ARG(X.Cp == 32 || X.Cp == 8);
    FVC_ARG((int64_t)X.B * X.H <= 65535);
    FVC_CUDA(launch_pdl(k_spynet_prep, dim3((unsigned)cdiv(X.W, 128), (unsigned)(X.B * X.H)), dim3(128), 0, s, im1, im2,
                        flow_prev, X, flow_up));
    g_launch_count++;
    FVC_CHECK_LAUNCH();
    return 0;
}

// ----------------------------------------------------------------------------------------------
// motion compensation input (net.py:64-66): warpframe = flow_warp(ref, mv); X = cat(warpframe, ref)
// ----------------------------------------------------------------------------------------------
__global__ void k_mc_prep(const float* __restrict__ ref, const float* __restrict__ mv, float* __restrict__ warpframe,
                          ActT X) {
    pdl_sync();
    const int H = X.H, W = X.W;
    // 2-D grid: x from blockIdx.x, (b, y) from blockIdx.y (no 64-bit divisions per thread)
    const int x = (int)(blockIdx.x * blockDim.x + threadIdx.x);
    const int y = (int)blockIdx.y % H, b = (int)blockIdx.y / H;
    if (x >= W) return;
    const int64_t i = ((int64_t)b * H + y) * W + x;
    float2 f = reinterpret_cast<const float2*>(mv)[i];
    WarpTaps t = warp_taps(x, y, f.x, f.y, W, H);
    size_t hw = (size_t)H * W;
    size_t po = (size_t)y * W + x;
    float v[6];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float* pl = ref + ((size_t)b * 3 + c) * hw;
        v[c] = warp_sample(pl, W, t);
        v[3 + c] = pl[po];
        warpframe[((size_t)b * 3 + c) * hw + po] = v[c];
    }
    store_record(X, X.p + act_pixel_offset(X, b, y, x), v, 6);
}
int launch_mc_prep(const float* ref, const float* mv, float* warpframe, ActT X, cudaStream_t s) {
    FVC_ARG(X.Cp == 32 || X.Cp == 8);
    FVC_ARG((int64_t)X.B * X.H <= 65535);
    FVC_CUDA(launch_pdl(k_mc_prep, dim3((unsigned)cdiv(X.W, 128), (unsigned)(X.B * X.H)), dim3(128), 0, s, ref, mv,
                        warpframe, X));
    g_launch_count++;
    FVC_CHECK_LAUNCH();
    return 0;
}

// prediction = warpnet(x) + warpframe (net.py:67); residual = input - prediction (net.py:81)
__global__ void k_mc_finish(const float* __restrict__ res, const float* __restrict__ warpframe,
                            const float* __restrict__ cur, float* __restrict__ pred, ActT R) {
    pdl_sync();
    const int H = R.H, W = R.W;
    // 2-D grid: x from blockIdx.x, (b, y) from blockIdx.y (no 64-bit divisions per thread)
    const int x = (int)(blockIdx.x * blockDim.x + threadIdx.x);
    const int y = (int)blockIdx.y % H, b = (int)blockIdx.y / H;
    if (x >= W) return;
    const int64_t i = ((int64_t)b * H + y) * W + x;
    size_t hw = (size_t)H * W;
    size_t po = (size_t)y * W + x;
    float v[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        size_t o = ((size_t)b * 3 + c) * hw + po;
        float p = res[i * 3 + c] + warpframe[o];
        pred[o] = p;
        v[c] = cur[o] - p;
    }
    store_record(R, R.p + act_pixel_offset(R, b, y, x), v, 3);
}
int launch_mc_finish(const float* res, const float* warpframe, const float* cur, float* pred, ActT R,
                     cudaStream_t s) {
    FVC_ARG(R.Cp == 32 || R.Cp == 8);
    FVC_ARG((int64_t)R.B * R.H <= 65535);
    FVC_CUDA(launch_pdl(k_mc_finish, dim3((unsigned)cdiv(R.W, 128), (unsigned)(R.B * R.H)), dim3(128), 0, s, res,
                        warpframe, cur, pred, R));
    g_launch_count++;
    FVC_CHECK_LAUNCH();
    return 0;
}

// ----------------------------------------------------------------------------------------------
// ACT helpers working on 8-channel groups (16-byte vectors of hi and of lo)
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void load8(const e16* rec, int Cp, int c8, float* v) {
    uint4 h = *reinterpret_cast<const uint4*>(rec + c8 * 8);
    uint4 l = *reinterpret_cast<const uint4*>(rec + Cp + c8 * 8);
    const uint32_t hh[4] = {h.x, h.y, h.z, h.w}, ll[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float a, b, c, d;
        e2f2(hh[j], a, b);
        e2f2(ll[j], c, d);
        v[2 * j] = a + c;
        v[2 * j + 1] = b + d;
    }
}
__device__ __forceinline__ void store8(e16* rec, int Cp, int c8, const float* v) {
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        e16 h0, l0, h1, l1;
        split16(v[2 * j], h0, l0);
        split16(v[2 * j + 1], h1, l1);
        hi[j] = pack16x2(h0, h1);
        lo[j] = pack16x2(l0, l1);
    }
    *reinterpret_cast<uint4*>(rec + c8 * 8) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(rec + Cp + c8 * 8) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

// AvgPool2d(2,2) on ACT (Warp_net c0_p / c1_p, endecoder.py:286-288) + optional relu copy
__global__ void k_pool_act(ActT in, ActT out, ActT out_relu) {
    pdl_sync();
    int G = out.Cp >> 3;
    int64_t n = (int64_t)out.B * out.H * out.W * G;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int g = (int)(i % G);
    int64_t pix = i / G;
    int x = (int)(pix % out.W);
    int y = (int)((pix / out.W) % out.H);
    int b = (int)(pix / ((int64_t)out.W * out.H));
    float a[8], bb[8], c[8], d[8], r[8];
    load8(in.p + act_pixel_offset(in, b, 2 * y, 2 * x), in.Cp, g, a);
    load8(in.p + act_pixel_offset(in, b, 2 * y, 2 * x + 1), in.Cp, g, bb);
    load8(in.p + act_pixel_offset(in, b, 2 * y + 1, 2 * x), in.Cp, g, c);
    load8(in.p + act_pixel_offset(in, b, 2 * y + 1, 2 * x + 1), in.Cp, g, d);
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = (((a[j] + bb[j]) + c[j]) + d[j]) / 4.0f;
    store8(out.p + act_pixel_offset(out, b, y, x), out.Cp, g, r);
    if (out_relu.p) {
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = fmaxf(r[j], 0.f);
        store8(out_relu.p + act_pixel_offset(out_relu, b, y, x), out_relu.Cp, g, r);
    }
}
int launch_pool_act(ActT in, ActT out, ActT out_relu, cudaStream_t s) {
    FVC_ARG(in.Cp == out.Cp && in.H == 2 * out.H && in.W == 2 * out.W);
    int64_t n = (int64_t)out.B * out.H * out.W * (out.Cp / 8);
    FVC_CUDA(launch_pdl(k_pool_act, LAUNCH_1D(n, 256), 0, s, in, out, out_relu));
    g_launch_count++;
    FVC_CHECK_LAUNCH();
    return 0;
}

// out = skip + bilinearupsacling2(low) (align_corners=True; endecoder.py:291, 293) + optional relu copy.
// Block = 32 output pixels of one image row x 8 channel groups (16-byte hi + 16-byte lo vectors): 8 consecutive
// lanes cover the 128-byte hi (and lo) half of a 64-channel record, so every warp access is four full lines.  Row
// and batch come from blockIdx (no 64-bit divisions: the first version spent more issue slots on index arithmetic
// than on the 14 memory instructions and reached 53 % of the HBM rate).
__global__ void __launch_bounds__(256) k_upadd_act(ActT low, ActT skip, ActT out, ActT out_relu) {
    pdl_sync();
    const int G = out.Cp >> 3;                       // 8-channel groups per record (8 for the 64-channel Warp_net)
    const int ppb = 256 / G;                         // pixels per block
    const int g = (int)threadIdx.x % G;
    const int x = (int)blockIdx.x * ppb + (int)threadIdx.x / G;
    const int y = (int)blockIdx.y % out.H;
    const int b = (int)blockIdx.y / out.H;
    if (x >= out.W) return;
    int x0, x1, y0, y1;
    float lx, ly;
    up2_index(x, low.W, 1, x0, x1, lx);
    up2_index(y, low.H, 1, y0, y1, ly);
    float v00[8], v01[8], v10[8], v11[8], sk[8], r[8];
    const e16* lrow0 = low.p + act_pixel_offset(low, b, y0, 0);
    const e16* lrow1 = low.p + act_pixel_offset(low, b, y1, 0);
    const size_t lrec = (size_t)2 * low.Cp;
    const size_t opix = act_pixel_offset(out, b, y, x);   // skip / out / out_relu share the geometry (checked on the host)
    load8(lrow0 + x0 * lrec, low.Cp, g, v00);
    load8(lrow0 + x1 * lrec, low.Cp, g, v01);
    load8(lrow1 + x0 * lrec, low.Cp, g, v10);
    load8(lrow1 + x1 * lrec, low.Cp, g, v11);
    load8(skip.p + opix, skip.Cp, g, sk);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        float up = (1.f - ly) * ((1.f - lx) * v00[j] + lx * v01[j]) + ly * ((1.f - lx) * v10[j] + lx * v11[j]);
        r[j] = sk[j] + up;
    }
    store8(out.p + opix, out.Cp, g, r);
    if (out_relu.p) {
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = fmaxf(r[j], 0.f);
        store8(out_relu.p + opix, out_relu.Cp, g, r);
    }
}
int launch_upadd_act(ActT low, ActT skip, ActT out, ActT out_relu, cudaStream_t s) {
    FVC_ARG(low.Cp == out.Cp && skip.Cp == out.Cp && out.H == 2 * low.H && out.W == 2 * low.W);
    FVC_ARG(!out.parity && !skip.parity && !low.parity && (!out_relu.p || (!out_relu.parity && out_relu.Cp == out.Cp)));
    FVC_ARG(skip.H == out.H && skip.W == out.W && (256 % (out.Cp / 8)) == 0 && (int64_t)out.B * out.H <= 65535);
    const int ppb = 256 / (out.Cp / 8);
    FVC_CUDA(launch_pdl(k_upadd_act, dim3((unsigned)cdiv(out.W, ppb), (unsigned)(out.B * out.H)), dim3(256), 0, s, low,
                        skip, out, out_relu));
    g_launch_count++;
    FVC_CHECK_LAUNCH();
    return 0;
}

// ----------------------------------------------------------------------------------------------
// GDN / IGDN (GDN.py:63-93) as a standalone kernel: one thread per pixel, C <= 64.
// ----------------------------------------------------------------------------------------------
__global__ void k_gdn_reparam(const float* __restrict__ beta, const float* __restrict__ gamma,
                              float* __restrict__ beta_eff, float* __restrict__ gamma_eff, int C) {
    const float ped = 1.4551915228366852e-11f;           // (2^-18)^2
    const float beta_bound = (float)1.0000072759311445e-03;  // sqrt(1e-6 + 2^-36)
    const float gamma_bound = 3.814697265625e-06f;        // 2^-18
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < C) {
        float b = fmaxf(beta[i], beta_bound);
        beta_eff[i] = b * b - ped;
    }
    if (i < C * C) {
        float g = fmaxf(gamma[i], gamma_bound);
        gamma_eff[i] = g * g - ped;
    }
}
int launch_gdn_reparam(const float* beta, const float* gamma, float* beta_eff, float* gamma_eff, int C,
                       cudaStream_t s) {
    k_gdn_reparam<<<LAUNCH_1D(C * C, 256), 0, s>>>(beta, gamma, beta_eff, gamma_eff, C);
    g_launch_count++;
    FVC_CHECK_LAUNCH();
    return 0;
}

__global__ void k_gdn_act(ActT in, int C, const float* __restrict__ beta, const float* __restrict__ gamma,
                          int inverse, ActT out) {
    extern __shared__ float sm[];  // gamma [C][C] then beta [C]
    for (int i = threadIdx.x; i < C * C; i += blockDim.x) sm[i] = gamma[i];
    for (int i = threadIdx.x; i < C; i += blockDim.x) sm[C * C + i] = beta[i];
    __syncthreads();
    int64_t n = (int64_t)in.B * in.H * in.W;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int x = (int)(i % in.W);
    int y = (int)((i / in.W) % in.H);
    int b = (int)(i / ((int64_t)in.W * in.H));
    const e16* rec = in.p + act_pixel_offset(in, b, y, x);
    e16* orec = out.p + act_pixel_offset(out, b, y, x);
    float v[64], sq[64];
#pragma unroll
    for (int g = 0; g < 8; ++g) {
        if (g * 8 < C) load8(rec, in.Cp, g, v + g * 8);
    }
#pragma unroll
    for (int c = 0; c < 64; ++c) sq[c] = v[c] * v[c];
    for (int g = 0; g < (out.Cp >> 3); ++g) {
        float r[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            int co = g * 8 + j;
            float val = 0.f;
            if (co < C) {
                float acc = 0.f;
                const float* gr = sm + co * C;
#pragma unroll
                for (int c = 0; c < 64; ++c)
                    if (c < C) acc = fmaf(gr[c], sq[c], acc);
                float nrm = sqrtf(acc + sm[C * C + co]);
                // v[co] with a runtime index would spill; recompute from the record instead
                float xv = e2f(rec[co]) + e2f(rec[in.Cp + co]);
                val = inverse ? xv * nrm : xv / nrm;
            }
            r[j] = val;
        }
        store8(orec, out.Cp, g, r);
    }
}
int launch_gdn_act(ActT in, int C, const float* beta_eff, const float* gamma_eff, int inverse, ActT out,
                   cudaStream_t s) {
    FVC_ARG(C <= 64 && C % 8 == 0 && in.Cp >= C && out.Cp >= C);
    int64_t n = (int64_t)in.B * in.H * in.W;
    size_t smem = (size_t)(C * C + C) * sizeof(float);
    k_gdn_act<<<LAUNCH_1D(n, 128), smem, s>>>(in, C, beta_eff, gamma_eff, inverse, out);
    g_launch_count++;
    FVC_CHECK_LAUNCH();
    return 0;
}

// ----------------------------------------------------------------------------------------------
// reconstruction + the three distortion terms (net.py:103-116), fused, with block partial sums
// ----------------------------------------------------------------------------------------------
// clip_mse = 0: first sum over the unclipped reconstruction (DVC, net.py:103-109); 1: over the clipped one
// (LSVC.forward, models.py:1383, 1400)
__global__ void k_recon_losses(const float* __restrict__ cur, const float* __restrict__ pred,
                               const float* __restrict__ warp, const float* __restrict__ res, int res_nhwc3, int B,
                               int HW, float* __restrict__ clipped, float* __restrict__ partials, int clip_mse) {
    pdl_sync();
    __shared__ float red[32];
    // one pixel (3 channels) per thread and step; (b, pixel) from a 32-bit index (B * HW < 2^31 checked on the host)
    const int npx = B * HW;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f;
    for (int i = (int)(blockIdx.x * blockDim.x + threadIdx.x); i < npx; i += (int)(gridDim.x * blockDim.x)) {
        const int b = i / HW, p = i - b * HW;
        const size_t plane0 = (size_t)b * 3 * HW + p;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const size_t o = plane0 + (size_t)c * HW;
            const float r = res_nhwc3 ? res[(size_t)i * 3 + c] : res[o];
            const float c0 = cur[o], p0 = pred[o], w0 = warp[o];
            const float rec = p0 + r;
            const float cl = fminf(fmaxf(rec, 0.f), 1.f);
            clipped[o] = cl;
            const float d0 = (clip_mse ? cl : rec) - c0, d1 = w0 - c0, d2 = p0 - c0;
            s0 = fmaf(d0, d0, s0);
            s1 = fmaf(d1, d1, s1);
            s2 = fmaf(d2, d2, s2);
        }
    }
    s0 = block_sum(s0, red);
    s1 = block_sum(s1, red);
    s2 = block_sum(s2, red);
    if (threadIdx.x == 0) {
        partials[blockIdx.x * 3 + 0] = s0;
        partials[blockIdx.x * 3 + 1] = s1;
        partials[blockIdx.x * 3 + 2] = s2;
    }
}
int launch_recon_losses(const float* cur, const float* pred, const float* warp, const float* res, int res_nhwc3,
                        int B, int HW, float* clipped, float* partials, int* nblocks_out, cudaStream_t s,
                        int clip_mse) {
    FVC_ARG((int64_t)B * HW < (1ll << 31));
    int64_t n = (int64_t)B * HW;
    int blocks = (int)std::min<int64_t>(cdiv64(n, 256 * 2), 148 * 8);
    if (blocks < 1) blocks = 1;
    FVC_CUDA(launch_pdl(k_recon_losses, blocks, 256, 0, s, cur, pred, warp, res, res_nhwc3, B, HW, clipped, partials, clip_mse));
    g_launch_count++;
    FVC_CHECK_LAUNCH();
    *nblocks_out = blocks;
    return 0;
}

// partials: [n][groups] -> out[g] = scale * sum_n (double accumulation, fixed order: deterministic)
__global__ void k_reduce_partials(const float* __restrict__ partials, int n, int groups, double scale,
                                  float* __restrict__ out) {
    pdl_sync();
    int g = blockIdx.x;
    __shared__ double sh[256];
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) acc += (double)partials[(size_t)i * groups + g];
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[g] = (float)(sh[0] * scale);
}
int launch_reduce_partials(const float* partials, int n, int groups, double scale, float* out, cudaStream_t s) {
    FVC_CUDA(launch_pdl(k_reduce_partials, groups, 256, 0, s, partials, n, groups, scale, out));
    g_launch_count++;
    FVC_CHECK_LAUNCH();
    return 0;
}

// ----------------------------------------------------------------------------------------------
// fused 3x3 tail convolutions (mvDecoder.deconv8, Warp_net conv6): weight regrouping and the tap sum
// ----------------------------------------------------------------------------------------------
__global__ void k_tapsplit_weights(const float* __restrict__ w, float* __restrict__ o, int Cin, int Cout, int rows) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * Cin) return;
    const int ci = i % Cin, q = i / Cin;          // q = tap * Cout + co
    float v = 0.f;
    if (q < 9 * Cout) {
        const int tap = q / Cout, co = q - tap * Cout;
        v = w[((size_t)co * Cin + ci) * 9 + tap];  // OIHW 3x3: tap = r * 3 + s
    }
    o[i] = v;
}
int launch_tapsplit_weights(const float* w3x3, float* w1x1, int Cin, int Cout, int rows, cudaStream_t s) {
    k_tapsplit_weights<<<cdiv(rows * Cin, 256), 256, 0, s>>>(w3x3, w1x1, Cin, Cout, rows);
    g_launch_count++;
    FVC_CHECK_LAUNCH();
    return 0;
}
// one thread per output pixel; a block covers 128 pixels of one image row, so the 3 x 130 partial-product records it
// gathers from are read once from HBM and served from L1 afterwards (each record is used by 9 output pixels)
template <int CO>
__global__ void __launch_bounds__(128) k_tapsum(const float* __restrict__ P, const float* __restrict__ bias,
                                                float* __restrict__ out, int H, int W, int cq) {
    pdl_sync();
    const int x = blockIdx.x * 128 + threadIdx.x, y = blockIdx.y % H, b = blockIdx.y / H;
    if (x >= W) return;
    float acc[CO];
#pragma unroll
    for (int c = 0; c < CO; ++c) acc[c] = 0.f;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const int yy = y + r - 1;
        if (yy < 0 || yy >= H) continue;
#pragma unroll
        for (int t = 0; t < 3; ++t) {
            const int xx = x + t - 1;
            if (xx < 0 || xx >= W) continue;
            const float* p = P + (((size_t)b * H + yy) * W + xx) * (size_t)cq + (r * 3 + t) * CO;
#pragma unroll
            for (int c = 0; c < CO; ++c) acc[c] += p[c];     // fixed order (r, s): deterministic
        }
    }
    float* o = out + (((size_t)b * H + y) * W + x) * CO;
#pragma unroll
    for (int c = 0; c < CO; ++c) o[c] = acc[c] + bias[c];
}
int launch_tapsum(const float* P, const float* bias, float* out, int B, int H, int W, int Cout, int cq, cudaStream_t s) {
    FVC_ARG((Cout == 2 || Cout == 3) && 9 * Cout <= cq && (int64_t)B * H <= 65535);
    dim3 grid((unsigned)cdiv(W, 128), (unsigned)(B * H));
    if (Cout == 2) FVC_CUDA(launch_pdl(k_tapsum<2>, grid, dim3(128), 0, s, P, bias, out, H, W, cq));
    else FVC_CUDA(launch_pdl(k_tapsum<3>, grid, dim3(128), 0, s, P, bias, out, H, W, cq));
    g_launch_count++;
    FVC_CHECK_LAUNCH();
    return 0;
}

// sums6 = mse, warploss, interloss, bits_feature, bits_z, bits_mv ->
// scalars7 = mse, warploss, interloss, bpp_feature, bpp_z, bpp_mv, bpp   (net.py:212-220)
__global__ void k_finalize_scalars(const float* __restrict__ sums6, float n_pix, float* __restrict__ out,
                                   const unsigned int* __restrict__ sat_count) {
    pdl_sync();
    if (threadIdx.x == 0) {
        if (sat_count && *sat_count) {
            // an activation left the fp16 operand-pair range and was clamped somewhere upstream: fail loudly
            for (int i = 0; i < 7; ++i) out[i] = __int_as_float(0x7fc00000);
            return;
        }
        out[0] = sums6[0];
        out[1] = sums6[1];
        out[2] = sums6[2];
        float bf = sums6[3] / n_pix, bz = sums6[4] / n_pix, bm = sums6[5] / n_pix;
        out[3] = bf;
        out[4] = bz;
        out[5] = bm;
        out[6] = (bf + bz) + bm;
    }
}
int launch_finalize_scalars(const float* sums6, float n_pix, float* scalars7, const unsigned int* sat_count,
                            cudaStream_t s) {
    FVC_CUDA(launch_pdl(k_finalize_scalars, 1, 32, 0, s, sums6, n_pix, scalars7, sat_count));
    g_launch_count++;
    FVC_CHECK_LAUNCH();
    return 0;
}

// ----------------------------------------------------------------------------------------------
// layout conversion
// ----------------------------------------------------------------------------------------------
__global__ void k_nchw_to_act(const float* __restrict__ x, ActT out, int C, int do_abs) {
    pdl_sync();
    int64_t n = (int64_t)out.B * out.H * out.W * out.Cp;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int c = (int)(i % out.Cp);
    int64_t pix = i / out.Cp;
    int xx = (int)(pix % out.W);
    int y = (int)((pix / out.W) % out.H);
    int b = (int)(pix / ((int64_t)out.W * out.H));
    float v = 0.f;
    if (c < C) v = x[(((size_t)b * C + c) * out.H + y) * out.W + xx];
    if (do_abs) v = fabsf(v);
    act_store(out, act_pixel_offset(out, b, y, xx), c, v);
}
int launch_nchw_to_act(const float* x, ActT out, int C, int do_abs, cudaStream_t s) {
    int64_t n = (int64_t)out.B * out.H * out.W * out.Cp;
    FVC_CUDA(launch_pdl(k_nchw_to_act, LAUNCH_1D(n, 256), 0, s, x, out, C, do_abs));
    g_launch_count++;
    FVC_CHECK_LAUNCH();
    return 0;
}
__global__ void k_nhwc_to_act(const float* __restrict__ x, ActT out, int C, int do_abs) {
    pdl_sync();
    int64_t n = (int64_t)out.B * out.H * out.W * out.Cp;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int c = (int)(i % out.Cp);
    int64_t pix = i / out.Cp;
    int xx = (int)(pix % out.W);
    int y = (int)((pix / out.W) % out.H);
    int b = (int)(pix / ((int64_t)out.W * out.H));
    float v = 0.f;
    if (c < C) v = x[(size_t)pix * C + c];
    if (do_abs) v = fabsf(v);
    act_store(out, act_pixel_offset(out, b, y, xx), c, v);
}
int launch_nhwc_to_act(const float* x, ActT out, int C, int do_abs, cudaStream_t s) {
    int64_t n = (int64_t)out.B * out.H * out.W * out.Cp;
    FVC_CUDA(launch_pdl(k_nhwc_to_act, LAUNCH_1D(n, 256), 0, s, x, out, C, do_abs));
    g_launch_count++;
    FVC_CHECK_LAUNCH();
    return 0;
}
__global__ void k_act_to_nchw(ActT in, int C, float* __restrict__ y) {
    int64_t n = (int64_t)in.B * C * in.H * in.W;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int xx = (int)(i % in.W);
    int yy = (int)((i / in.W) % in.H);
    int c = (int)((i / ((int64_t)in.W * in.H)) % C);
    int b = (int)(i / ((int64_t)in.W * in.H * C));
    y[i] = act_load(in, act_pixel_offset(in, b, yy, xx), c);
}
int launch_act_to_nchw(ActT in, int C, float* y, cudaStream_t s) {
    int64_t n = (int64_t)in.B * C * in.H * in.W;
    k_act_to_nchw<<<LAUNCH_1D(n, 256), 0, s>>>(in, C, y);
    g_launch_count++;
    FVC_CHECK_LAUNCH();
    return 0;
}
__global__ void k_nhwc_to_nchw(const float* __restrict__ x, float* __restrict__ y, int B, int C, int H, int W) {
    int64_t n = (int64_t)B * C * H * W;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int xx = (int)(i % W);
    int yy = (int)((i / W) % H);
    int c = (int)((i / ((int64_t)W * H)) % C);
    int b = (int)(i / ((int64_t)W * H * C));
    y[i] = x[(((size_t)b * H + yy) * W + xx) * C + c];
}
int launch_nhwc_to_nchw(const float* x, float* y, int B, int C, int H, int W, cudaStream_t s) {
    int64_t n = (int64_t)B * C * H * W;
    k_nhwc_to_nchw<<<LAUNCH_1D(n, 256), 0, s>>>(x, y, B, C, H, W);
    g_launch_count++;
    FVC_CHECK_LAUNCH();
    return 0;
}
__global__ void k_nchw_to_nhwc(const float* __restrict__ x, float* __restrict__ y, int B, int C, int H, int W) {
    int64_t n = (int64_t)B * C * H * W;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int c = (int)(i % C);
    int xx = (int)((i / C) % W);
    int yy = (int)((i / ((int64_t)C * W)) % H);
    int b = (int)(i / ((int64_t)C * W * H));
    y[i] = x[(((size_t)b * C + c) * H + yy) * W + xx];
}
int launch_nchw_to_nhwc(const float* x, float* y, int B, int C, int H, int W, cudaStream_t s) {
    int64_t n = (int64_t)B * C * H * W;
    k_nchw_to_nhwc<<<LAUNCH_1D(n, 256), 0, s>>>(x, y, B, C, H, W);
    g_launch_count++;
    FVC_CHECK_LAUNCH();
    return 0;
}

}  // namespace fvc
