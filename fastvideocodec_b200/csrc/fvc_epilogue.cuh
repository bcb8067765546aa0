// Shared convolution epilogue: bias, activation, residual, output formats.  Used by both engines.
#pragma once
#include "fvc_common.cuh"

namespace fvc {

__device__ __forceinline__ void ep_load8(const e16* rec, int Cp, int c0, float* v) {
    uint4 h = *reinterpret_cast<const uint4*>(rec + c0);
    uint4 l = *reinterpret_cast<const uint4*>(rec + Cp + c0);
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
__device__ __forceinline__ void ep_store8(e16* rec, int Cp, int c0, const float* v, bool relu) {
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float a = v[2 * j], b = v[2 * j + 1];
        if (relu) {
            a = fmaxf(a, 0.f);
            b = fmaxf(b, 0.f);
        }
        e16 h0, l0, h1, l1;
        split16(a, h0, l0);
        split16(b, h1, l1);
        hi[j] = pack16x2(h0, h1);
        lo[j] = pack16x2(l0, l1);
    }
    *reinterpret_cast<uint4*>(rec + c0) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(rec + Cp + c0) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}


// Packed split of 8 floats into hi/lo 16-bit pairs: one F2FP per two elements for hi, one for lo
// (6 instructions per pair instead of ~14 for the scalar split16 path).
__device__ __forceinline__ uint32_t ep_pack2(float a, float b) {
    uint32_t r;
#if FVC_SPLIT_FP16
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));   // low half = a
#else
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
#endif
    return r;
}
// Range tracking of the stored hi halves: satm holds the running max |hi| of both 16-bit lanes (one HMNMX2 per two
// elements).  A lane that reaches 0x7BFF (65504) was clamped by cvt.satfinite: the caller reports it (sat_count).
__device__ __forceinline__ uint32_t ep_sat_track(uint32_t satm, uint32_t hi) {
#if FVC_SPLIT_FP16 && !defined(FVC_NO_SAT)
    __half2 m = __hmax2(*reinterpret_cast<__half2*>(&satm), __habs2(*reinterpret_cast<__half2*>(&hi)));
    return *reinterpret_cast<uint32_t*>(&m);
#else
    return satm;
#endif
}
__device__ __forceinline__ bool ep_sat_hit(uint32_t satm) {
    return (satm & 0xffffu) >= 0x7bffu || (satm >> 16) >= 0x7bffu;
}
__device__ __forceinline__ void ep_store8_packed(e16* rec, int Cp, int c0, const float* v, bool relu, uint32_t& satm,
                                                 bool with_lo = true) {
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float a = v[2 * j], b = v[2 * j + 1];
        if (relu) {
            a = fmaxf(a, 0.f);
            b = fmaxf(b, 0.f);
        }
        hi[j] = ep_pack2(a, b);
        satm = ep_sat_track(satm, hi[j]);
        float ha, hb;
        e2f2(hi[j], ha, hb);
        lo[j] = ep_pack2(a - ha, b - hb);
    }
    *reinterpret_cast<uint4*>(rec + c0) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    if (with_lo) *reinterpret_cast<uint4*>(rec + Cp + c0) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

// 16 channels at once: one 32-byte (full L2 sector) store for the hi halves and one for the lo halves.
// (16-byte stores write half sectors: ncu showed DRAM reads ~= output size, i.e. read-for-ownership
// fills on every output sector.)  rec + c0 must be 32-byte aligned: c0 % 16 == 0.
__device__ __forceinline__ void ep_pack8(const float* v, bool relu, uint32_t* hi, uint32_t* lo, uint32_t& satm) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float a = v[2 * j], b = v[2 * j + 1];
        if (relu) {
            a = fmaxf(a, 0.f);
            b = fmaxf(b, 0.f);
        }
        hi[j] = ep_pack2(a, b);
        satm = ep_sat_track(satm, hi[j]);
        float ha, hb;
        e2f2(hi[j], ha, hb);
        lo[j] = ep_pack2(a - ha, b - hb);
    }
}
__device__ __forceinline__ void st_global_v8(void* p, const uint32_t* r) {
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(r[0]), "r"(r[1]), "r"(r[2]),
                 "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
__device__ __forceinline__ void ld_global_nc_v8(const void* p, uint32_t* r) {
    asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "l"(p));
}
__device__ __forceinline__ void ep_store16_packed(e16* rec, int Cp, int c0, const float* v16, bool relu, uint32_t& satm,
                                                  bool with_lo = true) {
    uint32_t hi[8], lo[8];
    ep_pack8(v16, relu, hi, lo, satm);
    ep_pack8(v16 + 8, relu, hi + 4, lo + 4, satm);
    st_global_v8(rec + c0, hi);
    if (with_lo) st_global_v8(rec + Cp + c0, lo);   // precision 'fast' keeps the (zero-initialised) lo halves untouched
}

__device__ __forceinline__ float ep_act(float v, int act) {
    switch (act) {
        case FVC_ACT_RELU: return fmaxf(v, 0.f);
        case FVC_ACT_LRELU01: return v > 0.f ? v : v * 0.1f;
        case FVC_ACT_LRELU001: return v > 0.f ? v : v * 0.01f;
        case FVC_ACT_EXP: return expf(v);
        default: return v;
    }
}

// v[NCH]: raw accumulators of output channels co0..co0+NCH-1 of output pixel (b,oy,ox).
// Applies bias + activation (+ residuals) in place and writes every configured output.
template <int NCH>
__device__ __forceinline__ void epilogue_apply(const Epilogue& ep, int Cout, int Hout, int Wout, int b, int oy,
                                               int ox, int co0, float* v, bool skip_bias_act = false) {
    if (!skip_bias_act) {
#pragma unroll
        for (int j = 0; j < NCH; ++j) {
            int c = co0 + j;
            v[j] = (c < Cout) ? ep_act(v[j] * ep.acc_scale + ep.bias[c], ep.act) : 0.f;
        }
    }
    size_t pix = ((size_t)b * Hout + oy) * Wout + ox;
    if (ep.res_act.p) {
        const e16* rec = ep.res_act.p + act_pixel_offset(ep.res_act, b, oy, ox);
#pragma unroll
        for (int g = 0; g < NCH / 8; ++g) {
            if (co0 + g * 8 < ep.res_act.Cp) {
                float r[8];
                ep_load8(rec, ep.res_act.Cp, co0 + g * 8, r);
#pragma unroll
                for (int j = 0; j < 8; ++j) v[g * 8 + j] += r[j];
            }
        }
    }
    if (ep.res_f32) {
#pragma unroll
        for (int j = 0; j < NCH; ++j)
            if (co0 + j < Cout) v[j] += ep.res_f32[pix * Cout + co0 + j];
    }
    if (ep.out_f32) {
        if ((Cout & 3) == 0) {
#pragma unroll
            for (int j = 0; j < NCH; j += 4)
                if (co0 + j < Cout)
                    *reinterpret_cast<float4*>(ep.out_f32 + pix * Cout + co0 + j) =
                        make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        } else {
#pragma unroll
            for (int j = 0; j < NCH; ++j)
                if (co0 + j < Cout) ep.out_f32[pix * Cout + co0 + j] = v[j];
        }
    }
    if (ep.out_act.p) {
        e16* rec = ep.out_act.p + act_pixel_offset(ep.out_act, b, oy, ox);
#pragma unroll
        for (int g = 0; g < NCH / 8; ++g)
            if (co0 + g * 8 < ep.out_act.Cp) ep_store8(rec, ep.out_act.Cp, co0 + g * 8, v + g * 8, false);
    }
    if (ep.out_act_relu.p) {
        e16* rec = ep.out_act_relu.p + act_pixel_offset(ep.out_act_relu, b, oy, ox);
#pragma unroll
        for (int g = 0; g < NCH / 8; ++g)
            if (co0 + g * 8 < ep.out_act_relu.Cp) ep_store8(rec, ep.out_act_relu.Cp, co0 + g * 8, v + g * 8, true);
    }
}

}  // namespace fvc
