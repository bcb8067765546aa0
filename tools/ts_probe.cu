// Probe (B200): A operand from TENSOR MEMORY (tcgen05 "TS" mode) for the hi/lo split products.
// The conv engine issues three MMAs per product (a_hi*w_hi, a_hi*w_lo, a_lo*w_hi) and every MMA re-reads its
// 128 x 16 A block (4 KB) from shared memory: the L1TEX data pipe is the binding unit.  Here the
// A block is copied ONCE into TMEM with tcgen05.cp.128x256b (same K-major SWIZZLE_128B descriptor as the MMA
// would use) and two MMAs read it from there.  Checks D = A*B^T against the host and measures cycles per
// (copy + 2 MMAs) against 2 SS-mode MMAs.
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_bin/ts_probe tools/ts_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cmath>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity) {
    for (uint32_t i = 0; i < (1u << 24); ++i) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0,1,0,p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) return true;
    }
    return false;
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc)
                 : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc)
                 : "memory");
}
__device__ __forceinline__ void cp_a(uint32_t a_tmem, uint64_t adesc) {
    asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(a_tmem), "l"(adesc) : "memory");
}
__device__ __forceinline__ void commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tma_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}

// mode 0: verify TS product; rate loops: variant 0 = SS (2 MMAs per k-step), 1 = TS (cp + 2 MMAs per k-step)
__global__ void __launch_bounds__(256, 1)
k_ts(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, int N, int iters, float* out,
     long long* status) {
    extern __shared__ uint8_t raw[];
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    const uint32_t a0 = base, b0 = base + 16 * 1024, bars = b0 + 32 * 1024, slot = bars + 64;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bars + 8 * i));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(slot) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t tmem;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem) : "r"(slot));
    const uint32_t dcol = tmem, acol = tmem + 256;     // D: columns 0..N-1; A staging: 8 columns per k-step
    if (warp == 0 && elect_one()) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bars), "r"(128 * 128 + N * 128) : "memory");
        tma_2d(a0, &mapA, bars, 0, 0);
        tma_2d(b0, &mapB, bars, 0, 0);
    }
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t ad = make_desc(a0, 1024u), bd = make_desc(b0, 1024u);
    if (warp == 1) {
        const bool ok = mbar_wait(bars, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (elect_one()) {
            if (!ok) status[1] = -1;
            for (int k = 0; k < 4; ++k) {
                cp_a(acol + 8 * k, ad + (uint64_t)(2 * k));
                mma_ts(dcol, acol + 8 * k, bd + (uint64_t)(2 * k), idesc, k ? 1u : 0u);
            }
            commit(bars + 8);
        }
    }
    if (warp >= 4) {
        const bool ok = mbar_wait(bars + 8, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int q = warp & 3;
        if (!ok && lane == 0) status[2] = -1;
        for (int c = 0; c < N; c += 8) {
            uint32_t v[8];
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                         : "r"(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)c) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            for (int j = 0; j < 8; ++j) out[((size_t)q * 32 + lane) * N + c + j] = __uint_as_float(v[j]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // ---- rate loops ------------------------------------------------------------------------------------
    if (warp == 1) {
        const bool lead = elect_one();
        for (int variant = 0; variant < 2; ++variant) {
            const long long t0 = clock64();
            for (int it = 0; it < iters; ++it) {
                if (lead) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        if (variant == 0) {
                            mma_ss(dcol, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, 1u);
                            mma_ss(dcol + 128, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, 1u);
                        } else {
                            const uint32_t ac = acol + 8 * (uint32_t)((4 * it + k) & 7);   // 8 staging slots in rotation
                            cp_a(ac, ad + (uint64_t)(2 * k));
                            mma_ts(dcol, ac, bd + (uint64_t)(2 * k), idesc, 1u);
                            mma_ts(dcol + 128, ac, bd + (uint64_t)(2 * k), idesc, 1u);
                        }
                    }
                }
                __syncwarp();
            }
            if (lead) commit(bars + 16 + 8 * variant);
            mbar_wait(bars + 16 + 8 * variant, 0);
            if (lead) status[4 + variant] = clock64() - t0;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
    }
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main() {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) return 2;
    PFN_encodeTiled encode = (PFN_encodeTiled)p;
    const int Ns[] = {32, 64, 128};
    for (int N : Ns) {
        std::vector<__half> hA(128 * 64), hB((size_t)N * 64);
        for (int i = 0; i < 128 * 64; ++i) hA[i] = __float2half((float)((i * 7 + (i >> 6)) % 13 - 6));
        for (int i = 0; i < N * 64; ++i) hB[i] = __float2half((float)((i * 5 + (i >> 6) * 3) % 11 - 5));
        __half *dA, *dB; float* dOut; long long* dSt;
        cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2);
        cudaMalloc(&dOut, 128 * (size_t)N * 4); cudaMalloc(&dSt, 64);
        cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
        cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
        cudaMemset(dOut, 0, 128 * (size_t)N * 4); cudaMemset(dSt, 0, 64);
        CUtensorMap mA, mB;
        cuuint64_t dimsA[2] = {64, 128}, str[1] = {128}, dimsB[2] = {64, (cuuint64_t)N};
        cuuint32_t boxA[2] = {64, 128}, boxB[2] = {64, (cuuint32_t)N}, es[2] = {1, 1};
        CUresult r1 = encode(&mA, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, dA, dimsA, str, boxA, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        CUresult r2 = encode(&mB, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, dB, dimsB, str, boxB, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r1 || r2) { printf("encode failed\n"); return 3; }
        const int smem = 64 * 1024, iters = 4000;
        cudaFuncSetAttribute(k_ts, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        k_ts<<<1, 256, smem>>>(mA, mB, N, iters, dOut, dSt);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("{\"N\": %d, \"error\": \"%s\"}\n", N, cudaGetErrorString(e)); return 1; }
        std::vector<float> hO(128 * (size_t)N);
        long long st[8];
        cudaMemcpy(hO.data(), dOut, hO.size() * 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(st, dSt, 64, cudaMemcpyDeviceToHost);
        double maxerr = 0; int bad = 0;
        for (int m = 0; m < 128; ++m)
            for (int n = 0; n < N; ++n) {
                float ref = 0;
                for (int k = 0; k < 64; ++k) ref += __half2float(hA[m * 64 + k]) * __half2float(hB[n * 64 + k]);
                const double d = fabs((double)ref - hO[(size_t)m * N + n]);
                if (d > maxerr) maxerr = d;
                if (d > 1e-3 && bad++ < 4) printf("  mismatch m=%d n=%d ref=%g got=%g\n", m, n, ref, hO[(size_t)m * N + n]);
            }
        printf("{\"N\": %d, \"ts_max_abs_err\": %g, \"mismatches\": %d, \"clk_per_kstep_2mma_ss\": %.1f, "
               "\"clk_per_kstep_cp_plus_2mma_ts\": %.1f, \"status\": [%lld, %lld]}\n",
               N, maxerr, bad, (double)st[4] / (4.0 * iters), (double)st[5] / (4.0 * iters), st[1], st[2]);
        cudaFree(dA); cudaFree(dB); cudaFree(dOut); cudaFree(dSt);
    }
    return 0;
}
