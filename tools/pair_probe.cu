// Probe (B200): mechanics and rate of the CTA-pair MMA (tcgen05 cta_group::2, M = 256) that the conv engine's
// pair mode relies on.  One cluster of 2 CTAs per SM pair:
//   * both CTAs TMA-load their own 128 rows of A and their own N/2 rows of B (cp.async.bulk.tensor cta_group::2,
//     completing on the LEADER's mbarrier), the leader issues tcgen05.mma.cta_group::2, a multicast commit wakes
//     both CTAs, each CTA reads its 128 x N block of D from its own TMEM; the peer acknowledges on the leader's
//     barrier through mapa + mbarrier.arrive.shared::cluster.   -> checked against a host product
//   * rate: cycles per pair MMA (K = 16) for N = 32..256 (shared-memory operand model: 4096 + 32*N/2 bytes / SM)
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_bin/pair_probe tools/pair_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
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
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0,1,0,p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity) {
    for (uint32_t i = 0; i < (1u << 24); ++i)
        if (mbar_try(bar, parity)) return true;
    return false;
}
__device__ __forceinline__ uint32_t cta_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mma2(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void commit2(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"((uint16_t)3)
                 : "memory");
}
__device__ __forceinline__ void tma2_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1)
        : "memory");
}

// out: [2 CTAs][128 rows][N] fp32; status[0..7]
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(256, 1)
k_pair(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, int N, int iters, float* out,
       long long* status, int commit_every, int rotate, int fill, int drain) {
    extern __shared__ uint8_t raw[];
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    const uint32_t a0 = base, b0 = base + 64 * 1024, bars = b0 + 64 * 1024, slot = bars + 64;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cta_rank();
    const uint32_t bar_full = bars, bar_done = bars + 8, bar_ack = bars + 16, bar_rate = bars + 24;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_full));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_done));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 8;" ::"r"(bar_ack));    // one arrive per reader warp of both CTAs
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_rate));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1000000;" ::"r"(bars + 32));   // never completes: commit sink
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(slot) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync();   // barriers of both CTAs initialised before any remote signal
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t tmem;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem) : "r"(slot));
    if (threadIdx.x == 0) status[4 + rank] = tmem;

    const uint32_t full_leader = mapa(bar_full, 0), ack_leader = mapa(bar_ack, 0);
    const uint32_t a_bytes = 128 * 128, bh_bytes = (uint32_t)(N / 2) * 128;
    // ---- producer: both CTAs load their halves; the bytes complete on the leader's barrier -------------
    if (warp == 0 && elect_one()) {
        if (rank == 0)
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_full), "r"(2 * (a_bytes + bh_bytes))
                         : "memory");
        tma2_2d(a0, &mapA, full_leader, 0, (int)rank * 128);
        tma2_2d(b0, &mapB, full_leader, 0, (int)rank * (N / 2));
    }
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
    // ---- leader: 4 k-steps (K = 64 = one 128-byte swizzle atom row) ----------------------------------
    if (warp == 1 && rank == 0) {
        const bool ok = mbar_wait(bar_full, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (elect_one()) {
            if (!ok) status[1] = -1;
            const uint64_t ad = make_desc(a0, 1024u), bd = make_desc(b0, 1024u);
            for (int k = 0; k < 4; ++k) mma2(tmem, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, k ? 1u : 0u);
            commit2(bar_done);
        }
    }
    // ---- both CTAs: wait for the multicast commit, read D, acknowledge to the leader -------------------
    if (warp >= 4) {
        const bool ok = mbar_wait(bar_done, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int q = warp & 3;
        if (!ok && lane == 0) status[2] = -1 - (long long)rank;
        for (int c = 0; c < N; c += 8) {
            uint32_t v[8];
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                         : "r"(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)c)
                         : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            for (int j = 0; j < 8; ++j) out[((size_t)rank * 128 + q * 32 + lane) * N + c + j] = __uint_as_float(v[j]);
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0)
            asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(ack_leader) : "memory");
        // optional background TMEM drains (both CTAs, 4 warps each): tcgen05.ld of 64 columns the MMAs do not write,
        // for ~drain * 1000 clk, as the conv engine's accumulator warps do while the next group is issued.
        // (measured: ~179 B/clk per CTA of drains, with or without TMA fills, leave the pair MMA at 43 / 64 clk)
        if (drain && N <= 128) {
            const long long t0 = clock64();
            long long nld = 0;
            float sink = 0.f;
            while (clock64() - t0 < (long long)drain * 1000) {
                for (int c = 0; c < 64; c += 8) {
                    uint32_t v[8];
                    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                                 : "r"(tmem + ((uint32_t)(q * 32) << 16) + 256u + (uint32_t)c)
                                 : "memory");
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    sink += __uint_as_float(v[0]);
                }
                nld += 8;
            }
            if (sink == 123.456f) status[3] = 2;
            if (warp == 4 && lane == 0 && rank == 0) { status[8] = nld * 4 /*warps*/ * 32 * 8 * 4; status[9] = clock64() - t0; }
        }
    }
    // ---- optional background TMA fills (both CTAs): weight-like boxes streamed into a 4-slot ring while the
    // rate loop runs, as the conv engine's producer does ------------------------------------------------
    if (fill && warp == 0 && elect_one()) {
        const uint32_t f0 = base + 129 * 1024, fb = bars + 128;
        const uint32_t box_bytes = (uint32_t)(N / 2) * 128;
        for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(fb + 8 * i));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const long long t0 = clock64();
        long long nfill = 0;
        uint32_t ph = 0;
        while (clock64() - t0 < (long long)fill * 1000) {
            for (int i = 0; i < 4; ++i) {
                if (nfill >= 4) mbar_wait(fb + 8 * i, ph ^ 1u);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(fb + 8 * i), "r"(box_bytes) : "memory");
                asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                             ::"r"(f0 + (uint32_t)i * 16384u), "l"(&mapB), "r"(fb + 8 * i), "r"(0), "r"(0) : "memory");
                ++nfill;
            }
            ph ^= 1u;
        }
        for (int i = 0; i < 4; ++i) mbar_wait(fb + 8 * i, ph ^ 1u);
        if (rank == 0) { status[6] = nfill * box_bytes; status[7] = clock64() - t0; }
    }
    // ---- leader: wait for both acknowledgements, then the rate loop ---------------------------------
    if (warp == 1 && rank == 0) {
        const bool ok = mbar_wait(bar_ack, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const bool lead = elect_one();
        if (lead && !ok) status[3] = -1;
        const long long t0 = clock64();
        const uint64_t ad = make_desc(a0, 1024u), bd = make_desc(b0, 1024u);
        for (int it = 0; it < iters; ++it) {
            if (lead) {
#pragma unroll
                // rotate: rate loop variant 1 walks A over 4 x 16 KB and B over 8 x 8 KB regions (no operand re-use
                // at short distance, as in the conv engine's weight stream); variant 0 re-uses one tile
                const uint64_t ao = rotate ? (uint64_t)(((it & 3) * 16384) >> 4) : 0ull;
                const uint64_t bo = rotate ? (uint64_t)(((it & 7) * 8192) >> 4) : 0ull;
                for (int k = 0; k < 4; ++k) mma2(tmem, ad + ao + (uint64_t)(2 * k), bd + bo + (uint64_t)(2 * k), idesc, 1u);
            }
            // a multicast commit every `commit_every` groups of 4 MMAs (the engine commits once per weight stage)
            if (lead && commit_every && (it & (commit_every - 1)) == commit_every - 1) commit2(bars + 32);   // power of two
            __syncwarp();
        }
        if (lead) commit2(bar_rate);
        mbar_wait(bar_rate, 0);
        if (lead) status[0] = clock64() - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
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
    const int Ns[] = {32, 64, 128, 256, 64, 128, 32, 64, 128, 64, 128, 64, 128, 64, 128};
    const int CEs[] = {0, 0, 0, 0, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2};
    const int ROT[] = {0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 1, 1, 1, 1};
    const int FILL[] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 3000, 3000, 0, 0, 3000, 3000};   // background TMA fills for ~3 M clk
    const int DRAIN[] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 3000, 3000, 3000, 3000};  // background tcgen05.ld drains
    for (int ci = 0; ci < 15; ++ci) {
        const int N = Ns[ci], ce = CEs[ci], rot = ROT[ci], fill = FILL[ci], drain = DRAIN[ci];
        std::vector<__half> hA(256 * 64), hB((size_t)N * 64);
        for (int i = 0; i < 256 * 64; ++i) hA[i] = __float2half((float)((i * 7 + (i >> 6)) % 13 - 6));
        for (int i = 0; i < N * 64; ++i) hB[i] = __float2half((float)((i * 5 + (i >> 6) * 3) % 11 - 5));
        __half *dA, *dB;
        float* dOut;
        long long* dSt;
        cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2);
        cudaMalloc(&dOut, 256 * (size_t)N * 4); cudaMalloc(&dSt, 128);
        cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
        cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
        cudaMemset(dOut, 0, 256 * (size_t)N * 4); cudaMemset(dSt, 0, 128);
        CUtensorMap mA, mB;
        cuuint64_t dimsA[2] = {64, 256}, strA[1] = {128}, dimsB[2] = {64, (cuuint64_t)N};
        cuuint32_t boxA[2] = {64, 128}, boxB[2] = {64, (cuuint32_t)(N / 2)}, es[2] = {1, 1};
        CUresult r1 = encode(&mA, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, dA, dimsA, strA, boxA, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        CUresult r2 = encode(&mB, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, dB, dimsB, strA, boxB, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r1 || r2) { printf("encode failed %d %d\n", (int)r1, (int)r2); return 3; }
        const int smem = 200 * 1024;
        cudaFuncSetAttribute(k_pair, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        const int iters = 4000;
        k_pair<<<2, 256, smem>>>(mA, mB, N, iters, dOut, dSt, ce, rot, fill, drain);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("{\"N\": %d, \"error\": \"%s\"}\n", N, cudaGetErrorString(e)); return 1; }
        std::vector<float> hO(256 * (size_t)N);
        long long st[16];
        cudaMemcpy(hO.data(), dOut, hO.size() * 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(st, dSt, 128, cudaMemcpyDeviceToHost);
        double maxerr = 0;
        int bad = 0;
        for (int m = 0; m < 256; ++m)
            for (int n = 0; n < N; ++n) {
                float ref = 0;
                for (int k = 0; k < 64; ++k) ref += __half2float(hA[m * 64 + k]) * __half2float(hB[n * 64 + k]);
                const double d = fabs((double)ref - hO[(size_t)m * N + n]);
                if (d > maxerr) maxerr = d;
                if (d > 1e-3 && bad++ < 4) printf("  mismatch m=%d n=%d ref=%g got=%g\n", m, n, ref, hO[(size_t)m * N + n]);
            }
        printf("{\"N\": %d, \"rotate\": %d, \"commit_every_4mma_groups\": %d, \"max_abs_err\": %g, \"mismatches\": %d, \"clk_per_pair_mma\": %.1f, \"status\": [%lld, %lld, %lld], "
               "\"tmem_base\": [%lld, %lld], \"fill_bytes_per_clk_per_cta\": %.1f, \"drain_bytes_per_clk_per_cta\": %.1f}\n",
               N, rot, ce, maxerr, bad, (double)st[0] / (4.0 * iters), st[1], st[2], st[3], st[4], st[5], st[7] ? (double)st[6] / (double)st[7] : 0.0, st[9] ? (double)st[8] / (double)st[9] : 0.0);
        cudaFree(dA); cudaFree(dB); cudaFree(dOut); cudaFree(dSt);
    }
    return 0;
}
