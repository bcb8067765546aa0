// Micro-benchmark (B200): cycles per tcgen05.mma (M=128, K=16, 16-bit operands, SS mode, SWIZZLE_128B
// K-major) as a function of N and of the number of issuing warps, plus tcgen05.ld drain bandwidth.
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/mma_rate tools/mma_rate.cu
// Grounds the tile-shape choices of fvc_conv_tc.cu ("tensor pipe model").
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

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
__device__ __forceinline__ void tc_mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc)
        : "memory");
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

// mode 0: every MMA re-uses the same A/B tile (4 k-steps inside a 128-B swizzle atom)
// mode 1: A start address also moves by 128 B per MMA group (tap shift), 48 KB A region
// mode 2: as 0, but the accumulator rotates over 4 column blocks every 4 MMAs (sub-tiles of the conv engine)
// mode 3: as 2, with the conv engine's strided A (8-row groups 38*128 B apart) and moving tap offsets
// mode 4: as 0, accumulator rotates over 4 column blocks every MMA
__global__ void __launch_bounds__(256, 1) k_rate(int N, int nissue, int iters, int mode, long long* out) {
    extern __shared__ uint8_t raw[];
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    const uint32_t a0 = base, b0 = base + 64 * 1024, bars = b0 + 32 * 1024 + 1024, slot = bars + 64;
    for (uint32_t i = threadIdx.x * 4; i < 100 * 1024; i += blockDim.x * 4)
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(base + i), "r"(0x3c003c00u));   // fp16 1.0
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 8; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bars + 8 * i));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1000000;" ::"r"(bars + 32));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(slot) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (threadIdx.x == 0) asm volatile("st.shared.b32 [%0], %1;" ::"r"(slot + 16), "r"(0u));
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t tmem;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem) : "r"(slot));
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    long long t0 = clock64();
    if (warp < nissue) {
        const uint32_t dcol = tmem + (uint32_t)((warp * N) % (512 - N + 1));
        const uint64_t ad = make_desc(a0 + warp * 1024u, mode == 3 ? 38u * 128u : 1024u), bd = make_desc(b0, 1024u);
        const bool lead = elect_one();
        for (int it = 0; it < iters; ++it) {
            const uint64_t ashift = (mode == 1 || mode == 3) ? (uint64_t)(((it % 40) * 128u) >> 4) : 0ull;
            const uint32_t drot = (mode == 2 || mode == 3) ? (uint32_t)((it & 3) * N) : 0u;
            if (lead) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    tc_mma(dcol + drot + (mode == 4 ? (uint32_t)(k * N) : 0u), ad + ashift + (uint64_t)(k * 2),
                           bd + (uint64_t)(k * 2), idesc, 1u);
            }
            // modes 5/6/7: a commit to a (never waited) mbarrier every 16 / 4 / 64 MMAs
            if (lead && ((mode == 5 && (it & 3) == 3) || mode == 6 || (mode == 7 && (it & 15) == 15)))
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bars + 32)
                             : "memory");
            __syncwarp();
        }
        if (lead)
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bars + 8 * warp)
                         : "memory");
        while (!mbar_try(bars + 8 * warp, 0)) {}
    }
    // modes 8/9: warps 4..7 read TMEM (other columns) while warp 0 issues MMAs: does tcgen05.ld slow the MMAs?
    if (mode >= 8 && warp >= 4) {
        volatile uint32_t* flag = reinterpret_cast<volatile uint32_t*>(raw + (slot - smem_u32(raw)) + 16);
        const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + 256u;
        float acc = 0.f;
        unsigned long long nld = 0;
        while (*flag == 0) {
            for (int c = 0; c < 32; c += 8) {
                uint32_t v[8];
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                             : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                             : "r"(taddr + c) : "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int q = 0; q < 8; ++q) acc += __uint_as_float(v[q]);
            }
            nld += 4;
            if (mode == 9) __nanosleep(400);
        }
        if (acc == 123.f) out[3] = 1;
        if (blockIdx.x == 0 && threadIdx.x == 128) out[1] = (long long)nld;
    }
    long long t1 = clock64();
    if (mode >= 8 && warp == 0) {
        volatile uint32_t* flag = reinterpret_cast<volatile uint32_t*>(raw + (slot - smem_u32(raw)) + 16);
        *flag = 1;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
    }
}

// TMEM drain: nwarps warps (multiple of 4) each read `cols` columns of their lane quarter, `iters` times
__global__ void __launch_bounds__(1024, 1) k_drain(int cols, int iters, long long* out, float* sink) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = slot;
    const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * cols) % 512u;
    float acc = 0.f;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        for (int c = 0; c < cols; c += 8) {
            uint32_t v[8];
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                         : "r"(taddr + c) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int q = 0; q < 8; ++q) acc += __uint_as_float(v[q]);
        }
    }
    long long t1 = clock64();
    if (acc == 123.f) sink[threadIdx.x] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
    }
}

int main() {
    long long* d;
    float* sink;
    cudaMalloc(&d, 64); cudaMemset(d, 0, 64);
    cudaMalloc(&sink, 4096);
    cudaFuncSetAttribute(k_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024);
    const int iters = 2000;
    for (int mode = 0; mode < 10; ++mode)
        for (int N : {16, 32, 64, 128, 256})
            for (int nissue : {1, 2, 4}) {
                if (nissue * N > 512) continue;
                if (mode >= 2 && (nissue > 1 || N > 128)) continue;
                long long h = 0;
                for (int rep = 0; rep < 2; ++rep) {
                    k_rate<<<148, 256, 110 * 1024>>>(N, nissue, iters, mode, d);
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                }
                long long hh[2] = {0, 0};
                cudaMemcpy(hh, d, 16, cudaMemcpyDeviceToHost);
                h = hh[0];
                double per = (double)h / ((double)iters * 4 * nissue);
                printf("{\"bench\":\"mma\",\"mode\":%d,\"N\":%d,\"issuers\":%d,\"clk_per_mma\":%.1f,\"floor\":%.1f,\"frac_of_floor\":%.3f,\"ldtm_bytes_per_clk\":%.1f}\n",
                       mode, N, nissue, per, 128.0 * N / 256.0, (128.0 * N / 256.0) / per,
                       mode >= 8 ? (double)hh[1] * 4 * 1024.0 / (double)h : 0.0);
            }
    for (int nw : {4, 8, 16})
        for (int cols : {32, 64, 128}) {
            long long h = 0;
            for (int rep = 0; rep < 2; ++rep) {
                k_drain<<<148, nw * 32, 0>>>(cols, 200, d, sink);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            }
            cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
            double bytes = (double)nw * 32 * cols * 4 * 200;
            printf("{\"bench\":\"drain\",\"warps\":%d,\"cols\":%d,\"bytes_per_clk\":%.1f}\n", nw, cols, bytes / (double)h);
        }
    return 0;
}
