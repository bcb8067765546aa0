// Probe (B200): what global-store rate does the conv engine's epilogue ACCESS PATTERN allow?
// The accumulator warps of k_conv_tc write ACT records (NHWC [hi Cp | lo Cp] fp16): a thread owns one pixel and 32 of
// its channels, i.e. 64 contiguous bytes of the hi half and 64 of the lo half, stored as four 32-byte st.global.v8 —
// every lane of a warp instruction hits a different 128-byte line.  Output-heavy layers (feature_ext, ResBlock conv1)
// run at ~8 B/clk/SM of stores = 2.1 TB/s.  This probe times, on a 534 MB buffer (one full-resolution 64-channel
// tensor), with a persistent 148 x 512-thread grid like the engine's:
//   pattern  : the engine's per-thread pattern (lane = pixel, 2 x 64 B runs per 256-byte record)
//   coalesced: the same bytes with each warp instruction writing 1 KB contiguous (what smem staging / TMA store gives)
//   memset   : cudaMemsetAsync
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_bin/store_probe tools/store_probe.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

__device__ __forceinline__ void st_v8(void* p, uint32_t v) {
    asm volatile("st.global.v8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"l"(p), "r"(v) : "memory");
}

// tile = 256 pixels (two 128-pixel sub-tiles), 512 threads: thread -> (pixel, channel half)
__global__ void __launch_bounds__(512) k_pattern(uint8_t* out, int ntiles, int work) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int q = warp & 3, cg = warp >> 2;          // TMEM lane quarter, column group (sub-tile, channel half)
    const int px = (cg >> 1) * 128 + q * 32 + lane, half = cg & 1;
    uint32_t v = threadIdx.x;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        for (int i = 0; i < work; ++i) v = v * 1664525u + 1013904223u;   // stand-in for the epilogue arithmetic
        uint8_t* rec = out + ((size_t)t * 256 + px) * 256 + half * 64;
        st_v8(rec, v);
        st_v8(rec + 32, v);
        st_v8(rec + 128, v);
        st_v8(rec + 160, v);
    }
}

__global__ void __launch_bounds__(512) k_coalesced(uint8_t* out, int ntiles, int work) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t v = threadIdx.x;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        for (int i = 0; i < work; ++i) v = v * 1664525u + 1013904223u;
        uint8_t* base = out + (size_t)t * 65536 + warp * 4096 + lane * 32;   // 4 KB per warp, 1 KB per instruction
#pragma unroll
        for (int j = 0; j < 4; ++j) st_v8(base + j * 1024, v);
    }
}

int main() {
    const int ntiles = 8160;                 // 1088 x 1920 pixels / 256
    const size_t bytes = (size_t)ntiles * 65536;
    uint8_t* buf;
    cudaMalloc(&buf, bytes);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    auto run = [&](const char* name, int mode, int grid, int work) {
        float best = 1e9f;
        for (int it = 0; it < 6; ++it) {
            cudaEventRecord(e0);
            if (mode == 0) k_pattern<<<grid, 512>>>(buf, ntiles, work);
            else if (mode == 1) k_coalesced<<<grid, 512>>>(buf, ntiles, work);
            else cudaMemsetAsync(buf, it, bytes);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            if (it >= 2 && ms < best) best = ms;
        }
        printf("{\"probe\": \"store\", \"kind\": \"%s\", \"grid\": %d, \"work\": %d, \"ms\": %.4f, \"GBps\": %.0f}\n", name, grid,
               work, best, bytes / best * 1e-6);
    };
    run("memset", 2, 0, 0);
    for (int grid : {148, 296, 8160})
        for (int work : {0, 64}) {
            run("pattern", 0, grid, work);
            run("coalesced", 1, grid, work);
        }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
