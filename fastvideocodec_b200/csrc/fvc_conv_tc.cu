// tcgen05 / TMEM / TMA implicit-GEMM convolution engine (sm_100a).  Design notes and measurements: DESIGN.md 4.1.
//
// GEMM view: M = 128 output pixels (a 16x8 sub-tile), N = output channels, K = taps x input channels.
//   * A (activations) is never im2col'ed.  A CTA loads ONE halo'ed input patch [PH][PW][pitch] per record
//     segment with a single 5-D TMA box (zero fill outside the image = conv padding) and every filter tap is
//     just a different start address into that patch: the 8-row core groups of the UMMA K-major swizzled
//     layout are 8 consecutive pixels of one image row and successive groups are successive image rows
//     (SBO = PW*pitch).  pitch = 128 B (SWIZZLE_128B) or 32 B (SWIZZLE_32B, narrow [hi 8 | lo 8] records).
//   * B (weights) is a pre-packed 16-bit stream in exactly the order the K loop consumes it, pulled by 2-D TMA
//     boxes of T tiles through a multi-stage mbarrier ring.
//   * Precision: activations and weights are hi/lo fp16 pairs (bf16 pairs with FVC_SPLIT=bf16); each product
//     is issued as a_hi*w_hi + a_hi*w_lo + a_lo*w_hi on kind::f16 MMAs with fp32 accumulation in TMEM.  For
//     Cout <= 32 the w_hi and w_lo rows sit side by side in one tile (MMA N = 2*Cout, "merged"): a_hi is read
//     once for two products and the epilogue adds the two column blocks.
//   * S sub-tiles (S*N accumulator columns in TMEM) share every weight stage, cutting the weight traffic per
//     pixel S-fold; stride-2 convs read the four parity planes of a parity-planar input; stride-2 transposed
//     convs run as 4 output-phase sub-convolutions.
//   * Accumulation: tcgen05 adds into its fp32 TMEM accumulator with TRUNCATION (measured: relative bias
//     -4.9e-8 per MMA in the chain, -2.9e-5 after the 588 MMAs of a 7x7x64 filter; tools/tc_bias.py).  So
//     TMEM only ever holds SHORT chains (<= 48 MMAs, a "group" of taps, possibly spanning the segment passes
//     of a tile): two or four partial-accumulator buffers rotate, and 16 accumulator warps drain each finished
//     partial with tcgen05.ld and add it to a running sum in registers in fp32 round-to-nearest, overlapped
//     with the MMAs of the next group.
//   * Warp roles (20 warps): warp 0 = TMA producer (polls the weight ring and the patch ring), warps 1-3 =
//     MMA issuers (sub-tile s belongs to issuer s mod 3; warp 1 owns the TMEM allocation), warps 4..19 = accumulator /
//     epilogue warps (drain -> running sum -> bias/activation/skip/GDN -> hi/lo split -> global).  setmaxnreg
//     gives the control warpgroup 64 and the accumulator warpgroups 104 registers.  Persistent over tiles.
//   * CTA pairs (cta_group::2): two x-neighbouring tiles form one M = 256 MMA issued by the leader CTA; each CTA
//     stages its own patch and half of every weight tile; commits are multicast to both CTAs.
//   * Up to four partial-accumulator buffers (CT <= 128) let the issuers run a whole tile ahead of an epilogue.
#include <cuda.h>
#include <atomic>
#include <type_traits>
#include <vector>
#include <cstring>
#include <cmath>

#include "fvc_kernels.cuh"
#include "fvc_epilogue.cuh"

namespace fvc {

// ----------------------------------------------------------------------------------------------
// device tables
// ----------------------------------------------------------------------------------------------
#define TC_MAX_PASS 16
#define TC_MAX_TAPS 64
#define TC_THREADS 640   // 20 warps = 5 warpgroups: warps 0-3 {producer, issuer, issuer, idle} shrink to 64 registers
                         // (setmaxnreg), the 16 accumulator warps 4-19 grow to 104: 128*64 + 512*104 = the CTA's launch
                         // allocation 640*96 (a larger request blocks forever: the pool is what the launch allocated)
#define TC_ACC_WARP0 4
#define TC_ISSUERS 3     // warps 1..3: the issue loop costs ~100 clk per MMA and thread (register -> uniform-register
                         // moves of the descriptors), so the sub-tiles of a tile are spread over three issuing threads
#define TC_REGS_CTRL 64
#define TC_REGS_ACC 104

struct TcPass {
    int8_t seg, plane;     // record segment (128 B unit) and parity plane (0 for stride-1 inputs)
    int8_t oy, ox;         // patch origin relative to the tile's q origin (input grid of that plane)
    int16_t tap_first, ntaps;
    int8_t nbt, ks0, ks1;  // weight tiles per tap and their k-step counts (16 channels per k-step)
    int8_t gtaps;          // taps per accumulation group (one TMEM chain, drained to registers)
    uint32_t btile_first;  // first weight tile of the pass in the stream
    // accumulation groups (TMEM chains of <= chain_max MMAs) may run across the segment passes of a tile:
    // bit t of gmask = a new group starts at tap t; gend = the group is closed after the last tap
    // (low-tap layers, e.g. stride-2 transposed convolutions, otherwise drain TMEM once per pass)
    uint32_t gmask_lo, gmask_hi;
    int32_t gend;
};
struct TcSub {
    int pass_first, npass, py, px;
    uint32_t btile_first;
    int ngroups;           // accumulation groups per tile
};
struct alignas(64) TcParams {
    CUtensorMap mapA;
    CUtensorMap mapB;
    TcSub sub[4];
    TcPass pass[TC_MAX_PASS];
    int32_t tap_off[TC_MAX_TAPS + 1];  // byte offset of the tap's window inside the patch (+1: the issuers prefetch
                                       // entry t + 1 without a bound check, see the issue loops)
    int nsub, S, SX, N, PW, PH, nst, CT;   // CT = S*N accumulator columns per partial buffer
    int Hq, Wq, tiles_x, tiles_y, B, os, Hout, Wout, Cout, nchunks;
    uint32_t patch_bytes, patch_tx, btile_bytes, stage_bytes, tmem_cols, idesc;
    int npb, T;       // patch buffers (1 or 2), weight tiles per stage
    int pair;         // CTA-pair mode: tcgen05 cta_group::2, M = 256 (two CTAs x 128 pixels), each CTA stages half of
                      // every weight tile; tiles_x then counts PAIRS of tiles along x
    int nab_log2;     // log2 of the number of partial-accumulator buffers in TMEM (2 or 4 buffers of CT columns)
    int fast;         // precision 'fast': hi halves only (one MMA per product); ACT outputs are written without lo
    int merged;       // Cout <= 16: weight rows interleave 8-row blocks of w_hi and w_lo (MMA N = 32, two
                      // products per A read); the epilogue adds the two column blocks of every channel chunk
    int pitch;        // bytes per pixel of a record segment in shared memory: 128 (SWIZZLE_128B), or 32 for the
                      // 8-channel records [hi 8 | lo 8] of the network's input layers (SWIZZLE_32B)
    uint32_t zero;    // always 0 (opaque to the compiler: used to build false dependencies)
    uint32_t acc_sleep_ns;
    unsigned long long* dbg;   // optional [8] cycle counters of block 0's MMA issuer (FVC_TC_DEBUG=1)
    int planes;       // 1 or 4 (parity-planar input)
    // TMA-store epilogue (per warp): every accumulator warp stages the 32 pixels x 32 channels it owns in its own 4 KB of
    // shared memory ([32 rows][64 B] hi block + lo block, SWIZZLE_64B) and writes them with two cp.async.bulk.tensor
    // stores; no synchronisation across warps
    CUtensorMap mapO;
    CUtensorMap mapG;  // fused GDN: the packed gamma stream (three 64-row tiles)
    int gdn;          // 1: fused (I)GDN epilogue (k_conv_tc<4, false, false, 2>)
    int tap;          // 1: fused tail convolution (k_conv_tc<4, RES, false, 3>); mapG then maps the regrouped weights
    uint32_t tap_idesc;   // instruction descriptor of the second MMA (M = 128, N = 32)
    int tmast;        // 0: per-lane st.global epilogue; 1: per-warp staged TMA stores (out_act only)
    int o_mode;       // 0: plain NHWC output; 1: parity-planar output (4 plane boxes); 2: stride-2 transposed
                      //    conv (output pixel = 2q + phase: boxes with element stride 2)
    int o_nseg;       // 64-byte pieces of the hi half of an output record (Cp / 32); the lo half follows
    uint32_t stg_bytes;
    Epilogue ep;
    int nab;          // number of partial buffers: 1 << nab_log2, or 3 (fused GDN / tail kernels: two buffers hold less
                      // than the three chains of a 3x3 tile, four do not fit beside the second MMA's accumulators).
                      // Last member: the offsets of everything above stay what the default kernels were tuned with
};

// ----------------------------------------------------------------------------------------------
// PTX wrappers
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// bounded wait: a protocol bug traps instead of hanging the GPU.  (No printf here: a call in the slow
// path makes the compiler spill every live accumulator around each wait.)
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 26)) __trap();
    }
}
// polite wait for warps that are off the critical path (16 accumulator warps polling at full speed
// take issue slots from the MMA issuer warp that shares their scheduler)
__device__ __forceinline__ void mbar_wait_sleep(uint32_t bar, uint32_t parity, uint32_t ns) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        __nanosleep(ns);
        if (++spins > (1u << 24)) __trap();
    }
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1,
                                            int c2, int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
// ---- CTA-pair (cta_group::2) variants ------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster) {
    // default semantics (release at CTA scope), as for a local arrive: a cluster-scope release costs an ERRBAR that
    // waits for the epilogue's outstanding global stores (ncu: 14 % of the kernel's stall samples)
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster) : "memory");
}
// TMA loads of a CTA pair: the bytes complete on the barrier `bar_cluster` (a shared::cluster address, here
// always the leader CTA's barrier)
__device__ __forceinline__ void tma_load_5d_pair(uint32_t dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1,
                                                 int c2, int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_store_5d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3, int c4) {
    asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];"
                 ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
                 : "memory");
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// 16 output channels (lc = 0 or 2: first or second chunk pair of the thread's 32 channels) of one pixel into the warp's
// staging blocks: [32 rows][64 B] for the hi halves at wstg, the lo halves 2 KB further; 16-byte chunk j of row r sits at
// chunk j ^ ((r >> 1) & 3) (SWIZZLE_64B, the layout the store's tensor map expects; the 32 lanes of the warp, 32 rows
// writing the same logical chunk, then hit 8 different 16-byte bank groups: 4 wavefronts per 512 bytes, the minimum).
__device__ __forceinline__ void stage_store16(uint32_t wstg, int row, int lc, const float* v16, uint32_t& satm,
                                              bool with_lo) {
    uint32_t hi[8], lo[8];
    ep_pack8(v16, false, hi, lo, satm);
    ep_pack8(v16 + 8, false, hi + 4, lo + 4, satm);
    const uint32_t sw = (uint32_t)(row >> 1) & 3u;
    const uint32_t r0 = wstg + ((uint32_t)row << 6);
    st_shared_v4(r0 + ((((uint32_t)lc) ^ sw) << 4), hi[0], hi[1], hi[2], hi[3]);
    st_shared_v4(r0 + ((((uint32_t)lc + 1u) ^ sw) << 4), hi[4], hi[5], hi[6], hi[7]);
    if (with_lo) {
        st_shared_v4(r0 + 2048u + ((((uint32_t)lc) ^ sw) << 4), lo[0], lo[1], lo[2], lo[3]);
        st_shared_v4(r0 + 2048u + ((((uint32_t)lc + 1u) ^ sw) << 4), lo[4], lo[5], lo[6], lo[7]);
    }
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// commit of the leader's MMAs, delivered to the barrier at the same offset in BOTH CTAs of the pair
__device__ __forceinline__ void tc_commit_pair(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"((uint16_t)3)
                 : "memory");
}
template <bool PAIR>
__device__ __forceinline__ void tc_commit_t(uint32_t bar) {
    if (PAIR) tc_commit_pair(bar);
    else tc_commit(bar);
}
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tc_ld8(uint32_t taddr, uint32_t* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
// one thread of a converged warp; the compiler then knows the branch is single-threaded and keeps
// descriptors in uniform registers (no per-instruction waterfall as with `lane == 0`)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tc_st8(uint32_t taddr, const uint32_t* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, sm100 version 1)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t sbo_bytes, uint32_t layout = 2u) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);                 // start address, bits [0,14)
    d |= (uint64_t)1 << 16;                                   // leading byte offset (ignored, K-major swizzled)
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;        // stride byte offset, bits [32,46)
    d |= (uint64_t)1 << 46;                                   // descriptor version 1 (Blackwell)
    d |= (uint64_t)layout << 61;                              // 2: SWIZZLE_128B, 6: SWIZZLE_32B
    return d;
}

// tcgen05.mma with the descriptors given as (low word, high word): the high words are loop constants
// and the low words move by small immediates, so the issuing thread does 32-bit adds only.
__device__ __forceinline__ void tc_mma2(uint32_t tmem_d, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi,
                                        uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "mov.b64 da, {%1, %2};\n\t"
        "mov.b64 db, {%3, %4};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
        ::"r"(tmem_d), "r"(alo), "r"(ahi), "r"(blo), "r"(bhi), "r"(idesc), "r"(acc)
        : "memory");
}

__device__ __forceinline__ void tc_mma2_pair(uint32_t tmem_d, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi,
                                             uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "mov.b64 da, {%1, %2};\n\t"
        "mov.b64 db, {%3, %4};\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n\t}"
        ::"r"(tmem_d), "r"(alo), "r"(ahi), "r"(blo), "r"(bhi), "r"(idesc), "r"(acc)
        : "memory");
}

// All MMAs of one weight tile: S sub-tiles x KS k-steps, straight-line (the issuing thread is the
// bottleneck otherwise: the MMA queue is shallow, every scalar instruction between MMAs shows up as
// tensor-pipe idle time).
template <int KS, bool PAIR>
__device__ __forceinline__ void issue_subtile(uint32_t dcol, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi,
                                              uint32_t idesc, uint32_t acc0) {
#pragma unroll
    for (int k = 0; k < KS; ++k) {
        if (PAIR) tc_mma2_pair(dcol, alo + (uint32_t)(2 * k), ahi, blo + (uint32_t)(2 * k), bhi, idesc, k == 0 ? acc0 : 1u);
        else tc_mma2(dcol, alo + (uint32_t)(2 * k), ahi, blo + (uint32_t)(2 * k), bhi, idesc, k == 0 ? acc0 : 1u);
    }
}
// An issuer warp owns the sub-tiles s = issuer, issuer + TC_ISSUERS (< S): at most two.  Their accumulator-column
// and A-address offsets (do*, so*) are computed once per kernel; `ns` (0..2) is warp-uniform, so the two `if`s are
// uniform branches.  (Written as a loop over all S sub-tiles with an ownership test, every issuer executed the
// descriptor arithmetic of ALL sub-tiles with predicated-off MMAs: ~340 clk per filter tap in the narrow loop.)
template <int KS, bool PAIR>
__device__ __forceinline__ void issue_stage(uint32_t dcol, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi,
                                            uint32_t idesc, uint32_t acc0, int ns, uint32_t so0, uint32_t do0,
                                            uint32_t so1, uint32_t do1) {
    if (ns > 0) issue_subtile<KS, PAIR>(dcol + do0, alo + so0, ahi, blo, bhi, idesc, acc0);
    if (ns > 1) issue_subtile<KS, PAIR>(dcol + do1, alo + so1, ahi, blo, bhi, idesc, acc0);
}

// ----------------------------------------------------------------------------------------------
// (tile, pass) iterator shared by the producer and the MMA issuer
// ----------------------------------------------------------------------------------------------
struct PassIter {
    int tile, pass;     // current tile id, pass index inside the tile's sub-convolution
    int b, sub, ty, tx;
    int xmul, xadd;     // CTA pairs: tile column = 2 * (pair column) + CTA rank
    __device__ __forceinline__ void decode(const TcParams& P) {
        // the sub-convolution (output phase of a stride-2 transposed convolution) varies fastest: the 4 phases of a
        // q-tile read the same input patch and now run at the same time on neighbouring CTAs, so three of the four
        // patch loads hit L2 (phase-major order re-read the whole input from DRAM once per phase: 1.07 GB instead of
        // 0.27 GB for mvDecoder.deconv7, ncu)
        // (rotated by the q-tile index: a persistent CTA strides by a multiple of 4 tiles and would otherwise always get
        // the same phase, and the phases have 1, 2, 2 and 4 taps)
        int t = tile;
        sub = (t + t / P.nsub) % P.nsub; t /= P.nsub;
        tx = (t % P.tiles_x) * xmul + xadd; t /= P.tiles_x;
        ty = t % P.tiles_y;
        b = t / P.tiles_y;
    }
    __device__ __forceinline__ bool valid(int ntiles) const { return tile < ntiles; }
    __device__ __forceinline__ void next(const TcParams& P, int stride) {
        if (++pass >= P.sub[sub].npass) {
            pass = 0;
            tile += stride;
            decode(P);
        }
    }
};

// Tile epilogue of one accumulator thread: NCH chunks of 8 consecutive accumulator columns.  Global
// loads (residual) of a batch of chunks are all issued before any of them is consumed and before the
// batch's stores, so a thread pays one memory round trip per batch instead of one per chunk
// (ncu: the per-chunk version was stall_long_sb-bound and starved the MMA pipe).
// LEAN: the instantiation for the common layer shape (ACT outputs only, no activation beyond ReLU / LeakyReLU(0.1), additive
// skip only): the rarely used branches are compiled out, the kernel's code shrinks (instruction fetch is a visible stall
// of the epilogue-bound layers, ncu)
template <int NCH, bool RES, bool STG = false, bool LEAN = false>
__device__ __forceinline__ void tile_epilogue(const TcParams& P, const float* __restrict__ bias_s, const float* run,
                                              int b, int sub, int ty, int tx, int th, int tw, uint32_t colbase,
                                              uint32_t stg) {
    const Epilogue& ep = P.ep;
    constexpr int HB = NCH > 4 ? 2 : NCH;   // chunks per batch (register budget: 96 regs at 576 threads)
    const int N = P.merged ? (P.N >> 1) : P.N, Cout = P.Cout, act = ep.act;   // channels per sub-tile
    const float acc_scale = ep.acc_scale;
    const int qy = ty * 16 + th;
    const int oy = qy * P.os + P.sub[sub].py;
    const bool row_ok = qy < P.Hq;
    int s0 = (int)colbase / N, c00 = (int)colbase - s0 * N;
    // per-pixel state, recomputed only when the chunk sequence moves to the next sub-tile
    int cur_s = -1;
    bool ok = false;
    e16 *rec_out = nullptr, *rec_relu = nullptr, *rec_sq = nullptr;
    const e16* rec_res = nullptr;
    size_t pixC = 0;   // pixel index * Cout (fp32 NHWC tensors)
#ifdef FVC_NO_FAST
    const bool wlo = true;
#else
    const bool wlo = P.fast == 0;   // precision 'fast': only the hi halves of ACT outputs are written
#endif
    // staging row of this thread's pixel inside its warp's block (TMA-store epilogue): the order in which the store
    // boxes enumerate the warp's 4 x 8 pixels (parity-planar outputs: plane-major, 2 x 4 pixels per plane)
    const int lth = th & 3;
    const int srow = P.o_mode == 1 ? ((((lth & 1) << 1) | (tw & 1)) * 8 + (lth >> 1) * 4 + (tw >> 1)) : (lth * 8 + tw);
    const int cA = (int)colbase - ((int)colbase / N) * N;      // first channel of this thread inside its sub-tile
    uint32_t satm = 0; // running max |hi| of the ACT values this thread stores (range check, see ep_sat_track)
    auto enter_pixel = [&](int s) {
        cur_s = s;
        const int qx = (tx * P.SX + s) * 8 + tw;
        const int ox = qx * P.os + P.sub[sub].px;
        ok = row_ok && qx < P.Wq;
        if (!ok) return;
        if (ep.out_act.p) rec_out = ep.out_act.p + act_pixel_offset(ep.out_act, b, oy, ox);
        if (ep.out_act_relu.p) rec_relu = ep.out_act_relu.p + act_pixel_offset(ep.out_act_relu, b, oy, ox);
        if (ep.out_act_sq.p) rec_sq = ep.out_act_sq.p + act_pixel_offset(ep.out_act_sq, b, oy, ox);
        if (RES) rec_res = ep.res_act.p + act_pixel_offset(ep.res_act, b, oy, ox);
        pixC = (((size_t)b * P.Hout + oy) * P.Wout + ox) * (size_t)Cout;
    };
#pragma unroll
    for (int h0 = 0; h0 < NCH; h0 += HB) {
        uint4 rh[HB], rl[HB];
        int cc[HB], ss[HB];
        // ---- phase 1: residual loads of the whole batch (one memory round trip per batch) ------------
#pragma unroll
        for (int j = 0; j < HB; ++j) {
            cc[j] = c00;
            ss[j] = s0;
            if (RES) {
                if (s0 != cur_s) enter_pixel(s0);
                if (!((NCH % 2 == 0) && (HB % 2 == 0)) || (j & 1) == 0) {
                    rh[j] = make_uint4(0, 0, 0, 0);
                    rl[j] = make_uint4(0, 0, 0, 0);
                    if ((NCH % 2 == 0) && (HB % 2 == 0)) {
                        rh[j + 1 < HB ? j + 1 : j] = make_uint4(0, 0, 0, 0);
                        rl[j + 1 < HB ? j + 1 : j] = make_uint4(0, 0, 0, 0);
                    }
                }
                if (ok && c00 < ep.res_act.Cp) {
                    if ((NCH % 2 == 0) && (HB % 2 == 0)) {
                        // chunk pairs: one 32-byte (full sector) load for the hi halves and one for the lo halves
                        if ((j & 1) == 0) {
                            uint32_t t8[8];
                            ld_global_nc_v8(rec_res + c00, t8);
                            rh[j] = make_uint4(t8[0], t8[1], t8[2], t8[3]);
                            rh[j + 1 < HB ? j + 1 : j] = make_uint4(t8[4], t8[5], t8[6], t8[7]);
                            ld_global_nc_v8(rec_res + ep.res_act.Cp + c00, t8);
                            rl[j] = make_uint4(t8[0], t8[1], t8[2], t8[3]);
                            rl[j + 1 < HB ? j + 1 : j] = make_uint4(t8[4], t8[5], t8[6], t8[7]);
                        }
                    } else {
                        rh[j] = __ldg(reinterpret_cast<const uint4*>(rec_res + c00));
                        rl[j] = __ldg(reinterpret_cast<const uint4*>(rec_res + ep.res_act.Cp + c00));
                    }
                }
            }
            c00 += 8;
            if (c00 >= N) { c00 -= N; ++s0; }
        }
        // ---- phase 2: bias, activation, residual, stores --------------------------------------------
        // (bias_s is zero beyond Cout and those accumulators are exact zeros: no per-element channel
        // masks; the activation is selected once per chunk, not per element).  Chunks are finished in
        // pairs so that ACT outputs go out as 32-byte (full-sector) stores.
        // N % 16 == 0 and an even chunk count per thread make the first chunk of every pair start at a
        // channel multiple of 16, so both chunks of a pair always lie in the same sub-tile / pixel.
        constexpr int PAIR = (NCH % 2 == 0) ? 2 : 1;
#pragma unroll
        for (int j0 = 0; j0 < HB; j0 += PAIR) {
            float v[PAIR * 8];
            if (ss[j0] != cur_s) enter_pixel(ss[j0]);
            if (!ok) continue;
#pragma unroll
            for (int jj = 0; jj < PAIR; ++jj) {
                const int j = j0 + jj;
                const int c0 = cc[j];
                float* w = v + jj * 8;
                {
                    const float4 b0 = *reinterpret_cast<const float4*>(bias_s + c0);
                    const float4 b1 = *reinterpret_cast<const float4*>(bias_s + c0 + 4);
                    const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                    for (int q = 0; q < 8; ++q) w[q] = fmaf(run[(h0 + j) * 8 + q], acc_scale, bb[q]);
                }
                if (act == FVC_ACT_RELU) {
#pragma unroll
                    for (int q = 0; q < 8; ++q) w[q] = fmaxf(w[q], 0.f);
                } else if (act == FVC_ACT_LRELU01) {
#pragma unroll
                    for (int q = 0; q < 8; ++q) w[q] = w[q] > 0.f ? w[q] : w[q] * 0.1f;
                } else if (!LEAN && act == FVC_ACT_EXP) {
#pragma unroll
                    for (int q = 0; q < 8; ++q) w[q] = expf(w[q]);
                } else if (!LEAN && act == FVC_ACT_LRELU001) {
#pragma unroll
                    for (int q = 0; q < 8; ++q) w[q] = w[q] > 0.f ? w[q] : w[q] * 0.01f;
                }
                if (RES) {
                    const uint32_t hh[4] = {rh[j].x, rh[j].y, rh[j].z, rh[j].w}, ll[4] = {rl[j].x, rl[j].y, rl[j].z, rl[j].w};
                    float r[8];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        float a0, a1, b0, b1;
                        e2f2(hh[q], a0, a1);
                        e2f2(ll[q], b0, b1);
                        r[2 * q] = a0 + b0;
                        r[2 * q + 1] = a1 + b1;
                    }
                    if (LEAN || ep.res_mode == 0) {
#pragma unroll
                        for (int q = 0; q < 8; ++q) w[q] += r[q];
                    } else if (ep.res_mode == 1) {   // GDN.py:88-93: x / sqrt(beta + gamma . x^2)
#pragma unroll
                        for (int q = 0; q < 8; ++q) w[q] = r[q] / sqrtf(w[q]);
                    } else {                         // IGDN: x * sqrt(.)
#pragma unroll
                        for (int q = 0; q < 8; ++q) w[q] = r[q] * sqrtf(w[q]);
                    }
                }
                if (!LEAN && ep.res_f32) {
#pragma unroll
                    for (int q = 0; q < 8; ++q)
                        if (c0 + q < Cout) w[q] += ep.res_f32[pixC + c0 + q];
                }
                if (!LEAN && ep.out_f32) {
                    if ((Cout & 7) == 0) {
                        // 8 channels = one 32-byte sector per store
                        if (c0 < Cout) st_global_v8(ep.out_f32 + pixC + c0, reinterpret_cast<const uint32_t*>(w));
                    } else if ((Cout & 3) == 0) {
#pragma unroll
                        for (int q = 0; q < 8; q += 4)
                            if (c0 + q < Cout)
                                *reinterpret_cast<float4*>(ep.out_f32 + pixC + c0 + q) =
                                    make_float4(w[q], w[q + 1], w[q + 2], w[q + 3]);
                    } else {
#pragma unroll
                        for (int q = 0; q < 8; ++q)
                            if (c0 + q < Cout) ep.out_f32[pixC + c0 + q] = w[q];
                    }
                }
            }
            const int c0 = cc[j0];
            if (PAIR == 2) {
                if (ep.out_act.p && c0 < ep.out_act.Cp) {
                    if (STG && stg) stage_store16(stg, srow, (c0 - cA) >> 3, v, satm, wlo);
                    else ep_store16_packed(rec_out, ep.out_act.Cp, c0, v, false, satm, wlo);
                }
                if (ep.out_act_relu.p && c0 < ep.out_act_relu.Cp) ep_store16_packed(rec_relu, ep.out_act_relu.Cp, c0, v, true, satm, wlo);
            } else {
                if (ep.out_act.p && c0 < ep.out_act.Cp) ep_store8_packed(rec_out, ep.out_act.Cp, c0, v, false, satm, wlo);
                if (ep.out_act_relu.p && c0 < ep.out_act_relu.Cp) ep_store8_packed(rec_relu, ep.out_act_relu.Cp, c0, v, true, satm, wlo);
            }
            if (!LEAN && ep.out_act_sq.p && c0 < ep.out_act_sq.Cp) {   // squares for the following (I)GDN
#pragma unroll
                for (int q = 0; q < PAIR * 8; ++q) v[q] = v[q] * v[q] * ep.sq_scale;
                if (PAIR == 2) ep_store16_packed(rec_sq, ep.out_act_sq.Cp, c0, v, false, satm, wlo);
                else ep_store8_packed(rec_sq, ep.out_act_sq.Cp, c0, v, false, satm, wlo);
            }
        }
    }
    if (ep.sat_count && ep_sat_hit(satm)) atomicAdd(ep.sat_count, 1u);
}

// Fused (I)GDN tile epilogue (GDN.py:63-93) of one accumulator thread: 32 output channels of one pixel (CT = 128).
//   1. x = acc * scale + bias (kept in `run`); x^2 * 2^-6 as fp16 hi/lo into the shared-memory staging tile, laid out as
//      the K-major SWIZZLE_128B A operand of an MMA: per sub-tile a hi and a lo block of [128 pixel rows][64 ch = 128 B];
//   2. all 16 accumulator warps meet; one thread issues, per sub-tile, the 12 MMAs
//      norm = sq_hi*g_hi + sq_hi*g_lo + sq_lo*g_hi (same tiles, same order, same single TMEM chain as the stand-alone
//      1x1 "norm" convolution this replaces: bit-identical), D in TMEM columns behind the two partial buffers;
//   3. every thread reads its 32 norm columns back, y = x / sqrt(beta + norm) (IGDN: x * sqrt(.)) with x rounded to
//      the 22-bit hi/lo record precision exactly as the unfused path read it back from memory, and stores y.
// Saves, per (I)GDN: the raw and the squared ACT tensors (written and read back) and one launch.
__device__ __forceinline__ void gdn_tile_epilogue(const TcParams& P, const float* __restrict__ bias_s,
                                                  const float* __restrict__ gbeta_s, float* run, int b, int sub, int ty,
                                                  int tx, int th, int tw, uint32_t colbase, uint32_t stgA, uint32_t gdnB,
                                                  uint32_t tmem_base, int quarter, uint32_t bar_gw, uint32_t bar_gdone,
                                                  uint32_t& gphase) {
    const Epilogue& ep = P.ep;
    const int N = P.N;                                    // 64
    const int s0 = (int)colbase / N, cA = (int)colbase - s0 * N;
    const int qy = ty * 16 + th, qx = (tx * P.SX + s0) * 8 + tw;
    const int oy = qy * P.os + P.sub[sub].py, ox = qx * P.os + P.sub[sub].px;
    const bool ok = qy < P.Hq && qx < P.Wq;
    const bool wlo = P.fast == 0;
    uint32_t satm = 0;
    // ---- 1. x in place, squares into the A-operand staging tile ------------------------------------------------
#pragma unroll
    for (int i = 0; i < 32; i += 4) {   // bias as 16-byte shared-memory loads (cA is a multiple of 32)
        const float4 bq = *reinterpret_cast<const float4*>(bias_s + cA + i);
        run[i] = fmaf(run[i], ep.acc_scale, bq.x);
        run[i + 1] = fmaf(run[i + 1], ep.acc_scale, bq.y);
        run[i + 2] = fmaf(run[i + 2], ep.acc_scale, bq.z);
        run[i + 3] = fmaf(run[i + 3], ep.acc_scale, bq.w);
    }
    {
        const int row = th * 8 + tw;
        const uint32_t sw = (uint32_t)(row & 7);
        const uint32_t blk_hi = stgA + (uint32_t)(s0 * 2) * 16384u + ((uint32_t)row << 7), blk_lo = blk_hi + 16384u;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            float sq[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) sq[q] = run[16 * j + q] * run[16 * j + q] * ep.sq_scale;
            uint32_t hi[8], lo[8];
            ep_pack8(sq, false, hi, lo, satm);
            ep_pack8(sq + 8, false, hi + 4, lo + 4, satm);
            const uint32_t ch = (uint32_t)(cA + 16 * j) >> 3;   // 16-byte chunk of the 128-byte row (even)
            st_shared_v4(blk_hi + ((ch ^ sw) << 4), hi[0], hi[1], hi[2], hi[3]);
            st_shared_v4(blk_hi + (((ch + 1u) ^ sw) << 4), hi[4], hi[5], hi[6], hi[7]);
            st_shared_v4(blk_lo + ((ch ^ sw) << 4), lo[0], lo[1], lo[2], lo[3]);
            st_shared_v4(blk_lo + (((ch + 1u) ^ sw) << 4), lo[4], lo[5], lo[6], lo[7]);
        }
    }
    // ---- 2. second MMA: norm = gamma . x^2 ----------------------------------------------------------------------
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    asm volatile("bar.sync 1, 512;" ::: "memory");
    if (threadIdx.x == TC_ACC_WARP0 * 32) {
        tc_fence_after();
        mbar_wait(bar_gw, 0);                             // gamma tiles resident (completes once per kernel)
        const uint64_t d0 = make_desc(0, 1024u, 2u);      // K-major SWIZZLE_128B, 8-row groups 1024 B apart
        for (int s = 0; s < P.S; ++s) {
            const uint32_t a_hi = stgA + (uint32_t)(s * 2) * 16384u, a_lo = a_hi + 16384u;
            const uint32_t dcol = tmem_base + (uint32_t)(P.nab * P.CT) + (uint32_t)(s * N);
            const uint32_t aa[3] = {a_hi, a_hi, a_lo};    // x tiles [g_hi][g_lo][g_hi] of the packed stream
#pragma unroll
            for (int t = 0; t < 3; ++t)
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    tc_mma(dcol, d0 | (uint64_t)(((aa[t] + 32u * k) & 0x3FFFFu) >> 4),
                           d0 | (uint64_t)(((gdnB + 8192u * t + 32u * k) & 0x3FFFFu) >> 4), P.idesc, (t | k) ? 1u : 0u);
        }
        tc_commit(bar_gdone);
    }
    mbar_wait(bar_gdone, gphase);
    gphase ^= 1u;
    tc_fence_after();
    // ---- 3. normalise and store ---------------------------------------------------------------------------------
    e16* rec_out = ok ? ep.out_act.p + act_pixel_offset(ep.out_act, b, oy, ox) : nullptr;
    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(P.nab * P.CT) + colbase;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        uint32_t d[16];
        tc_ld16(taddr + 16 * j, d);
        tc_wait_ld();
        float y[16], gb[16];
#pragma unroll
        for (int q = 0; q < 16; q += 4) {
            const float4 bq = *reinterpret_cast<const float4*>(gbeta_s + cA + 16 * j + q);
            gb[q] = bq.x; gb[q + 1] = bq.y; gb[q + 2] = bq.z; gb[q + 3] = bq.w;
        }
#pragma unroll
        for (int q = 0; q < 16; q += 2) {
            // x as the unfused path read it back from its hi/lo record
            const float x0 = run[16 * j + q], x1 = run[16 * j + q + 1];
            const uint32_t h = ep_pack2(x0, x1);
            float h0, h1, l0, l1;
            e2f2(h, h0, h1);
            e2f2(ep_pack2(x0 - h0, x1 - h1), l0, l1);
            const float r0 = h0 + l0, r1 = h1 + l1;
            const float n0 = fmaf(__uint_as_float(d[q]), ep.gdn_scale, gb[q]);
            const float n1 = fmaf(__uint_as_float(d[q + 1]), ep.gdn_scale, gb[q + 1]);
            y[q] = ep.gdn_inverse ? r0 * sqrtf(n0) : r0 / sqrtf(n0);
            y[q + 1] = ep.gdn_inverse ? r1 * sqrtf(n1) : r1 / sqrtf(n1);
        }
        if (ok) ep_store16_packed(rec_out, ep.out_act.Cp, cA + 16 * j, y, false, satm, wlo);
    }
    tc_fence_before();   // the norm columns are read: the next tile's MMAs may overwrite them after the next barrier
    if (ep.sat_count && ep_sat_hit(satm)) atomicAdd(ep.sat_count, 1u);
}

// Fused tail convolution, tile epilogue of one accumulator thread (32 output channels of one pixel, CT = 128):
//   1. y = act(acc * scale + bias) (+ skip), rounded to the hi/lo record exactly as the tail convolution would have read
//      it from memory, written as fp16 hi/lo into the A-operand staging tile: per sub-tile and 64-channel segment a hi
//      and a lo block of [128 pixel rows][128 B] (K-major SWIZZLE_128B);
//   2. all 16 accumulator warps meet; one thread issues per sub-tile y_hi*w_hi + y_hi*w_lo (per segment) + y_lo*w_hi
//      against the 32-row weight tiles (rows = (tap, channel) of the tail's 3x3 kernel), D in TMEM behind the partials;
//   3. every thread reads its share of the 32 columns and stores P (fp32) for its pixel.
template <bool RES>
__device__ __forceinline__ void tap_stage_issue(const TcParams& P, const float* __restrict__ bias_s, float* run, int b,
                                                int sub, int ty, int tx, int th, int tw, uint32_t colbase,
                                                uint32_t stgA, uint32_t tapB, uint32_t tmem_base, uint32_t bar_gw,
                                                uint32_t bar_gdone) {
    const Epilogue& ep = P.ep;
    const int N = P.N;                                    // 64 or 128 = channels of y per pixel
    const int nseg = N >> 6;                              // 64-channel segments
    const int s0 = (int)colbase / N, cA = (int)colbase - s0 * N;
    const int qy = ty * 16 + th, qx = (tx * P.SX + s0) * 8 + tw;
    const int oy = qy * P.os + P.sub[sub].py, ox = qx * P.os + P.sub[sub].px;
    const bool ok = qy < P.Hq && qx < P.Wq;
    uint32_t satm = 0;
    // ---- 1. y -> staging ------------------------------------------------------------------------------------------
    {
        const int row = th * 8 + tw;
        const uint32_t sw = (uint32_t)(row & 7);
        // blocks of sub-tile s0: [seg][hi, lo]; this thread's 32 channels lie in segment cA / 64
        const uint32_t blk_hi = stgA + (uint32_t)((s0 * nseg + (cA >> 6)) * 2) * 16384u + ((uint32_t)row << 7);
        const uint32_t blk_lo = blk_hi + 16384u;
        const e16* rec_res = nullptr;
        if (RES && ok) rec_res = ep.res_act.p + act_pixel_offset(ep.res_act, b, oy, ox);
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            float v[16];
            uint32_t rh[8], rl[8];
            if (RES) {
#pragma unroll
                for (int q = 0; q < 8; ++q) rh[q] = rl[q] = 0u;
                if (ok) {
                    ld_global_nc_v8(rec_res + cA + 16 * j, rh);
                    ld_global_nc_v8(rec_res + ep.res_act.Cp + cA + 16 * j, rl);
                }
            }
            // bias as four 16-byte shared-memory loads (cA is a multiple of 32); the activation is selected once per
            // 16 channels; LeakyReLU as max(w, 0.1 w): the same bits as the select, one instruction less
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
                const float4 bq = *reinterpret_cast<const float4*>(bias_s + cA + 16 * j + 4 * q4);
                v[4 * q4 + 0] = fmaf(run[16 * j + 4 * q4 + 0], ep.acc_scale, bq.x);
                v[4 * q4 + 1] = fmaf(run[16 * j + 4 * q4 + 1], ep.acc_scale, bq.y);
                v[4 * q4 + 2] = fmaf(run[16 * j + 4 * q4 + 2], ep.acc_scale, bq.z);
                v[4 * q4 + 3] = fmaf(run[16 * j + 4 * q4 + 3], ep.acc_scale, bq.w);
            }
            if (ep.act == FVC_ACT_RELU) {
#pragma unroll
                for (int q = 0; q < 16; ++q) v[q] = fmaxf(v[q], 0.f);
            } else if (ep.act == FVC_ACT_LRELU01) {
#pragma unroll
                for (int q = 0; q < 16; ++q) v[q] = fmaxf(v[q], v[q] * 0.1f);
            }
            if (RES) {
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    float a0, a1, b0, b1;
                    e2f2(rh[q], a0, a1);
                    e2f2(rl[q], b0, b1);
                    v[2 * q] += a0 + b0;
                    v[2 * q + 1] += a1 + b1;
                }
            }
            uint32_t hi[8], lo[8];
            ep_pack8(v, false, hi, lo, satm);
            ep_pack8(v + 8, false, hi + 4, lo + 4, satm);
            const uint32_t ch = (uint32_t)((cA & 63) + 16 * j) >> 3;
            st_shared_v4(blk_hi + ((ch ^ sw) << 4), hi[0], hi[1], hi[2], hi[3]);
            st_shared_v4(blk_hi + (((ch + 1u) ^ sw) << 4), hi[4], hi[5], hi[6], hi[7]);
            st_shared_v4(blk_lo + ((ch ^ sw) << 4), lo[0], lo[1], lo[2], lo[3]);
            st_shared_v4(blk_lo + (((ch + 1u) ^ sw) << 4), lo[4], lo[5], lo[6], lo[7]);
        }
    }
    // ---- 2. second MMA: P = y . W' ----------------------------------------------------------------------------------
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    asm volatile("bar.sync 1, 512;" ::: "memory");
    if (threadIdx.x == TC_ACC_WARP0 * 32) {
        tc_fence_after();
        mbar_wait(bar_gw, 0);                             // weight tiles resident (completes once per kernel)
        const uint64_t d0 = make_desc(0, 1024u, 2u);
        for (int s = 0; s < P.S; ++s) {
            const uint32_t dcol = tmem_base + (uint32_t)(P.nab * P.CT) + (uint32_t)(s * 32);
            uint32_t first = 0u;
            // stream order: per segment [w_hi][w_lo] (against y_hi), then per segment [w_hi] (against y_lo)
            for (int t = 0; t < 3 * nseg; ++t) {
                const int g = t < 2 * nseg ? (t >> 1) : (t - 2 * nseg);
                const uint32_t a = stgA + (uint32_t)((s * nseg + g) * 2 + (t < 2 * nseg ? 0 : 1)) * 16384u;
                const uint32_t w = tapB + 4096u * (uint32_t)t;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    tc_mma(dcol, d0 | (uint64_t)(((a + 32u * k) & 0x3FFFFu) >> 4),
                           d0 | (uint64_t)(((w + 32u * k) & 0x3FFFFu) >> 4), P.tap_idesc, first);
                    first = 1u;
                }
            }
        }
        tc_commit(bar_gdone);
    }
    if (ep.sat_count && ep_sat_hit(satm)) atomicAdd(ep.sat_count, 1u);
}

// Second half of the fused tail epilogue, one tile later (software pipeline: the accumulators of the next tile are
// drained while this tile's second MMA runs): wait for the MMA, read the thread's share of the 32 columns, store P.
__device__ __forceinline__ void tap_finish(const TcParams& P, int b, int sub, int ty, int tx, int th, int tw,
                                           uint32_t colbase, uint32_t tmem_base, int quarter, uint32_t bar_gdone,
                                           uint32_t& gphase) {
    const Epilogue& ep = P.ep;
    const int N = P.N;
    const int s0 = (int)colbase / N, cA = (int)colbase - s0 * N;
    const int qy = ty * 16 + th, qx = (tx * P.SX + s0) * 8 + tw;
    const int oy = qy * P.os + P.sub[sub].py, ox = qx * P.os + P.sub[sub].px;
    const bool ok = qy < P.Hq && qx < P.Wq;
    mbar_wait(bar_gdone, gphase);
    gphase ^= 1u;
    tc_fence_after();
    // ---- 3. P out: the sub-tile's 32 columns are shared by its N / 32 threads per pixel ------------------------------
    {
        const int share = 1024 / N;                        // 16 (N = 64) or 8 (N = 128) columns per thread
        const int c0 = (cA >> 5) * share;                  // first column of this thread
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(P.nab * P.CT) + (uint32_t)(s0 * 32 + c0);
        float* dst = ok ? ep.tap_out + (((size_t)b * P.Hout + oy) * P.Wout + ox) * (size_t)ep.tap_cq : nullptr;
        for (int q0 = 0; q0 < share; q0 += 8) {
            uint32_t d[8];
            tc_ld8(taddr + q0, d);
            tc_wait_ld();
            if (ok) {
#pragma unroll
                for (int q = 0; q < 8; q += 4)
                    if (c0 + q0 + q < ep.tap_cq)
                        *reinterpret_cast<float4*>(dst + c0 + q0 + q) =
                            make_float4(__uint_as_float(d[q]) * ep.tap_scale, __uint_as_float(d[q + 1]) * ep.tap_scale,
                                        __uint_as_float(d[q + 2]) * ep.tap_scale, __uint_as_float(d[q + 3]) * ep.tap_scale);
            }
        }
    }
    tc_fence_before();
}

// Warp roles: 0 = TMA producer (weight stream + patches), 1, 2 = MMA issuers (warp 1 owns the TMEM
// allocation), 3..18 = accumulator / epilogue warps (any 16 consecutive warps cover every TMEM lane
// quarter 4 times).
// NCH: 8-column chunks of the running sum each accumulator thread owns (CT/4 = 8*NCH);
// RES: the epilogue adds an ACT-format residual (ResBlock skip connection)
// MODE 1: TMA-store epilogue, MODE 2: fused (I)GDN epilogue (CT = 128 kernels only; separate instantiations so that the
// default kernels carry none of their code: compiled into the same kernel the staged stores cost the skip-connection
// variant 64 bytes of spills)
template <int NCH, bool RES, bool PAIR, int MODE = 0>
__global__ void __launch_bounds__(TC_THREADS, 1) k_conv_tc(const __grid_constant__ TcParams P) {
    constexpr bool STG = MODE == 1;    // TMA-store epilogue
    constexpr bool GDN = MODE == 2;    // fused (I)GDN epilogue
    constexpr bool TAP = MODE == 3;    // fused 3x3 tail convolution (tap-split second MMA)
    constexpr bool LEAN = MODE == 4;   // MODE 0 with the rare epilogue branches compiled out (tile_epilogue)
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment: SWIZZLE_128B atoms
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t patch0 = base;                                   // npb patch buffers
    const uint32_t stg0 = base + P.npb * P.patch_bytes;             // staging tile of the TMA-store epilogue (or empty)
    const uint32_t bst0 = stg0 + P.stg_bytes;                       // nst weight stages of T tiles
    const uint32_t bars = bst0 + P.nst * P.stage_bytes;             // mbarriers (8 B each)
    const uint32_t bar_pfull = bars, bar_pempty = bars + 16;        // [npb <= 2] each
    const uint32_t bar_bfull = bars + 32, bar_bempty = bars + 32 + 8 * 8;   // [nst <= 8] each
    const uint32_t bar_afull = bars + 32 + 16 * 8, bar_aempty = bar_afull + 32;  // partial buffers [<= 4] each
    const uint32_t tmem_slot = bar_aempty + 32;
    const uint32_t bar_gw = tmem_slot + 8, bar_gdone = tmem_slot + 16;   // fused GDN: gamma tiles loaded / norm MMAs done
    // fused GDN: three 8 KB gamma tiles after the A-operand staging tile (hi + lo block per sub-tile); fused tail
    // convolution: 3 * N/64 weight tiles of 4 KB after the staging tile (hi + lo block per sub-tile and segment)
    const uint32_t gdnB = stg0 + (TAP ? (uint32_t)P.S * (uint32_t)(P.N >> 6) * 32768u : (uint32_t)P.S * 32768u);
    float* bias_s = reinterpret_cast<float*>(smem_raw + (bars - smem_u32(smem_raw)) + 256);   // [N <= 128]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ntiles = P.B * P.nsub * P.tiles_y * P.tiles_x;   // PAIR: pairs of tiles (x-neighbours)
    // CTA pair: both CTAs walk the same (pair-tile, pass) sequence; rank 0 ("leader") issues every MMA for both
    const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
    const int tile0 = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int tstride = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    const int xmul = PAIR ? 2 : 1, xadd = (int)rank;

    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(bar_pfull + 8 * i, 1);
            mbar_init(bar_pempty + 8 * i, TC_ISSUERS);   // one commit per MMA issuer warp
        }
        for (int i = 0; i < 4; ++i) {
            mbar_init(bar_afull + 8 * i, TC_ISSUERS);
            mbar_init(bar_aempty + 8 * i, PAIR ? 32 : 16);  // one arrive per accumulator warp (of both CTAs)
        }
        for (int i = 0; i < P.nst; ++i) {
            mbar_init(bar_bfull + 8 * i, 1);
            mbar_init(bar_bempty + 8 * i, TC_ISSUERS);
        }
        if (GDN || TAP) {
            mbar_init(bar_gw, 1);
            mbar_init(bar_gdone, 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        if (PAIR) {   // executed by one warp of each CTA of the pair
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot),
                         "r"(P.tmem_cols)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot),
                         "r"(P.tmem_cols)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    if (threadIdx.x >= 128 && (int)threadIdx.x - 128 < P.N) {
        const int c = (int)threadIdx.x - 128;
        bias_s[c] = c < P.Cout ? P.ep.bias[c] : 0.f;
        if (GDN && c < 64) bias_s[128 + c] = P.ep.gdn_beta[c];   // beta_eff behind the (<= 128) bias values
    }
    tc_fence_before();
    __syncthreads();
    if (PAIR) cluster_sync_all();   // the peer's barriers are initialised before any remote signal
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
    // everything above (barriers, TMEM allocation, bias staging: parameters only) overlapped the previous kernel's
    // tail; from here on activations written by earlier kernels are read and buffers they read are overwritten
#ifndef FVC_NO_PDL_SYNC
    pdl_sync();
#endif

    // register re-allocation per warpgroup: the control warps need few registers, the accumulator warps many
    if (warp < TC_ACC_WARP0) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(TC_REGS_CTRL));
    if (warp == 0) {
        // ================= TMA producer: weight-stream ring + patch ring, polled by one thread ==========
        // Two independent iterators; whichever ring has a free slot is served (a blocking wait on one
        // ring would starve the other: the patch of the next pass frees only when the current pass ends).
        if (elect_one()) {
            if (GDN) {   // the (I)GDN's gamma tiles stay resident for the whole kernel
                mbar_expect_tx(bar_gw, 3u * 8192u);
                for (int t = 0; t < 3; ++t) tma_load_2d(gdnB + 8192u * (uint32_t)t, &P.mapG, bar_gw, 0, t * 64);
            }
            if (TAP) {   // regrouped weights of the fused tail convolution
                const int nt = 3 * (P.N >> 6);
                mbar_expect_tx(bar_gw, (uint32_t)nt * 4096u);
                for (int t = 0; t < nt; ++t) tma_load_2d(gdnB + 4096u * (uint32_t)t, &P.mapG, bar_gw, 0, t * 32);
            }
            PassIter wc, pc;
            wc.xmul = xmul; wc.xadd = xadd;
            wc.tile = tile0; wc.pass = 0; wc.decode(P);
            pc = wc;
            // PAIR: all loads complete on the LEADER's "full" barriers, which expect the bytes of both CTAs
            const uint32_t pfull_l = PAIR ? mapa_u32(bar_pfull, 0) : bar_pfull;
            const uint32_t bfull_l = PAIR ? mapa_u32(bar_bfull, 0) : bar_bfull;
            const int Nh = P.N >> 1;
            uint32_t st = 0, stph = 0;   // weight-stage ring position and its phase bit
            uint32_t set = 0, pph = 0;   // patch ring
            const uint32_t nst = (uint32_t)P.nst, npb = (uint32_t)P.npb;
            const int T = P.T;
            int wi = 0;                  // next weight tile of wc's pass
            uint32_t spins = 0;
            while (wc.valid(ntiles) || pc.valid(ntiles)) {
                bool progress = false;
                if (pc.valid(ntiles) && mbar_try_wait(bar_pempty + 8 * set, pph ^ 1u)) {
                    const TcPass& ps = P.pass[P.sub[pc.sub].pass_first + pc.pass];
                    const int x0 = pc.tx * 8 * P.SX + ps.ox, y0 = pc.ty * 16 + ps.oy;
                    if (PAIR) {
                        if (rank == 0) mbar_expect_tx(bar_pfull + 8 * set, 2u * P.patch_tx);
                        tma_load_5d_pair(patch0 + set * P.patch_bytes, &P.mapA, pfull_l + 8 * set, 0, ps.seg, x0, y0,
                                         pc.b * P.planes + ps.plane);
                    } else {
                        mbar_expect_tx(bar_pfull + 8 * set, P.patch_tx);
                        tma_load_5d(patch0 + set * P.patch_bytes, &P.mapA, bar_pfull + 8 * set, 0, ps.seg, x0, y0,
                                    pc.b * P.planes + ps.plane);
                    }
                    if (++set == npb) { set = 0; pph ^= 1u; }
                    pc.next(P, tstride);
                    progress = true;
                }
                if (wc.valid(ntiles) && mbar_try_wait(bar_bempty + 8 * st, stph ^ 1u)) {
                    const TcPass& ps = P.pass[P.sub[wc.sub].pass_first + wc.pass];
                    // T weight tiles per stage; the last stage of a pass over-reads (the stream is padded)
                    if (PAIR) {
                        // this CTA stages rows [rank * N/2, (rank + 1) * N/2) of each of the stage's T tiles
                        if (rank == 0) mbar_expect_tx(bar_bfull + 8 * st, 2u * P.stage_bytes);
                        for (int u = 0; u < T; ++u)
                            tma_load_2d_pair(bst0 + st * P.stage_bytes + (uint32_t)u * P.btile_bytes, &P.mapB, bfull_l + 8 * st, 0,
                                             (int)((ps.btile_first + (uint32_t)(wi + u)) * (uint32_t)P.N) + (int)rank * Nh);
                    } else {
                        mbar_expect_tx(bar_bfull + 8 * st, P.stage_bytes);
                        tma_load_2d(bst0 + st * P.stage_bytes, &P.mapB, bar_bfull + 8 * st, 0,
                                    (int)((ps.btile_first + (uint32_t)wi) * (uint32_t)P.N));
                    }
                    if (++st == nst) { st = 0; stph ^= 1u; }
                    wi += T;
                    if (wi >= ps.ntaps * ps.nbt) { wi = 0; wc.next(P, tstride); }
                    progress = true;
                }
                if (progress) spins = 0;
                else if (++spins > (1u << 25)) __trap();
            }
        }
    } else if (warp >= 1 && warp <= TC_ISSUERS && rank == 0) {
        // ================================ MMA issuers ============================================
        // Two warps run the same loop nest and issue the MMAs of the even / odd sub-tiles (disjoint
        // accumulator columns, so no ordering hazard): the tensor-pipe queue is shallow and one warp's
        // scalar work between MMA batches (descriptor arithmetic, barrier polls) is hidden by the other's
        // MMAs.  Every "done" barrier counts one commit per issuer warp.
        // The whole warp runs the loop nest (warp-uniform control flow and address arithmetic stay in
        // uniform registers: 1-2 scalar instructions per MMA); only the elected lane issues tcgen05 ops.
        {
            const bool lead = elect_one();
            const int s_first = warp - 1;
            const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);   // provably warp-uniform copy
            const bool dbg_on = P.dbg != nullptr && blockIdx.x == 0 && warp == 1;
            long long w_pfull = 0, w_aempty = 0, w_bfull = 0;
            const long long t_begin = dbg_on ? clock64() : 0;
            PassIter cur;
            cur.xmul = xmul; cur.xadd = xadd;
            cur.tile = tile0; cur.pass = 0; cur.decode(P);
            uint32_t gg = 0;                    // accumulation-group counter
            bool gopen = false;                 // a group (TMEM chain) is open; it may span segment passes
            uint32_t gpb = 0, gdcol = 0, gacc0 = 0;
            uint32_t set = 0, pph = 0;          // patch ring
            uint32_t st = 0, stph = 0;          // weight-stage ring position and its phase bit
            // descriptors: everything but the start address is fixed; per MMA only the low word moves
            const uint32_t layout = P.pitch == 32 ? 6u : 2u;
            const uint64_t adesc0 = make_desc(0, (uint32_t)(P.PW * P.pitch), layout);
            const uint64_t bdesc0 = make_desc(0, (uint32_t)(8 * P.pitch), layout);
            const uint32_t sstep = (uint32_t)(8 * P.pitch) >> 4;   // next sub-tile: 8 pixels further
            const uint32_t ahi = (uint32_t)(adesc0 >> 32), bhi = (uint32_t)(bdesc0 >> 32);
            const uint32_t lo0 = (uint32_t)adesc0;            // LBO field; the address bits are added below
            const int S = P.S, T = P.T;
            const uint32_t N = (uint32_t)P.N, nst = (uint32_t)P.nst, CT = (uint32_t)P.CT, npb = (uint32_t)P.npb;
            const uint32_t idesc = P.idesc, btile16 = P.btile_bytes >> 4, stage16 = P.stage_bytes >> 4;
            const uint32_t bst16 = lo0 + (bst0 >> 4);
            // sub-tiles owned by this issuer: s_first and s_first + TC_ISSUERS
            const int ns = S > s_first ? (S - 1 - s_first) / TC_ISSUERS + 1 : 0;
            const uint32_t so0 = (uint32_t)s_first * sstep, do0 = (uint32_t)s_first * N;
            const uint32_t so1 = (uint32_t)(s_first + TC_ISSUERS) * sstep, do1 = (uint32_t)(s_first + TC_ISSUERS) * N;
            const uint32_t nabs = (uint32_t)P.nab_log2, nabm = (1u << nabs) - 1u;
            // fused GDN / tail kernels may run a ring of three buffers; every other instantiation keeps the mask
            // arithmetic (and its exact code: ptxas' allocation of the skip-connection variant is fragile)
            const bool nab3 = (GDN || TAP) && P.nab == 3;
            while (cur.valid(ntiles)) {
                const TcSub& sb = P.sub[cur.sub];
                const TcPass& ps = P.pass[sb.pass_first + cur.pass];
                { const long long c0 = dbg_on ? clock64() : 0;
                  mbar_wait(bar_pfull + 8 * set, pph);
                  if (dbg_on) w_pfull += clock64() - c0; }
                const uint32_t pa16 = lo0 + ((patch0 + set * P.patch_bytes) >> 4);
                const int ntaps = ps.ntaps, gtaps = ps.gtaps;
                const bool two = ps.nbt == 2, short2 = ps.ks1 != 4;   // second tile per tap / with 2 k-steps
                const bool short1 = ps.ks0 == 2;                      // fast mode, 32-channel records: hi half only
                const bool narrow = ps.ks0 == 1;                      // 8-channel records: one k-step per tile
                const int32_t* toffp = P.tap_off + ps.tap_first;
                int slot = 0;                   // tile index inside the current weight stage
                uint32_t blo = 0;
                int t = 0;
                if (narrow) {
                    // 8-channel records: a tile is a single k-step, so whole weight stages (T/2 taps x 2 tiles x S
                    // sub-tiles) are issued per loop iteration; per-tile loop control would dominate otherwise
                    while (t < ntaps) {
                        const int t0 = t, t1 = min(t + gtaps, ntaps);
                        uint32_t pb = gg & nabm, rph = (gg >> nabs) & 1u;
                        if constexpr (GDN || TAP) {
                            if (nab3) { const uint32_t q3 = gg / 3u; pb = gg - 3u * q3; rph = q3 & 1u; }
                        }
                        { const long long c0 = dbg_on ? clock64() : 0;
                          mbar_wait(bar_aempty + 8 * pb, rph ^ 1u);
                          if (dbg_on) w_aempty += clock64() - c0; }
                        tc_fence_after();
                        const uint32_t dcol = tmem_u + pb * CT;
                        while (t < t1) {
                            if (slot == 0) {
                                const long long c0 = dbg_on ? clock64() : 0;
                                mbar_wait(bar_bfull + 8 * st, stph);
                                if (dbg_on) w_bfull += clock64() - c0;
                                blo = bst16 + st * stage16;
                            }
                            const int tpt = two ? 2 : 1;       // tiles per tap (1 with merged hi/lo rows)
                            const int nt = min((T - slot) >> (tpt - 1), t1 - t);   // tpt is 1 or 2: no division
                            if (lead) {
                                uint32_t toff = (uint32_t)toffp[t] >> 4;
                                for (int u = 0; u < nt; ++u) {
                                    const uint32_t alo = pa16 + toff;
                                    toff = (uint32_t)toffp[t + u + 1] >> 4;   // (no min(): it would leave the uniform datapath)
                                    const uint32_t b0 = blo + (uint32_t)(tpt * u) * btile16;
                                    issue_stage<1, PAIR>(dcol, alo, ahi, b0, bhi, idesc, (t + u == t0) ? 0u : 1u, ns, so0, do0, so1, do1);
                                    if (two) issue_stage<1, PAIR>(dcol, alo, ahi, b0 + btile16, bhi, idesc, 1u, ns, so0, do0, so1, do1);
                                }
                            }
                            t += nt;
                            slot += tpt * nt;
                            blo += (uint32_t)(tpt * nt) * btile16;
                            if (slot >= T || t == ntaps) {
                                if (lead) tc_commit_t<PAIR>(bar_bempty + 8 * st);
                                if (++st == nst) { st = 0; stph ^= 1u; }
                                slot = 0;
                            }
                        }
                        if (lead) tc_commit_t<PAIR>(bar_afull + 8 * pb);
                        ++gg;
                    }
                }
                uint32_t toff = (uint32_t)toffp[0] >> 4;
                const unsigned long long gmask = ((unsigned long long)ps.gmask_hi << 32) | ps.gmask_lo;
                // The tap loop, instantiated for the k-step count of a tap's first weight tile: 4, or 2 for the hi-only
                // half tile of 32-channel records in precision 'fast'.  (A run-time test of that count inside the loop
                // cost the issue-bound 3x3 layers 5 %: the issue loop's instruction count is what bounds them.)
                auto tap_loop = [&](auto ks0_tag) {
                    constexpr int KS0 = decltype(ks0_tag)::value;
                    for (; t < ntaps; ++t) {
                        if ((gmask >> t) & 1ull) {
                            // a new accumulation group starts here: hand the finished chain to the accumulator
                            // warps; the next partial buffer must have been drained by them
                            if (gopen) {
                                if (lead) tc_commit_t<PAIR>(bar_afull + 8 * gpb);
                                ++gg;
                            }
                            uint32_t rph;
                            if constexpr (GDN || TAP) {
                                if (nab3) { const uint32_t q3 = gg / 3u; gpb = gg - 3u * q3; rph = q3 & 1u; }
                                else { gpb = gg & nabm; rph = (gg >> nabs) & 1u; }
                            } else {
                                gpb = gg & nabm;
                                rph = (gg >> nabs) & 1u;
                            }
                            { const long long c0 = dbg_on ? clock64() : 0;
                              mbar_wait(bar_aempty + 8 * gpb, rph ^ 1u);
                              if (dbg_on) w_aempty += clock64() - c0; }
                            tc_fence_after();
                            gdcol = tmem_u + gpb * CT;
                            gacc0 = 0u;
                            gopen = true;
                        }
                        const uint32_t alo = pa16 + toff;
                        toff = (uint32_t)toffp[t + 1] >> 4;   // prefetched for the next tap (table has one spare entry)
    #pragma unroll
                        for (int j = 0; j < 2; ++j) {
                            if (j == 1 && !two) break;
                            if (slot == 0) {
                                const long long c0 = dbg_on ? clock64() : 0;
                                mbar_wait(bar_bfull + 8 * st, stph);
                                if (dbg_on) w_bfull += clock64() - c0;
                                blo = bst16 + st * stage16;
                            }
                            if (lead) {
                                // every weight tile has 4 k-steps except the [lo | 0] tile of 32-channel records (2)
                                if (j == 1 && short2) issue_stage<2, PAIR>(gdcol, alo, ahi, blo, bhi, idesc, gacc0, ns, so0, do0, so1, do1);
                                else issue_stage<KS0, PAIR>(gdcol, alo, ahi, blo, bhi, idesc, gacc0, ns, so0, do0, so1, do1);
                            }
                            gacc0 = 1u;
                            blo += btile16;
                            if (++slot == T || (t == ntaps - 1 && (j == 1 || !two))) {
                                if (lead) tc_commit_t<PAIR>(bar_bempty + 8 * st);   // frees the stage when these MMAs retire
                                if (++st == nst) { st = 0; stph ^= 1u; }
                                slot = 0;
                            }
                        }
                    }
                };
                if (short1) tap_loop(std::integral_constant<int, 2>{});
                else tap_loop(std::integral_constant<int, 4>{});
                if (ps.gend && gopen) {
                    if (lead) tc_commit_t<PAIR>(bar_afull + 8 * gpb);      // chain complete -> accumulator warps
                    ++gg;
                    gopen = false;
                }
                if (lead) tc_commit_t<PAIR>(bar_pempty + 8 * set);    // patch buffer reusable
                __syncwarp();
                if (++set == npb) { set = 0; pph ^= 1u; }
                cur.next(P, tstride);
            }
            if (dbg_on && lead) {
                P.dbg[0] = (unsigned long long)(clock64() - t_begin);
                P.dbg[1] = (unsigned long long)w_pfull;
                P.dbg[2] = (unsigned long long)w_aempty;
                P.dbg[3] = (unsigned long long)w_bfull;
            }
        }
    }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(TC_REGS_ACC));
        // ======================= accumulator / epilogue warps (16: warps 3..18) ====================
        const int quarter = warp & 3;                     // TMEM lane quarter this warp may access
        const int part = (warp - TC_ACC_WARP0) >> 2;                 // which quarter of the CT columns
        const int row = quarter * 32 + lane;              // accumulator row = pixel inside the sub-tile
        const int th = row >> 3, tw = row & 7;
        const uint32_t colbase = (uint32_t)part * (uint32_t)(NCH * 8);
        constexpr bool PARK = NCH >= 6;
        const uint32_t nabs = (uint32_t)P.nab_log2, nabm = (1u << nabs) - 1u;
        const bool nab3 = (GDN || TAP) && P.nab == 3;
        const uint32_t aempty_l = PAIR ? mapa_u32(bar_aempty, 0) : bar_aempty;   // the leader's issuers wait on it
        uint32_t gg = 0;
        float run[NCH * 8];
        // this warp's staging blocks (TMA-store epilogue: STG instantiations, CT = 128 only)
        const uint32_t stg = (STG && P.tmast) ? stg0 + (uint32_t)(warp - TC_ACC_WARP0) * 4096u : 0u;
        uint32_t gphase = 0;      // phase of the fused GDN's "norm MMAs done" barrier
        bool tap_pending = false; // fused tail: a tile whose second MMA is in flight (finished one tile later)
        int pb_b = 0, pb_sub = 0, pb_ty = 0, pb_tx = 0;
#ifdef FVC_TC_ACCDBG
        const bool adbg = P.dbg != nullptr && blockIdx.x == 0 && warp == TC_ACC_WARP0;
        long long a_wait = 0, a_drain = 0, a_epi = 0;
#endif
        for (int tile = tile0; tile < ntiles; tile += tstride) {
            const int ng = P.sub[(tile + tile / P.nsub) % P.nsub].ngroups;
            if (RES) {
                // pull this thread's skip-connection records into L2 while the tile's MMAs run (the epilogue's
                // loads then hit L2 instead of paying DRAM latency in the middle of the store stream)
                int t = tile;
                const int sub = (t + t / P.nsub) % P.nsub; t /= P.nsub;
                const int tx = (t % P.tiles_x) * xmul + xadd; t /= P.tiles_x;
                const int ty = t % P.tiles_y;
                const int b = t / P.tiles_y;
                const int Nv = P.merged ? (P.N >> 1) : P.N;
                const int cb = P.merged ? (int)(colbase >> 1) : (int)colbase;
                const int nc = P.merged ? NCH * 4 : NCH * 8;
                const int qy = ty * 16 + th;
                if (qy < P.Hq) {
                    for (int s = cb / Nv; s <= (cb + nc - 1) / Nv; ++s) {
                        const int qx = (tx * P.SX + s) * 8 + tw;
                        if (qx >= P.Wq) continue;
                        const e16* rec = P.ep.res_act.p + act_pixel_offset(P.ep.res_act, b, qy * P.os + P.sub[sub].py,
                                                                         qx * P.os + P.sub[sub].px);
                        const int c0 = max(cb - s * Nv, 0);
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(rec + c0));
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(rec + P.ep.res_act.Cp + c0));
                    }
                }
            }
            for (int g = 0; g < ng; ++g, ++gg) {
                uint32_t pb, rph;
                if constexpr (GDN || TAP) {
                    if (nab3) { const uint32_t q3 = gg / 3u; pb = gg - 3u * q3; rph = q3 & 1u; }
                    else { pb = gg & nabm; rph = (gg >> nabs) & 1u; }
                } else {
                    pb = gg & nabm;
                    rph = (gg >> nabs) & 1u;
                }
#ifdef FVC_TC_ACCDBG
                const long long c0 = adbg ? clock64() : 0;
#endif
                mbar_wait_sleep(bar_afull + 8 * pb, rph, P.acc_sleep_ns);
                tc_fence_after();
#ifdef FVC_TC_ACCDBG
                const long long c1 = adbg ? clock64() : 0;
                a_wait += c1 - c0;
#endif
                const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + pb * P.CT + colbase;
#pragma unroll
                for (int i = 0; i < NCH; ++i) {
                    if constexpr (GDN || TAP) {
                        // first chain of a tile: straight into the running sums (these kernels' tiles are one to three
                        // chains; the generic path's copy costs 32 moves per thread and tile)
                        if (g == 0) {
                            tc_ld8(taddr + i * 8, reinterpret_cast<uint32_t*>(run + i * 8));
                            tc_wait_ld();
                            continue;
                        }
                    }
                    float v[8];
                    // At most two TMEM loads in flight: the address of load i carries a (zero) dependency on
                    // the sums of chunk i-2.  Unconstrained, ptxas issues 4+ loads back to back and then
                    // spills the 64 running sums around every group (seen in SASS).
                    uint32_t dep = 0;
                    if (NCH > 4 && i >= 2) dep = __float_as_uint(run[(i - 2) * 8]) & P.zero;
                    tc_ld8(taddr + i * 8 + dep, reinterpret_cast<uint32_t*>(v));
                    tc_wait_ld();
                    if (g == 0) {
#pragma unroll
                        for (int q = 0; q < 8; ++q) run[i * 8 + q] = v[q];
                    } else {
#pragma unroll
                        for (int q = 0; q < 8; ++q) run[i * 8 + q] += v[q];   // fp32 round-to-nearest
                    }
                }
                if (!PARK || g + 1 < ng) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) {
                        if (PAIR) mbar_arrive_cluster(aempty_l + 8 * pb);
                        else mbar_arrive(bar_aempty + 8 * pb);
                    }
                }
#ifdef FVC_TC_ACCDBG
                if (adbg) a_drain += clock64() - c1;
#endif
            }
#ifdef FVC_TC_ACCDBG
            const long long e0 = adbg ? clock64() : 0;
#endif
            // tile coordinates are decoded only now (laundered through an empty asm) so that the epilogue's
            // address arithmetic cannot be hoisted above the drain loop, where it would spill `run`
            int t = tile;
            asm volatile("" : "+r"(t));
            const int sub = (t + t / P.nsub) % P.nsub; t /= P.nsub;
            const int tx = (t % P.tiles_x) * xmul + xadd; t /= P.tiles_x;
            const int ty = t % P.tiles_y;
            const int b = t / P.tiles_y;
            if (STG && stg) {
                // the warp's staging blocks are free once its previous bulk stores have READ them (a tile ago)
                if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                __syncwarp();
            }
            if constexpr (GDN) {
                gdn_tile_epilogue(P, bias_s, bias_s + 128, run, b, sub, ty, tx, th, tw, colbase, stg0, gdnB, tmem_base,
                                  quarter, bar_gw, bar_gdone, gphase);
            } else if constexpr (TAP) {
                // software pipeline over tiles: this tile's accumulators are already drained (above) while the previous
                // tile's second MMA was running; now finish the previous tile, then stage and issue this one
                if (tap_pending) tap_finish(P, pb_b, pb_sub, pb_ty, pb_tx, th, tw, colbase, tmem_base, quarter, bar_gdone, gphase);
                tap_stage_issue<RES>(P, bias_s, run, b, sub, ty, tx, th, tw, colbase, stg0, gdnB, tmem_base, bar_gw, bar_gdone);
                tap_pending = true; pb_b = b; pb_sub = sub; pb_ty = ty; pb_tx = tx;
            } else if constexpr (!PARK) {
                if constexpr (NCH % 2 == 0) {
                    if (P.merged) {
                        // columns come in 8-wide blocks [w_hi products | w_lo products]: fold them
#pragma unroll
                        for (int i = 0; i < NCH / 2; ++i)
#pragma unroll
                            for (int q = 0; q < 8; ++q) run[i * 8 + q] = run[2 * i * 8 + q] + run[(2 * i + 1) * 8 + q];
                        tile_epilogue<NCH / 2, RES, false, LEAN>(P, bias_s, run, b, sub, ty, tx, th, tw, colbase >> 1, stg);
                    } else {
                        tile_epilogue<NCH, RES, STG, LEAN>(P, bias_s, run, b, sub, ty, tx, th, tw, colbase, stg);
                    }
                } else {
                    tile_epilogue<NCH, RES, false, LEAN>(P, bias_s, run, b, sub, ty, tx, th, tw, colbase, stg);
                }
            } else {
                // 48/64 running sums + the epilogue state do not fit 96 registers (ncu: spill reloads
                // queued behind the epilogue's global stores were its main stall).  Park the upper half of
                // the sums in the partial buffer that was just drained (it stays ours until we arrive on
                // its "empty" barrier), finish the lower half from registers, then fetch the rest back.
                constexpr int NH = ((NCH / 2) + 1) & ~1;
                const uint32_t pb = (gg - 1u) & nabm;
                const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + pb * P.CT + colbase;
#pragma unroll
                for (int i = NH; i < NCH; ++i) tc_st8(taddr + i * 8, reinterpret_cast<const uint32_t*>(run + i * 8));
                tc_wait_st();
                tile_epilogue<NH, RES, false, LEAN>(P, bias_s, run, b, sub, ty, tx, th, tw, colbase, stg);
#pragma unroll
                for (int i = NH; i < NCH; ++i) {
                    tc_ld8(taddr + i * 8, reinterpret_cast<uint32_t*>(run + (i - NH) * 8));
                    tc_wait_ld();
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if (PAIR) mbar_arrive_cluster(aempty_l + 8 * pb);
                    else mbar_arrive(bar_aempty + 8 * pb);
                }
                tile_epilogue<NCH - NH, RES, false, LEAN>(P, bias_s, run, b, sub, ty, tx, th, tw, colbase + NH * 8, stg);
            }
            if (STG && stg) {
                // generic-proxy writes of the staging blocks -> visible to the async proxy; then one lane issues the
                // warp's stores: hi and lo piece of its 4 x 8 pixels (x 4 planes for parity-planar outputs)
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) {
                    const int Nv = P.N;
                    const int sI = (int)colbase / Nv, pc = ((int)colbase - sI * Nv) >> 5;    // sub-tile, 32-channel piece
                    const int qx0 = (tx * P.SX + sI) * 8, qy0 = ty * 16 + quarter * 4;
                    if (qx0 < P.Wq && qy0 < P.Hq) {
                        const int nhalf = P.fast ? 1 : 2;
                        for (int hl = 0; hl < nhalf; ++hl) {
                            const uint32_t src = stg + (uint32_t)hl * 2048u;
                            const int piece = pc + hl * P.o_nseg;
                            if (P.o_mode == 0) {
                                tma_store_5d(&P.mapO, src, 0, piece, qx0, qy0, b);
                            } else if (P.o_mode == 1) {
#pragma unroll
                                for (int pl = 0; pl < 4; ++pl)
                                    tma_store_5d(&P.mapO, src + (uint32_t)pl * 512u, 0, piece, qx0 >> 1, qy0 >> 1, b * 4 + pl);
                            } else {
                                tma_store_5d(&P.mapO, src, 0, piece, qx0 * 2 + P.sub[sub].px, qy0 * 2 + P.sub[sub].py, b);
                            }
                        }
                    }
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
            }
#ifdef FVC_TC_ACCDBG
            if (adbg) a_epi += clock64() - e0;
#endif
        }

        if (TAP && tap_pending) tap_finish(P, pb_b, pb_sub, pb_ty, pb_tx, th, tw, colbase, tmem_base, quarter, bar_gdone, gphase);
        if (STG && stg && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
#ifdef FVC_TC_ACCDBG
        if (adbg && lane == 0) { P.dbg[4] = (unsigned long long)a_wait; P.dbg[5] = (unsigned long long)a_drain; P.dbg[6] = (unsigned long long)a_epi; }
#endif
    }
    tc_fence_before();
    __syncthreads();
    if (PAIR) cluster_sync_all();   // neither CTA may retire while the other still reads its shared memory / signals it
    if (warp == 1) {
        __syncwarp();
        tc_fence_after();
        if (PAIR)
            asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(P.tmem_cols)
                         : "memory");
        else
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(P.tmem_cols)
                         : "memory");
    }
}

// ----------------------------------------------------------------------------------------------
// weight stream packing
// ----------------------------------------------------------------------------------------------
struct BTileInfo {
    int8_t r, s;       // kernel tap
    uint8_t kind;      // 0: hi(c0..c0+63)  1: lo(c0..c0+63)  2: [hi(0..31) | hi(0..31)]  3: [lo(0..31) | 0]
                       // 4: [hi(0..7) | hi(0..7)]  5: [lo(0..7) | 0]   (16-element rows)
                       // 6, 7, 8: merged hi/lo row blocks (see k_tc_pack)
    uint8_t c0;
};
__global__ void k_absmax(const float* __restrict__ w, size_t n, float* __restrict__ out) {
    __shared__ float red[32];
    float m = 0.f;
    for (size_t i = threadIdx.x; i < n; i += blockDim.x) m = fmaxf(m, fabsf(w[i]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x < 32) {
        m = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (threadIdx.x == 0) *out = m;
    }
}

__global__ void k_tc_pack(const float* __restrict__ w, e16* __restrict__ out, const BTileInfo* __restrict__ info,
                          int ntiles, int N, int Cin, int Cout, int k, int transposed, float wscale, int rowlen) {
    // rowlen: 16-bit elements per weight-tile row (64 = 128-byte rows; 16 = 32-byte rows of the 8-channel records)
    size_t n = (size_t)ntiles * N * rowlen;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int col = (int)(i % rowlen);
    int row = (int)((i / rowlen) % N);
    int tile = (int)(i / ((size_t)N * rowlen));
    BTileInfo bi = info[tile];
    int ci;
    bool lo;
    bool zero = false;
    if (bi.kind <= 1) {
        ci = bi.c0 + col;
        lo = bi.kind == 1;
    } else if (bi.kind == 2) {
        ci = col & 31;
        lo = false;
    } else if (bi.kind == 3) {
        ci = col & 31;
        lo = true;
        zero = col >= 32;
    } else if (bi.kind == 4) {   // [hi(0..7) | hi(0..7)] against [a_hi | a_lo]
        ci = col & 7;
        lo = false;
    } else if (bi.kind == 5) {   // [lo(0..7) | 0]
        ci = col & 7;
        lo = true;
        zero = col >= 8;
    } else {
        // merged tiles (MMA N = 32 for <= 16 output channels): 8-row blocks alternate w_hi / w_lo rows
        const bool lo_block = (row >> 3) & 1;
        row = ((row >> 4) << 3) | (row & 7);          // output channel of this row
        lo = lo_block;
        if (bi.kind == 9) {                           // narrow records: A = [a_hi(0..7) | a_lo(0..7)]
            ci = col & 7;
            zero = lo_block && col >= 8;
        } else if (bi.kind == 6) {                    // A = a_hi(c0..c0+63): w_hi and w_lo rows
            ci = bi.c0 + col;
        } else if (bi.kind == 7) {                    // A = a_lo(c0..c0+63): w_hi rows only
            ci = bi.c0 + col;
            zero = lo_block;
        } else {                                      // 8: A = [a_hi(0..31) | a_lo(0..31)]
            ci = col & 31;
            zero = lo_block && col >= 32;
        }
    }
    float v = 0.f;
    if (!zero && ci < Cin && row < Cout) {
        v = transposed ? w[(((size_t)ci * Cout + row) * k + bi.r) * k + bi.s]
                       : w[(((size_t)row * Cin + ci) * k + bi.r) * k + bi.s];
    }
    v *= wscale;  // power of two: exact
    e16 hi = f2e(v);
    out[i] = lo ? f2e(v - e2f(hi)) : hi;
}

// ----------------------------------------------------------------------------------------------
// host: plan
// ----------------------------------------------------------------------------------------------
struct TcPlan {
    bool lean = false;     // the layer qualifies for the LEAN instantiation (k_conv_tc<.., 4>)
    unsigned long long* dbg = nullptr;
    TcParams P;
    e16* wstream = nullptr;
    size_t smem = 0;
    int grid = 0;
};

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled get_encode() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (PFN_encodeTiled)p;
    }
    return fn;
}

bool tc_supported(const ConvLayer& L, int CinP) {
    if (!(CinP == 8 || CinP == 32 || CinP == 64 || CinP == 128)) return false;
    if (L.k > 7 || L.Cout > 128) return false;
    return true;
}

static int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return v ? atoi(v) : dflt;
}

int tc_plan_create(const ConvLayer& L, const float* w_ref, ActT in, int Hout, int Wout, const Epilogue& ep,
                   TcPlan** out, cudaStream_t s, bool fast, bool no_merge) {
    FVC_ARG(tc_supported(L, in.Cp));
    const bool gdn = ep.gdn_beta != nullptr;   // (I)GDN fused into this convolution's epilogue
    if (gdn) {
        FVC_ARG(ep.gdn_gamma && L.Cout == 64 && ep.out_act.p && ep.out_act.Cp == 64 && !ep.res_act.p && !ep.res_f32 &&
                !ep.out_f32 && !ep.out_act_relu.p && !ep.out_act_sq.p && ep.act == FVC_ACT_NONE);
    }
    FVC_ARG((L.st == 2) == (in.parity != 0));
    PFN_encodeTiled encode = get_encode();
    if (!encode) {
        set_error("cuTensorMapEncodeTiled is not available from the driver");
        return FVC_ERR_CUDA;
    }
    TcPlan* plan = new TcPlan();
    TcParams& P = plan->P;
    memset(&P, 0, sizeof(P));
    const int Cp = in.Cp;
    // output channels covered by the MMA: every channel of an ACT output record must be written
    int chans = L.Cout;
    if (ep.out_act.p) chans = std::max(chans, ep.out_act.Cp);
    if (ep.out_act_relu.p) chans = std::max(chans, ep.out_act_relu.Cp);
    if (ep.out_act_sq.p) chans = std::max(chans, ep.out_act_sq.Cp);
    // Cout <= 16 ("merged"): padded channels of the output records are never written (the buffers are
    // zero-initialised) and the MMA N = 32 carries w_hi and w_lo row blocks side by side.
    const int merge_max = env_int("FVC_TC_MERGED", 32);
    const bool tap = ep.tap_w != nullptr;      // fused tail convolution (second MMA against regrouped weights)
    if (tap) {
        FVC_ARG(!gdn && ep.tap_out && (L.Cout == 64 || L.Cout == 128) && ep.tap_cq >= 4 && ep.tap_cq <= 32 &&
                (ep.tap_cq & 3) == 0 && !ep.out_act.p && !ep.out_f32 && !ep.res_f32 && !ep.out_act_relu.p &&
                !ep.out_act_sq.p && (!ep.res_act.p || ep.res_mode == 0));
    }
    const bool merged = !fast && !gdn && !no_merge && ((Cp >= 32 && L.Cout <= merge_max) || (Cp == 8 && L.Cout <= env_int("FVC_TC_MERGED_NARROW", 32)));
    if (merged) chans = cdiv(L.Cout, 16) * 16;
    const int N = merged ? 2 * chans : std::max(16, cdiv(chans, 16) * 16);   // MMA N
    P.N = N;
    P.merged = merged ? 1 : 0;
    P.fast = fast ? 1 : 0;
    P.nchunks = 0;
    P.Cout = L.Cout;
    P.nsub = L.nsub;
    P.os = L.os;
    P.Hout = Hout; P.Wout = Wout;
    P.Hq = Hout / L.os; P.Wq = Wout / L.os;
    P.B = in.B;
    P.planes = (L.st == 2) ? 4 : 1;
    P.ep = ep;

    // ---- passes: (segment, plane) x taps ---------------------------------------------------------
    struct HostTap { int ey, ex, r, s; };
    std::vector<BTileInfo> binfo;
    int npass = 0, ntapent = 0;
    int gymin = 127, gymax = -127, gxmin = 127, gxmax = -127;   // extents over all (sub, plane) groups
    struct Group { int sub, plane; std::vector<HostTap> taps; int eymin, exmin, eymax, exmax; int tap_first; };
    std::vector<Group> groups;
    for (int si = 0; si < L.nsub; ++si) {
        const SubConv& S = L.sub[si];
        for (int pl = 0; pl < P.planes; ++pl) {
            Group g;
            g.sub = si; g.plane = pl; g.tap_first = 0;
            g.eymin = g.exmin = 127; g.eymax = g.exmax = -127;
            for (int t = 0; t < S.ntaps; ++t) {
                int dy = S.dy[t], dx = S.dx[t];
                int ey = dy, ex = dx;
                if (L.st == 2) {
                    int py = dy & 1, px = dx & 1;
                    if (py * 2 + px != pl) continue;
                    ey = (dy - py) / 2; ex = (dx - px) / 2;
                }
                g.taps.push_back({ey, ex, S.r[t], S.s[t]});
                g.eymin = std::min(g.eymin, ey); g.eymax = std::max(g.eymax, ey);
                g.exmin = std::min(g.exmin, ex); g.exmax = std::max(g.exmax, ex);
            }
            if (g.taps.empty()) continue;
            gymin = std::min(gymin, g.eymin); gymax = std::max(gymax, g.eymax);
            gxmin = std::min(gxmin, g.exmin); gxmax = std::max(gxmax, g.exmax);
            groups.push_back(g);
        }
    }
    int max_ext_y = 0, max_ext_x = 0;
    for (auto& g : groups) {
        max_ext_y = std::max(max_ext_y, g.eymax - g.eymin);
        max_ext_x = std::max(max_ext_x, g.exmax - g.exmin);
    }
    // ---- tile shape -------------------------------------------------------------------------------
    const int pw_align = env_int("FVC_TC_PW_ALIGN", 1);
    const int smem_cap = 232448 - 1024 /*alignment*/ - 1024 /*barriers + bias*/ - env_int("FVC_TC_SMEM_RESERVE", 0);
    const int pitch = Cp == 8 ? 32 : 128;
    P.pitch = pitch;
    // CTA pairs (cta_group::2): two x-neighbouring tiles are one M = 256 MMA; each CTA stages half of every weight
    // tile, which halves the weight bytes staged through shared memory and the B-operand reads per SM.
    // Measured at 1080p (tools/layer_ab.py): pairs win or tie on every layer.
    // (the fused GDN epilogue issues cta_group::1 MMAs of its own: such a kernel cannot mix in cta_group::2)
    const bool pair = N % 16 == 0 && env_int("FVC_TC_PAIR", 1) != 0 && !gdn && !tap;
    P.pair = pair ? 1 : 0;
    const int tile_bytes = N * pitch / (pair ? 2 : 1);   // per CTA
    // TMA-store epilogue (FVC_TC_TMAST: 0 off [default], 1 on for every eligible layer): each accumulator warp stages
    // its 32 pixels x 32 channels in shared memory and writes them with bulk tensor stores.  Eligible: an ACT output whose
    // record width equals the MMA N (every channel computed), not merged, CT = 128 (each thread owns 32 channels of one
    // sub-tile), outputs of stride-2 transposed convolutions only when not parity-planar.
    const int tmast_env = env_int("FVC_TC_TMAST", 0);
    int o_mode = 0;
    bool tmast = tmast_env != 0 && ep.out_act.p != nullptr && !merged && ep.out_act.Cp == N && N >= 32 && N <= 128;
    if (tmast) {
        if (L.os == 2) {
            o_mode = 2;
            if (ep.out_act.parity) tmast = false;
        } else if (ep.out_act.parity) {
            o_mode = 1;
            if ((Hout & 1) || (Wout & 1)) tmast = false;
        }
    }
    int o_nseg = tmast ? ep.out_act.Cp / 32 : 0;
    int SX = 0, nst = 0, PW = 0, PH = 16 + max_ext_y, npb = 2, T = 1;
    // S sub-tiles of 128 pixels share every weight stage; CT = S*N accumulator columns per partial buffer,
    // CT/32 in {1,2,3,4,6,8} (template instantiations), CT <= 256.  Two patch buffers when that still leaves room
    // for a deep weight ring; otherwise one (big-halo 7x7 layers: the exposed patch load is a few % of a pass, a
    // starved weight ring costs far more).
    // Measured per layer class at 1080p (chain 48, tools/layer_ab.py): CT = 128 with FOUR partial buffers wins for
    // every N >= 64 class (N = 128: S = 1, N = 64: S = 2): the MMA issuers can then be a whole tile ahead of an
    // epilogue, and each issuer warp owns exactly one sub-tile.  The 7x7 N = 64 layer (SpyNet conv2) keeps S = 3
    // (one sub-tile per issuer, two partial buffers), the N <= 32 layers S = 4.
    const int ct_cap = N > 96 ? env_int("FVC_TC_CTMAX128", 128)
                              : ((N > 32 && L.k < 7) ? env_int("FVC_TC_CTMAX64", 128)
                                                     : env_int("FVC_TC_CTMAX", N > 32 ? 192 : 256));   // 7x7 N=64: S = 3,
                                                                                                        // one sub-tile per issuer
    const int sx_max = std::min(env_int("FVC_TC_SX", 4),
                                std::max(1, std::min(merged ? 128 : ct_cap, ep.res_act.p ? env_int("FVC_TC_CTRES", 128) : 256) / N));
    int tmax = std::max(1, std::min(env_int("FVC_TC_T", 8), (pair ? env_int("FVC_TC_TPAIR", 512) : 256) / N));   // <= 32 KB per stage and CTA
    // layers with a skip-connection / GDN operand: stages of 2 weight tiles (measured at 1080p, tools/layer_ab.py,
    // profiles/r02_layer_times_T.txt: ResBlock conv2 0.508 -> 0.454 ms, 0.517 -> 0.471 ms; every other class prefers 8)
    if (ep.res_act.p) tmax = std::min(tmax, std::max(1, env_int("FVC_TC_T_RES", 2)));
    // S trades weight re-reads / per-tile overhead (cost ~ one sub-tile's worth per tile, calibrated on
    // SpyNet level 0/1) against filling the 148 SMs: score = wave efficiency * S / (S + 1).
    int sms = 148;
    {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    }
    for (int attempt = 0; attempt < 2 && !SX; ++attempt) {
    if (attempt == 1) {
        if (!tmast) break;
        tmast = false;      // no shape leaves room for the staging tile: per-lane stores for this layer
    }
    double best_eff = -1.0;
    for (int sx = std::max(1, sx_max); sx >= 1; --sx) {
        const int ct32 = sx * N / 32;
        if ((sx * N) % 32 != 0 || !(ct32 == 1 || ct32 == 2 || ct32 == 3 || ct32 == 4 || ct32 == 6 || ct32 == 8)) continue;
        if (merged && (ct32 & 1)) continue; // a thread must own both column blocks (hi, lo) of its channel chunks
        if ((tmast || gdn || tap) && ct32 != 4) continue;   // staging: every thread owns 32 channels of one sub-tile
        // TMA stores: 16 accumulator warps x (2 KB hi + 2 KB lo); fused GDN: A-operand tile (hi + lo block per sub-tile)
        // + three 8 KB gamma tiles
        // fused tail: A-operand tile (hi + lo block per sub-tile and 64-channel segment) + 3 * N/64 weight tiles of 4 KB
        const long stage_need = tmast ? 16L * 4096L
                                      : (gdn ? (long)sx * 32768L + 24576L
                                             : (tap ? (long)sx * (N / 64) * 32768L + 3L * (N / 64) * 4096L : 0L));
        int pw = 8 * sx + max_ext_x;
        pw = cdiv(pw, pw_align) * pw_align;
        size_t patch = (size_t)PH * pw * pitch;
        patch = (patch + 1023) & ~(size_t)1023;
        const long nt = (long)in.B * L.nsub * cdiv(P.Hq, 16) * (pair ? cdiv(cdiv(P.Wq, 8 * sx), 2) : cdiv(P.Wq, 8 * sx));
        const int units = pair ? sms / 2 : sms;
        const double eff = (double)nt / (double)(cdiv64(nt, units) * units) * (double)sx / (double)(sx + 1);
        if (eff <= best_eff + 0.02) continue;   // a smaller S must buy a real gain
        bool fits = false;
        for (int nb = 2; nb >= 1 && !fits; --nb) {
            const long wroom = (long)smem_cap - (long)nb * (long)patch - stage_need;
            const long want = nb == 2 ? std::min<long>(48 * 1024, 6L * tile_bytes) : 2L * tile_bytes;
            if (wroom < want) continue;
            int t = (int)std::max<long>(1, std::min<long>(tmax, wroom / (4L * tile_bytes)));
            if (pitch == 32 && !merged) t = std::max(2, t & ~1);   // narrow records: stages hold whole taps (2 tiles each)
            int st = (int)std::min<long>(8, wroom / ((long)t * tile_bytes));
            if (st < 2) continue;
            SX = sx; nst = st; PW = pw; npb = nb; T = t;
            fits = true;
        }
        if (!fits) continue;
        best_eff = eff;
    }
    }
    if (!SX) {
        set_error("tc_plan_create: no tile shape fits shared memory");
        delete plan;
        return FVC_ERR_STATE;
    }
    P.S = SX; P.SX = SX; P.PW = PW; P.PH = PH; P.nst = nst; P.npb = npb; P.T = T;
    if (!tmast) o_nseg = 0;
    P.tmast = tmast ? 1 : 0;
    P.o_mode = o_mode;
    P.o_nseg = o_nseg;
    P.stg_bytes = tmast ? 16u * 4096u : (gdn ? (uint32_t)SX * 32768u + 24576u : 0u);
    if (tap) P.stg_bytes = (((uint32_t)SX * (uint32_t)(N / 64) * 32768u + 3u * (uint32_t)(N / 64) * 4096u) + 1023u) & ~1023u;
    P.gdn = gdn ? 1 : 0;
    P.tap = tap ? 1 : 0;
    P.CT = SX * N;
    P.patch_bytes = (uint32_t)((((size_t)PH * PW * pitch) + 1023) & ~(size_t)1023);
    P.patch_tx = (uint32_t)((size_t)PH * PW * pitch);
    P.btile_bytes = (uint32_t)tile_bytes;             // per CTA (half of the tile's rows in pair mode)
    P.stage_bytes = (uint32_t)T * P.btile_bytes;
    // partial-accumulator buffers: 4 when they fit the 512 TMEM columns (the MMA issuers then run up to a whole
    // tile ahead of an epilogue), else 2
    P.nab_log2 = (4 * P.CT <= 512 && env_int("FVC_TC_NAB", 4) >= 4 && !gdn && !tap) ? 2 : 1;
    P.nab = 1 << P.nab_log2;
    // fused GDN / tail kernels: a third buffer when it fits beside the second MMA's accumulators (a 3x3 tile of 64-channel
    // records is three chains: with two buffers the issuers stall on the tile's own epilogue)
    if ((gdn || tap) && 3 * P.CT + (gdn ? SX * 64 : SX * 32) <= 512 && env_int("FVC_TC_NAB3", 1) != 0) P.nab = 3;
    uint32_t cols = 32;
    // fused GDN / tail: the second MMA's accumulators (S x 64 / S x 32 columns) live behind the two partial buffers
    while (cols < (uint32_t)(P.nab * P.CT + (gdn ? SX * 64 : (tap ? SX * 32 : 0)))) cols <<= 1;
    P.tmem_cols = cols;
    P.tiles_x = pair ? cdiv(cdiv(P.Wq, 8 * SX), 2) : cdiv(P.Wq, 8 * SX);
    P.tiles_y = cdiv(P.Hq, 16);
    // instruction descriptor: D=f32, A=B=f16 (0) or bf16 (1), K-major both, N, M=128
    const uint32_t fmt = FVC_SPLIT_FP16 ? 0u : 1u;
    P.idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) |
              ((uint32_t)((pair ? 256 : 128) >> 4) << 24);

    for (auto& g : groups) {
        g.tap_first = ntapent;
        if (ntapent + (int)g.taps.size() > TC_MAX_TAPS) {
            set_error("tap table overflow");
            delete plan;
            return FVC_ERR_STATE;
        }
        for (auto& tp : g.taps) P.tap_off[ntapent++] = ((tp.ey - g.eymin) * PW + (tp.ex - g.exmin)) * pitch;
    }
    // ---- segment passes -----------------------------------------------------------------------------
    struct SegPass { int seg, nbt, ks0, ks1, kind0, kind1, c0; };
    std::vector<SegPass> segp;
    if (fast) {
        // hi halves only.  Narrow records: the single K = 16 step [a_hi | a_lo] x [w_hi | w_hi] (a_lo * w_hi rides along
        // for free); 32-channel records: the first two k-steps of the [hi | hi] tile, i.e. a_hi * w_hi.
        if (Cp == 8) segp.push_back({0, 1, 1, 0, 4, 0, 0});
        else if (Cp == 32) segp.push_back({0, 1, 2, 0, 2, 0, 0});
        else if (Cp == 64) segp.push_back({0, 1, 4, 0, 0, 0, 0});
        else { segp.push_back({0, 1, 4, 0, 0, 0, 0}); segp.push_back({1, 1, 4, 0, 0, 0, 64}); }
    } else if (merged) {
        if (Cp == 8) segp.push_back({0, 1, 1, 0, 9, 0, 0});
        else if (Cp == 32) segp.push_back({0, 1, 4, 0, 8, 0, 0});
        else if (Cp == 64) { segp.push_back({0, 1, 4, 0, 6, 0, 0}); segp.push_back({1, 1, 4, 0, 7, 0, 0}); }
        else {
            segp.push_back({0, 1, 4, 0, 6, 0, 0}); segp.push_back({1, 1, 4, 0, 6, 0, 64});
            segp.push_back({2, 1, 4, 0, 7, 0, 0}); segp.push_back({3, 1, 4, 0, 7, 0, 64});
        }
    } else if (Cp == 8) segp.push_back({0, 2, 1, 1, 4, 5, 0});
    else if (Cp == 32) segp.push_back({0, 2, 4, 2, 2, 3, 0});
    else if (Cp == 64) { segp.push_back({0, 2, 4, 4, 0, 1, 0}); segp.push_back({1, 1, 4, 0, 0, 0, 0}); }
    else {
        segp.push_back({0, 2, 4, 4, 0, 1, 0}); segp.push_back({1, 2, 4, 4, 0, 1, 64});
        segp.push_back({2, 1, 4, 0, 0, 0, 0}); segp.push_back({3, 1, 4, 0, 0, 0, 64});
    }
    for (int si = 0; si < L.nsub; ++si) {
        TcSub& sb = P.sub[si];
        sb.pass_first = npass;
        sb.py = L.sub[si].py; sb.px = L.sub[si].px;
        sb.btile_first = (uint32_t)binfo.size();
        // tap tables per group of this sub
        for (auto& sp : segp) {
            for (auto& g : groups) {
                if (g.sub != si) continue;
                if (npass >= TC_MAX_PASS) { set_error("too many passes"); delete plan; return FVC_ERR_STATE; }
                TcPass& ps = P.pass[npass++];
                ps.seg = (int8_t)sp.seg; ps.plane = (int8_t)g.plane;
                ps.oy = (int8_t)g.eymin; ps.ox = (int8_t)g.exmin;
                ps.ntaps = (int16_t)g.taps.size();
                ps.nbt = (int8_t)sp.nbt; ps.ks0 = (int8_t)sp.ks0; ps.ks1 = (int8_t)sp.ks1;
                {
                    // taps per accumulation group: keep every TMEM chain <= chain_max MMAs
                    const int per_tap = sp.ks0 + (sp.nbt == 2 ? sp.ks1 : 0);
                    const int chain_max = fast ? env_int("FVC_TC_CHAIN_FAST", 4096) : env_int("FVC_TC_CHAIN", 48);
                    ps.gtaps = (int8_t)std::max(1, std::min(127, chain_max / per_tap));
                }
                ps.tap_first = (int16_t)g.tap_first;
                ps.btile_first = (uint32_t)binfo.size();
                for (auto& tp : g.taps) {
                    binfo.push_back({(int8_t)tp.r, (int8_t)tp.s, (uint8_t)sp.kind0, (uint8_t)sp.c0});
                    if (sp.nbt == 2) binfo.push_back({(int8_t)tp.r, (int8_t)tp.s, (uint8_t)sp.kind1, (uint8_t)sp.c0});
                }
            }
        }
        sb.npass = npass - sb.pass_first;
        sb.ngroups = 0;
        {
            const int chain_max = fast ? env_int("FVC_TC_CHAIN_FAST", 4096) : env_int("FVC_TC_CHAIN", 48);
            const bool span = env_int("FVC_TC_SPAN", 1) != 0;
            int chain = 0;
            for (int q = sb.pass_first; q < npass; ++q) {
                TcPass& ps = P.pass[q];
                const int per_tap = ps.ks0 + (ps.nbt == 2 ? ps.ks1 : 0);
                unsigned long long mask = 0;
                const bool narrow_pass = ps.ks0 == 1;   // the narrow-record loop keeps per-pass groups of gtaps taps
                for (int t = 0; t < ps.ntaps; ++t) {
                    bool start;
                    if (narrow_pass || !span) start = (t % ps.gtaps) == 0;
                    else start = (q == sb.pass_first && t == 0) || chain + per_tap > chain_max;
                    if (start) { mask |= 1ull << t; chain = 0; sb.ngroups++; }
                    chain += per_tap;
                }
                ps.gmask_lo = (uint32_t)mask;
                ps.gmask_hi = (uint32_t)(mask >> 32);
                ps.gend = (narrow_pass || !span || q == npass - 1) ? 1 : 0;
                if (ps.gend) chain = 0;
            }
        }
    }

    // ---- weight stream -------------------------------------------------------------------------------
    const int nbt_total = (int)binfo.size();
    // padded by T tiles: the last weight stage of a pass always loads a full T-tile box
    const int rowlen = pitch / 2;
    size_t wbytes = (size_t)(nbt_total + T) * N * rowlen * sizeof(e16);
    BTileInfo* dinfo = nullptr;
    cudaError_t ce = cudaMalloc(&plan->wstream, wbytes);
    if (ce == cudaSuccess) ce = cudaMemsetAsync(plan->wstream, 0, wbytes, s);
    if (ce == cudaSuccess) ce = cudaMalloc(&dinfo, binfo.size() * sizeof(BTileInfo));
    if (ce == cudaSuccess)
        ce = cudaMemcpyAsync(dinfo, binfo.data(), binfo.size() * sizeof(BTileInfo), cudaMemcpyHostToDevice, s);
    if (ce != cudaSuccess) {
        delete plan;
        return cuda_fail(ce, "tc weight stream alloc", __FILE__, __LINE__);
    }
    {
        // power-of-two weight scale: in half mode it lifts hi AND lo of the weights into the normal
        // range (max |w| -> [2^11, 2^12)); the epilogue multiplies the accumulators by 1/scale.
        float wscale = 1.f;
#if FVC_SPLIT_FP16
        float* dmax = nullptr;
        float hmax = 0.f;
        ce = cudaMalloc(&dmax, 4);
        if (ce == cudaSuccess) {
            k_absmax<<<1, 1024, 0, s>>>(w_ref, (size_t)L.Cin * L.Cout * L.k * L.k, dmax);
            g_launch_count++;
            ce = cudaMemcpyAsync(&hmax, dmax, 4, cudaMemcpyDeviceToHost, s);
            if (ce == cudaSuccess) ce = cudaStreamSynchronize(s);
            cudaFree(dmax);
        }
        if (ce != cudaSuccess) {
            cudaFree(dinfo);
            cudaFree(plan->wstream);
            delete plan;
            return cuda_fail(ce, "k_absmax", __FILE__, __LINE__);
        }
        if (hmax > 0.f && std::isfinite(hmax)) {
            int e = 0;
            std::frexp(hmax, &e);              // hmax = m * 2^e, m in [0.5, 1)
            wscale = std::ldexp(1.f, 12 - e);  // max |w| * scale in [2^11, 2^12)
        }
#endif
        P.ep.acc_scale = ep.acc_scale / wscale;
        size_t n = (size_t)nbt_total * N * rowlen;
        k_tc_pack<<<(unsigned)cdiv64((int64_t)n, 256), 256, 0, s>>>(w_ref, plan->wstream, dinfo, nbt_total, N, L.Cin,
                                                                    L.Cout, L.k, L.transposed, wscale, rowlen);
        g_launch_count++;
        ce = cudaGetLastError();
        if (ce == cudaSuccess) ce = cudaStreamSynchronize(s);   // binfo (host vector) must outlive the copy
        cudaFree(dinfo);
        if (ce != cudaSuccess) {
            cudaFree(plan->wstream);
            delete plan;
            return cuda_fail(ce, "k_tc_pack", __FILE__, __LINE__);
        }
    }

    // ---- tensor maps -----------------------------------------------------------------------------------
    {
        const cuuint64_t rec = (cuuint64_t)Cp * 4;   // bytes per pixel record
        const int nseg = std::max(1, Cp / 32);
        const int Wd = in.parity ? in.W / 2 : in.W, Hd = in.parity ? in.H / 2 : in.H;
        const CUtensorMapSwizzle swz = pitch == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_128B;
        cuuint64_t dims[5] = {(cuuint64_t)rowlen, (cuuint64_t)nseg, (cuuint64_t)Wd, (cuuint64_t)Hd,
                              (cuuint64_t)(in.B * P.planes)};
        cuuint64_t strides[4] = {(cuuint64_t)pitch, rec, rec * Wd, rec * Wd * Hd};
        cuuint32_t box[5] = {(cuuint32_t)rowlen, 1, (cuuint32_t)PW, (cuuint32_t)PH, 1};
        cuuint32_t estr[5] = {1, 1, 1, 1, 1};
        CUresult r = encode(&P.mapA, (FVC_SPLIT_FP16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16), 5, (void*)in.p, dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            set_error("cuTensorMapEncodeTiled(A) failed: %d (Cp=%d W=%d H=%d PW=%d PH=%d)", (int)r, Cp, Wd, Hd, PW, PH);
            cudaFree(plan->wstream);
            delete plan;
            return FVC_ERR_CUDA;
        }
        cuuint64_t bdims[2] = {(cuuint64_t)rowlen, (cuuint64_t)(nbt_total + T) * N};
        cuuint64_t bstr[1] = {(cuuint64_t)pitch};
        cuuint32_t bbox[2] = {(cuuint32_t)rowlen, (cuuint32_t)(pair ? N / 2 : T * N)};
        cuuint32_t bes[2] = {1, 1};
        r = encode(&P.mapB, (FVC_SPLIT_FP16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16), 2, (void*)plan->wstream, bdims, bstr, bbox, bes,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            set_error("cuTensorMapEncodeTiled(B) failed: %d", (int)r);
            cudaFree(plan->wstream);
            delete plan;
            return FVC_ERR_CUDA;
        }
    }
    if (gdn) {
        cuuint64_t gd[2] = {64, 3 * 64};
        cuuint64_t gs[1] = {128};
        cuuint32_t gb[2] = {64, 64};
        cuuint32_t ge[2] = {1, 1};
        CUresult r = encode(&P.mapG, (FVC_SPLIT_FP16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16), 2,
                            (void*)ep.gdn_gamma, gd, gs, gb, ge, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            set_error("cuTensorMapEncodeTiled(gamma) failed: %d", (int)r);
            cudaFree(plan->wstream);
            delete plan;
            return FVC_ERR_CUDA;
        }
    }
    if (tap) {
        const int nt = 3 * (N / 64);
        cuuint64_t gd[2] = {64, (cuuint64_t)(nt * 32)};
        cuuint64_t gs[1] = {128};
        cuuint32_t gb[2] = {64, 32};
        cuuint32_t ge[2] = {1, 1};
        CUresult r = encode(&P.mapG, (FVC_SPLIT_FP16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16), 2,
                            (void*)ep.tap_w, gd, gs, gb, ge, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            set_error("cuTensorMapEncodeTiled(tail weights) failed: %d", (int)r);
            cudaFree(plan->wstream);
            delete plan;
            return FVC_ERR_CUDA;
        }
        const uint32_t fmt = FVC_SPLIT_FP16 ? 0u : 1u;
        P.tap_idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(32 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    }
    if (tmast) {
        const ActT& o = ep.out_act;
        const cuuint64_t rec = (cuuint64_t)o.Cp * 4;
        const int Wd = o.parity ? o.W / 2 : o.W, Hd = o.parity ? o.H / 2 : o.H;
        // records as 64-byte pieces: [hi: Cp/32 pieces][lo: Cp/32 pieces]
        cuuint64_t dims[5] = {32, (cuuint64_t)(2 * o_nseg), (cuuint64_t)Wd, (cuuint64_t)Hd, (cuuint64_t)(o.B * (o.parity ? 4 : 1))};
        cuuint64_t strides[4] = {64, rec, rec * Wd, rec * Wd * Hd};
        cuuint32_t box[5] = {32, 1, 8, 4, 1};                        // a warp's 4 x 8 pixels
        cuuint32_t estr[5] = {1, 1, 1, 1, 1};
        if (o_mode == 1) { box[2] = 4; box[3] = 2; }
        if (o_mode == 2) { box[2] = 16; box[3] = 8; estr[2] = 2; estr[3] = 2; }   // 8 x 4 pixels at stride 2
        CUresult r = encode(&P.mapO, (FVC_SPLIT_FP16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16), 5,
                            (void*)o.p, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                            CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            set_error("cuTensorMapEncodeTiled(O) failed: %d (Cp=%d W=%d H=%d mode=%d)", (int)r, o.Cp, Wd, Hd, o_mode);
            cudaFree(plan->wstream);
            delete plan;
            return FVC_ERR_CUDA;
        }
    }
    P.acc_sleep_ns = (uint32_t)env_int("FVC_TC_ACC_SLEEP", 100);
    if (env_int("FVC_TC_DEBUG", 0)) {
        if (cudaMalloc(&plan->dbg, 64) == cudaSuccess) cudaMemset(plan->dbg, 0, 64);
        P.dbg = plan->dbg;
    }
    plan->lean = env_int("FVC_TC_LEAN", 1) != 0 && !gdn && !tap && !tmast && ep.out_act.p && !ep.out_f32 && !ep.res_f32 &&
                 !ep.out_act_sq.p && (ep.act == FVC_ACT_NONE || ep.act == FVC_ACT_RELU || ep.act == FVC_ACT_LRELU01) &&
                 (!ep.res_act.p || ep.res_mode == 0);
    plan->smem = 1024 + (size_t)npb * P.patch_bytes + (size_t)P.stg_bytes + (size_t)nst * P.stage_bytes + 1024;
    int ntiles = P.B * P.nsub * P.tiles_y * P.tiles_x;
    plan->grid = pair ? 2 * std::max(1, std::min(ntiles, sms / 2)) : std::max(1, std::min(ntiles, sms));
    *out = plan;
    return 0;
}

template <int NCH, bool RES, bool PAIR, int MODE = 0>
static int tc_launch_t3(TcPlan* plan, cudaStream_t s) {
    // the attribute is per device (and this function may run on several host threads): one bit per device ordinal
    static std::atomic<unsigned long long> attr_set{0};
    int dev = 0;
    FVC_CUDA(cudaGetDevice(&dev));
    const unsigned long long bit = 1ull << (dev & 63);
    if (!(attr_set.load(std::memory_order_acquire) & bit)) {
        FVC_CUDA(cudaFuncSetAttribute(k_conv_tc<NCH, RES, PAIR, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
        attr_set.fetch_or(bit, std::memory_order_release);
    }
    {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)plan->grid);
        cfg.blockDim = dim3(TC_THREADS);
        cfg.dynamicSmemBytes = plan->smem;
        cfg.stream = s;
        cudaLaunchAttribute at[2];
        int na = 0;
        if (PAIR) {
            at[na].id = cudaLaunchAttributeClusterDimension;   // CTA pair = one cluster (same TPC)
            at[na].val.clusterDim.x = 2; at[na].val.clusterDim.y = 1; at[na].val.clusterDim.z = 1;
            ++na;
        }
        if (pdl_enabled()) {
            at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // see pdl_sync()
            at[na].val.programmaticStreamSerializationAllowed = 1;
            ++na;
        }
        cfg.attrs = at;
        cfg.numAttrs = na;
        FVC_CUDA(cudaLaunchKernelEx(&cfg, k_conv_tc<NCH, RES, PAIR, MODE>, plan->P));
    }
    g_launch_count++;
    FVC_CHECK_LAUNCH();
    if (plan->dbg) {   // debugging aid: where block 0's MMA issuer spent its cycles
        unsigned long long h[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        FVC_CUDA(cudaStreamSynchronize(s));
        FVC_CUDA(cudaMemcpy(h, plan->dbg, sizeof(h), cudaMemcpyDeviceToHost));
        const TcParams& Q = plan->P;
        fprintf(stderr, "tcdbg pair=%d N=%d S=%d T=%d nst=%d npb=%d PW=%d PH=%d Cout=%d Hout=%d total=%llu pfull=%llu aempty=%llu bfull=%llu | acc wait=%llu drain=%llu epi=%llu res=%d ntiles=%d\n",
                Q.pair, Q.N, Q.S, Q.T, Q.nst, Q.npb, Q.PW, Q.PH, Q.Cout, Q.Hout, h[0], h[1], h[2], h[3], h[4], h[5], h[6],
                Q.ep.res_act.p ? 1 : 0, Q.B * Q.nsub * Q.tiles_y * Q.tiles_x);
    }
    return 0;
}

template <int NCH, bool RES>
static int tc_launch_t2(TcPlan* plan, cudaStream_t s) {
    if constexpr (NCH == 4) {
        if (plan->P.tmast)
            return plan->P.pair ? tc_launch_t3<NCH, RES, true, 1>(plan, s) : tc_launch_t3<NCH, RES, false, 1>(plan, s);
        if constexpr (!RES) {
            if (plan->P.gdn) return tc_launch_t3<NCH, false, false, 2>(plan, s);
        }
        if (plan->P.tap) return tc_launch_t3<NCH, RES, false, 3>(plan, s);
    }
    if constexpr (NCH >= 2 && NCH <= 6) {   // the accumulator widths the frame's lean layers use (compile time: 16 kernels)
        if (plan->lean) return plan->P.pair ? tc_launch_t3<NCH, RES, true, 4>(plan, s) : tc_launch_t3<NCH, RES, false, 4>(plan, s);
    }
    return plan->P.pair ? tc_launch_t3<NCH, RES, true>(plan, s) : tc_launch_t3<NCH, RES, false>(plan, s);
}

template <int NCH>
static int tc_launch_t(TcPlan* plan, cudaStream_t s) {
    return plan->P.ep.res_act.p ? tc_launch_t2<NCH, true>(plan, s) : tc_launch_t2<NCH, false>(plan, s);
}

int tc_plan_launch(TcPlan* plan, cudaStream_t s) {
    FVC_ARG(plan != nullptr);
    switch (plan->P.CT / 32) {
        case 1: return tc_launch_t<1>(plan, s);
        case 2: return tc_launch_t<2>(plan, s);
        case 3: return tc_launch_t<3>(plan, s);
        case 4: return tc_launch_t<4>(plan, s);
        case 6: return tc_launch_t<6>(plan, s);
        case 8: return tc_launch_t<8>(plan, s);
    }
    set_error("tc_plan_launch: unsupported accumulator width %d", plan->P.CT);
    return FVC_ERR_STATE;
}

// packed weight stream and accumulator scale of a plan (the fused GDN epilogue of the producing convolution uses the
// stand-alone norm convolution's gamma stream: tiles [hi][lo][hi] of 64 rows x 128 B)
const e16* tc_plan_wstream(const TcPlan* plan) { return plan ? plan->wstream : nullptr; }
float tc_plan_acc_scale(const TcPlan* plan) { return plan ? plan->P.ep.acc_scale : 0.f; }
// layout the fused tail convolution expects of its regrouped 1x1 weights: N = 32 rows, not merged, 128-byte rows
bool tc_plan_is_tap_layout(const TcPlan* plan) {
    return plan && !plan->P.merged && plan->P.N == 32 && plan->P.pitch == 128 && plan->P.fast == 0;
}
bool tc_plan_is_gdn_norm_layout(const TcPlan* plan) {
    return plan && !plan->P.merged && plan->P.N == 64 && plan->P.pitch == 128 && plan->P.fast == 0;
}

void tc_plan_destroy(TcPlan* plan) {
    if (!plan) return;
    if (plan->wstream) cudaFree(plan->wstream);
    if (plan->dbg) cudaFree(plan->dbg);
    delete plan;
}

}  // namespace fvc
