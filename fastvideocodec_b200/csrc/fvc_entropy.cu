// Real entropy coding of the three quantised latents on the GPU: the `calrealbits` branch of
// VideoCompressor.forward, reference DVC/net.py:123-138 (feature, Laplace(0, sigma)), 155-168 (z, BitEstimator),
// 183-195 (mv, BitEstimator).
//
// The reference builds, per element, a float CDF table with 2*mxrange entries cdf[i] = F(i - mxrange - 0.5), hands it
// to torchac (un-vendored dependency), which converts it to 16-bit integers
//     Q[i] = round(cdf[i] * (2^16 - (Lp - 1))) + i      (Lp = 2*mxrange; the "+ i" makes every symbol codable)
// with the end of the last symbol pinned to 2^16, range-codes the symbols s = q + mxrange and counts
// len(byte_stream) * 8 as the real bits.  Here:
//   * the model is the same integer CDF, evaluated ON THE FLY for the coded symbol only (start = Q[s],
//     freq = Q[s+1] - Q[s]): no [n, 300] table ever reaches HBM (that is a 300x expansion: 1.1 GB per 1080p frame).
//     BitEstimator CDFs do not depend on the position: one [C, 2R] table per forward, built by one small kernel;
//     the Laplace CDF is closed form in sigma.  k_cdf_table_* materialise full tables only for the parity tests;
//   * the coder is rANS (32-bit state, 16-bit renormalisation, probabilities in 1/2^16) instead of torchac's
//     arithmetic coder: same model, code length within 32 bits per lane of the same ideal sum(-log2(freq / 2^16)),
//     and it parallelises: the symbols are cut into lanes of L consecutive symbols (default 8192), every lane is an
//     independent rANS stream coded by one thread; phase 1 (model evaluation, the expensive transcendental part)
//     is fully parallel, phase 2 (the serial state update) is integer-only.
//   * container: "FVR1" | n | L | nlanes | u16 words-per-lane[nlanes] (padded to 4 bytes) | lanes back to back, each
//     lane = final state (2 words, high first) followed by its renormalisation words in DECODING order.
// Symbols are taken in the engine's NHWC order (pixel-major, channel fastest).
#include <cstring>

#include "fvc_kernels.cuh"
#include "fvc_bits.cuh"

namespace fvc {

#define RANS_L (1u << 16)
#define FVR_MAGIC 0x31525646u   // "FVR1"

// ----------------------------------------------------------------------------------------------
// integer CDF model
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t torchac_quant(float cdf, int i, int Lp) {
    // torchac._convert_to_int_and_normalize: round(cdf * (2^16 - (Lp - 1))) + arange(Lp), 16-bit wrap-around
    const float scaled = __fmul_rn(cdf, (float)(65536 - (Lp - 1)));
    return ((uint32_t)(int)rintf(scaled) + (uint32_t)i) & 0xffffu;
}
// Laplace(0, sigma): Q[i] for table index i (value v = i - R); shared by the encoder and the decoder, never inlined,
// so both evaluate the very same instruction sequence (a differently contracted FMA would desynchronise them)
__device__ __noinline__ uint32_t laplace_q(int i, int R, float sigma) {
    const float v = (float)(i - R) - 0.5f;
    return torchac_quant(laplace_cdf(v, sigma), i, 2 * R);
}
// start and end of symbol s under Q (the end of the last symbol is 2^16: torchac's max_symbol rule)
__device__ __forceinline__ void laplace_interval(int s, int R, float sigma, uint32_t& start, uint32_t& end) {
    start = laplace_q(s, R, sigma);
    end = (s == 2 * R - 2) ? 65536u : laplace_q(s + 1, R, sigma);
}

// T[c][i], i in [0, 2R): Q[i] for i < 2R-1, 2^16 at i = 2R-1; made strictly increasing (float evaluation of the
// saturated CDF tails may wobble by one unit).  One block per channel.
__global__ void k_cdf_table_factorized(FactorizedParams prm, int C, int R, uint32_t* __restrict__ table) {
    pdl_sync();
    const int c = blockIdx.x;
    const int Lp = 2 * R;
    __shared__ ChanParams cp;
    if (threadIdx.x == 0) cp = make_chan_params(prm, c);
    __syncthreads();
    uint32_t* T = table + (size_t)c * Lp;
    for (int i = threadIdx.x; i < Lp; i += blockDim.x) {
        const float v = (float)(i - R) - 0.5f;
        T[i] = (i == Lp - 1) ? 65536u : torchac_quant(factorized_cdf(v, cp), i, Lp);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < Lp - 1; ++i)
            if (T[i] <= T[i - 1]) T[i] = T[i - 1] + 1;
        // (cannot run into 2^16: Q[i] <= 65536 - (Lp - 1) + i)
    }
}

// full per-element Laplace tables [n][2R] (parity tests only: this is the 300x expansion the coder avoids)
__global__ void k_cdf_table_laplace(const float* __restrict__ sigma, int64_t n, int R, uint32_t* __restrict__ table) {
    const int Lp = 2 * R;
    const int64_t total = n * Lp;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (int64_t)gridDim.x * blockDim.x) {
        const int i = (int)(k % Lp);
        const float sg = fminf(fmaxf(sigma[k / Lp], 1e-5f), 1e10f);   // net.py:142
        table[k] = (i == Lp - 1) ? 65536u : laplace_q(i, R, sg);
    }
}

// ----------------------------------------------------------------------------------------------
// phase 1: symbol -> two words: start | freq << 16 (freq <= 65536 - (Lp - 2) fits 16 bits), floor((2^32 - 1) / freq)
// err[0] counts symbols outside [-R, R-2] (torchac check_input_bounds would raise), err[1] empty intervals
// ----------------------------------------------------------------------------------------------
// Exact x / freq, x % freq for x < 2^32, 1 <= freq <= 2^16 from m = floor((2^32 - 1) / freq) = 2^32 / freq - e,
// 0 < e <= 1: mulhi(x, m) = floor(x / freq - x e / 2^32) and x e / 2^32 < 1, so the estimate is the quotient or one
// less: one upward correction, never a downward one.  m depends on the symbol only, not on the coder state.
__device__ __forceinline__ uint32_t rans_rcp(uint32_t freq) { return 0xffffffffu / freq; }
__device__ __forceinline__ void rans_divmod(uint32_t x, uint32_t freq, uint32_t rcp, uint32_t& q, uint32_t& r) {
    q = __umulhi(x, rcp);
    r = x - q * freq;
    if (r >= freq) { ++q; r -= freq; }
}

__device__ __forceinline__ int symbol_of(float x, int R, unsigned int* err) {
    int s = (int)rintf(x) + R;                       // net.py:76/91/100 round, then x + mxrange
    if (s < 0 || s > 2 * R - 2) {
        atomicAdd(err, 1u);
        s = min(max(s, 0), 2 * R - 2);
    }
    return s;
}
__global__ void k_sym_factorized(const float* __restrict__ x, int64_t n, int C, int R,
                                 const uint32_t* __restrict__ table, uint32_t* __restrict__ packed,
                                 unsigned int* __restrict__ err) {
    pdl_sync();
    const int Lp = 2 * R;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const int s = symbol_of(x[i], R, err);
        const uint32_t start = table[(size_t)c * Lp + s], end = table[(size_t)c * Lp + s + 1];
        reinterpret_cast<uint2*>(packed)[i] = make_uint2(start | ((end - start) << 16), rans_rcp(end - start));
    }
}
__global__ void k_sym_laplace(const float* __restrict__ x, const float* __restrict__ sigma, int64_t n, int R,
                              uint32_t* __restrict__ packed, unsigned int* __restrict__ err) {
    pdl_sync();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int s = symbol_of(x[i], R, err);
        const float sg = fminf(fmaxf(sigma[i], 1e-5f), 1e10f);
        uint32_t start, end;
        laplace_interval(s, R, sg, start, end);
        if (end <= start) {                          // non-monotonic float CDF: not codable (torchac would break too)
            atomicAdd(err + 1, 1u);
            end = start + 1;
        }
        reinterpret_cast<uint2*>(packed)[i] = make_uint2(start | ((end - start) << 16), rans_rcp(end - start));
    }
}

// ----------------------------------------------------------------------------------------------
// phase 2: one rANS stream per lane of L symbols.  words: [nlanes][L + 2] 16-bit, filled from the END of the lane's
// slot backwards, so that the finished lane reads forward as: state_hi, state_lo, then the words the decoder pulls.
// ----------------------------------------------------------------------------------------------
__global__ void k_rans_encode(const uint32_t* __restrict__ packed, int64_t n, int L, uint16_t* __restrict__ words,
                              uint32_t* __restrict__ lane_words) {
    pdl_sync();
    const int64_t nlanes = (n + L - 1) / L;
    const int64_t lane = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (lane >= nlanes) return;
    const int64_t first = lane * L;
    const int cnt = (int)min((int64_t)L, n - first);
    uint16_t* slot = words + lane * (int64_t)(L + 2);
    int w = L + 2;                                    // write cursor (exclusive)
    uint32_t x = RANS_L;
    // The state update is a serial chain over the lane's symbols and this thread issues in order: everything that does
    // not depend on the state was done in phase 1 (interval and the reciprocal m = floor((2^32 - 1) / freq) that turns
    // the division into a multiply-high), the next symbol's record is fetched while this one is coded.
    const uint2* rec = reinterpret_cast<const uint2*>(packed) + first;
    uint2 p = cnt > 0 ? rec[cnt - 1] : make_uint2(0x10000u, 0xffffffffu);
    for (int k = cnt - 1; k >= 0; --k) {              // rANS codes backwards, decodes forwards
        const uint32_t start = p.x & 0xffffu, freq = p.x >> 16, rcur = p.y;
        if (k > 0) p = rec[k - 1];
        if (x >= (freq << 16)) {                      // x_max = ((RANS_L >> 16) << 16) * freq
            slot[--w] = (uint16_t)(x & 0xffffu);
            x >>= 16;
        }
        uint32_t q, r;
        rans_divmod(x, freq, rcur, q, r);
        x = (q << 16) + r + start;
    }
    slot[--w] = (uint16_t)(x & 0xffffu);
    slot[--w] = (uint16_t)(x >> 16);
    lane_words[lane] = (uint32_t)(L + 2 - w);
}

// container assembly: one block per lane (block 0 also writes the header)
__global__ void k_rans_pack(const uint16_t* __restrict__ words, const uint32_t* __restrict__ lane_words, int64_t n,
                            int L, int slot_words, uint8_t* __restrict__ out, uint32_t* __restrict__ total_bytes) {
    pdl_sync();
    const int nlanes = (int)((n + L - 1) / L);
    const int lane = blockIdx.x;
    __shared__ uint32_t off_s;
    if (threadIdx.x == 0) {
        uint32_t off = 0;
        for (int j = 0; j < lane; ++j) off += lane_words[j];
        off_s = off;
    }
    __syncthreads();
    const uint32_t hdr = 16u + (((uint32_t)nlanes * 2u + 3u) & ~3u);
    uint16_t* payload = reinterpret_cast<uint16_t*>(out + hdr);
    const uint32_t nw = lane_words[lane];
    const uint16_t* src = words + (int64_t)lane * slot_words + (slot_words - nw);
    for (uint32_t k = threadIdx.x; k < nw; k += blockDim.x) payload[off_s + k] = src[k];
    if (lane == 0) {
        uint32_t* h = reinterpret_cast<uint32_t*>(out);
        if (threadIdx.x == 0) { h[0] = FVR_MAGIC; h[1] = (uint32_t)n; h[2] = (uint32_t)L; h[3] = (uint32_t)nlanes; }
        uint16_t* lw = reinterpret_cast<uint16_t*>(out + 16);
        for (int j = threadIdx.x; j < ((nlanes + 1) & ~1); j += blockDim.x) lw[j] = j < nlanes ? (uint16_t)lane_words[j] : 0;
    }
    if (lane == nlanes - 1 && threadIdx.x == 0) *total_bytes = hdr + 2u * (off_s + nw);
}

// ----------------------------------------------------------------------------------------------
// decoding: one thread per lane
// ----------------------------------------------------------------------------------------------
struct LaneReader {
    const uint16_t* w;
    uint32_t x;
    __device__ __forceinline__ bool open(const uint8_t* stream, int64_t nbytes, int64_t n_expect, int lane, int L,
                                         int& cnt) {
        if (nbytes < 16) return false;
        const uint32_t* h = reinterpret_cast<const uint32_t*>(stream);
        if (h[0] != FVR_MAGIC || (int64_t)h[1] != n_expect || (int)h[2] != L) return false;
        const int nlanes = (int)h[3];
        if (nlanes != (int)((n_expect + L - 1) / L) || lane >= nlanes) return false;
        const uint16_t* lw = reinterpret_cast<const uint16_t*>(stream + 16);
        const uint32_t hdr = 16u + (((uint32_t)nlanes * 2u + 3u) & ~3u);
        uint32_t off = 0;
        for (int j = 0; j < lane; ++j) off += lw[j];
        if ((int64_t)hdr + 2 * ((int64_t)off + lw[lane]) > nbytes || lw[lane] < 2) return false;
        w = reinterpret_cast<const uint16_t*>(stream + hdr) + off;
        x = ((uint32_t)w[0] << 16) | w[1];
        w += 2;
        cnt = (int)min((int64_t)L, n_expect - (int64_t)lane * L);
        return true;
    }
    __device__ __forceinline__ void advance(uint32_t start, uint32_t freq) {
        x = freq * (x >> 16) + (x & 0xffffu) - start;
        if (x < RANS_L) x = (x << 16) | *w++;
    }
};

// q_out: fp32 NHWC [n] (value = symbol - R); err[2] counts lanes that could not be opened
__global__ void k_rans_decode_factorized(const uint8_t* __restrict__ stream, int64_t nbytes, int64_t n, int C, int R,
                                         const uint32_t* __restrict__ table, float* __restrict__ q_out,
                                         unsigned int* __restrict__ err, int L) {
    const int lane = blockIdx.x * blockDim.x + threadIdx.x;
    if (lane >= (int)((n + L - 1) / L)) return;
    LaneReader rd;
    int cnt = 0;
    if (!rd.open(stream, nbytes, n, lane, L, cnt)) {
        atomicAdd(err + 2, 1u);
        return;
    }
    const int Lp = 2 * R;
    const int64_t first = (int64_t)lane * L;
    for (int k = 0; k < cnt; ++k) {
        const int c = (int)((first + k) % C);
        const uint32_t* T = table + (size_t)c * Lp;
        const uint32_t slot = rd.x & 0xffffu;
        int lo = 0, hi = Lp - 2;                      // largest s with T[s] <= slot
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (T[mid] <= slot) lo = mid; else hi = mid - 1;
        }
        rd.advance(T[lo], T[lo + 1] - T[lo]);
        q_out[first + k] = (float)(lo - R);
    }
}
__global__ void k_rans_decode_laplace(const uint8_t* __restrict__ stream, int64_t nbytes, int64_t n, int R,
                                      const float* __restrict__ sigma, float* __restrict__ q_out,
                                      unsigned int* __restrict__ err, int L) {
    const int lane = blockIdx.x * blockDim.x + threadIdx.x;
    if (lane >= (int)((n + L - 1) / L)) return;
    LaneReader rd;
    int cnt = 0;
    if (!rd.open(stream, nbytes, n, lane, L, cnt)) {
        atomicAdd(err + 2, 1u);
        return;
    }
    const int64_t first = (int64_t)lane * L;
    for (int k = 0; k < cnt; ++k) {
        const float sg = fminf(fmaxf(sigma[first + k], 1e-5f), 1e10f);
        const uint32_t slot = rd.x & 0xffffu;
        int lo = 0, hi = 2 * R - 2;                   // largest s with Q[s] <= slot
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (laplace_q(mid, R, sg) <= slot) lo = mid; else hi = mid - 1;
        }
        uint32_t start, end;
        laplace_interval(lo, R, sg, start, end);
        if (end <= start) end = start + 1;
        rd.advance(start, end - start);
        q_out[first + k] = (float)(lo - R);
    }
}

// ----------------------------------------------------------------------------------------------
// Indexed-table coder: the model CompressAI's EntropyModel.compress / decompress hands its range coder
// (reference entropy_models.py:80-93, 237-247 call it through entropy_bottleneck / gaussian_conditional): per element
// an index into a set of quantised CDF tables (`_quantized_cdf` [ntab][stride], `_cdf_length`, `_offset`, written by
// update()), symbols outside a table's range escaped through its last (tail-mass) bin followed by the raw value in
// 4-bit bypass digits:
//     v = symbol - offset[idx];  max = cdf_length[idx] - 2
//     v < 0    -> raw = -2v - 1, v = max;      v >= max -> raw = 2 (v - max), v = max
//     code v under the table; if v == max: nb = number of 4-bit digits of raw, coded as digits of 15 then the rest
//     (while nb >= 15: 15, nb -= 15; then nb), then the nb digits of raw, least significant first.
// Same lanes and container as above; a bypass digit d is the interval [d << 12, (d + 1) << 12).  One thread per lane.
// err[0]: index outside [0, ntab); err[1]: empty interval (malformed table); err[2]: lane could not be opened.
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void rans_put(uint32_t& x, uint16_t* slot, int& w, uint32_t start, uint32_t freq) {
    if ((uint64_t)x >= ((uint64_t)freq << 16)) {     // freq may be 2^16 (a one-symbol table): compare in 64 bits
        slot[--w] = (uint16_t)(x & 0xffffu);
        x >>= 16;
    }
    uint32_t q, r;
    rans_divmod(x, freq, rans_rcp(freq), q, r);       // the reciprocal does not depend on the state: off the chain
    x = (q << 16) + r + start;
}

__global__ void k_rans_encode_indexed(const int32_t* __restrict__ symbols, const int32_t* __restrict__ indexes, int64_t n,
                                      int L, int slot_words, const int32_t* __restrict__ cdf, int ntab, int stride,
                                      const int32_t* __restrict__ cdf_len, const int32_t* __restrict__ offset,
                                      uint16_t* __restrict__ words, uint32_t* __restrict__ lane_words,
                                      unsigned int* __restrict__ err) {
    const int64_t nlanes = (n + L - 1) / L;
    const int64_t lane = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (lane >= nlanes) return;
    const int64_t first = lane * L;
    const int cnt = (int)min((int64_t)L, n - first);
    uint16_t* slot = words + lane * (int64_t)slot_words;
    int w = slot_words;
    uint32_t x = RANS_L;
    for (int k = cnt - 1; k >= 0; --k) {              // elements backwards, and inside an element its digits backwards
        int idx = indexes[first + k];
        if (idx < 0 || idx >= ntab) { atomicAdd(err, 1u); idx = 0; }
        const int32_t* T = cdf + (size_t)idx * stride;
        const int maxv = cdf_len[idx] - 2;
        int v = symbols[first + k] - offset[idx];
        uint32_t raw = 0;
        if (v < 0) { raw = (uint32_t)(-2 * (int64_t)v - 1); v = maxv; }
        else if (v >= maxv) { raw = (uint32_t)(2 * ((int64_t)v - maxv)); v = maxv; }
        if (v == maxv) {
            int nb = 0;
            while (nb < 8 && (raw >> (nb * 4)) != 0) ++nb;
            for (int j = nb - 1; j >= 0; --j) rans_put(x, slot, w, ((raw >> (j * 4)) & 15u) << 12, 1u << 12);
            const int n15 = nb / 15, last = nb - 15 * n15;   // decode order: n15 digits of 15, then `last`
            rans_put(x, slot, w, (uint32_t)last << 12, 1u << 12);
            for (int j = 0; j < n15; ++j) rans_put(x, slot, w, 15u << 12, 1u << 12);
        }
        const uint32_t start = (uint32_t)T[v];
        uint32_t freq = (uint32_t)T[v + 1] - start;
        if ((int32_t)freq <= 0 || T[v + 1] > 65536) { atomicAdd(err + 1, 1u); freq = 1; }
        rans_put(x, slot, w, start, freq);
    }
    slot[--w] = (uint16_t)(x & 0xffffu);
    slot[--w] = (uint16_t)(x >> 16);
    lane_words[lane] = (uint32_t)(slot_words - w);
}

__global__ void k_rans_decode_indexed(const uint8_t* __restrict__ stream, int64_t nbytes, int64_t n,
                                      const int32_t* __restrict__ indexes, const int32_t* __restrict__ cdf, int ntab,
                                      int stride, const int32_t* __restrict__ cdf_len, const int32_t* __restrict__ offset,
                                      int32_t* __restrict__ symbols, unsigned int* __restrict__ err, int L) {
    const int lane = blockIdx.x * blockDim.x + threadIdx.x;
    if (lane >= (int)((n + L - 1) / L)) return;
    LaneReader rd;
    int cnt = 0;
    if (!rd.open(stream, nbytes, n, lane, L, cnt)) {
        atomicAdd(err + 2, 1u);
        return;
    }
    const int64_t first = (int64_t)lane * L;
    for (int k = 0; k < cnt; ++k) {
        int idx = indexes[first + k];
        if (idx < 0 || idx >= ntab) { atomicAdd(err, 1u); idx = 0; }
        const int32_t* T = cdf + (size_t)idx * stride;
        const int maxv = cdf_len[idx] - 2;
        const int32_t slot = (int32_t)(rd.x & 0xffffu);
        int lo = 0, hi = maxv;                        // largest v with T[v] <= slot
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (T[mid] <= slot) lo = mid; else hi = mid - 1;
        }
        rd.advance((uint32_t)T[lo], (uint32_t)(T[lo + 1] - T[lo]));
        int v = lo;
        if (v == maxv) {
            auto digit = [&]() {
                const uint32_t d = (rd.x & 0xffffu) >> 12;
                rd.advance(d << 12, 1u << 12);
                return (int)d;
            };
            int d = digit(), nb = d;
            while (d == 15) { d = digit(); nb += d; }
            uint32_t raw = 0;
            for (int j = 0; j < nb; ++j) raw |= (uint32_t)digit() << (4 * (j & 7));
            v = (int)(raw >> 1);
            if (raw & 1u) v = -v - 1; else v += maxv;
        }
        symbols[first + k] = v + offset[idx];
    }
}

// bits = 8 * bytes, as floats, for the bpp bookkeeping (net.py:136: len(byte_stream) * 8)
__global__ void k_bytes_to_bits(const uint32_t* __restrict__ nbytes, const unsigned int* __restrict__ err,
                                float* __restrict__ bits) {
    pdl_sync();
    // an uncodable symbol (outside +-mxrange, or an empty interval) makes the stream invalid: NaN instead of a number
    if (threadIdx.x == 0) *bits = (err[0] | err[1]) ? __int_as_float(0x7fc00000) : 8.0f * (float)(*nbytes);
}

// ----------------------------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------------------------
static int grid_for(int64_t n, int threads) {
    return (int)std::min<int64_t>(std::max<int64_t>(cdiv64(n, threads * 4), 1), 148 * 8);
}

size_t entropy_stream_capacity(int64_t n, int L) {
    const int64_t nlanes = cdiv64(std::max<int64_t>(n, 1), L);
    return (size_t)(16 + ((nlanes * 2 + 3) & ~(int64_t)3) + 2 * (n + 2 * nlanes));
}
size_t entropy_words_capacity(int64_t n, int L) { return (size_t)cdiv64(std::max<int64_t>(n, 1), L) * (size_t)(L + 2); }

int launch_cdf_table_factorized(FactorizedParams prm, int C, int R, uint32_t* table, cudaStream_t s) {
    FVC_ARG(C >= 1 && R >= 2 && R <= 16384);
    FVC_CUDA(launch_pdl(k_cdf_table_factorized, C, 128, 0, s, prm, C, R, table));
    g_launch_count++;
    FVC_CHECK_LAUNCH();
    return 0;
}
int launch_cdf_table_laplace(const float* sigma, int64_t n, int R, uint32_t* table, cudaStream_t s) {
    FVC_ARG(R >= 2 && R <= 16384);
    if (n == 0) return 0;
    k_cdf_table_laplace<<<grid_for(n * 2 * R, 256), 256, 0, s>>>(sigma, n, R, table);
    g_launch_count++;
    FVC_CHECK_LAUNCH();
    return 0;
}
int launch_sym_factorized(const float* x, int64_t n, int C, int R, const uint32_t* table, uint32_t* packed,
                          unsigned int* err, cudaStream_t s) {
    FVC_CUDA(launch_pdl(k_sym_factorized, grid_for(n, 256), 256, 0, s, x, n, C, R, table, packed, err));
    g_launch_count++;
    FVC_CHECK_LAUNCH();
    return 0;
}
int launch_sym_laplace(const float* x, const float* sigma, int64_t n, int R, uint32_t* packed, unsigned int* err,
                       cudaStream_t s) {
    FVC_CUDA(launch_pdl(k_sym_laplace, grid_for(n, 256), 256, 0, s, x, sigma, n, R, packed, err));
    g_launch_count++;
    FVC_CHECK_LAUNCH();
    return 0;
}
int launch_rans_encode(const uint32_t* packed, int64_t n, int L, uint16_t* words, uint32_t* lane_words, uint8_t* out,
                       uint32_t* total_bytes, cudaStream_t s) {
    FVC_ARG(n >= 1 && L >= 1 && L <= 65533);
    const int64_t nlanes = cdiv64(n, L);
    FVC_ARG(nlanes <= 65535);
    FVC_CUDA(launch_pdl(k_rans_encode, (unsigned)cdiv64(nlanes, 32), 32, 0, s, packed, n, L, words, lane_words));
    g_launch_count++;
    FVC_CHECK_LAUNCH();
    FVC_CUDA(launch_pdl(k_rans_pack, (unsigned)nlanes, 128, 0, s, (const uint16_t*)words, (const uint32_t*)lane_words, n,
                        L, L + 2, out, total_bytes));
    g_launch_count++;
    FVC_CHECK_LAUNCH();
    return 0;
}
int launch_rans_decode_factorized(const uint8_t* stream, int64_t nbytes, int64_t n, int L, int C, int R,
                                  const uint32_t* table, float* q_out, unsigned int* err, cudaStream_t s) {
    FVC_ARG(n >= 1 && L >= 1);
    const int nl = (int)cdiv64(n, L);
    k_rans_decode_factorized<<<cdiv(nl, 32), 32, 0, s>>>(stream, nbytes, n, C, R, table, q_out, err, L);
    g_launch_count++;
    FVC_CHECK_LAUNCH();
    return 0;
}
int launch_rans_decode_laplace(const uint8_t* stream, int64_t nbytes, int64_t n, int L, int R, const float* sigma,
                               float* q_out, unsigned int* err, cudaStream_t s) {
    FVC_ARG(n >= 1 && L >= 1);
    const int nl = (int)cdiv64(n, L);
    k_rans_decode_laplace<<<cdiv(nl, 32), 32, 0, s>>>(stream, nbytes, n, R, sigma, q_out, err, L);
    g_launch_count++;
    FVC_CHECK_LAUNCH();
    return 0;
}
// indexed-table coder: 4-bit bypass digits grow the state by 4 bits each, a table symbol by at most 16: at most
// (16 + 9 * 4) / 16 = 3.25 words per element
int entropy_indexed_slot_words(int L) { return (13 * L + 3) / 4 + 3; }
size_t entropy_stream_capacity_indexed(int64_t n, int L) {
    const int64_t nlanes = cdiv64(std::max<int64_t>(n, 1), L);
    return (size_t)(16 + ((nlanes * 2 + 3) & ~(int64_t)3) + 2 * nlanes * (int64_t)entropy_indexed_slot_words(L));
}
int launch_rans_encode_indexed(const int32_t* symbols, const int32_t* indexes, int64_t n, int L, const int32_t* cdf,
                               int ntab, int stride, const int32_t* cdf_len, const int32_t* offset, uint16_t* words,
                               uint32_t* lane_words, uint8_t* out, uint32_t* total_bytes, unsigned int* err,
                               cudaStream_t s) {
    FVC_ARG(n >= 1 && L >= 1 && entropy_indexed_slot_words(L) <= 65535 && ntab >= 1 && stride >= 2);
    const int64_t nlanes = cdiv64(n, L);
    FVC_ARG(nlanes <= 65535);
    const int slot = entropy_indexed_slot_words(L);
    k_rans_encode_indexed<<<(unsigned)cdiv64(nlanes, 32), 32, 0, s>>>(symbols, indexes, n, L, slot, cdf, ntab, stride,
                                                                      cdf_len, offset, words, lane_words, err);
    g_launch_count++;
    FVC_CHECK_LAUNCH();
    FVC_CUDA(launch_pdl(k_rans_pack, (unsigned)nlanes, 128, 0, s, (const uint16_t*)words, (const uint32_t*)lane_words, n,
                        L, slot, out, total_bytes));
    g_launch_count++;
    FVC_CHECK_LAUNCH();
    return 0;
}
int launch_rans_decode_indexed(const uint8_t* stream, int64_t nbytes, int64_t n, int L, const int32_t* indexes,
                               const int32_t* cdf, int ntab, int stride, const int32_t* cdf_len, const int32_t* offset,
                               int32_t* symbols, unsigned int* err, cudaStream_t s) {
    FVC_ARG(n >= 1 && L >= 1 && ntab >= 1 && stride >= 2);
    const int nl = (int)cdiv64(n, L);
    k_rans_decode_indexed<<<cdiv(nl, 32), 32, 0, s>>>(stream, nbytes, n, indexes, cdf, ntab, stride, cdf_len, offset,
                                                      symbols, err, L);
    g_launch_count++;
    FVC_CHECK_LAUNCH();
    return 0;
}
int launch_bytes_to_bits(const uint32_t* nbytes, const unsigned int* err, float* bits, cudaStream_t s) {
    FVC_CUDA(launch_pdl(k_bytes_to_bits, 1, 32, 0, s, nbytes, err, bits));
    g_launch_count++;
    FVC_CHECK_LAUNCH();
    return 0;
}

}  // namespace fvc
