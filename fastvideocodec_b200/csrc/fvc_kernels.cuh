// Internal launch wrappers (host side) for the memory-bound kernels and both conv engines.
#pragma once
#include "fvc_common.cuh"

namespace fvc {

// ---- geometry / elementwise (fvc_geom.cu) ----------------------------------------------------
int launch_avg_pool2_planar(const float* x, float* y, int planes, int H, int W, cudaStream_t s);
int launch_u8hwc_to_f32chw(const uint8_t* src, float* dst, int n, int H, int W, cudaStream_t s);
int launch_upsample2x_planar(const float* x, float* y, int planes, int H, int W, int align_corners, float scale,
                             cudaStream_t s);
int launch_flow_warp_nchw(const float* img, const float* flow, float* out, int B, int C, int H, int W,
                          cudaStream_t s);
// SpyNet level input: X = [im1(3), warp(im2, up)(3), up(2), 0...] as ACT(Cp=32); up = 2*bilinear_up(flow_prev)
// flow_prev: fp32 NHWC [B,H/2,W/2,2] or nullptr (zero flow).  flow_up: fp32 NHWC [B,H,W,2].
int launch_spynet_prep(const float* im1, const float* im2, const float* flow_prev, ActT X, float* flow_up,
                       cudaStream_t s);
// Motion compensation input: warpframe (planar) = warp(ref, mv); X = [warp(3), ref(3), 0...] ACT(Cp=32)
int launch_mc_prep(const float* ref, const float* mv_nhwc2, float* warpframe, ActT X, cudaStream_t s);
// prediction = res + warpframe (planar out); residual = cur - prediction -> ACT (Cp=32, ch 0..2)
int launch_mc_finish(const float* res_nhwc3, const float* warpframe, const float* cur, float* prediction, ActT R,
                     cudaStream_t s);
int launch_pool_act(ActT in, ActT out, ActT out_relu, cudaStream_t s);
int launch_upadd_act(ActT low, ActT skip, ActT out, ActT out_relu, cudaStream_t s);
int launch_gdn_act(ActT in, int C, const float* beta_eff, const float* gamma_eff, int inverse, ActT out,
                   cudaStream_t s);
int launch_gdn_reparam(const float* beta, const float* gamma, float* beta_eff, float* gamma_eff, int C,
                       cudaStream_t s);
// losses.  res_nhwc3 != 0: res is fp32 NHWC with 3 channels, everything else planar [B,3,H,W].
int launch_recon_losses(const float* cur, const float* pred, const float* warp, const float* res, int res_nhwc3,
                        int B, int HW, float* clipped, float* partials, int* nblocks_out, cudaStream_t s,
                        int clip_mse = 0);
// fused 3x3 tail convolutions: w1x1[(r*3+s)*Cout + co][ci] = w[co][ci][r][s] (rows beyond 9*Cout zero), and the sum of
// the 9 shifted per-pixel partial products: out[b,y,x,co] = bias[co] + sum_{r,s} P[b,y+r-1,x+s-1,(r*3+s)*Cout+co]
int launch_tapsplit_weights(const float* w3x3, float* w1x1, int Cin, int Cout, int rows, cudaStream_t s);
int launch_tapsum(const float* P, const float* bias, float* out, int B, int H, int W, int Cout, int cq, cudaStream_t s);
int launch_finalize_scalars(const float* sums6, float n_pix, float* scalars7, const unsigned int* sat_count,
                            cudaStream_t s);
int launch_reduce_partials(const float* partials, int n, int groups, double scale, float* out, cudaStream_t s);
// layout conversion
int launch_nchw_to_act(const float* x, ActT out, int C, int do_abs, cudaStream_t s);
int launch_act_to_nchw(ActT in, int C, float* y, cudaStream_t s);
int launch_nhwc_to_nchw(const float* x, float* y, int B, int C, int H, int W, cudaStream_t s);
int launch_nchw_to_nhwc(const float* x, float* y, int B, int C, int H, int W, cudaStream_t s);
int launch_nhwc_to_act(const float* x, ActT out, int C, int do_abs, cudaStream_t s);

// ---- entropy-model bit estimation (fvc_bits.cu) ----------------------------------------------
struct FactorizedParams {
    const float* p[11];  // f1.h f1.b f1.a f2.h f2.b f2.a f3.h f3.b f3.a f4.h f4.b, each [C]
};
// x: fp32, NHWC [npix,C] (nhwc=1) or NCHW [B,C,HW] (nhwc=0).  q_f32 (same layout, optional),
// q_act (optional).  partials: per-block bit sums (>= grid entries); returns grid size in *nblocks.
int launch_quant_bits_factorized(const float* x, int nhwc, int B, int C, int HW, FactorizedParams prm, float* q_f32,
                                 ActT q_act, float* partials, int* nblocks, cudaStream_t s);
int launch_quant_bits_laplace(const float* x, const float* sigma, int64_t n, int C, float* q_f32, ActT q_act,
                              float* partials, int* nblocks, cudaStream_t s);
int launch_eb_forward(const float* x, const float* packed, const float* medians, float* xhat, float* lik,
                      float* partials, int* nblocks, int B, int C, int HW, cudaStream_t s);
int launch_gaussian_forward(const float* x, const float* scales, const float* means, float* xhat, float* lik,
                            float* partials, int* nblocks, int64_t n, cudaStream_t s);
int bits_max_blocks();

// ---- entropy coding of the quantised latents (fvc_entropy.cu; net.py:123-138, 155-168, 183-195) ------------
// R = mxrange (150): symbols s = q + R in [0, 2R-2]; tables are uint32 [C][2R] / [n][2R]; L = symbols per rANS lane
size_t entropy_stream_capacity(int64_t n, int L);   // bytes
size_t entropy_words_capacity(int64_t n, int L);    // 16-bit words of scratch
int entropy_indexed_slot_words(int L);               // indexed-table coder: 16-bit words of scratch per lane
size_t entropy_stream_capacity_indexed(int64_t n, int L);
int launch_rans_encode_indexed(const int32_t* symbols, const int32_t* indexes, int64_t n, int L, const int32_t* cdf,
                               int ntab, int stride, const int32_t* cdf_len, const int32_t* offset, uint16_t* words,
                               uint32_t* lane_words, uint8_t* out, uint32_t* total_bytes, unsigned int* err,
                               cudaStream_t s);
int launch_rans_decode_indexed(const uint8_t* stream, int64_t nbytes, int64_t n, int L, const int32_t* indexes,
                               const int32_t* cdf, int ntab, int stride, const int32_t* cdf_len, const int32_t* offset,
                               int32_t* symbols, unsigned int* err, cudaStream_t s);
int launch_cdf_table_factorized(FactorizedParams prm, int C, int R, uint32_t* table, cudaStream_t s);
int launch_cdf_table_laplace(const float* sigma, int64_t n, int R, uint32_t* table, cudaStream_t s);
int launch_sym_factorized(const float* x, int64_t n, int C, int R, const uint32_t* table, uint32_t* packed,
                          unsigned int* err, cudaStream_t s);
int launch_sym_laplace(const float* x, const float* sigma, int64_t n, int R, uint32_t* packed, unsigned int* err,
                       cudaStream_t s);
int launch_rans_encode(const uint32_t* packed, int64_t n, int L, uint16_t* words, uint32_t* lane_words, uint8_t* out,
                       uint32_t* total_bytes, cudaStream_t s);
int launch_rans_decode_factorized(const uint8_t* stream, int64_t nbytes, int64_t n, int L, int C, int R,
                                  const uint32_t* table, float* q_out, unsigned int* err, cudaStream_t s);
int launch_rans_decode_laplace(const uint8_t* stream, int64_t nbytes, int64_t n, int L, int R, const float* sigma,
                               float* q_out, unsigned int* err, cudaStream_t s);
int launch_bytes_to_bits(const uint32_t* nbytes, const unsigned int* err, float* bits, cudaStream_t s);

// ---- convolution engines ---------------------------------------------------------------------
// SIMT: packed fp32 weights [sub][tap][CinP][CoutS]
struct SimtWeights {
    float* w = nullptr;     // device
    size_t sub_stride = 0;  // floats between sub-convolutions
    int CinP = 0, CoutS = 0;
};
int simt_pack_weights(const ConvLayer& L, const float* w_ref, int CinP, int CoutS, SimtWeights* out,
                      cudaStream_t s);
int launch_conv_simt(const ConvLayer& L, const SimtWeights& W, ActT in, int Hout, int Wout, const Epilogue& ep,
                     cudaStream_t s);

// fp32 CUDA-core path for convolutions with 2-3 output channels (fvc_conv_few.cu)
bool few_supported(const ConvLayer& L, int CinP);
int few_pack_weights(const ConvLayer& L, const float* w_ref, float** out, cudaStream_t s);
int launch_conv_few(const ConvLayer& L, const float* w_packed, const float* bias, ActT in, int Hout, int Wout,
                    const Epilogue& ep, cudaStream_t s);

// TC (tcgen05): packed bf16 hi/lo weight stream + TMA descriptors (fvc_conv_tc.cu)
struct TcPlan;  // opaque
// fast = 1: one MMA per product on the hi halves only (fp16 operands, ~11 significant bits), ACT outputs carry hi only
int tc_plan_create(const ConvLayer& L, const float* w_ref, ActT in, int Hout, int Wout, const Epilogue& ep,
                   TcPlan** plan, cudaStream_t s, bool fast = false, bool no_merge = false);
int tc_plan_launch(TcPlan* plan, cudaStream_t s);
void tc_plan_destroy(TcPlan* plan);
const e16* tc_plan_wstream(const TcPlan* plan);
float tc_plan_acc_scale(const TcPlan* plan);
bool tc_plan_is_gdn_norm_layout(const TcPlan* plan);
bool tc_plan_is_tap_layout(const TcPlan* plan);
bool tc_supported(const ConvLayer& L, int CinP);

}  // namespace fvc
