/*
 * fvc_b200.h — C-ABI of libfvc_b200.so: the B200 (sm_100a) implementation of the DVC P-frame
 * coding hot path of bochen-sysnet/FastVideoCodec.
 *
 * The reference has no FFI: its seam is the Python nn.Module API (DVC/net.py::VideoCompressor).
 * This library sits directly below that seam; fastvideocodec_b200/_lib.py binds it with ctypes
 * (INTEGRATION.md shows the stub).  Each entry point cites the reference code it replaces
 * (paths relative to the reference checkout).
 *
 * Conventions
 *   - plain C types only; every pointer is a DEVICE pointer unless the name ends in _host;
 *   - tensors are dense fp32 NCHW exactly as the reference holds them, unless stated;
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream); calls are
 *     stream-ordered and never synchronise unless stated;
 *   - return 0 on success, <0 on error; fvc_last_error() gives the thread-local message;
 *   - there is no CPU fallback: without a CUDA device every compute call fails with FVC_ERR_CUDA.
 */
#ifndef FVC_B200_H_
#define FVC_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define FVC_API __attribute__((visibility("default")))
#else
#define FVC_API
#endif

#define FVC_OK 0
#define FVC_ERR_ARG (-1)
#define FVC_ERR_CUDA (-2)
#define FVC_ERR_STATE (-3)

/* activation codes for fvc_conv2d */
#define FVC_ACT_NONE 0
#define FVC_ACT_RELU 1
#define FVC_ACT_LRELU01 2 /* LeakyReLU(0.1) — analysis_mv.py:17 */
#define FVC_ACT_EXP 3     /* synthesis_prior.py:57 */
#define FVC_ACT_LRELU001 4 /* nn.LeakyReLU() default slope 0.01 — entropy_models.py:166-188 */

/* convolution engines */
#define FVC_IMPL_SIMT 0 /* fp32 CUDA-core implicit GEMM (checker / bring-up path) */
#define FVC_IMPL_TC 1   /* tcgen05 + TMEM + TMA, fp16 hi/lo operand pairs (3 MMAs per product), fp32 accumulate:
                           precision 'exact' (element-level parity with the fp32 reference) */
#define FVC_IMPL_TC_FAST 2 /* the same engine, precision 'fast': ONE fp16 MMA per product (hi halves only), fp32
                              accumulate, long TMEM chains; metric-level parity only (SURVEY 7.2-1, configs[1]) */

FVC_API int fvc_version(void);
FVC_API const char* fvc_last_error(void);

/* ---------------------------------------------------------------------------------------------
 * Op-level entry points (each mirrors one reference op; used by the parity tests and by
 * fastvideocodec_b200.ops).
 * ------------------------------------------------------------------------------------------- */

/* F.avg_pool2d(x, 2, 2) — endecoder.py:344-346, AvgPool2d endecoder.py:273,275.
 * x: [planes,H,W] -> y: [planes,H/2,W/2]. */
FVC_API int fvc_avg_pool2(const float* x, float* y, int planes, int H, int W, void* stream);

/* F.interpolate(x, (2H,2W), mode='bilinear', align_corners) * scale — endecoder.py:173-184, 353.
 * x: [planes,H,W] -> y: [planes,2H,2W]. */
FVC_API int fvc_upsample2x_bilinear(const float* x, float* y, int planes, int H, int W, int align_corners,
                            float scale, void* stream);

/* flow_warp / torch_warp — endecoder.py:52-67, 116-119 (grid_sample bilinear, border,
 * align_corners=False on an align_corners=True style grid).
 * img: [B,C,H,W], flow: [B,2,H,W] (ch0 = x, ch1 = y) -> out: [B,C,H,W]. */
FVC_API int fvc_flow_warp(const float* img, const float* flow, float* out, int B, int C, int H, int W, void* stream);

/* transforms.ToTensor() — the reference's frame ingest (dataset.py:75, models.py:425): uint8 HWC [n,H,W,3] ->
 * fp32 CHW [n,3,H,W] in [0,1], x / 255 with IEEE division (bit-identical to torchvision).  Device pointers. */
FVC_API int fvc_u8hwc_to_f32chw(const uint8_t* src, float* dst, int n, int H, int W, void* stream);

/* nn.Conv2d(Cin,Cout,k,stride,padding=k//2) [transposed=0; weight [Cout,Cin,k,k]] or
 * nn.ConvTranspose2d(Cin,Cout,k,stride,padding=k//2,output_padding=stride-1)
 * [transposed=1; weight [Cin,Cout,k,k]], + bias, + activation.  stride in {1,2}.
 * x: [B,Cin,H,W] -> y: [B,Cout,Ho,Wo] with Ho = H/stride (conv) or H*stride (transposed).
 * impl selects the engine (FVC_IMPL_*). */
FVC_API int fvc_conv2d(const float* x, const float* w, const float* bias, float* y, int B, int Cin, int H, int W,
               int Cout, int k, int stride, int transposed, int act, int impl, void* stream);

/* The same convolution as a reusable handle, for callers that apply one layer many times (the reference's nn.Conv2d /
 * nn.ConvTranspose2d modules outside VideoCompressor: entropy_models.py:160-190 hyper-prior stacks, one call per
 * forward).  fvc_conv_op_create packs the weights and builds the engine plan once (the only step with a host
 * synchronisation); fvc_conv_op_run converts x, launches and converts back, asynchronously on `stream`.  The handle
 * copies the bias, owns its staging tensors and is specific to the input shape [B,Cin,H,W] and to the device it was
 * created on; weights changed afterwards need a new handle.  Destroy only after the last run has finished. */
typedef struct fvc_conv_op fvc_conv_op;
FVC_API int fvc_conv_op_create(fvc_conv_op** out, const float* w, const float* bias, int B, int Cin, int H, int W,
                               int Cout, int k, int stride, int transposed, int act, int impl, void* stream);
FVC_API int fvc_conv_op_run(fvc_conv_op* op, const float* x, float* y, void* stream);
FVC_API void fvc_conv_op_destroy(fvc_conv_op* op);

/* GDN.forward — GDN.py:63-93.  beta: [C], gamma: [C,C] are the RAW parameters (the lower bounds
 * and the square-minus-pedestal reparametrisation of GDN.py:73-79 are applied inside). */
FVC_API int fvc_gdn(const float* x, const float* beta, const float* gamma, float* y, int B, int C, int H, int W,
            int inverse, void* stream);

/* round + BitEstimator likelihood + clamp-log2 sum — net.py:76/91, 153-178, 181-205,
 * bitEstimator.py:20-42.  params: 11 device pointers to [C] vectors in the order
 * f1.h f1.b f1.a f2.h f2.b f2.a f3.h f3.b f3.a f4.h f4.b.
 * x: [B,C,H,W] -> q_out: [B,C,H,W] (torch.round(x)), bits_out: one float (total bits). */
FVC_API int fvc_quant_bits_factorized(const float* x, const float* const* params, float* q_out, float* bits_out,
                              int B, int C, int H, int W, void* stream);

/* round + Laplace(0, clamp(sigma,1e-5,1e10)) likelihood + clamp-log2 sum — net.py:100, 121-151. */
FVC_API int fvc_quant_bits_laplace(const float* x, const float* sigma, float* q_out, float* bits_out, int64_t n,
                           void* stream);

/* recon = pred + res; clipped = clamp(recon,0,1); sums_out[3] = mean((recon-cur)^2),
 * mean((warp-cur)^2), mean((pred-cur)^2) — net.py:103-116.  All [n] fp32. */
FVC_API int fvc_recon_losses(const float* cur, const float* pred, const float* warp, const float* res,
                     float* clipped_out, float* means_out, int64_t n, void* stream);

/* CompressAI-compatible likelihoods used by entropy_models.py:55-68, 202-219 (RecProbModel,
 * MeanScaleHyperPriors) + get_estimate_bits (74-78).  Parity at this boundary is UNPINNED
 * (CompressAI is not vendored by the reference).
 * fvc_eb_forward: EntropyBottleneck eval forward with filters (3,3,3,3):
 *   matrices/biases/factors: packed per channel, medians: [C].
 *   x: [B,C,H,W] -> xhat_out, lik_out: [B,C,H,W]; bits_out: sum clamp(-log2(lik+1e-5),0,50). */
FVC_API int fvc_eb_forward(const float* x, const float* packed_params, const float* medians, float* xhat_out,
                   float* lik_out, float* bits_out, int B, int C, int H, int W, void* stream);
/* fvc_gaussian_forward: GaussianConditional eval forward (scale_bound 0.11, likelihood_bound 1e-9). */
FVC_API int fvc_gaussian_forward(const float* x, const float* scales, const float* means, float* xhat_out,
                         float* lik_out, float* bits_out, int64_t n, void* stream);

/* Real entropy coding of the quantised latents — the `calrealbits` branch of net.py:123-138 (feature under
 * Laplace(0, sigma)), 155-168 (z under bitEstimator_z), 183-195 (mv under bitEstimator_mv).
 * Model = the integer CDF the reference hands to torchac (un-vendored; its published conversion restated):
 *     Q[i] = round(F(i - mxrange - 0.5) * (2^16 - (2*mxrange - 1))) + i,  i in [0, 2*mxrange),  end of the last symbol 2^16;
 * symbol s = q + mxrange in [0, 2*mxrange - 2].  Coder = rANS (32-bit state, 16-bit words), one independent stream per
 * lane of lane_len consecutive symbols (NHWC order); container "FVR1" (fvc_entropy.cu).  The code length is within
 * 32 bits per lane + header of the ideal sum(-log2(freq / 2^16)) that torchac's arithmetic coder also attains.
 * Tables are uint32 [C][2*mxrange] (factorized, per channel) or [n][2*mxrange] (Laplace, per element: test use only). */
FVC_API int fvc_cdf_table_factorized(const float* const* params, int C, int mxrange, uint32_t* table_out, void* stream);
FVC_API int fvc_cdf_table_laplace(const float* sigma, int64_t n, int mxrange, uint32_t* table_out, void* stream);
/* x: the PRE-round latent, fp32, n values in coding order with channel = index % C; q = round(x) is coded.
 * stream_out (device, capacity >= fvc_entropy_stream_capacity(n, lane_len)); nbytes_out: device uint32;
 * err_out: device uint32[3] = symbols outside [-mxrange, mxrange-2], empty intervals, unreadable lanes. */
FVC_API int fvc_entropy_encode_factorized(const float* x, int64_t n, int C, const uint32_t* table, int mxrange,
                                          int lane_len, void* stream_out, int64_t capacity, uint32_t* nbytes_out,
                                          uint32_t* err_out, void* stream);
FVC_API int fvc_entropy_encode_laplace(const float* x, const float* sigma, int64_t n, int mxrange, int lane_len,
                                       void* stream_out, int64_t capacity, uint32_t* nbytes_out, uint32_t* err_out,
                                       void* stream);
FVC_API int fvc_entropy_decode_factorized(const void* stream_in, int64_t nbytes, int64_t n, int C, const uint32_t* table,
                                          int mxrange, int lane_len, float* q_out, uint32_t* err_out, void* stream);
FVC_API int fvc_entropy_decode_laplace(const void* stream_in, int64_t nbytes, int64_t n, const float* sigma, int mxrange,
                                       int lane_len, float* q_out, uint32_t* err_out, void* stream);
FVC_API int64_t fvc_entropy_stream_capacity(int64_t n, int lane_len);

/* Indexed-table coder: the model CompressAI's EntropyModel.compress / decompress codes under, which the reference's
 * RecProbModel / MeanScaleHyperPriors call (entropy_models.py:80-93, 237-247 -> entropy_bottleneck.compress,
 * gaussian_conditional.compress(x, indexes, means)).  CompressAI is an un-vendored dependency: its published
 * encode_with_indexes rule is restated (parity unpinned): element i uses table indexes[i] of cdf [ntab][cdf_stride]
 * (int32, 16-bit precision, cdf_length[t] valid entries, last bin = tail mass), v = symbols[i] - offset[t]; values outside
 * [0, cdf_length[t] - 2) are escaped through the last bin followed by the raw value in 4-bit bypass digits.  Coder and
 * container as above (rANS lanes, "FVR1"); symbols in the caller's order.  All pointers are device memory.
 * err_out: device uint32[3] = indexes outside [0, ntab), empty intervals (malformed table), unreadable lanes. */
FVC_API int fvc_entropy_encode_indexed(const int32_t* symbols, const int32_t* indexes, int64_t n, const int32_t* cdf,
                                       int ntab, int cdf_stride, const int32_t* cdf_length, const int32_t* offset,
                                       int lane_len, void* stream_out, int64_t capacity, uint32_t* nbytes_out,
                                       uint32_t* err_out, void* stream);
FVC_API int fvc_entropy_decode_indexed(const void* stream_in, int64_t nbytes, int64_t n, const int32_t* indexes,
                                       const int32_t* cdf, int ntab, int cdf_stride, const int32_t* cdf_length,
                                       const int32_t* offset, int lane_len, int32_t* symbols_out, uint32_t* err_out,
                                       void* stream);
FVC_API int64_t fvc_entropy_stream_capacity_indexed(int64_t n, int lane_len);

/* ---------------------------------------------------------------------------------------------
 * Whole-path context: VideoCompressor.forward — net.py:70-220.
 * ------------------------------------------------------------------------------------------- */
typedef struct fvc_ctx fvc_ctx;

/* Creates a context for frames [B,3,H,W]; H and W must be multiples of 64.
 * levels = SpyNet pyramid depth (reference: 4, endecoder.py:318).  impl = FVC_IMPL_*. */
FVC_API fvc_ctx* fvc_ctx_create(int B, int H, int W, int levels, int impl);
FVC_API void fvc_ctx_destroy(fvc_ctx* ctx);

/* Hands one reference state_dict entry (key as in SURVEY 8b, e.g. "mvEncoder.conv3.weight") to the
 * context.  The data is read (and packed) inside the call, stream-ordered; the caller keeps
 * ownership.  numel is checked against the expected shape. */
FVC_API int fvc_ctx_set_param(fvc_ctx* ctx, const char* key, const float* data, int64_t numel, void* stream);
/* Number of keys still missing (0 = ready). */
FVC_API int fvc_ctx_missing_params(fvc_ctx* ctx);

/* One P-frame.  cur, ref: [B,3,H,W] in [0,1].  recon_out: [B,3,H,W] (clamped reconstruction =
 * next reference).  scalars_out: 7 floats = mse_loss, warploss, interloss, bpp_feature, bpp_z,
 * bpp_mv, bpp (net.py:220).  Stream-ordered, no host synchronisation. */
FVC_API int fvc_pframe_forward(fvc_ctx* ctx, const float* cur, const float* ref, float* recon_out, float* scalars_out,
                       void* stream);

/* Copies a named intermediate of the last fvc_pframe_forward as fp32 NCHW (testing/inspection):
 * estmv mvfeature quant_mv mv_hat warpframe prediction feature z z_hat sigma feat_hat recon_res.
 * Returns the element count, or <0. */
FVC_API int64_t fvc_ctx_get_tensor(fvc_ctx* ctx, const char* name, float* out, int64_t capacity, void* stream);

/* Closed-loop GOP from HOST memory — models.py:368-383 (parallel_compression, 'DVC-pretrained'):
 * frames_host: [G,B,3,H,W] (pinned or pageable), frame 0 is the decoded I-frame.
 * recon_host: [G-1,B,3,H,W] or NULL.  scalars_host: [G-1,7].  Copies H2D, runs G-1 P-frames,
 * copies results D2H and synchronises the stream before returning. */
FVC_API int fvc_gop_forward_host(fvc_ctx* ctx, const float* frames_host, int G, float* recon_host, float* scalars_host,
                         void* stream);
/* The same GOP call fed with the frames as the reference's loader holds them BEFORE transforms.ToTensor()
 * (dataset.py:68-75: decoded image -> PIL -> ToTensor): frames_host_u8 [G,B,H,W,3], uint8, channel-interleaved.
 * Uploads 1 byte per sample instead of 4 and applies ToTensor (HWC -> CHW, x / 255) on the device. */
FVC_API int fvc_gop_forward_host_u8(fvc_ctx* ctx, const uint8_t* frames_host_u8, int G, float* recon_host,
                                    float* scalars_host, void* stream);

/* Intra frame on the device.  The reference codes frame 0 of a GOP by shelling out to bpgenc / bpgdec
 * through temporary JPEG files (I_compression, models.py:412-429: external binaries, no learned intra codec).  This is
 * the replacement the survey proposes, "a hyperprior image codec from the same conv engine": the residual branch of
 * VideoCompressor.forward (net.py:86-116: resEncoder, respriorEncoder, respriorDecoder, resDecoder, both bit
 * estimators) applied to the frame itself, i.e. the P-frame forward with a zero prediction and no motion branch; same
 * weights, same kernels.  scalars_out[7] keeps the P-frame layout: mse, mean(x^2), mean(x^2), bpp_feature, bpp_z, 0, bpp.
 * With fvc_ctx_set_realbits the two latents are entropy-coded (streams 0 and 1 of fvc_ctx_get_bitstream) and
 * fvc_iframe_decode_bitstreams reproduces recon_out from them bit for bit. */
FVC_API int fvc_iframe_forward(fvc_ctx* ctx, const float* frame, float* recon_out, float* scalars_out, void* stream);
FVC_API int fvc_iframe_decode_bitstreams(fvc_ctx* ctx, const void* feat_stream, int64_t feat_bytes, const void* z_stream,
                                         int64_t z_bytes, float* recon_out, void* stream);

/* LSVC (reference models.py:1157-1411, non-attention "-128" variants: the same sub-networks as DVC) codes the
 * P-frames of a GOP in two phases (LSVC.forward, models.py:1344-1411):
 *   phase A: opticFlow + mv_codec on ALL frames at once against their ORIGINAL reference frames
 *            cur, ref: [B,3,H,W] -> mv_hat_out: [B,2,H,W] (decoded motion), bits_mv_out: 1 float (sum over B);
 *   phase B: per tree layer, motioncompensation + res_codec against the RECONSTRUCTED references
 *            cur, ref, mv_hat -> com_out = clip(MC + res_hat), mc_out = MC frames, warp_out = warped frames
 *            (all [B,3,H,W]); sums_out: 5 floats = sum((com-cur)^2), sum((warp-cur)^2), sum((mc-cur)^2),
 *            bits_feature, bits_z (sums over B).
 * B is the context's batch; use one context per batch size. */
FVC_API int fvc_lsvc_mv_forward(fvc_ctx* ctx, const float* cur, const float* ref, float* mv_hat_out,
                                float* bits_mv_out, void* stream);
FVC_API int fvc_lsvc_mc_res_forward(fvc_ctx* ctx, const float* cur, const float* ref, const float* mv_hat,
                                    float* com_out, float* mc_out, float* warp_out, float* sums_out, void* stream);

/* Decoder half of VideoCompressor.forward — net.py:77-80 (mvDecoder, motioncompensation) and net.py:101-105
 * (resDecoder, recon = prediction + recon_res, clamp): what a receiver computes from the entropy-decoded latents.
 * ref: [B,3,H,W]; quant_mv: [B,128,H/16,W/16]; feat_hat: [B,96,H/16,W/16] (integer-valued fp32, NCHW)
 * -> recon_out: [B,3,H,W] (clamped).  Stream-ordered. */
FVC_API int fvc_decode_from_latents(fvc_ctx* ctx, const float* ref, const float* quant_mv, const float* feat_hat,
                                    float* recon_out, void* stream);

/* Teacher forcing for tests / inspection: until cleared (NULL), every following forward replaces the output of the
 * three quantisers (net.py:76 quant_mv [B,128,H/16,W/16], net.py:91 z_hat [B,64,H/64,W/64], net.py:100 feat_hat
 * [B,96,H/16,W/16]; fp32 NCHW device tensors owned by the caller) by the given tensors; the bit estimates are
 * still those of the free-running quantisers. */
FVC_API int fvc_ctx_force_latents(fvc_ctx* ctx, const float* quant_mv, const float* z_hat, const float* feat_hat);

/* Range check of the tcgen05 engine's fp16 operand pairs: number of epilogue tiles (since creation / last reset)
 * that stored an activation with |v| >= 65504 (clamped: the reference is fp32, so the result is then wrong).
 * While it is non-zero, fvc_pframe_forward writes NaN into scalars_out and fvc_gop_forward_host fails with
 * FVC_ERR_STATE.  Synchronises the stream. */
FVC_API int64_t fvc_ctx_saturation_count(fvc_ctx* ctx, int reset, void* stream);

/* calrealbits (net.py:57): when enabled, every following fvc_pframe_forward entropy-codes the three quantised latents
 * on the GPU (stream-ordered, no host synchronisation) and bpp_feature / bpp_z / bpp_mv / bpp in scalars_out are the
 * REAL bits (8 x stream bytes, net.py:136) instead of the estimates.  mxrange = VideoCompressor.mxrange (150). */
FVC_API int fvc_ctx_set_realbits(fvc_ctx* ctx, int enable, int mxrange);
/* Byte stream of the last forward: which = 0 feature, 1 z, 2 mv.  out_host may be NULL (size query).  Returns the
 * byte count; fails with FVC_ERR_STATE if a symbol was not codable.  Synchronises the stream. */
FVC_API int64_t fvc_ctx_get_bitstream(fvc_ctx* ctx, int which, void* out_host, int64_t capacity, void* stream);
/* The receiver: the three streams (DEVICE pointers) + the reference frame -> clamped reconstruction; decodes z,
 * runs respriorDecoder for sigma, decodes feature and mv, then the path of fvc_decode_from_latents. */
FVC_API int fvc_decode_bitstreams(fvc_ctx* ctx, const float* ref, const void* feat_stream, int64_t feat_bytes,
                                  const void* z_stream, int64_t z_bytes, const void* mv_stream, int64_t mv_bytes,
                                  float* recon_out, void* stream);

/* Launch statistics since creation: kernels launched by this library through ctx. */
FVC_API int64_t fvc_ctx_launch_count(fvc_ctx* ctx);
/* Dominant-kernel timing hook for bench.py: seconds spent in convolution kernels during the last
 * fvc_pframe_forward when FVC_PROFILE=1 was set at create time (uses CUDA events; else -1). */
FVC_API double fvc_ctx_last_conv_seconds(fvc_ctx* ctx);
/* Per-layer convolution times of the last profiled forward: lines "<layer> <ms>" (FVC_PROFILE=1). */
FVC_API const char* fvc_ctx_profile_text(fvc_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* FVC_B200_H_ */
