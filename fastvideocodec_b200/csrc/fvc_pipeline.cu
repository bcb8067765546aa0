// Whole-path context: VideoCompressor.forward (reference DVC/net.py:70-220) as a stream-ordered
// sequence of sm_100a kernels over library-owned buffers.  The context consumes the reference
// state_dict layout directly and packs weights on the device.
#include <map>
#include <string>
#include <vector>
#include <cstdarg>
#include <cstring>

#include "fvc_kernels.cuh"

namespace fvc {

thread_local std::string g_err;
thread_local int64_t g_launch_count = 0;

bool pdl_enabled() {
    // opt-in: measured neutral to -1 % on the power-capped 1080p frame (profiles/r02_pdl_ab.txt)
    static const bool on = [] { const char* v = getenv("FVC_PDL"); return v && v[0] == '1'; }();
    return on;
}

void set_error(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
}
int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
    set_error("CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
    return FVC_ERR_CUDA;
}

// ----------------------------------------------------------------------------------------------
struct ConvRt {
    std::string name;
    ConvLayer L;
    int CinP = 0;      // channel count of the input records
    int CoutS = 0;     // padded output-channel count of the packed weights
    int act = FVC_ACT_NONE;
    float* w_raw = nullptr;  // device copy of the reference-layout weight
    float* bias = nullptr;   // device [Cout]
    bool have_w = false, have_b = false;
    SimtWeights simt;
    TcPlan* tc = nullptr;
    float* few_w = nullptr;  // packed weights of the few-output-channel fp32 path
};

struct GdnRt {
    std::string name;
    int C = 64, inverse = 0;
    float *beta_raw = nullptr, *gamma_raw = nullptr, *beta_eff = nullptr, *gamma_eff = nullptr;
    bool have_b = false, have_g = false;
};

struct BitEstRt {
    std::string name;
    int C = 0;
    float* p[11] = {nullptr};
    bool have[11] = {false};
};

static int pad_c(int c) { return c <= 32 ? 32 : (c <= 64 ? 64 : 128); }

}  // namespace fvc

using namespace fvc;

struct fvc_ctx {
    int B, H, W, levels, impl;
    int cp_narrow = 32;
    std::vector<void*> allocs;
    std::map<std::string, ConvRt> conv;
    std::map<std::string, GdnRt> gdn;
    BitEstRt be_z, be_mv;
    int64_t launches = 0;
    bool profile = false;
    bool use_few = true;     // FVC_FEW=0 routes the 2-3 output-channel layers through the tensor-core engine
    bool gdn_fused = true;   // FVC_GDN_FUSED=0: (I)GDN as a separate 1x1 "norm" convolution launch (round-1 form)
    bool tail_fused = true;  // FVC_TAIL_FUSED=0: mvDecoder.deconv8 / warpnet.conv6 as their own 3x3 convolution launches
    bool tail_fused_warpnet = true;
    float* taps_buf = nullptr;   // per-pixel partial products of a fused tail convolution, fp32 [B,H,W,<=28]
    double last_conv_seconds = -1.0;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> conv_events;
    std::vector<std::string> conv_event_names;
    size_t conv_event_used = 0;
    std::string profile_text;

    // ---- buffers ---------------------------------------------------------------------------
    std::vector<float*> pyr1, pyr2;                 // planar pyramids, scale 1..levels-1 (scale 0 = user ptr)
    std::vector<ActT> sx, sa1, sa2, sa3, sa4;       // per SpyNet level
    std::vector<float*> sflow_up, sflow;            // fp32 NHWC2
    ActT estmv_act;                                  // parity layout, input of mvEncoder.conv1
    ActT e[8];                                       // mvEncoder activations e[1..7]
    float* mvfeature = nullptr;                      // fp32 NHWC128
    ActT quant_mv;
    ActT d[8];                                       // mvDecoder activations d[1..7]
    float* mv_hat = nullptr;                         // fp32 NHWC2
    float* warpframe = nullptr;                      // planar
    ActT xmc;
    ActT wf, wt0, wc0, wc0p, wc0p_r, wt1, wc1, wc1p, wc1p_r, wt2, wc2, wc2_r, wt3, wc3, wc3u, wc3u_r, wt4, wc4,
        wc4u, wc4u_r, wt5, wc5;
    float* wres = nullptr;                           // fp32 NHWC3
    float* prediction = nullptr;                     // planar
    ActT residual;                                   // parity
    ActT r_raw[3], r[3];                             // resEncoder: conv output, GDN output
    ActT r_sq[3], g_sq[3];                           // squares of the conv outputs (tcgen05 GDN path)
    float* feature = nullptr;                        // fp32 NHWC96
    ActT featabs;
    ActT p1, p2;
    float* z = nullptr;                              // fp32 NHWC64
    ActT z_hat;
    ActT s1, s2;
    float* sigma = nullptr;                          // fp32 NHWC96
    ActT feat_hat;
    ActT g_raw[3], g[3];
    float* recon_res = nullptr;                      // fp32 NHWC3
    float* loss_partials = nullptr;
    float* bits_partials = nullptr;
    int* bits_counts = nullptr;                      // device [3]
    float* scalars = nullptr;                        // device [7] (internal copy)
    float *stage_cur = nullptr, *stage_ref = nullptr, *stage_rec = nullptr;  // GOP driver staging
    cudaStream_t copy_stream = nullptr;              // H2D of the GOP's frames, overlapped with the computation
    // real entropy coding runs beside the rest of the frame: the coders are serial chains on a handful of warps (0.8 ms
    // each at 1080p) whose inputs are final when they are launched and whose outputs are only read after the frame
    cudaStream_t ent_stream = nullptr;
    cudaEvent_t ent_fork = nullptr, ent_done = nullptr;
    bool ent_pending = false;
    cudaEvent_t copy_fence = nullptr;
    std::vector<cudaEvent_t> copy_events;
    int stage_G = 0, stage_G_u8 = 0;
    uint8_t* stage_u8 = nullptr;          // fvc_gop_forward_host_u8: the uploaded uint8 HWC frames
    float* stage_frames = nullptr;
    float* stage_scalars = nullptr;
    // teacher forcing (tests / inspection): device fp32 NCHW tensors that replace the quantised latents
    // quant_mv, z_hat, feat_hat right after the quantisers (nullptr = free running)
    const float* force_q[3] = {nullptr, nullptr, nullptr};
    // number of epilogue tiles that produced an activation at or beyond the fp16 pair range (|v| >= 65504):
    // those values were clamped, the results are wrong; scalars turn NaN and the host entry points fail
    unsigned int* sat_count = nullptr;
    // real entropy coding (calrealbits, net.py:123-138 / 155-168 / 183-195): allocated on first use
    int realbits = 0, mxrange = 150, rans_L = 8192;
    uint32_t *cdf_tab_mv = nullptr, *cdf_tab_z = nullptr;   // [C][2R] integer CDFs of the two BitEstimators
    uint32_t* sym_packed = nullptr;                          // 2 words per symbol: start | freq << 16, reciprocal (largest latent)
    uint16_t* rans_words = nullptr;
    uint32_t* rans_lane_words = nullptr;
    uint8_t* stream[3] = {nullptr, nullptr, nullptr};        // 0: feature, 1: z, 2: mv
    size_t stream_cap[3] = {0, 0, 0};
    uint32_t* stream_bytes = nullptr;                        // device [3]
    unsigned int* ent_err = nullptr;                         // device [3]: out-of-range symbols, empty intervals, bad lanes
    int ent_R = 0, ent_L = 0;

    template <typename T>
    int alloc(T** p, size_t bytes) {
        void* q = nullptr;
        FVC_CUDA(cudaMalloc(&q, bytes ? bytes : 16));
        FVC_CUDA(cudaMemset(q, 0, bytes ? bytes : 16));   // padded channels of ACT records stay finite (zero)
        allocs.push_back(q);
        *p = reinterpret_cast<T*>(q);
        return 0;
    }
    // geometry only, storage borrowed from `like` (a tensor that is never written or read through this handle: it only
    // shapes the TMA descriptors of a plan that is created for its packed weights and never launched)
    void alias_act(ActT* t, int h, int w, int Cp, int parity, void* like) {
        t->B = B; t->H = h; t->W = w; t->Cp = Cp; t->parity = parity;
        t->p = reinterpret_cast<e16*>(like);
    }
    int alloc_act(ActT* t, int h, int w, int Cp, int parity) {
        t->B = B; t->H = h; t->W = w; t->Cp = Cp; t->parity = parity;
        return alloc(&t->p, act_bytes(B, h, w, Cp));
    }
};

namespace fvc {

static void add_conv(fvc_ctx* c, const std::string& name, int Cin, int Cout, int k, int stride, int transposed,
                     int act, int CinP) {
    ConvRt r;
    r.name = name;
    make_conv_layer(r.L, Cin, Cout, k, stride, transposed);
    r.CinP = CinP;
    r.CoutS = std::max(pad_c(Cout), 32);
    r.act = act;
    c->conv[name] = r;
}

static int build_layers(fvc_ctx* c) {
    char buf[128];
    // records of the network's input layers (<= 8 real channels): narrow [hi 8 | lo 8] records on the
    // tensor-core engine (one 16-element k-step per tap and product pair instead of a 32-channel padded one)
    const int cin_narrow = c->impl != FVC_IMPL_SIMT ? 8 : 32;
    c->cp_narrow = cin_narrow;
    const int sp[5][2] = {{8, 32}, {32, 64}, {64, 32}, {32, 16}, {16, 2}};
    for (int l = 0; l < c->levels; ++l)
        for (int i = 0; i < 5; ++i) {
            snprintf(buf, sizeof(buf), "opticFlow.moduleBasic.%d.conv%d", l, i + 1);
            add_conv(c, buf, sp[i][0], sp[i][1], 7, 1, 0, i < 4 ? FVC_ACT_RELU : FVC_ACT_NONE,
                     i == 0 ? cin_narrow : pad_c(sp[i][0]));
        }
    for (int i = 1; i <= 8; ++i) {
        snprintf(buf, sizeof(buf), "mvEncoder.conv%d", i);
        add_conv(c, buf, i == 1 ? 2 : 128, 128, 3, (i & 1) ? 2 : 1, 0, i < 8 ? FVC_ACT_LRELU01 : FVC_ACT_NONE,
                 i == 1 ? cin_narrow : 128);
    }
    for (int i = 1; i <= 8; ++i) {
        snprintf(buf, sizeof(buf), "mvDecoder.deconv%d", i);
        if (i & 1) add_conv(c, buf, 128, 128, 3, 2, 1, FVC_ACT_LRELU01, 128);
        else add_conv(c, buf, 128, i == 8 ? 2 : 128, 3, 1, 0, i < 8 ? FVC_ACT_LRELU01 : FVC_ACT_NONE, 128);
    }
    add_conv(c, "warpnet.feature_ext", 6, 64, 3, 1, 0, FVC_ACT_RELU, cin_narrow);
    for (int i = 0; i < 6; ++i) {
        snprintf(buf, sizeof(buf), "warpnet.conv%d.conv1", i);
        add_conv(c, buf, 64, 64, 3, 1, 0, FVC_ACT_RELU, 64);  // relu2 fused into conv1's epilogue
        snprintf(buf, sizeof(buf), "warpnet.conv%d.conv2", i);
        add_conv(c, buf, 64, 64, 3, 1, 0, FVC_ACT_NONE, 64);
    }
    add_conv(c, "warpnet.conv6", 64, 3, 3, 1, 0, FVC_ACT_NONE, 64);
    add_conv(c, "resEncoder.conv1", 3, 64, 5, 2, 0, FVC_ACT_NONE, cin_narrow);
    add_conv(c, "resEncoder.conv2", 64, 64, 5, 2, 0, FVC_ACT_NONE, 64);
    add_conv(c, "resEncoder.conv3", 64, 64, 5, 2, 0, FVC_ACT_NONE, 64);
    add_conv(c, "resEncoder.conv4", 64, 96, 5, 2, 0, FVC_ACT_NONE, 64);
    add_conv(c, "resDecoder.deconv1", 96, 64, 5, 2, 1, FVC_ACT_NONE, 128);
    add_conv(c, "resDecoder.deconv2", 64, 64, 5, 2, 1, FVC_ACT_NONE, 64);
    add_conv(c, "resDecoder.deconv3", 64, 64, 5, 2, 1, FVC_ACT_NONE, 64);
    add_conv(c, "resDecoder.deconv4", 64, 3, 5, 2, 1, FVC_ACT_NONE, 64);
    add_conv(c, "respriorEncoder.conv1", 96, 64, 3, 1, 0, FVC_ACT_RELU, 128);
    add_conv(c, "respriorEncoder.conv2", 64, 64, 5, 2, 0, FVC_ACT_RELU, 64);
    add_conv(c, "respriorEncoder.conv3", 64, 64, 5, 2, 0, FVC_ACT_NONE, 64);
    add_conv(c, "respriorDecoder.deconv1", 64, 64, 5, 2, 1, FVC_ACT_RELU, 64);
    add_conv(c, "respriorDecoder.deconv2", 64, 64, 5, 2, 1, FVC_ACT_RELU, 64);
    add_conv(c, "respriorDecoder.deconv3", 64, 96, 3, 1, 1, FVC_ACT_EXP, 64);
    for (int i = 1; i <= 3; ++i) {
        GdnRt g;
        snprintf(buf, sizeof(buf), "resEncoder.gdn%d", i);
        g.name = buf; g.inverse = 0;
        c->gdn[buf] = g;
        snprintf(buf, sizeof(buf), "resDecoder.igdn%d", i);
        g.name = buf; g.inverse = 1;
        c->gdn[buf] = g;
        // tcgen05 engine: (I)GDN norm = 1x1 convolution of the squares with gamma_eff, bias beta_eff
        snprintf(buf, sizeof(buf), "resEncoder.gdn%d#norm", i);
        add_conv(c, buf, 64, 64, 1, 1, 0, FVC_ACT_NONE, 64);
        snprintf(buf, sizeof(buf), "resDecoder.igdn%d#norm", i);
        add_conv(c, buf, 64, 64, 1, 1, 0, FVC_ACT_NONE, 64);
    }
    // fused 3x3 tails (tcgen05 engine): the tail's weights regrouped as a 1x1 convolution with (tap, channel) outputs
    add_conv(c, "mvDecoder.deconv8#taps", 128, 32, 1, 1, 0, FVC_ACT_NONE, 128);
    add_conv(c, "warpnet.conv6#taps", 64, 32, 1, 1, 0, FVC_ACT_NONE, 64);
    c->be_z.name = "bitEstimator_z"; c->be_z.C = 64;
    c->be_mv.name = "bitEstimator_mv"; c->be_mv.C = 128;
    // device storage for parameters
    for (auto& kv : c->conv) {
        ConvRt& r = kv.second;
        size_t nw = (size_t)r.L.Cin * r.L.Cout * r.L.k * r.L.k;
        if (c->alloc(&r.w_raw, nw * sizeof(float))) return FVC_ERR_CUDA;
        if (c->alloc(&r.bias, (size_t)r.L.Cout * sizeof(float))) return FVC_ERR_CUDA;
    }
    for (const char* tn : {"mvDecoder.deconv8#taps", "warpnet.conv6#taps"}) c->conv[tn].have_b = true;   // zero bias
    for (auto& kv : c->gdn) {
        GdnRt& g = kv.second;
        if (c->alloc(&g.beta_raw, 64 * 4) || c->alloc(&g.gamma_raw, 64 * 64 * 4) || c->alloc(&g.beta_eff, 64 * 4) ||
            c->alloc(&g.gamma_eff, 64 * 64 * 4))
            return FVC_ERR_CUDA;
    }
    for (BitEstRt* be : {&c->be_z, &c->be_mv})
        for (int i = 0; i < 11; ++i)
            if (c->alloc(&be->p[i], (size_t)be->C * 4)) return FVC_ERR_CUDA;
    return 0;
}

static int build_buffers(fvc_ctx* c) {
    const int B = c->B, H = c->H, W = c->W, L = c->levels;
#define A(expr) do { if (expr) return FVC_ERR_CUDA; } while (0)
    c->pyr1.assign(L, nullptr);
    c->pyr2.assign(L, nullptr);
    for (int s = 1; s < L; ++s) {
        A(c->alloc(&c->pyr1[s], (size_t)B * 3 * (H >> s) * (W >> s) * 4));
        A(c->alloc(&c->pyr2[s], (size_t)B * 3 * (H >> s) * (W >> s) * 4));
    }
    c->sx.resize(L); c->sa1.resize(L); c->sa2.resize(L); c->sa3.resize(L); c->sa4.resize(L);
    c->sflow_up.assign(L, nullptr); c->sflow.assign(L, nullptr);
    for (int i = 0; i < L; ++i) {
        int h = H >> (L - 1 - i), w = W >> (L - 1 - i);
        A(c->alloc_act(&c->sx[i], h, w, c->cp_narrow, 0));
        A(c->alloc_act(&c->sa1[i], h, w, 32, 0));
        A(c->alloc_act(&c->sa2[i], h, w, 64, 0));
        A(c->alloc_act(&c->sa3[i], h, w, 32, 0));
        A(c->alloc_act(&c->sa4[i], h, w, 32, 0));
        A(c->alloc(&c->sflow_up[i], (size_t)B * h * w * 2 * 4));
        A(c->alloc(&c->sflow[i], (size_t)B * h * w * 2 * 4));
    }
    A(c->alloc_act(&c->estmv_act, H, W, c->cp_narrow, 1));
    // mvEncoder: conv i output at H >> ceil(i/2); parity layout when the consumer is a stride-2 conv
    for (int i = 1; i <= 7; ++i) {
        int sh = (i + 1) / 2;
        A(c->alloc_act(&c->e[i], H >> sh, W >> sh, 128, (i % 2 == 0) ? 1 : 0));
    }
    A(c->alloc(&c->mvfeature, (size_t)B * (H / 16) * (W / 16) * 128 * 4));
    A(c->alloc_act(&c->quant_mv, H / 16, W / 16, 128, 0));
    for (int i = 1; i <= 7; ++i) {
        int sh = 4 - (i + 1) / 2;  // deconv1 -> H/8, deconv2 -> H/8, deconv3 -> H/4, ...
        if (i == 7 && c->tail_fused) continue;   // deconv7's output stays in the SM (fused deconv8): 1.07 GB at 1080p
        A(c->alloc_act(&c->d[i], H >> sh, W >> sh, 128, 0));
    }
    A(c->alloc(&c->mv_hat, (size_t)B * H * W * 2 * 4));
    A(c->alloc(&c->warpframe, (size_t)B * 3 * H * W * 4));
    A(c->alloc_act(&c->xmc, H, W, c->cp_narrow, 0));
    A(c->alloc_act(&c->wf, H, W, 64, 0));
    A(c->alloc_act(&c->wt0, H, W, 64, 0));
    A(c->alloc_act(&c->wc0, H, W, 64, 0));
    A(c->alloc_act(&c->wc0p, H / 2, W / 2, 64, 0));
    A(c->alloc_act(&c->wc0p_r, H / 2, W / 2, 64, 0));
    A(c->alloc_act(&c->wt1, H / 2, W / 2, 64, 0));
    A(c->alloc_act(&c->wc1, H / 2, W / 2, 64, 0));
    A(c->alloc_act(&c->wc1p, H / 4, W / 4, 64, 0));
    A(c->alloc_act(&c->wc1p_r, H / 4, W / 4, 64, 0));
    A(c->alloc_act(&c->wt2, H / 4, W / 4, 64, 0));
    A(c->alloc_act(&c->wc2, H / 4, W / 4, 64, 0));
    A(c->alloc_act(&c->wc2_r, H / 4, W / 4, 64, 0));
    A(c->alloc_act(&c->wt3, H / 4, W / 4, 64, 0));
    A(c->alloc_act(&c->wc3, H / 4, W / 4, 64, 0));
    A(c->alloc_act(&c->wc3u, H / 2, W / 2, 64, 0));
    A(c->alloc_act(&c->wc3u_r, H / 2, W / 2, 64, 0));
    A(c->alloc_act(&c->wt4, H / 2, W / 2, 64, 0));
    A(c->alloc_act(&c->wc4, H / 2, W / 2, 64, 0));
    A(c->alloc_act(&c->wc4u, H, W, 64, 0));
    A(c->alloc_act(&c->wc4u_r, H, W, 64, 0));
    A(c->alloc_act(&c->wt5, H, W, 64, 0));
    if (!c->tail_fused_warpnet) A(c->alloc_act(&c->wc5, H, W, 64, 0));   // else: stays in the SM (fused conv6)
    A(c->alloc(&c->wres, (size_t)B * H * W * 3 * 4));
    A(c->alloc(&c->prediction, (size_t)B * 3 * H * W * 4));
    A(c->alloc_act(&c->residual, H, W, c->cp_narrow, 1));
    for (int i = 0; i < 3; ++i) {
        A(c->alloc_act(&c->r[i], H >> (i + 1), W >> (i + 1), 64, 1));
        if (c->gdn_fused) {   // raw conv output and its squares never leave the SM: geometry only
            c->alias_act(&c->r_raw[i], H >> (i + 1), W >> (i + 1), 64, 0, c->r[i].p);
            c->alias_act(&c->r_sq[i], H >> (i + 1), W >> (i + 1), 64, 0, c->r[i].p);
        } else {
            A(c->alloc_act(&c->r_raw[i], H >> (i + 1), W >> (i + 1), 64, 0));
            A(c->alloc_act(&c->r_sq[i], H >> (i + 1), W >> (i + 1), 64, 0));
        }
    }
    A(c->alloc(&c->feature, (size_t)B * (H / 16) * (W / 16) * 96 * 4));
    A(c->alloc_act(&c->featabs, H / 16, W / 16, 128, 0));
    A(c->alloc_act(&c->p1, H / 16, W / 16, 64, 1));
    A(c->alloc_act(&c->p2, H / 32, W / 32, 64, 1));
    A(c->alloc(&c->z, (size_t)B * (H / 64) * (W / 64) * 64 * 4));
    A(c->alloc_act(&c->z_hat, H / 64, W / 64, 64, 0));
    A(c->alloc_act(&c->s1, H / 32, W / 32, 64, 0));
    A(c->alloc_act(&c->s2, H / 16, W / 16, 64, 0));
    A(c->alloc(&c->sigma, (size_t)B * (H / 16) * (W / 16) * 96 * 4));
    A(c->alloc_act(&c->feat_hat, H / 16, W / 16, 128, 0));
    for (int i = 0; i < 3; ++i) {
        A(c->alloc_act(&c->g[i], H >> (3 - i), W >> (3 - i), 64, 0));
        if (c->gdn_fused) {
            c->alias_act(&c->g_raw[i], H >> (3 - i), W >> (3 - i), 64, 0, c->g[i].p);
            c->alias_act(&c->g_sq[i], H >> (3 - i), W >> (3 - i), 64, 0, c->g[i].p);
        } else {
            A(c->alloc_act(&c->g_raw[i], H >> (3 - i), W >> (3 - i), 64, 0));
            A(c->alloc_act(&c->g_sq[i], H >> (3 - i), W >> (3 - i), 64, 0));
        }
    }
    A(c->alloc(&c->recon_res, (size_t)B * H * W * 3 * 4));
    A(c->alloc(&c->loss_partials, (size_t)148 * 8 * 3 * 4));
    A(c->alloc(&c->bits_partials, (size_t)bits_max_blocks() * 3 * 4));
    A(c->alloc(&c->bits_counts, 3 * sizeof(int)));
    A(c->alloc(&c->scalars, 8 * 4));
    A(c->alloc(&c->sat_count, 4));
    if (c->impl != FVC_IMPL_SIMT) A(c->alloc(&c->taps_buf, (size_t)B * H * W * 28 * 4));
    if (c->tail_fused) {   // geometry of the two tensors that stay in the SM (they shape never-launched plans)
        c->alias_act(&c->d[7], H, W, 128, 0, c->taps_buf);
        if (c->tail_fused_warpnet) c->alias_act(&c->wc5, H, W, 64, 0, c->taps_buf);
    }
#undef A
    return 0;
}

// ----------------------------------------------------------------------------------------------
// FVC_PROFILE=1: every launch of the frame is bracketed by CUDA events on the launching stream.  Convolutions are
// listed under their layer name, the memory-bound kernels under "@<kernel>[:<tag>]".
struct ProfScope {
    fvc_ctx* c;
    cudaStream_t s;
    cudaEvent_t e1 = nullptr;
    ProfScope(fvc_ctx* ctx, const std::string& name, cudaStream_t st) : c(ctx), s(st) {
        if (!c->profile) return;
        if (c->conv_event_used == c->conv_events.size()) {
            cudaEvent_t a = nullptr, b = nullptr;
            if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) return;
            c->conv_events.push_back({a, b});
        }
        cudaEvent_t e0 = c->conv_events[c->conv_event_used].first;
        e1 = c->conv_events[c->conv_event_used].second;
        if (c->conv_event_names.size() <= c->conv_event_used) c->conv_event_names.resize(c->conv_event_used + 1);
        c->conv_event_names[c->conv_event_used] = name;
        c->conv_event_used++;
        cudaEventRecord(e0, s);
    }
    ~ProfScope() {
        if (e1) cudaEventRecord(e1, s);
    }
};
// memory-bound kernel launch with profiling scope (uses the enclosing function's `rc`, `c`, `s`)
#define PK(tag, expr) do { ProfScope _ps(c, tag, s); rc = (expr); if (rc) return rc; } while (0)

static ActT no_act() {
    ActT t;
    t.p = nullptr; t.B = t.H = t.W = t.Cp = t.parity = 0;
    return t;
}
static Epilogue make_ep(const ConvRt& r) {
    Epilogue ep;
    memset(&ep, 0, sizeof(ep));
    ep.bias = r.bias;
    ep.acc_scale = 1.f;
    ep.act = r.act;
    ep.res_act = no_act();
    ep.out_act = no_act();
    ep.out_act_relu = no_act();
    ep.out_act_sq = no_act();
    ep.sq_scale = 1.f;
    ep.res_mode = 0;
    return ep;
}

static int run_conv(fvc_ctx* c, const std::string& name, ActT in, int Hout, int Wout, Epilogue ep,
                    cudaStream_t s) {
    auto it = c->conv.find(name);
    if (it == c->conv.end()) {
        set_error("unknown layer %s", name.c_str());
        return FVC_ERR_STATE;
    }
    ConvRt& r = it->second;
    if (!r.have_w || !r.have_b) {
        set_error("parameters of %s not set", name.c_str());
        return FVC_ERR_STATE;
    }
    ep.bias = r.bias;
    ep.act = r.act;
    ep.sat_count = c->sat_count;
    ProfScope prof(c, name, s);
    int rc;
    if (c->impl != FVC_IMPL_SIMT && c->use_few && r.few_w && !in.parity && ep.act == FVC_ACT_NONE && !ep.res_act.p &&
        !ep.out_act_relu.p && !ep.out_act_sq.p && ep.out_f32) {
        rc = launch_conv_few(r.L, r.few_w, r.bias, in, Hout, Wout, ep, s);
    } else if (c->impl != FVC_IMPL_SIMT && tc_supported(r.L, r.CinP)) {
        if (!r.tc) {
            rc = tc_plan_create(r.L, r.w_raw, in, Hout, Wout, ep, &r.tc, s, c->impl == FVC_IMPL_TC_FAST);
            if (rc) return rc;
        }
        rc = tc_plan_launch(r.tc, s);
    } else {
        if (ep.gdn_beta) {
            set_error("fused GDN requested on the SIMT engine (%s)", name.c_str());
            return FVC_ERR_STATE;
        }
        rc = launch_conv_simt(r.L, r.simt, in, Hout, Wout, ep, s);
    }
    return rc;
}

// conv followed by (I)GDN: conv kernel writes the raw activations, a second kernel normalises
// (fusing the 64x64 GDN matvec into the tcgen05 epilogue is future work; these layers are 3 % of the FLOPs)
static int run_conv_gdn(fvc_ctx* c, const std::string& conv, const std::string& gdn, ActT in, ActT raw, ActT sq,
                        ActT out, cudaStream_t s) {
    GdnRt& g = c->gdn[gdn];
    if (!g.have_b || !g.have_g) {
        set_error("parameters of %s not set", gdn.c_str());
        return FVC_ERR_STATE;
    }
    ConvRt& r = c->conv[conv];
    Epilogue ep = make_ep(r);
    ep.out_act = raw;
    const bool tc = c->impl != FVC_IMPL_SIMT;
    // squares are stored scaled by 2^-6 so that |x| up to ~2000 stays inside the fp16 range of the hi half
    const float sq_scale = 1.f / 64.f;
    if (tc && c->gdn_fused) {
        // (I)GDN inside the producing convolution's epilogue (north_star; GDN.py:63-93): the squared tile goes to shared
        // memory as the A operand of a second MMA against gamma_eff, the stand-alone "#norm" convolution's packed
        // weight stream (always the 3-product hi/lo form, also in precision 'fast').  The plan of that convolution
        // is created (never launched) to own the stream and its scale.
        const std::string nname = gdn + "#norm";
        ConvRt& n = c->conv[nname];
        if (!n.tc) {
            Epilogue en = make_ep(n);
            en.bias = n.bias;
            en.act = FVC_ACT_NONE;
            en.acc_scale = 1.f / sq_scale;
            en.res_act = raw;
            en.res_mode = g.inverse ? 2 : 1;
            en.out_act = out;
            int rc = tc_plan_create(n.L, n.w_raw, sq, out.H, out.W, en, &n.tc, s, false);
            if (rc) return rc;
        }
        if (tc_plan_is_gdn_norm_layout(n.tc)) {
            Epilogue ef = make_ep(r);
            ef.out_act = out;
            ef.sq_scale = sq_scale;
            ef.gdn_beta = n.bias;
            ef.gdn_gamma = tc_plan_wstream(n.tc);
            ef.gdn_scale = tc_plan_acc_scale(n.tc);
            ef.gdn_inverse = g.inverse;
            return run_conv(c, conv, in, out.H, out.W, ef, s);
        }
    }
    if (tc) {
        ep.out_act_sq = sq;
        ep.sq_scale = sq_scale;
    }
    int rc = run_conv(c, conv, in, out.H, out.W, ep, s);
    if (rc) return rc;
    if (!tc) return launch_gdn_act(raw, g.C, g.beta_eff, g.gamma_eff, g.inverse, out, s);
    // norm_i = beta_i + sum_j gamma_ij x_j^2 as a 1x1 tcgen05 convolution; y = x / sqrt(norm) or x * sqrt(norm)
    const std::string nname = gdn + "#norm";
    ConvRt& n = c->conv[nname];
    Epilogue en = make_ep(n);
    en.acc_scale = 1.f / sq_scale;
    en.res_act = raw;
    en.res_mode = g.inverse ? 2 : 1;
    en.out_act = out;
    return run_conv(c, nname, sq, out.H, out.W, en, s);
}

// Fused 3x3 tail (SURVEY 7.1-8 / VERDICT r1 #4b,c): fills the producer's epilogue with the tail's regrouped weights.
// `tail` = name of the 3x3 convolution with 2-3 outputs, `y` = geometry of the producer's output (never written).
static int tail_epilogue(fvc_ctx* c, const std::string& tail, ActT y, Epilogue& ep, cudaStream_t s) {
    ConvRt& t = c->conv[tail + "#taps"];
    ConvRt& tl = c->conv[tail];
    if (!t.have_w) {
        set_error("parameters of %s not set", tail.c_str());
        return FVC_ERR_STATE;
    }
    if (!t.tc) {   // created (never launched): owns the packed 32-row weight tiles and their scale
        Epilogue et = make_ep(t);
        et.bias = t.bias;
        et.out_f32 = c->taps_buf;
        int rc = tc_plan_create(t.L, t.w_raw, y, y.H, y.W, et, &t.tc, s, false, true);
        if (rc) return rc;
    }
    if (!tc_plan_is_tap_layout(t.tc)) {
        set_error("unexpected weight layout for the fused tail %s", tail.c_str());
        return FVC_ERR_STATE;
    }
    ep.tap_w = tc_plan_wstream(t.tc);
    ep.tap_scale = tc_plan_acc_scale(t.tc);
    ep.tap_out = c->taps_buf;
    ep.tap_cq = (9 * tl.L.Cout + 3) & ~3;
    return 0;
}

static int res_block(fvc_ctx* c, int idx, ActT x_relu, ActT x_skip, ActT tmp, ActT out, ActT out_relu,
                     cudaStream_t s, const char* fused_tail = nullptr) {
    char n1[64], n2[64];
    snprintf(n1, sizeof(n1), "warpnet.conv%d.conv1", idx);
    snprintf(n2, sizeof(n2), "warpnet.conv%d.conv2", idx);
    Epilogue ep = make_ep(c->conv[n1]);
    ep.out_act = tmp;  // relu(conv1(relu(x)))
    int rc = run_conv(c, n1, x_relu, tmp.H, tmp.W, ep, s);
    if (rc) return rc;
    ep = make_ep(c->conv[n2]);
    ep.res_act = x_skip;
    if (fused_tail) {   // the block's output only feeds a 3x3 tail: it stays in the SM (second MMA), P goes out
        rc = tail_epilogue(c, fused_tail, out, ep, s);
        if (rc) return rc;
    } else {
        ep.out_act = out;
        ep.out_act_relu = out_relu;
    }
    return run_conv(c, n2, tmp, out.H, out.W, ep, s);
}

// ---- real entropy coding --------------------------------------------------------------------------------
static int64_t latent_count(fvc_ctx* c, int which) {   // 0: feature [B,96,H/16,W/16], 1: z [B,64,H/64,W/64], 2: mv [B,128,..]
    const int64_t p16 = (int64_t)c->B * (c->H / 16) * (c->W / 16), p64 = (int64_t)c->B * (c->H / 64) * (c->W / 64);
    return which == 0 ? p16 * 96 : (which == 1 ? p64 * 64 : p16 * 128);
}
static int ensure_entropy_buffers(fvc_ctx* c) {
    if (c->sym_packed && c->ent_R == c->mxrange && c->ent_L == c->rans_L) return 0;
    if (c->sym_packed) {
        set_error("mxrange / lane length of a context cannot change once entropy coding was used");
        return FVC_ERR_STATE;
    }
    const int R = c->mxrange, L = c->rans_L;
    const int64_t nmax = latent_count(c, 2);
    if (c->alloc(&c->cdf_tab_mv, (size_t)128 * 2 * R * 4) || c->alloc(&c->cdf_tab_z, (size_t)64 * 2 * R * 4) ||
        c->alloc(&c->sym_packed, (size_t)nmax * 8) || c->alloc(&c->rans_words, entropy_words_capacity(nmax, L) * 2) ||
        c->alloc(&c->rans_lane_words, (size_t)cdiv64(nmax, L) * 4) || c->alloc(&c->stream_bytes, 16) ||
        c->alloc(&c->ent_err, 16))
        return FVC_ERR_CUDA;
    for (int k = 0; k < 3; ++k) {
        c->stream_cap[k] = entropy_stream_capacity(latent_count(c, k), L);
        if (c->alloc(&c->stream[k], c->stream_cap[k])) return FVC_ERR_CUDA;
    }
    c->ent_R = R; c->ent_L = L;
    return 0;
}
static FactorizedParams be_params(const BitEstRt& be) {
    FactorizedParams prm;
    for (int i = 0; i < 11; ++i) prm.p[i] = be.p[i];
    return prm;
}
// The coder's stream: everything queued on `main` so far (the coder's inputs) happens before what follows on it.
static int ent_fork(fvc_ctx* c, cudaStream_t main, cudaStream_t* out) {
    static const bool side = !(getenv("FVC_ENT_STREAM") && getenv("FVC_ENT_STREAM")[0] == '0');
    if (!side) { *out = main; return 0; }
    if (!c->ent_stream) {
        FVC_CUDA(cudaStreamCreateWithFlags(&c->ent_stream, cudaStreamNonBlocking));
        FVC_CUDA(cudaEventCreateWithFlags(&c->ent_fork, cudaEventDisableTiming));
        FVC_CUDA(cudaEventCreateWithFlags(&c->ent_done, cudaEventDisableTiming));
    }
    FVC_CUDA(cudaEventRecord(c->ent_fork, main));
    FVC_CUDA(cudaStreamWaitEvent(c->ent_stream, c->ent_fork, 0));
    c->ent_pending = true;
    *out = c->ent_stream;
    return 0;
}
// `main` continues only after the coders queued so far have finished (before their byte counts / streams are read and
// before the next frame overwrites their inputs)
static int ent_join(fvc_ctx* c, cudaStream_t main) {
    if (!c->ent_pending) return 0;
    FVC_CUDA(cudaEventRecord(c->ent_done, c->ent_stream));
    FVC_CUDA(cudaStreamWaitEvent(main, c->ent_done, 0));
    c->ent_pending = false;
    return 0;
}
// which = 1 (z) or 2 (mv): x is the pre-round latent, fp32 NHWC
static int encode_factorized(fvc_ctx* c, int which, const float* x, cudaStream_t main) {
    int rc;
    cudaStream_t s;
    rc = ent_fork(c, main, &s);
    if (rc) return rc;
    const BitEstRt& be = which == 1 ? c->be_z : c->be_mv;
    uint32_t* tab = which == 1 ? c->cdf_tab_z : c->cdf_tab_mv;
    const int64_t n = latent_count(c, which);
    PK(which == 1 ? "@entropy_model:z" : "@entropy_model:mv", launch_cdf_table_factorized(be_params(be), be.C, c->mxrange, tab, s));
    rc = launch_sym_factorized(x, n, be.C, c->mxrange, tab, c->sym_packed, c->ent_err, s);
    if (rc) return rc;
    PK(which == 1 ? "@rans_encode:z" : "@rans_encode:mv",
       launch_rans_encode(c->sym_packed, n, c->rans_L, c->rans_words, c->rans_lane_words, c->stream[which], c->stream_bytes + which, s));
    return 0;
}
static int encode_laplace(fvc_ctx* c, cudaStream_t main) {
    int rc;
    cudaStream_t s;
    rc = ent_fork(c, main, &s);
    if (rc) return rc;
    const int64_t n = latent_count(c, 0);
    PK("@entropy_model:feature", launch_sym_laplace(c->feature, c->sigma, n, c->mxrange, c->sym_packed, c->ent_err, s));
    PK("@rans_encode:feature",
       launch_rans_encode(c->sym_packed, n, c->rans_L, c->rans_words, c->rans_lane_words, c->stream[0], c->stream_bytes + 0, s));
    return 0;
}

// mvDecoder (synthesis_mv.py:71-79): c->quant_mv (ACT) -> c->mv_hat (fp32 NHWC2)
static int run_mv_decoder(fvc_ctx* c, cudaStream_t s) {
    const int H = c->H, W = c->W;
    char nm[96];
    ActT in = c->quant_mv;
    for (int i = 1; i <= 8; ++i) {
        snprintf(nm, sizeof(nm), "mvDecoder.deconv%d", i);
        Epilogue ep = make_ep(c->conv[nm]);
        int sh = 4 - (i + 1) / 2;
        int rc;
        if (i == 7 && c->tail_fused) {
            // deconv7 -> LeakyReLU -> deconv8 (synthesis_mv.py:77-79) in one kernel + the tap sum: the 128-channel
            // full-resolution tensor (1.07 GB at 1080p) is never written
            rc = tail_epilogue(c, "mvDecoder.deconv8", c->d[7], ep, s);
            if (!rc) rc = run_conv(c, nm, in, H, W, ep, s);
            if (rc) return rc;
            ConvRt& t8 = c->conv["mvDecoder.deconv8"];
            PK("@k_tapsum:mv", launch_tapsum(c->taps_buf, t8.bias, c->mv_hat, c->B, H, W, 2, ep.tap_cq, s));
            break;
        }
        if (i < 8) ep.out_act = c->d[i];
        else ep.out_f32 = c->mv_hat;
        rc = run_conv(c, nm, in, H >> sh, W >> sh, ep, s);
        if (rc) return rc;
        if (i < 8) in = c->d[i];
    }
    return 0;
}

// Phase A of the path: opticFlow + mvEncoder + quantiser/bits + mvDecoder (net.py:71-77; LSVC: models.py:1350-1351,
// 1333-1342).  Leaves mv_hat in c->mv_hat (fp32 NHWC2) and the mv bit partials in c->bits_partials[2].
static int forward_mv(fvc_ctx* c, const float* cur, const float* ref, int* nb_mv_out, cudaStream_t s) {
    const int B = c->B, H = c->H, W = c->W, L = c->levels;
    char nm[96];
    int rc;
#define R(expr) do { rc = (expr); if (rc) return rc; } while (0)
    // ---- SpyNet (endecoder.py:337-356) --------------------------------------------------------
    const float* p1 = cur;
    const float* p2 = ref;
    std::vector<const float*> im1(L), im2(L);
    im1[0] = cur; im2[0] = ref;
    for (int sc = 1; sc < L; ++sc) {
        PK("@k_avg_pool2_planar", launch_avg_pool2_planar(im1[sc - 1], c->pyr1[sc], B * 3, H >> (sc - 1), W >> (sc - 1), s));
        PK("@k_avg_pool2_planar", launch_avg_pool2_planar(im2[sc - 1], c->pyr2[sc], B * 3, H >> (sc - 1), W >> (sc - 1), s));
        im1[sc] = c->pyr1[sc];
        im2[sc] = c->pyr2[sc];
    }
    (void)p1; (void)p2;
    for (int i = 0; i < L; ++i) {
        int sc = L - 1 - i;
        int h = H >> sc, w = W >> sc;
        PK("@k_spynet_prep:L" + std::to_string(i), launch_spynet_prep(im1[sc], im2[sc], i == 0 ? nullptr : c->sflow[i - 1], c->sx[i], c->sflow_up[i], s));
        ActT chain_in[5] = {c->sx[i], c->sa1[i], c->sa2[i], c->sa3[i], c->sa4[i]};
        for (int k = 0; k < 5; ++k) {
            snprintf(nm, sizeof(nm), "opticFlow.moduleBasic.%d.conv%d", i, k + 1);
            Epilogue ep = make_ep(c->conv[nm]);
            if (k < 4) {
                ep.out_act = chain_in[k + 1];
            } else {
                ep.res_f32 = c->sflow_up[i];
                ep.out_f32 = c->sflow[i];
                if (i == L - 1) ep.out_act = c->estmv_act;
            }
            R(run_conv(c, nm, chain_in[k], h, w, ep, s));
        }
    }
    // ---- mvEncoder (analysis_mv.py:58-66) ----------------------------------------------------
    {
        ActT in = c->estmv_act;
        for (int i = 1; i <= 8; ++i) {
            snprintf(nm, sizeof(nm), "mvEncoder.conv%d", i);
            Epilogue ep = make_ep(c->conv[nm]);
            int sh = (i + 1) / 2;
            if (i < 8) ep.out_act = c->e[i];
            else ep.out_f32 = c->mvfeature;
            R(run_conv(c, nm, in, H >> sh, W >> sh, ep, s));
            if (i < 8) in = c->e[i];
        }
    }
    int nb_mv = 0;
    const int maxb = bits_max_blocks();
    {
        FactorizedParams prm;
        for (int i = 0; i < 11; ++i) prm.p[i] = c->be_mv.p[i];
        PK("@k_quant_bits_factorized:mv", launch_quant_bits_factorized(c->mvfeature, 1, B, 128, (H / 16) * (W / 16), prm, nullptr, c->quant_mv,
                                       c->bits_partials + 2 * maxb, &nb_mv, s));
    }
    if (c->realbits) R(encode_factorized(c, 2, c->mvfeature, s));                     // net.py:183-195
    if (c->force_q[0]) R(launch_nchw_to_act(c->force_q[0], c->quant_mv, 128, 0, s));   // teacher forcing
    *nb_mv_out = nb_mv;
    R(run_mv_decoder(c, s));
#undef R
    return 0;
}

// motion compensation (net.py:64-68, endecoder.py:282-296): c->mv_hat, ref -> c->warpframe, c->prediction;
// with `cur` also the residual record (net.py:81); cur == nullptr (decoder): the residual is not formed
static int run_motion_comp(fvc_ctx* c, const float* cur, const float* ref, cudaStream_t s) {
    const int H = c->H, W = c->W;
    int rc;
#define R(expr) do { rc = (expr); if (rc) return rc; } while (0)
    PK("@k_mc_prep", launch_mc_prep(ref, c->mv_hat, c->warpframe, c->xmc, s));
    {
        Epilogue ep = make_ep(c->conv["warpnet.feature_ext"]);
        ep.out_act = c->wf;
        R(run_conv(c, "warpnet.feature_ext", c->xmc, H, W, ep, s));
    }
    R(res_block(c, 0, c->wf, c->wf, c->wt0, c->wc0, no_act(), s));          // f >= 0: relu(f) == f
    PK("@k_pool_act:full", launch_pool_act(c->wc0, c->wc0p, c->wc0p_r, s));
    R(res_block(c, 1, c->wc0p_r, c->wc0p, c->wt1, c->wc1, no_act(), s));
    PK("@k_pool_act:half", launch_pool_act(c->wc1, c->wc1p, c->wc1p_r, s));
    R(res_block(c, 2, c->wc1p_r, c->wc1p, c->wt2, c->wc2, c->wc2_r, s));
    R(res_block(c, 3, c->wc2_r, c->wc2, c->wt3, c->wc3, no_act(), s));
    PK("@k_upadd_act:half", launch_upadd_act(c->wc3, c->wc1, c->wc3u, c->wc3u_r, s));
    R(res_block(c, 4, c->wc3u_r, c->wc3u, c->wt4, c->wc4, no_act(), s));
    PK("@k_upadd_act:full", launch_upadd_act(c->wc4, c->wc0, c->wc4u, c->wc4u_r, s));
    if (c->tail_fused_warpnet) {
        // ResBlock 5's conv2 (+ skip) -> conv6 (endecoder.py:294-295) in one kernel + the tap sum
        R(res_block(c, 5, c->wc4u_r, c->wc4u, c->wt5, c->wc5, no_act(), s, "warpnet.conv6"));
        ConvRt& t6 = c->conv["warpnet.conv6"];
        PK("@k_tapsum:warpnet", launch_tapsum(c->taps_buf, t6.bias, c->wres, c->B, H, W, 3, 28, s));
    } else {
        R(res_block(c, 5, c->wc4u_r, c->wc4u, c->wt5, c->wc5, no_act(), s));
        Epilogue ep = make_ep(c->conv["warpnet.conv6"]);
        ep.out_f32 = c->wres;
        R(run_conv(c, "warpnet.conv6", c->wc5, H, W, ep, s));
    }
    PK("@k_mc_finish", launch_mc_finish(c->wres, c->warpframe, cur ? cur : ref, c->prediction, c->residual, s));
#undef R
    return 0;
}

// hyper-prior decoder (synthesis_prior.py:42-58): c->z_hat (ACT) -> c->sigma (fp32 NHWC96)
static int run_prior_decoder(fvc_ctx* c, cudaStream_t s) {
    const int H = c->H, W = c->W;
    int rc;
    Epilogue ep = make_ep(c->conv["respriorDecoder.deconv1"]);
    ep.out_act = c->s1;
    rc = run_conv(c, "respriorDecoder.deconv1", c->z_hat, H / 32, W / 32, ep, s);
    if (rc) return rc;
    ep = make_ep(c->conv["respriorDecoder.deconv2"]);
    ep.out_act = c->s2;
    rc = run_conv(c, "respriorDecoder.deconv2", c->s1, H / 16, W / 16, ep, s);
    if (rc) return rc;
    ep = make_ep(c->conv["respriorDecoder.deconv3"]);
    ep.out_f32 = c->sigma;
    return run_conv(c, "respriorDecoder.deconv3", c->s2, H / 16, W / 16, ep, s);
}

// residual decoder (synthesis.py:54-58): c->feat_hat (ACT) -> c->recon_res (fp32 NHWC3)
static int run_res_decoder(fvc_ctx* c, cudaStream_t s) {
    int rc;
#define R(expr) do { rc = (expr); if (rc) return rc; } while (0)
    R(run_conv_gdn(c, "resDecoder.deconv1", "resDecoder.igdn1", c->feat_hat, c->g_raw[0], c->g_sq[0], c->g[0], s));
    R(run_conv_gdn(c, "resDecoder.deconv2", "resDecoder.igdn2", c->g[0], c->g_raw[1], c->g_sq[1], c->g[1], s));
    R(run_conv_gdn(c, "resDecoder.deconv3", "resDecoder.igdn3", c->g[1], c->g_raw[2], c->g_sq[2], c->g[2], s));
    Epilogue ep = make_ep(c->conv["resDecoder.deconv4"]);
    ep.out_f32 = c->recon_res;
    R(run_conv(c, "resDecoder.deconv4", c->g[2], c->H, c->W, ep, s));
#undef R
    return 0;
}

// Phase B: motion compensation + residual codec + reconstruction (net.py:79-116; LSVC: models.py:1375-1383,
// 1300-1331) with the motion field in c->mv_hat.  sums: 3 loss partial sums reduced with `loss_scale`,
// bits_feature, bits_z in c->scalars[0..4].
static int forward_mc_res(fvc_ctx* c, const float* cur, const float* ref, float* recon_out, double loss_scale,
                          int clip_mse, cudaStream_t s, bool intra = false) {
    const int B = c->B, H = c->H, W = c->W;
    int rc;
    int nb_z = 0, nb_f = 0;
    const int maxb = bits_max_blocks();
#define R(expr) do { rc = (expr); if (rc) return rc; } while (0)
    if (intra) {
        // intra frame: no reference, the prediction is zero and the "residual" is the frame itself
        const size_t n = (size_t)B * 3 * H * W * 4;
        FVC_CUDA(cudaMemsetAsync(c->prediction, 0, n, s));
        FVC_CUDA(cudaMemsetAsync(c->warpframe, 0, n, s));
        PK("@k_nchw_to_act:intra", launch_nchw_to_act(cur, c->residual, 3, 0, s));
    } else {
        R(run_motion_comp(c, cur, ref, s));
    }
    // ---- residual encoder (analysis.py:44-48) ---------------------------------------------------
    R(run_conv_gdn(c, "resEncoder.conv1", "resEncoder.gdn1", c->residual, c->r_raw[0], c->r_sq[0], c->r[0], s));
    R(run_conv_gdn(c, "resEncoder.conv2", "resEncoder.gdn2", c->r[0], c->r_raw[1], c->r_sq[1], c->r[1], s));
    R(run_conv_gdn(c, "resEncoder.conv3", "resEncoder.gdn3", c->r[1], c->r_raw[2], c->r_sq[2], c->r[2], s));
    {
        Epilogue ep = make_ep(c->conv["resEncoder.conv4"]);
        ep.out_f32 = c->feature;
        R(run_conv(c, "resEncoder.conv4", c->r[2], H / 16, W / 16, ep, s));
    }
    // ---- hyper-prior (analysis_prior.py:40-56, synthesis_prior.py:42-58) ------------------------
    PK("@k_nhwc_to_act", launch_nhwc_to_act(c->feature, c->featabs, 96, 1, s));
    {
        Epilogue ep = make_ep(c->conv["respriorEncoder.conv1"]);
        ep.out_act = c->p1;
        R(run_conv(c, "respriorEncoder.conv1", c->featabs, H / 16, W / 16, ep, s));
        ep = make_ep(c->conv["respriorEncoder.conv2"]);
        ep.out_act = c->p2;
        R(run_conv(c, "respriorEncoder.conv2", c->p1, H / 32, W / 32, ep, s));
        ep = make_ep(c->conv["respriorEncoder.conv3"]);
        ep.out_f32 = c->z;
        R(run_conv(c, "respriorEncoder.conv3", c->p2, H / 64, W / 64, ep, s));
    }
    {
        FactorizedParams prm;
        for (int i = 0; i < 11; ++i) prm.p[i] = c->be_z.p[i];
        PK("@k_quant_bits_factorized:z", launch_quant_bits_factorized(c->z, 1, B, 64, (H / 64) * (W / 64), prm, nullptr, c->z_hat,
                                       c->bits_partials + 1 * maxb, &nb_z, s));
    }
    if (c->realbits) R(encode_factorized(c, 1, c->z, s));                           // net.py:155-168
    if (c->force_q[1]) R(launch_nchw_to_act(c->force_q[1], c->z_hat, 64, 0, s));   // teacher forcing
    R(run_prior_decoder(c, s));
    PK("@k_quant_bits_laplace", launch_quant_bits_laplace(c->feature, c->sigma, (int64_t)B * (H / 16) * (W / 16) * 96, 96, nullptr,
                                c->feat_hat, c->bits_partials + 0 * maxb, &nb_f, s));
    if (c->realbits) R(encode_laplace(c, s));                                         // net.py:123-138
    if (c->force_q[2]) R(launch_nchw_to_act(c->force_q[2], c->feat_hat, 96, 0, s));   // teacher forcing
    R(run_res_decoder(c, s));
    // ---- reconstruction, distortion, rate (net.py:103-116, 207-217) -------------------------------
    int nloss = 0;
    PK("@k_recon_losses", launch_recon_losses(cur, c->prediction, c->warpframe, c->recon_res, 1, B, H * W, recon_out, c->loss_partials,
                          &nloss, s, clip_mse));
    R(launch_reduce_partials(c->loss_partials, nloss, 3, loss_scale, c->scalars, s));
    R(launch_reduce_partials(c->bits_partials + 0 * maxb, nb_f, 1, 1.0, c->scalars + 3, s));
    R(launch_reduce_partials(c->bits_partials + 1 * maxb, nb_z, 1, 1.0, c->scalars + 4, s));
#undef R
    return 0;
}

// VideoCompressor.forward (net.py:70-220) = phase A + phase B + bpp bookkeeping
static int forward(fvc_ctx* c, const float* cur, const float* ref, float* recon_out, float* scalars_out,
                   cudaStream_t s) {
    const int B = c->B, H = c->H, W = c->W;
    c->conv_event_used = 0;
    int nb_mv = 0;
    int rc = forward_mv(c, cur, ref, &nb_mv, s);
    if (rc) return rc;
    rc = forward_mc_res(c, cur, ref, recon_out, 1.0 / ((double)B * 3 * H * W), 0, s);
    if (rc) return rc;
    rc = launch_reduce_partials(c->bits_partials + 2 * bits_max_blocks(), nb_mv, 1, 1.0, c->scalars + 5, s);
    if (rc) return rc;
    rc = ent_join(c, s);
    if (rc) return rc;
    if (c->realbits) {   // total_bits = real_bits (net.py:147-149, 172-174, 200-202): 8 x bytes of the three streams
        for (int k = 0; k < 3 && !rc; ++k) rc = launch_bytes_to_bits(c->stream_bytes + k, c->ent_err, c->scalars + 3 + k, s);
        if (rc) return rc;
    }
    return launch_finalize_scalars(c->scalars, (float)((double)B * H * W), scalars_out, c->sat_count, s);
}

}  // namespace fvc

// ==================================================================================================
// C ABI: context
// ==================================================================================================
extern "C" {

const char* fvc_last_error(void) { return fvc::g_err.c_str(); }
int fvc_version(void) { return 200; }

fvc_ctx* fvc_ctx_create(int B, int H, int W, int levels, int impl) {
    if (B < 1 || H < 64 || W < 64 || (H % 64) || (W % 64) || levels < 1 || levels > 6 ||
        (impl != FVC_IMPL_SIMT && impl != FVC_IMPL_TC && impl != FVC_IMPL_TC_FAST)) {
        set_error("fvc_ctx_create: bad geometry B=%d H=%d W=%d levels=%d impl=%d (H, W must be multiples of 64)", B,
                  H, W, levels, impl);
        return nullptr;
    }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        set_error("fvc_ctx_create: no CUDA device (%s); this library has no CPU fallback", cudaGetErrorString(e));
        return nullptr;
    }
    fvc_ctx* c = new fvc_ctx();
    c->B = B; c->H = H; c->W = W; c->levels = levels; c->impl = impl;
    const char* prof = getenv("FVC_PROFILE");
    c->profile = prof && prof[0] == '1';
    const char* few = getenv("FVC_FEW");
    c->use_few = !(few && few[0] == '0');

    const char* gf0 = getenv("FVC_GDN_FUSED");
    c->gdn_fused = impl != FVC_IMPL_SIMT && !(gf0 && gf0[0] == '0');
    // symbols per rANS lane of the real entropy coder: a lane is one serial chain (~85 ns per symbol), so the coder's
    // time is proportional to it, and each lane costs 32 bits of final state: 8192 -> +0.1 % size, 3 x 0.8 ms per 1080p
    // frame; 2048 -> +0.4 %, 3 x 0.2 ms.  Encoder and decoder must agree (the container records it).
    if (const char* rl = getenv("FVC_RANS_L")) {
        const int v = atoi(rl);
        if (v >= 64 && v <= 16384) c->rans_L = v;
    }
    const char* tf0 = getenv("FVC_TAIL_FUSED");   // 0: off, 1 (default): mvDecoder.deconv8 and warpnet.conv6, 2: deconv8 only
    c->tail_fused = impl != FVC_IMPL_SIMT && !(tf0 && tf0[0] == '0');
    c->tail_fused_warpnet = c->tail_fused && !(tf0 && tf0[0] == '2');
    if (build_layers(c) || build_buffers(c)) {
        fvc_ctx_destroy(c);
        return nullptr;
    }
    return c;
}

void fvc_ctx_destroy(fvc_ctx* c) {
    if (!c) return;
    cudaDeviceSynchronize();
    for (auto& kv : c->conv) {
        if (kv.second.simt.w) cudaFree(kv.second.simt.w);
        if (kv.second.tc) tc_plan_destroy(kv.second.tc);
        if (kv.second.few_w) cudaFree(kv.second.few_w);
    }
    for (void* p : c->allocs) cudaFree(p);
    for (auto& ev : c->conv_events) {
        cudaEventDestroy(ev.first);
        cudaEventDestroy(ev.second);
    }
    if (c->stage_frames) cudaFree(c->stage_frames);
    if (c->stage_u8) cudaFree(c->stage_u8);
    if (c->stage_rec) cudaFree(c->stage_rec);
    if (c->stage_scalars) cudaFree(c->stage_scalars);
    for (cudaEvent_t e : c->copy_events) cudaEventDestroy(e);
    if (c->copy_fence) cudaEventDestroy(c->copy_fence);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    if (c->ent_fork) cudaEventDestroy(c->ent_fork);
    if (c->ent_done) cudaEventDestroy(c->ent_done);
    if (c->ent_stream) cudaStreamDestroy(c->ent_stream);
    delete c;
}

int fvc_ctx_set_param(fvc_ctx* c, const char* key_c, const float* data, int64_t numel, void* stream) {
    FVC_ARG(c && key_c && data);
    cudaStream_t s = (cudaStream_t)stream;
    std::string key(key_c);
    size_t dot = key.rfind('.');
    FVC_ARG(dot != std::string::npos);
    std::string mod = key.substr(0, dot), leaf = key.substr(dot + 1);
    auto ci = c->conv.find(mod);
    if (ci != c->conv.end()) {
        ConvRt& r = ci->second;
        if (leaf == "weight") {
            int64_t n = (int64_t)r.L.Cin * r.L.Cout * r.L.k * r.L.k;
            if (numel != n) { set_error("%s: expected %lld elements, got %lld", key_c, (long long)n, (long long)numel); return FVC_ERR_ARG; }
            FVC_CUDA(cudaMemcpyAsync(r.w_raw, data, n * 4, cudaMemcpyDeviceToDevice, s));
            int rc = simt_pack_weights(r.L, r.w_raw, r.CinP, r.CoutS, &r.simt, s);
            if (rc) return rc;
            if (few_supported(r.L, r.CinP)) {
                rc = few_pack_weights(r.L, r.w_raw, &r.few_w, s);
                if (rc) return rc;
            }
            if (r.tc) { tc_plan_destroy(r.tc); r.tc = nullptr; }
            r.have_w = true;
            // a fused tail's regrouped weights follow its 3x3 weights; the producer's plan points at their stream
            const char* prod = mod == "mvDecoder.deconv8" ? "mvDecoder.deconv7" : (mod == "warpnet.conv6" ? "warpnet.conv5.conv2" : nullptr);
            if (prod) {
                ConvRt& t = c->conv[mod + "#taps"];
                rc = launch_tapsplit_weights(r.w_raw, t.w_raw, r.L.Cin, r.L.Cout, 32, s);
                if (rc) return rc;
                if (t.tc) { tc_plan_destroy(t.tc); t.tc = nullptr; }
                t.have_w = true;
                ConvRt& pr = c->conv[prod];
                if (pr.tc) { tc_plan_destroy(pr.tc); pr.tc = nullptr; }
            }
            return 0;
        }
        if (leaf == "bias") {
            if (numel != r.L.Cout) { set_error("%s: expected %d elements", key_c, r.L.Cout); return FVC_ERR_ARG; }
            FVC_CUDA(cudaMemcpyAsync(r.bias, data, numel * 4, cudaMemcpyDeviceToDevice, s));
            r.have_b = true;
            return 0;
        }
    }
    auto gi = c->gdn.find(mod);
    if (gi != c->gdn.end()) {
        GdnRt& g = gi->second;
        if (leaf == "beta") {
            FVC_ARG(numel == g.C);
            FVC_CUDA(cudaMemcpyAsync(g.beta_raw, data, numel * 4, cudaMemcpyDeviceToDevice, s));
            g.have_b = true;
        } else if (leaf == "gamma") {
            FVC_ARG(numel == (int64_t)g.C * g.C);
            FVC_CUDA(cudaMemcpyAsync(g.gamma_raw, data, numel * 4, cudaMemcpyDeviceToDevice, s));
            g.have_g = true;
        } else {
            set_error("unknown parameter %s", key_c);
            return FVC_ERR_ARG;
        }
        if (g.have_b && g.have_g) {
            int rc = launch_gdn_reparam(g.beta_raw, g.gamma_raw, g.beta_eff, g.gamma_eff, g.C, s);
            if (rc) return rc;
            ConvRt& n = c->conv[mod + "#norm"];   // gamma_eff is [out][in] = a 1x1 OIHW kernel
            FVC_CUDA(cudaMemcpyAsync(n.w_raw, g.gamma_eff, (size_t)g.C * g.C * 4, cudaMemcpyDeviceToDevice, s));
            FVC_CUDA(cudaMemcpyAsync(n.bias, g.beta_eff, (size_t)g.C * 4, cudaMemcpyDeviceToDevice, s));
            rc = simt_pack_weights(n.L, n.w_raw, n.CinP, n.CoutS, &n.simt, s);
            if (rc) return rc;
            if (n.tc) { tc_plan_destroy(n.tc); n.tc = nullptr; }
            n.have_w = n.have_b = true;
            // the producing convolution's plan refers to the norm plan's gamma stream (fused epilogue): rebuild it
            std::string prod = mod;   // resEncoder.gdnK -> resEncoder.convK, resDecoder.igdnK -> resDecoder.deconvK
            size_t pg = prod.find("igdn");
            if (pg != std::string::npos) prod.replace(pg, 4, "deconv");
            else if ((pg = prod.find("gdn")) != std::string::npos) prod.replace(pg, 3, "conv");
            auto pi = c->conv.find(prod);
            if (pi != c->conv.end() && pi->second.tc) { tc_plan_destroy(pi->second.tc); pi->second.tc = nullptr; }
        }
        return 0;
    }
    // bitEstimator_{z,mv}.f{1..4}.{h,b,a}
    for (BitEstRt* be : {&c->be_z, &c->be_mv}) {
        if (key.compare(0, be->name.size() + 1, be->name + ".") == 0) {
            std::string rest = key.substr(be->name.size() + 1);  // f1.h
            if (rest.size() == 4 && rest[0] == 'f' && rest[2] == '.') {
                int fi = rest[1] - '1';
                int pi = rest[3] == 'h' ? 0 : (rest[3] == 'b' ? 1 : (rest[3] == 'a' ? 2 : -1));
                if (fi >= 0 && fi < 4 && pi >= 0 && !(fi == 3 && pi == 2)) {
                    FVC_ARG(numel == be->C);
                    int idx = fi * 3 + pi;
                    FVC_CUDA(cudaMemcpyAsync(be->p[idx], data, numel * 4, cudaMemcpyDeviceToDevice, s));
                    be->have[idx] = true;
                    return 0;
                }
            }
        }
    }
    set_error("unknown parameter %s", key_c);
    return FVC_ERR_ARG;
}

int fvc_ctx_missing_params(fvc_ctx* c) {
    if (!c) return -1;
    int m = 0;
    for (auto& kv : c->conv) m += (!kv.second.have_w) + (!kv.second.have_b);
    for (auto& kv : c->gdn) m += (!kv.second.have_b) + (!kv.second.have_g);
    for (BitEstRt* be : {&c->be_z, &c->be_mv})
        for (int i = 0; i < 11; ++i) m += !be->have[i];
    return m;
}

int fvc_pframe_forward(fvc_ctx* c, const float* cur, const float* ref, float* recon_out, float* scalars_out,
                       void* stream) {
    FVC_ARG(c && cur && ref && recon_out && scalars_out);
    if (fvc_ctx_missing_params(c) != 0) {
        set_error("fvc_pframe_forward: %d parameters not set", fvc_ctx_missing_params(c));
        return FVC_ERR_STATE;
    }
    int64_t before = g_launch_count;
    int rc = forward(c, cur, ref, recon_out, scalars_out, (cudaStream_t)stream);
    c->launches += g_launch_count - before;
    if (rc == 0 && c->profile) {
        FVC_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
        double tot = 0;
        c->profile_text.clear();
        char line[192];
        for (size_t i = 0; i < c->conv_event_used; ++i) {
            float ms = 0;
            cudaEventElapsedTime(&ms, c->conv_events[i].first, c->conv_events[i].second);
            if (c->conv_event_names[i][0] != '@') tot += ms * 1e-3;   // convolution launches only
            snprintf(line, sizeof(line), "%s %.4f\n", c->conv_event_names[i].c_str(), ms);
            c->profile_text += line;
        }
        c->last_conv_seconds = tot;
    }
    return rc;
}

/* Intra frame through the residual branch alone ("hyperprior image codec from the same conv engine"):
 * y = resEncoder(x), z = respriorEncoder(|y|), sigma = respriorDecoder(round z), x_hat = clamp(resDecoder(round y)),
 * i.e. VideoCompressor.forward (net.py:86-116) with a zero prediction and no motion branch.  The reference has no
 * learned intra codec (models.py:412-429 shells out to bpgenc / bpgdec); this reuses its modules and weights. */
int fvc_iframe_forward(fvc_ctx* c, const float* frame, float* recon_out, float* scalars_out, void* stream) {
    FVC_ARG(c && frame && recon_out && scalars_out);
    if (fvc_ctx_missing_params(c) != 0) {
        set_error("fvc_iframe_forward: %d parameters not set", fvc_ctx_missing_params(c));
        return FVC_ERR_STATE;
    }
    cudaStream_t s = (cudaStream_t)stream;
    const int B = c->B, H = c->H, W = c->W;
    int64_t before = g_launch_count;
    c->conv_event_used = 0;
    int rc = forward_mc_res(c, frame, nullptr, recon_out, 1.0 / ((double)B * 3 * H * W), 0, s, true);
    if (rc) return rc;
    FVC_CUDA(cudaMemsetAsync(c->scalars + 5, 0, 4, s));   // no motion stream
    rc = ent_join(c, s);
    if (rc) return rc;
    if (c->realbits)
        for (int k = 0; k < 2 && !rc; ++k) rc = launch_bytes_to_bits(c->stream_bytes + k, c->ent_err, c->scalars + 3 + k, s);
    if (!rc) rc = launch_finalize_scalars(c->scalars, (float)((double)B * H * W), scalars_out, c->sat_count, s);
    c->launches += g_launch_count - before;
    return rc;
}

/* Decoder of fvc_iframe_forward's two streams (fvc_ctx_get_bitstream 0 and 1). */
int fvc_iframe_decode_bitstreams(fvc_ctx* c, const void* feat_stream, int64_t feat_bytes, const void* z_stream,
                                 int64_t z_bytes, float* recon_out, void* stream) {
    FVC_ARG(c && feat_stream && z_stream && recon_out);
    if (fvc_ctx_missing_params(c) != 0) {
        set_error("fvc_iframe_decode_bitstreams: %d parameters not set", fvc_ctx_missing_params(c));
        return FVC_ERR_STATE;
    }
    cudaStream_t s = (cudaStream_t)stream;
    int64_t before = g_launch_count;
    c->conv_event_used = 0;
    int rc = ensure_entropy_buffers(c);
    const int R = c->mxrange, L = c->rans_L;
    const size_t n = (size_t)c->B * 3 * c->H * c->W * 4;
    int nloss = 0;
    if (!rc) rc = launch_cdf_table_factorized(be_params(c->be_z), 64, R, c->cdf_tab_z, s);
    if (!rc) rc = launch_rans_decode_factorized((const uint8_t*)z_stream, z_bytes, latent_count(c, 1), L, 64, R, c->cdf_tab_z, c->z, c->ent_err, s);
    if (!rc) rc = launch_nhwc_to_act(c->z, c->z_hat, 64, 0, s);
    if (!rc) rc = run_prior_decoder(c, s);
    if (!rc) rc = launch_rans_decode_laplace((const uint8_t*)feat_stream, feat_bytes, latent_count(c, 0), L, R, c->sigma, c->feature, c->ent_err, s);
    if (!rc) rc = launch_nhwc_to_act(c->feature, c->feat_hat, 96, 0, s);
    if (!rc) rc = run_res_decoder(c, s);
    if (rc) return rc;
    FVC_CUDA(cudaMemsetAsync(c->prediction, 0, n, s));
    FVC_CUDA(cudaMemsetAsync(c->warpframe, 0, n, s));
    // clamp(0 + recon_res, 0, 1); the distortion sums (against the zero prediction) are discarded
    rc = launch_recon_losses(c->prediction, c->prediction, c->warpframe, c->recon_res, 1, c->B, c->H * c->W, recon_out,
                             c->loss_partials, &nloss, s, 0);
    c->launches += g_launch_count - before;
    if (rc) return rc;
    unsigned int err[3] = {0, 0, 0};
    FVC_CUDA(cudaMemcpyAsync(err, c->ent_err, 12, cudaMemcpyDeviceToHost, s));
    FVC_CUDA(cudaStreamSynchronize(s));
    if (err[2]) {
        set_error("fvc_iframe_decode_bitstreams: %u lanes could not be opened (wrong container, size or geometry)", err[2]);
        cudaMemsetAsync(c->ent_err, 0, 12, s);
        return FVC_ERR_ARG;
    }
    return 0;
}

/* LSVC two-phase use of the path (reference models.py:1344-1411): phase A on a batch of frames against their
 * ORIGINAL reference frames, phase B per tree layer against RECONSTRUCTED references. */
int fvc_lsvc_mv_forward(fvc_ctx* c, const float* cur, const float* ref, float* mv_hat_out, float* bits_mv_out,
                        void* stream) {
    FVC_ARG(c && cur && ref && mv_hat_out && bits_mv_out);
    if (fvc_ctx_missing_params(c) != 0) {
        set_error("fvc_lsvc_mv_forward: %d parameters not set", fvc_ctx_missing_params(c));
        return FVC_ERR_STATE;
    }
    cudaStream_t s = (cudaStream_t)stream;
    int64_t before = g_launch_count;
    c->conv_event_used = 0;
    int nb_mv = 0;
    int rc = forward_mv(c, cur, ref, &nb_mv, s);
    if (!rc) rc = launch_reduce_partials(c->bits_partials + 2 * bits_max_blocks(), nb_mv, 1, 1.0, bits_mv_out, s);
    if (!rc) rc = launch_nhwc_to_nchw(c->mv_hat, mv_hat_out, c->B, 2, c->H, c->W, s);
    if (!rc) rc = ent_join(c, s);
    c->launches += g_launch_count - before;
    return rc;
}

int fvc_lsvc_mc_res_forward(fvc_ctx* c, const float* cur, const float* ref, const float* mv_hat, float* com_out,
                            float* mc_out, float* warp_out, float* sums_out, void* stream) {
    FVC_ARG(c && cur && ref && mv_hat && com_out && mc_out && warp_out && sums_out);
    if (fvc_ctx_missing_params(c) != 0) {
        set_error("fvc_lsvc_mc_res_forward: %d parameters not set", fvc_ctx_missing_params(c));
        return FVC_ERR_STATE;
    }
    cudaStream_t s = (cudaStream_t)stream;
    int64_t before = g_launch_count;
    c->conv_event_used = 0;
    const size_t n = (size_t)c->B * 3 * c->H * c->W;
    int rc = launch_nchw_to_nhwc(mv_hat, c->mv_hat, c->B, 2, c->H, c->W, s);
    if (!rc) rc = forward_mc_res(c, cur, ref, com_out, 1.0, 1, s);   // sums (not means) of the squared errors
    if (!rc) rc = ent_join(c, s);
    if (!rc) {
        FVC_CUDA(cudaMemcpyAsync(mc_out, c->prediction, n * 4, cudaMemcpyDeviceToDevice, s));
        FVC_CUDA(cudaMemcpyAsync(warp_out, c->warpframe, n * 4, cudaMemcpyDeviceToDevice, s));
        FVC_CUDA(cudaMemcpyAsync(sums_out, c->scalars, 5 * 4, cudaMemcpyDeviceToDevice, s));
    }
    c->launches += g_launch_count - before;
    return rc;
}

/* Decoder half of the path (net.py:77-80 mvDecoder + motioncompensation, 101-105 resDecoder, reconstruction and
 * clamp): what a receiver runs on the entropy-decoded latents. */
int fvc_decode_from_latents(fvc_ctx* c, const float* ref, const float* quant_mv, const float* feat_hat,
                            float* recon_out, void* stream) {
    FVC_ARG(c && ref && quant_mv && feat_hat && recon_out);
    if (fvc_ctx_missing_params(c) != 0) {
        set_error("fvc_decode_from_latents: %d parameters not set", fvc_ctx_missing_params(c));
        return FVC_ERR_STATE;
    }
    cudaStream_t s = (cudaStream_t)stream;
    int64_t before = g_launch_count;
    c->conv_event_used = 0;
    int nloss = 0;
    int rc = launch_nchw_to_act(quant_mv, c->quant_mv, 128, 0, s);
    if (!rc) rc = run_mv_decoder(c, s);
    if (!rc) rc = run_motion_comp(c, nullptr, ref, s);
    if (!rc) rc = launch_nchw_to_act(feat_hat, c->feat_hat, 96, 0, s);
    if (!rc) rc = run_res_decoder(c, s);
    // clamp(prediction + recon_res, 0, 1); the distortion sums are formed against `ref` and discarded
    if (!rc) rc = launch_recon_losses(ref, c->prediction, c->warpframe, c->recon_res, 1, c->B, c->H * c->W, recon_out,
                                      c->loss_partials, &nloss, s, 0);
    c->launches += g_launch_count - before;
    return rc;
}

int fvc_ctx_force_latents(fvc_ctx* c, const float* quant_mv, const float* z_hat, const float* feat_hat) {
    FVC_ARG(c != nullptr);
    c->force_q[0] = quant_mv;
    c->force_q[1] = z_hat;
    c->force_q[2] = feat_hat;
    return 0;
}

int64_t fvc_ctx_saturation_count(fvc_ctx* c, int reset, void* stream) {
    if (!c) { set_error("fvc_ctx_saturation_count: null context"); return FVC_ERR_ARG; }
    cudaStream_t s = (cudaStream_t)stream;
    unsigned int h = 0;
    FVC_CUDA(cudaMemcpyAsync(&h, c->sat_count, 4, cudaMemcpyDeviceToHost, s));
    FVC_CUDA(cudaStreamSynchronize(s));
    if (reset) FVC_CUDA(cudaMemsetAsync(c->sat_count, 0, 4, s));
    return (int64_t)h;
}

/* calrealbits (net.py:57, 123-138, 155-168, 183-195) */
int fvc_ctx_set_realbits(fvc_ctx* c, int enable, int mxrange) {
    FVC_ARG(c != nullptr && mxrange >= 2 && mxrange <= 16384);
    if (!c->sym_packed) c->mxrange = mxrange;          // fixed once the coder's buffers exist
    else if (mxrange != c->mxrange) {
        set_error("fvc_ctx_set_realbits: mxrange of a context cannot change (%d -> %d)", c->mxrange, mxrange);
        return FVC_ERR_STATE;
    }
    if (enable) {
        int rc = ensure_entropy_buffers(c);
        if (rc) return rc;
    }
    c->realbits = enable ? 1 : 0;
    return 0;
}

int64_t fvc_ctx_get_bitstream(fvc_ctx* c, int which, void* out_host, int64_t capacity, void* stream) {
    if (!c || which < 0 || which > 2 || !c->sym_packed) {
        set_error("fvc_ctx_get_bitstream: no bitstream (enable fvc_ctx_set_realbits and run a forward first)");
        return FVC_ERR_STATE;
    }
    cudaStream_t s = (cudaStream_t)stream;
    uint32_t nb = 0;
    unsigned int err[3] = {0, 0, 0};
    FVC_CUDA(cudaMemcpyAsync(&nb, c->stream_bytes + which, 4, cudaMemcpyDeviceToHost, s));
    FVC_CUDA(cudaMemcpyAsync(err, c->ent_err, 12, cudaMemcpyDeviceToHost, s));
    FVC_CUDA(cudaStreamSynchronize(s));
    if (err[0] || err[1]) {
        set_error("entropy coding: %u symbols outside [-mxrange, mxrange-2] (the reference's torchac call raises on these), "
                  "%u empty probability intervals", err[0], err[1]);
        cudaMemsetAsync(c->ent_err, 0, 12, s);
        return FVC_ERR_STATE;
    }
    if (out_host) {
        if ((int64_t)nb > capacity) { set_error("fvc_ctx_get_bitstream: capacity %lld < %u bytes", (long long)capacity, nb); return FVC_ERR_ARG; }
        FVC_CUDA(cudaMemcpyAsync(out_host, c->stream[which], nb, cudaMemcpyDeviceToHost, s));
        FVC_CUDA(cudaStreamSynchronize(s));
    }
    return (int64_t)nb;
}

/* The decoder of the codec: the three byte streams (device pointers) -> latents -> fvc_decode_from_latents' path.
 * Order as a receiver must do it: z (factorized) -> sigma = respriorDecoder(z_hat) -> feature (Laplace(0, sigma)) ; mv. */
int fvc_decode_bitstreams(fvc_ctx* c, const float* ref, const void* feat_stream, int64_t feat_bytes,
                          const void* z_stream, int64_t z_bytes, const void* mv_stream, int64_t mv_bytes, float* recon_out,
                          void* stream) {
    FVC_ARG(c && ref && feat_stream && z_stream && mv_stream && recon_out);
    if (fvc_ctx_missing_params(c) != 0) {
        set_error("fvc_decode_bitstreams: %d parameters not set", fvc_ctx_missing_params(c));
        return FVC_ERR_STATE;
    }
    cudaStream_t s = (cudaStream_t)stream;
    int64_t before = g_launch_count;
    c->conv_event_used = 0;
    int rc = ensure_entropy_buffers(c);
    const int R = c->mxrange, L = c->rans_L;
    int nloss = 0;
    if (!rc) rc = launch_cdf_table_factorized(be_params(c->be_z), 64, R, c->cdf_tab_z, s);
    if (!rc) rc = launch_cdf_table_factorized(be_params(c->be_mv), 128, R, c->cdf_tab_mv, s);
    // z_hat (the pre-round buffers are reused to hold the decoded integer values)
    if (!rc) rc = launch_rans_decode_factorized((const uint8_t*)z_stream, z_bytes, latent_count(c, 1), L, 64, R, c->cdf_tab_z, c->z, c->ent_err, s);
    if (!rc) rc = launch_nhwc_to_act(c->z, c->z_hat, 64, 0, s);
    if (!rc) rc = run_prior_decoder(c, s);
    if (!rc) rc = launch_rans_decode_laplace((const uint8_t*)feat_stream, feat_bytes, latent_count(c, 0), L, R, c->sigma, c->feature, c->ent_err, s);
    if (!rc) rc = launch_nhwc_to_act(c->feature, c->feat_hat, 96, 0, s);
    if (!rc) rc = launch_rans_decode_factorized((const uint8_t*)mv_stream, mv_bytes, latent_count(c, 2), L, 128, R, c->cdf_tab_mv, c->mvfeature, c->ent_err, s);
    if (!rc) rc = launch_nhwc_to_act(c->mvfeature, c->quant_mv, 128, 0, s);
    if (!rc) rc = run_mv_decoder(c, s);
    if (!rc) rc = run_motion_comp(c, nullptr, ref, s);
    if (!rc) rc = run_res_decoder(c, s);
    if (!rc) rc = launch_recon_losses(ref, c->prediction, c->warpframe, c->recon_res, 1, c->B, c->H * c->W, recon_out,
                                      c->loss_partials, &nloss, s, 0);
    c->launches += g_launch_count - before;
    if (rc) return rc;
    unsigned int err[3] = {0, 0, 0};
    FVC_CUDA(cudaMemcpyAsync(err, c->ent_err, 12, cudaMemcpyDeviceToHost, s));
    FVC_CUDA(cudaStreamSynchronize(s));
    if (err[2]) {
        set_error("fvc_decode_bitstreams: %u lanes could not be opened (wrong container, size or geometry)", err[2]);
        cudaMemsetAsync(c->ent_err, 0, 12, s);
        return FVC_ERR_ARG;
    }
    return 0;
}

int64_t fvc_ctx_launch_count(fvc_ctx* c) { return c ? c->launches : -1; }
const char* fvc_ctx_profile_text(fvc_ctx* c) { return c ? c->profile_text.c_str() : ""; }
double fvc_ctx_last_conv_seconds(fvc_ctx* c) { return c ? c->last_conv_seconds : -1.0; }

int64_t fvc_ctx_get_tensor(fvc_ctx* c, const char* name_c, float* out, int64_t capacity, void* stream) {
    if (!c || !name_c || !out) { set_error("fvc_ctx_get_tensor: null argument"); return FVC_ERR_ARG; }
    cudaStream_t s = (cudaStream_t)stream;
    std::string n(name_c);
    const int B = c->B, H = c->H, W = c->W;
    const int h16 = H / 16, w16 = W / 16;
    struct F32 { const float* p; int C, h, w; };
    std::map<std::string, F32> f32 = {
        {"estmv", {c->sflow[c->levels - 1], 2, H, W}}, {"mvfeature", {c->mvfeature, 128, h16, w16}},
        {"mv_hat", {c->mv_hat, 2, H, W}},               {"feature", {c->feature, 96, h16, w16}},
        {"z", {c->z, 64, H / 64, W / 64}},              {"sigma", {c->sigma, 96, h16, w16}},
        {"recon_res", {c->recon_res, 3, H, W}},         {"warpnet_res", {c->wres, 3, H, W}}};
    std::map<std::string, std::pair<ActT, int>> act = {{"quant_mv", {c->quant_mv, 128}},
                                                       {"z_hat", {c->z_hat, 64}},
                                                       {"feat_hat", {c->feat_hat, 96}},
                                                       {"residual", {c->residual, 3}},
                                                       {"warpnet_c0", {c->wc0, 64}},
                                                       {"warpnet_c5", {c->wc5, 64}},
                                                       {"mvenc_e1", {c->e[1], 128}},
                                                       {"mvdec_d7", {c->d[7], 128}},
                                                       {"resenc_r0", {c->r[0], 64}},
                                                       {"resdec_g2", {c->g[2], 64}}};
    std::map<std::string, const float*> planar = {{"warpframe", c->warpframe}, {"prediction", c->prediction}};
    auto fi = f32.find(n);
    if (fi != f32.end()) {
        int64_t cnt = (int64_t)B * fi->second.C * fi->second.h * fi->second.w;
        if (cnt > capacity) { set_error("capacity too small"); return FVC_ERR_ARG; }
        int rc = launch_nhwc_to_nchw(fi->second.p, out, B, fi->second.C, fi->second.h, fi->second.w, s);
        return rc ? rc : cnt;
    }
    if ((c->tail_fused && n == "mvdec_d7") || (c->tail_fused_warpnet && n == "warpnet_c5")) {
        set_error("%s is not materialised: it stays in the SM (fused tail convolution; FVC_TAIL_FUSED=0 restores it)", name_c);
        return FVC_ERR_STATE;
    }
    auto ai = act.find(n);
    if (ai != act.end()) {
        ActT t = ai->second.first;
        int64_t cnt = (int64_t)B * ai->second.second * t.H * t.W;
        if (cnt > capacity) { set_error("capacity too small"); return FVC_ERR_ARG; }
        int rc = launch_act_to_nchw(t, ai->second.second, out, s);
        return rc ? rc : cnt;
    }
    auto pi = planar.find(n);
    if (pi != planar.end()) {
        int64_t cnt = (int64_t)B * 3 * H * W;
        if (cnt > capacity) { set_error("capacity too small"); return FVC_ERR_ARG; }
        if (cudaMemcpyAsync(out, pi->second, cnt * 4, cudaMemcpyDeviceToDevice, s) != cudaSuccess) return FVC_ERR_CUDA;
        return cnt;
    }
    set_error("unknown tensor %s", name_c);
    return FVC_ERR_ARG;
}

static int gop_forward_host_impl(fvc_ctx* c, const void* frames_host_any, bool u8, int G, float* recon_host,
                                 float* scalars_host, void* stream) {
    FVC_ARG(c && frames_host_any && scalars_host && G >= 2);
    cudaStream_t s = (cudaStream_t)stream;
    const size_t fsz = (size_t)c->B * 3 * c->H * c->W;
    const float* frames_host = u8 ? nullptr : static_cast<const float*>(frames_host_any);
    const uint8_t* frames_u8 = u8 ? static_cast<const uint8_t*>(frames_host_any) : nullptr;
    if (u8 && c->stage_G_u8 < G) {
        if (c->stage_u8) cudaFree(c->stage_u8);
        FVC_CUDA(cudaMalloc(&c->stage_u8, fsz * G));
        c->stage_G_u8 = G;
    }
    if (c->stage_G < G) {
        if (c->stage_frames) cudaFree(c->stage_frames);
        if (c->stage_rec) cudaFree(c->stage_rec);
        if (c->stage_scalars) cudaFree(c->stage_scalars);
        FVC_CUDA(cudaMalloc(&c->stage_frames, fsz * G * 4));
        FVC_CUDA(cudaMalloc(&c->stage_rec, fsz * (G - 1) * 4));
        FVC_CUDA(cudaMalloc(&c->stage_scalars, (size_t)(G - 1) * 7 * 4));
        c->stage_G = G;
    }
    // Frames are uploaded one by one on a copy stream; P-frame i only waits for frames i-1 and i, so the
    // upload of the rest of the GOP overlaps the computation (10 x 25 MB at 1080p = ~5 ms over PCIe otherwise).
    if (!c->copy_stream) {
        FVC_CUDA(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
        FVC_CUDA(cudaEventCreateWithFlags(&c->copy_fence, cudaEventDisableTiming));
    }
    while ((int)c->copy_events.size() < G) {
        cudaEvent_t e;
        FVC_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        c->copy_events.push_back(e);
    }
    // the staging buffer may still be read by work queued on `s` from a previous call
    FVC_CUDA(cudaEventRecord(c->copy_fence, s));
    FVC_CUDA(cudaStreamWaitEvent(c->copy_stream, c->copy_fence, 0));
    for (int i = 0; i < G; ++i) {
        if (u8) {
            // 1 byte per sample over PCIe; ToTensor (HWC -> CHW, / 255) on the device, on the copy stream as well
            FVC_CUDA(cudaMemcpyAsync(c->stage_u8 + (size_t)i * fsz, frames_u8 + (size_t)i * fsz, fsz, cudaMemcpyHostToDevice,
                                     c->copy_stream));
            int rc = launch_u8hwc_to_f32chw(c->stage_u8 + (size_t)i * fsz, c->stage_frames + (size_t)i * fsz, c->B, c->H, c->W,
                                            c->copy_stream);
            if (rc) return rc;
        } else {
            FVC_CUDA(cudaMemcpyAsync(c->stage_frames + (size_t)i * fsz, frames_host + (size_t)i * fsz, fsz * 4,
                                     cudaMemcpyHostToDevice, c->copy_stream));
        }
        FVC_CUDA(cudaEventRecord(c->copy_events[i], c->copy_stream));
    }
    FVC_CUDA(cudaStreamWaitEvent(s, c->copy_events[0], 0));
    const float* prev = c->stage_frames;  // decoded I-frame (models.py:370)
    for (int i = 1; i < G; ++i) {
        float* rec = c->stage_rec + (size_t)(i - 1) * fsz;
        FVC_CUDA(cudaStreamWaitEvent(s, c->copy_events[i], 0));
        int rc = fvc_pframe_forward(c, c->stage_frames + (size_t)i * fsz, prev, rec, c->stage_scalars + (i - 1) * 7,
                                    stream);
        if (rc) return rc;
        prev = rec;  // x_prev = clipped_recon (models.py:372-375)
    }
    if (recon_host)
        FVC_CUDA(cudaMemcpyAsync(recon_host, c->stage_rec, fsz * (G - 1) * 4, cudaMemcpyDeviceToHost, s));
    FVC_CUDA(cudaMemcpyAsync(scalars_host, c->stage_scalars, (size_t)(G - 1) * 7 * 4, cudaMemcpyDeviceToHost, s));
    unsigned int sat = 0;
    FVC_CUDA(cudaMemcpyAsync(&sat, c->sat_count, 4, cudaMemcpyDeviceToHost, s));
    FVC_CUDA(cudaStreamSynchronize(s));
    if (sat) {
        set_error("fvc_gop_forward_host: activations reached the fp16 operand-pair range (|v| >= 65504) in %u epilogue "
                  "tiles and were clamped; results are invalid for these weights (see fvc_ctx_saturation_count)", sat);
        return FVC_ERR_STATE;
    }
    return 0;
}

int fvc_gop_forward_host(fvc_ctx* c, const float* frames_host, int G, float* recon_host, float* scalars_host,
                         void* stream) {
    return gop_forward_host_impl(c, frames_host, false, G, recon_host, scalars_host, stream);
}

int fvc_gop_forward_host_u8(fvc_ctx* c, const uint8_t* frames_host_u8, int G, float* recon_host, float* scalars_host,
                            void* stream) {
    return gop_forward_host_impl(c, frames_host_u8, true, G, recon_host, scalars_host, stream);
}

}  // extern "C"
