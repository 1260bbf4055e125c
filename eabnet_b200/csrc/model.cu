// Model plan, parameter table, weight packing and the C ABI (include/eabnet_b200.h).
//
// The architecture walk below restates EaBNet.__init__/forward (EaBNet.py:9-125 and the sub-modules at
// :157-624) as a list of layer descriptors that (a) declare the reference state_dict entries in the reference's
// registration order, (b) pack them into the kernels' layouts, and (c) launch the kernels.  Activations are
// channels-last [B,T,F,C] fp32 and are stored RAW next to their (sum, sumsq) statistics; normalisation and
// PReLU are applied by the consumer while it stages its operand (see common.cuh Xform).
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include <cuda_fp16.h>

#include "../../include/eabnet_b200.h"
#include "common.cuh"

namespace eab {

// ------------------------------------------------------------------------------------------------ errors
static thread_local std::string g_err;
static thread_local int g_launches = 0;
void set_error(const std::string& m) { g_err = m; }
int fail(const std::string& m) { g_err = m; return 1; }
int check_cuda(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return 0;
    g_err = std::string(what) + ": " + cudaGetErrorString(e);
    return 1;
}
namespace {
struct DevState { int sms = 0; std::map<const void*, int> smem; };
std::mutex g_dev_mu;
std::map<int, DevState> g_dev;
}  // namespace
int device_sm_count(int* sms) {
    int dev = 0;
    EAB_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(g_dev_mu);
    DevState& d = g_dev[dev];
    if (!d.sms) EAB_CUDA(cudaDeviceGetAttribute(&d.sms, cudaDevAttrMultiProcessorCount, dev));
    *sms = d.sms;
    return 0;
}
int ensure_dynamic_smem(const void* kernel, int bytes) {
    int dev = 0;
    EAB_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(g_dev_mu);
    int& cur = g_dev[dev].smem[kernel];
    if (bytes > cur) {
        EAB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
        cur = bytes;
    }
    return 0;
}
void count_launch(int n) { g_launches += n; }

// ------------------------------------------------------------------------------------------------ profiler
struct ProfRec { const char* cat; double flops, bytes, moved; cudaEvent_t a, b; };
struct Profiler {
    bool on = false;
    bool detail = false;
    std::vector<ProfRec> recs;
    std::vector<cudaEvent_t> pool;
    cudaEvent_t get() {
        cudaEvent_t e;
        if (!pool.empty()) { e = pool.back(); pool.pop_back(); return e; }
        cudaEventCreate(&e);
        return e;
    }
    void clear() {
        for (auto& r : recs) { pool.push_back(r.a); pool.push_back(r.b); }
        recs.clear();
    }
};
static thread_local Profiler g_prof;

ProfScope::ProfScope(const char* category, double flops, double bytes, cudaStream_t s, double moved) : rec(nullptr), st(s) {
    if (!g_prof.on) return;
    ProfRec r{category, flops, bytes, moved < 0 ? bytes : moved, g_prof.get(), g_prof.get()};
    cudaEventRecord(r.a, st);
    g_prof.recs.push_back(r);
    rec = reinterpret_cast<void*>(g_prof.recs.size());      // index + 1
}
ProfScope::~ProfScope() {
    if (!rec) return;
    cudaEventRecord(g_prof.recs[reinterpret_cast<size_t>(rec) - 1].b, st);
}
bool g_prof_on() { return g_prof.on; }
int launch_count() { return g_launches; }
void reset_launch_count() { g_launches = 0; }

namespace {

inline int ceil64(int x) { return (x + 63) / 64 * 64; }
inline int pad_n(int x) { int n = ceil64(x); return n == 192 ? 256 : n; }

struct Param {
    std::string name;
    int ndim = 0;
    int64_t shape[4] = {1, 1, 1, 1};
    int kind = 0;
    int fan_in = 1;
    std::vector<float> host;
    bool set = false;
    int64_t numel() const { int64_t n = 1; for (int i = 0; i < ndim; ++i) n *= shape[i]; return n; }
};

// norm (optional) + PReLU following a conv (2-D: conv -> norm -> PReLU; TCM: PReLU -> norm)
struct NormAct {
    bool has_norm = false;
    int C = 0;
    int gamma = -1, beta = -1, mean = -1, var = -1, alpha = -1;   // param indices
    size_t off_scale = 0, off_shift = 0, off_alpha = 0;           // floats into the device blob
    bool alpha01 = false;                                         // every PReLU slope lies in [0, 1]: PReLU(z) == max(z, a z)
};

struct ConvLayer {
    int w = -1, b = -1;                 // param indices
    int cin = 0, cout = 0, kt = 1, kf = 1;
    bool deconv = false, gated = false;
    bool perm_ri = false;               // first layer: reference channel ri*M+m  ->  memory order m*2+ri
    int M = 0;
    int N = 0, gate_off = 0;
    // packed variants: conv -> 1, deconv -> 2 (output parity)
    int nvar = 1;
    int ntaps[2] = {0, 0};
    int dt[2][kMaxTaps], df[2][kMaxTaps];
    size_t off_w[2] = {0, 0}, off_b = 0;
    // tcgen05 path: swizzled TF32 weight images (hi / lo), per variant
    bool umma_ok = false, wide = false;
    int zone = 0;                       // 0 encoder, 1 decoder, 2 encoder inner U-Nets (precision policy)
    int u_nslab = 0, u_kwidth = 0, u_N = 0, u_gate_off = 0;
    int u_ntaps[2] = {0, 0};
    int u_dt[2][kMaxTaps], u_df[2][kMaxTaps];
    size_t off_whi[2] = {0, 0}, off_wlo[2] = {0, 0}, off_ub = 0;
    // first layer, "pair" layout: a plane row holds two adjacent frequency positions (2 x cin values, one 64-wide slab); an
    // output position reads ceil(kf / 2) consecutive rows, so taps = kt x ceil(kf / 2) row shifts (stride-2 conv only)
    bool pair_ok = false;
    int p_ntaps = 0, p_dt[kMaxTaps], p_ds[kMaxTaps];
    size_t off_phi = 0, off_plo = 0;
    NormAct na;
};

struct UnetModule {
    ConvLayer in_conv;
    std::vector<ConvLayer> enco, deco;
};

// weight images of one pointwise / dilated GEMM on the tcgen05 path (columns split in chunks of <= 128)
struct UmmaW {
    bool ok = false;
    int ntaps = 1, nslab = 0, gate_off = 0;
    int nsplit = 1, ncol = 0;            // columns (N) per split
    int cout = 0;                        // stored channels per split
    int ld = 0;                          // row stride of the output tensor (channels)
    size_t off_hi[4] = {0, 0, 0, 0}, off_lo[4] = {0, 0, 0, 0}, off_bias[4] = {0, 0, 0, 0};
    bool has_bias = false;
};

struct TcmLayer {
    int dilation = 1;
    bool single = false;                 // GaGNet's SqueezedTCM: one dilated branch, no gate (GaGNet.py:285-326)
    bool perm = true;                    // residual stream in bottleneck order f*64+c (EaBNet); false = reference order
    UmmaW u_in, u_dil, u_out;
    UmmaW u_dl, u_dr;                    // gated TCM: the two dilated branches as separate [kd][64][64] image sets (tcm_chain.cu)
    int w_in = -1, w_left = -1, w_right = -1, w_out = -1;
    NormAct na_left, na_right, na_out;
    size_t off_in = 0, off_dil = 0, off_out = 0;
    int dt[kMaxTaps];
};

// GaGNet glance / gaze blocks (GaGNet.py:136-259)
struct GagIn {                       // in_conv_main(cat) * sigmoid(in_conv_gate(cat)) as d_feat/64 gated column splits
    int w_main = -1, b_main = -1, w_gate = -1, b_gate = -1;
    int nsplit = 0, K = 0, SW = 64;      // SW: value (= gate) columns per split
    size_t off_dense[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    UmmaW u[8];
};
struct GagLin {                      // Conv1d(d_feat, F, 1) with bias
    int w = -1, b = -1, N = 0;
    size_t off_w = 0, off_b = 0;
    UmmaW u;
};
struct GagModule {
    GagIn in_g, in_z;
    std::vector<TcmLayer> tcn_g, tcm_r, tcm_i;       // is_squeezed: tcm_i empty, tcm_r holds `tcm_ri`
    GagLin lin_g, lin_r, lin_i;
};

struct Act {                 // an activation tensor as seen by a consumer
    float* data = nullptr;
    int F = 0, C = 0;
    Xform xf = xform_identity();
    // lazy residual sum: value = xf(data) + xf2(data2) when data2 != null (never materialised on the staged path)
    float* data2 = nullptr;
    Xform xf2 = xform_identity();
    int RT = 0;              // streaming: frames in this tensor's ring (0 = offline)
};

inline void set_src(ConvSrc& s, const Act& a) {
    s.x = a.data; s.C = a.C; s.xf = a.xf; s.x2 = a.data2; s.xf2 = a.xf2; s.RT = a.RT;
}

struct Tap { Act act; int B = 0, T = 0; };

}  // namespace
}  // namespace eab

using namespace eab;

struct eab_model {
    eab_config cfg;
    int kind = 0;                                   // 0 EaBNet, 1 GaGNet post-filter (eab_gag_create)
    eab_gag_config gcfg;
    std::vector<GagModule> gags;
    std::vector<Param> params;
    std::unordered_map<std::string, int> index;

    // architecture
    std::vector<UnetModule> en_mod, de_mod;         // U2 variants
    std::vector<ConvLayer> en_plain, de_plain;      // U-Net variants (and the U2 last convs at the back)
    ConvLayer en_last, de_last;
    std::vector<TcmLayer> tcms;                     // q*p
    int Fb = 0;                                     // bottleneck F
    std::vector<int> Fchain;                        // encoder F sizes: F0 (input) .. F5
    // head
    int rnn[2][4] = {{-1, -1, -1, -1}, {-1, -1, -1, -1}};
    int dnn_w[2] = {-1, -1}, dnn_b[2] = {-1, -1}, ln_g = -1, ln_b = -1, cnn_w = -1, cnn_b = -1;
    size_t off_rnn[2][3] = {{0, 0, 0}, {0, 0, 0}}, off_dnn_w[2] = {0, 0}, off_dnn_b[2] = {0, 0}, off_ln_g = 0,
           off_ln_b = 0, off_cnn_w = 0, off_cnn_b = 0;
    int dnn_N[2] = {0, 0}, cnn_N = 0;
    UmmaW u_dnn[2], u_cnn;
    size_t off_rnn_img[2] = {0, 0}, off_rnn_ubias[2] = {0, 0};
    bool rnn_umma_ok = false;

    // device state
    float* blob = nullptr;
    size_t blob_floats = 0;
    bool dirty = true;
    int last_launches = 0;
    std::map<std::string, Tap> taps;
    void* scratch = nullptr;      // eab_enhance_host / eab_enhance_host_batches
    size_t scratch_bytes = 0;
    cudaStream_t s_in = nullptr, s_out = nullptr;     // copy streams of the pipelined host front door
    cudaStream_t s_comp = nullptr;                    // its compute stream when the caller passes the legacy default stream (not capturable)
    cudaStream_t s_comp2 = nullptr;                   // second compute stream: odd batches (option dual_stream)
    int opt_dual_stream = 1;
    // the per-slot step of the host front door as a CUDA graph (captured on the slot's second use, replayed afterwards)
    struct SlotGraph { cudaGraphExec_t exec = nullptr; const void* in = nullptr; void* out = nullptr; void* ws = nullptr; int B = 0, L = 0;
                       unsigned long long version = 0, mode = 0; int launches = 0; };
    SlotGraph slot_graph[2];
    unsigned long long param_version = 0;           // bumped by every commit
    int opt_host_graph = 1;
    cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_comp[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr};
    // options (eab_set_option)
    int opt_umma = 1;             // tcgen05 path for eligible layers
    int opt_enc_passes = 3;       // 3xTF32 in the encoder (single-pass TF32 there costs 4.8e-4 of the 1e-3 budget)
    int opt_dec_passes = 1;       // single-pass TF32 in the decoder
    int opt_inner_passes = 3;     // inner U-Nets of the encoder modules
    int opt_first_passes = 3;     // the first gated conv (2M input channels, tap-window rows: 3.5x the input bytes per pass-plane)
    int opt_staged = 1;           // stage_kernel + TMA-fed conv kernel (layers conv_raw does not take)
    int opt_raw = 1;              // conv_raw_kernel: raw fp32 tiles normalised in shared memory, no stage pass (preferred)
    int opt_raw_grid = 0;         // diagnostics / tests: cap on conv_raw's grid size (0 = one CTA per SM)
    int opt_fused_head = 1;       // w_dnn + filter-and-sum as one kernel
    int opt_head_w_tap = 0;       // also write the beam weights (debug tap "w") from the fused head kernel
    int opt_lstm_exp = 0;         // diagnostics (EAB_LSTM_EXPERIMENT builds)
    int opt_stream_tcm = 1;       // streaming: the whole TCM stack as one launch (0 = per-layer kernels)
    int opt_stream_umma = 1;      // streaming: the per-layer convs on the tcgen05 gather kernel (rows = streams x F); 0 = CUDA cores
    int opt_tcm_chain = 1;        // TCM stacks as single launches (tcm_chain.cu): a GaGNet module's three stacks / an EaBNet group; 2: one chain per launch, 3: cooperative grid-barrier form
    int opt_norm_log = 0;         // record where every InstanceNorm's (sum, sum of squares) of a forward live (eab_norm_stats)
    struct NormLog { int gamma; const double* stats; int C, count, B; };
    std::vector<NormLog> norm_log;
    int opt_lazy = 1;             // module residual sums are summed by the consumers' stage kernels, never materialised
    int opt_dbg_launch = -1;      // diagnostics: instrument the n-th tcgen05 conv launch of a forward
    int umma_launch_idx = 0;
    unsigned long long* dbg_buf = nullptr;
};

namespace eab {
namespace {

// ================================================================================================ declare
struct Builder {
    eab_model* m;
    int add(const std::string& name, std::initializer_list<int64_t> shape, int kind, int fan_in) {
        Param p;
        p.name = name;
        p.ndim = (int)shape.size();
        int i = 0;
        for (auto s : shape) p.shape[i++] = s;
        p.kind = kind;
        p.fan_in = fan_in;
        m->index[name] = (int)m->params.size();
        m->params.push_back(p);
        return (int)m->params.size() - 1;
    }
    void norm(const std::string& pfx, int C, NormAct& na) {
        na.has_norm = true;
        na.C = C;
        na.gamma = add(pfx + ".norm.weight", {C}, EAB_P_NORM_G, C);
        na.beta = add(pfx + ".norm.bias", {C}, EAB_P_NORM_B, C);
        if (m->cfg.norm_type == 1) {
            na.mean = add(pfx + ".norm.running_mean", {C}, EAB_P_BN_MEAN, C);
            na.var = add(pfx + ".norm.running_var", {C}, EAB_P_BN_VAR, C);
            add(pfx + ".norm.num_batches_tracked", {}, EAB_P_BN_COUNT, 1);
        }
    }
    // Sequential(gated (de)conv, [norm], PReLU)   (EaBNet.py:185-189, 214-231, 267-271, 351-358)
    ConvLayer gated(const std::string& pfx, int cin, int cout, int kt, int kf, bool deconv, bool with_norm) {
        ConvLayer L;
        L.cin = cin; L.cout = cout; L.kt = kt; L.kf = kf; L.deconv = deconv; L.gated = true;
        const std::string sub = kt > 1 ? (deconv ? ".0.conv.0" : ".0.conv.1") : ".0.conv";
        const int fan = (deconv ? 2 * cout : cin) * kt * kf;      // torch: weight.size(1) * receptive field
        if (deconv) L.w = add(pfx + sub + ".weight", {cin, 2 * cout, kt, kf}, EAB_P_CONV_W, fan);
        else        L.w = add(pfx + sub + ".weight", {2 * cout, cin, kt, kf}, EAB_P_CONV_W, fan);
        L.b = add(pfx + sub + ".bias", {2 * cout}, EAB_P_CONV_B, fan);
        L.na.C = cout;
        if (with_norm) {
            norm(pfx + ".1", cout, L.na);
            L.na.alpha = add(pfx + ".2.weight", {cout}, EAB_P_PRELU, cout);
        } else {
            L.na.alpha = add(pfx + ".1.weight", {cout}, EAB_P_PRELU, cout);
        }
        return L;
    }
    // Conv2dunit / Deconv2dunit (EaBNet.py:391-431)
    ConvLayer unit(const std::string& pfx, int cin, int cout, int kt, int kf, bool deconv) {
        ConvLayer L;
        L.cin = cin; L.cout = cout; L.kt = kt; L.kf = kf; L.deconv = deconv; L.gated = false;
        const int fan = (deconv ? cout : cin) * kt * kf;
        if (deconv) L.w = add(pfx + ".0.weight", {cin, cout, kt, kf}, EAB_P_CONV_W, fan);
        else        L.w = add(pfx + ".0.weight", {cout, cin, kt, kf}, EAB_P_CONV_W, fan);
        L.b = add(pfx + ".0.bias", {cout}, EAB_P_CONV_B, fan);
        norm(pfx + ".1", cout, L.na);
        L.na.alpha = add(pfx + ".2.weight", {cout}, EAB_P_PRELU, cout);
        return L;
    }
    UnetModule module(const std::string& pfx, int cin, int kt, int kf, int scale, bool deconv) {
        const eab_config& c = m->cfg;
        UnetModule U;
        U.in_conv = gated(pfx + ".in_conv", cin, c.c, kt, kf, deconv, true);
        for (int i = 0; i < scale; ++i)
            U.enco.push_back(unit(pfx + ".enco." + std::to_string(i) + ".conv", c.c, c.c, c.k2_t, c.k2_f, false));
        for (int i = 0; i < scale; ++i) {
            const int cin_d = (i == 0 || c.intra_connect == 1) ? c.c : 2 * c.c;
            U.deco.push_back(unit(pfx + ".deco." + std::to_string(i) + ".deconv", cin_d, c.c, c.k2_t, c.k2_f, true));
        }
        return U;
    }
};

// element (n, k) of a [N][64] fp16 K-major tile with the 128-byte swizzle the tensor core expects (index in halves)
inline size_t sw128_index_h(int n, int k) { return (size_t)n * 64 + (size_t)((((k >> 3) ^ (n & 7)) << 3) | (k & 7)); }

int conv_out_f(int Fin, int kf) { return Fin < kf ? -1 : (Fin - kf) / 2 + 1; }
int deconv_out_f(int Fin, int kf) { return 2 * (Fin - 1) + kf; }

int build(eab_model* m) {
    const eab_config& c = m->cfg;
    if (c.c < 1 || c.c > 128) return fail("c must be in 1..128");
    if (c.embed_dim < 1 || c.embed_dim > 128) return fail("embed_dim must be in 1..128");
    if (c.M < 1 || c.M > 64) return fail("M must be in 1..64");
    if (c.kd1 < 1 || c.kd1 > kMaxTaps) return fail("kd1 must be in 1..16");
    if (c.k1_t < 1 || c.k1_t > 2 || c.k2_t < 1 || c.k2_t > 2) return fail("temporal kernel sizes above 2 are not supported");
    if (c.k1_t * c.k1_f > kMaxTaps || c.k2_t * c.k2_f > kMaxTaps || c.k1_f < 1 || c.k2_f < 1) return fail("kernel too large");
    if (c.cd1 < 1 || c.cd1 > 128) return fail("cd1 must be in 1..128");
    if (c.p < 1 || c.q < 1 || c.q > 3) return fail("p >= 1 and 1 <= q <= 3 required");
    if (c.norm_type != 0 && c.norm_type != 1)
        return fail("norm_type 'cLN' cannot be constructed in the reference either (EaBNet.py:689,691)");
    Builder bd{m};
    // encoder F chain
    m->Fchain.clear();
    m->Fchain.push_back(c.n_freq);
    {
        int F = conv_out_f(c.n_freq, 5);
        m->Fchain.push_back(F);
        for (int i = 0; i < 4; ++i) { F = F > 0 ? conv_out_f(F, c.k1_f) : -1; m->Fchain.push_back(F); }
        if (F < 1) return fail("n_freq too small for the five stride-2 encoder stages");
        m->Fb = F;
    }
    if (c.d_feat != 64 * m->Fb)
        return fail("d_feat must equal 64 * bottleneck_F (the reference fails at run time otherwise, EaBNet.py:100,549)");
    if (c.is_u2) {
        m->en_mod.push_back(bd.module("en.meta_unet_list.0", 2 * c.M, 2, 5, 4, false));
        for (int i = 1; i < 4; ++i)
            m->en_mod.push_back(bd.module("en.meta_unet_list." + std::to_string(i), c.c, c.k1_t, c.k1_f, 4 - i, false));
        m->en_mod[0].in_conv.perm_ri = true;
        m->en_mod[0].in_conv.M = c.M;
        m->en_mod[0].in_conv.zone = 3;
        m->en_last = bd.gated("en.last_conv", c.c, 64, c.k1_t, c.k1_f, false, true);
        m->de_mod.push_back(bd.module("de.meta_unet_list.0", 128, c.k1_t, c.k1_f, 1, true));
        for (int i = 1; i < 4; ++i)
            m->de_mod.push_back(bd.module("de.meta_unet_list." + std::to_string(i), 2 * c.c, c.k1_t, c.k1_f, i + 1, true));
        m->de_last = bd.gated("de.last_conv", 2 * c.c, c.embed_dim, 2, 5, true, true);
        for (auto& U : m->de_mod) { U.in_conv.zone = 1; for (auto& L : U.enco) L.zone = 1; for (auto& L : U.deco) L.zone = 1; }
        m->de_last.zone = 1;
        // zone 2: the inner U-Nets of the encoder modules (their result is the residual branch of x0 + y)
        for (auto& U : m->en_mod) { for (auto& L : U.enco) L.zone = 2; for (auto& L : U.deco) L.zone = 2; }
    } else {
        m->en_plain.push_back(bd.gated("en.unet_list.0", 2 * c.M, c.c, 2, 5, false, true));
        m->en_plain[0].perm_ri = true;
        m->en_plain[0].M = c.M;
        m->en_plain.push_back(bd.gated("en.unet_list.1", c.c, c.c, c.k1_t, c.k1_f, false, false));
        m->en_plain.push_back(bd.gated("en.unet_list.2", c.c, c.c, c.k1_t, c.k1_f, false, false));
        m->en_plain.push_back(bd.gated("en.unet_list.3", c.c, c.c, c.k1_t, c.k1_f, false, true));
        m->en_plain.push_back(bd.gated("en.unet_list.4", c.c, 64, c.k1_t, c.k1_f, false, true));
        m->de_plain.push_back(bd.gated("de.unet_list.0", 128, c.c, c.k1_t, c.k1_f, true, true));
        for (int i = 1; i < 4; ++i)
            m->de_plain.push_back(bd.gated("de.unet_list." + std::to_string(i), 2 * c.c, c.c, c.k1_t, c.k1_f, true, true));
        m->de_plain.push_back(bd.gated("de.unet_list.4", 2 * c.c, c.embed_dim, 2, 5, true, true));
        for (auto& L : m->de_plain) L.zone = 1;
    }
    // head (EaBNet.py:75-81, 581-598)
    if (c.topo_type == 0 && c.bf_type == 0) {
        const int H = 64;
        for (int r = 0; r < 2; ++r) {
            const std::string p = std::string("bf_map.rnn") + (r ? "2" : "1");
            const int cin = r ? H : c.embed_dim;
            m->rnn[r][0] = bd.add(p + ".weight_ih_l0", {4 * H, cin}, EAB_P_LSTM, H);
            m->rnn[r][1] = bd.add(p + ".weight_hh_l0", {4 * H, H}, EAB_P_LSTM, H);
            m->rnn[r][2] = bd.add(p + ".bias_ih_l0", {4 * H}, EAB_P_LSTM, H);
            m->rnn[r][3] = bd.add(p + ".bias_hh_l0", {4 * H}, EAB_P_LSTM, H);
        }
        m->dnn_w[0] = bd.add("bf_map.w_dnn.0.weight", {H, H}, EAB_P_LIN_W, H);
        m->dnn_b[0] = bd.add("bf_map.w_dnn.0.bias", {H}, EAB_P_LIN_B, H);
        m->dnn_w[1] = bd.add("bf_map.w_dnn.2.weight", {2 * c.M, H}, EAB_P_LIN_W, H);
        m->dnn_b[1] = bd.add("bf_map.w_dnn.2.bias", {2 * c.M}, EAB_P_LIN_B, H);
        m->ln_g = bd.add("bf_map.norm.weight", {c.embed_dim}, EAB_P_NORM_G, c.embed_dim);
        m->ln_b = bd.add("bf_map.norm.bias", {c.embed_dim}, EAB_P_NORM_B, c.embed_dim);
    } else {
        const int n = c.topo_type == 0 ? 2 * c.M : 2;
        m->cnn_w = bd.add("bf_map.weight", {n, c.embed_dim, 1, 1}, EAB_P_CONV_W, c.embed_dim);
        m->cnn_b = bd.add("bf_map.bias", {n}, EAB_P_CONV_B, c.embed_dim);
    }
    // squeezed TCMs (EaBNet.py:83-86, 506-571)
    for (int g = 0; g < c.q; ++g)
        for (int i = 0; i < c.p; ++i) {
            if (i > 24) return fail("p too large (dilation 2^i overflows)");
            TcmLayer t;
            t.dilation = 1 << i;
            const std::string p = "stcns." + std::to_string(g) + ".tcm_list." + std::to_string(i);
            t.w_in = bd.add(p + ".in_conv.weight", {c.cd1, c.d_feat, 1}, EAB_P_CONV_W, c.d_feat);
            t.na_left.C = t.na_right.C = t.na_out.C = c.cd1;
            t.na_left.alpha = bd.add(p + ".left_conv.0.weight", {c.cd1}, EAB_P_PRELU, c.cd1);
            bd.norm(p + ".left_conv.1", c.cd1, t.na_left);
            t.w_left = bd.add(p + ".left_conv.3.weight", {c.cd1, c.cd1, c.kd1}, EAB_P_CONV_W, c.cd1 * c.kd1);
            t.na_right.alpha = bd.add(p + ".right_conv.0.weight", {c.cd1}, EAB_P_PRELU, c.cd1);
            bd.norm(p + ".right_conv.1", c.cd1, t.na_right);
            t.w_right = bd.add(p + ".right_conv.3.weight", {c.cd1, c.cd1, c.kd1}, EAB_P_CONV_W, c.cd1 * c.kd1);
            t.na_out.alpha = bd.add(p + ".out_conv.0.weight", {c.cd1}, EAB_P_PRELU, c.cd1);
            bd.norm(p + ".out_conv.1", c.cd1, t.na_out);
            t.w_out = bd.add(p + ".out_conv.2.weight", {c.d_feat, c.cd1, 1}, EAB_P_CONV_W, c.cd1);
            const int span = (c.kd1 - 1) * t.dilation;
            if (!c.is_causal && (span & 1)) return fail("non-causal TCM needs an even (kd1-1)*dilation");
            const int pad_left = c.is_causal ? span : span / 2;
            for (int k = 0; k < c.kd1; ++k) t.dt[k] = pad_left - k * t.dilation;
            m->tcms.push_back(t);
        }
    return 0;
}

// GaGNet.__init__ (GaGNet.py:69-73): encoder on cat(inpt, pre_x), then q glance-gaze modules.  m->cfg carries the
// settings the shared builder / packer / runner code reads (M = cin makes the first layer's 2M input channels the
// reference's cin*2 and its (ri, m) -> (m, ri) weight permutation the one gag_pack_kernel's channel order needs).
int build_gag(eab_model* m) {
    const eab_gag_config& g = m->gcfg;
    eab_config& c = m->cfg;
    if (g.cin != 2) return fail("GaGNet: cin must be 2 (the reference's glance / gaze 1x1 convs take 2*(fft_num/2+1) + d_feat channels, GaGNet.py:160,224)");
    if (g.n_dilas < 1 || g.n_dilas > 8) return fail("GaGNet: 1..8 dilation rates");
    if (g.fft_num < 2 || (g.fft_num & 1)) return fail("GaGNet: fft_num must be even");
    if (g.acti_type < 0 || g.acti_type > 2) return fail("GaGNet: a activation function must be assigned! (GaGNet.py:171-172)");
    memset(&c, 0, sizeof(c));
    c.k1_t = g.k1_t; c.k1_f = g.k1_f; c.k2_t = g.k2_t; c.k2_f = g.k2_f; c.c = g.c; c.M = g.cin; c.embed_dim = 64;
    c.kd1 = g.kd1; c.cd1 = g.cd1; c.d_feat = g.d_feat; c.p = g.p; c.q = g.q; c.is_causal = g.is_causal; c.is_u2 = g.is_u2;
    c.intra_connect = g.intra_connect; c.norm_type = g.norm_type; c.n_freq = g.fft_num / 2 + 1;
    if (c.c < 1 || c.c > 128) return fail("c must be in 1..128");
    if (c.kd1 < 1 || c.kd1 > kMaxTaps) return fail("kd1 must be in 1..16");
    if (c.k1_t < 1 || c.k1_t > 2 || c.k2_t < 1 || c.k2_t > 2) return fail("temporal kernel sizes above 2 are not supported");
    if (c.k1_t * c.k1_f > kMaxTaps || c.k2_t * c.k2_f > kMaxTaps || c.k1_f < 1 || c.k2_f < 1) return fail("kernel too large");
    if (c.cd1 < 1 || c.cd1 > 128) return fail("cd1 must be in 1..128");
    if (c.p < 1 || c.q < 1 || c.q > 16) return fail("p >= 1 and 1 <= q <= 16 required");
    if (c.norm_type != 0 && c.norm_type != 1) return fail("norm_type must be 'IN' or 'BN'");
    Builder bd{m};
    m->Fchain.clear();
    m->Fchain.push_back(c.n_freq);
    {
        int F = conv_out_f(c.n_freq, 5);
        m->Fchain.push_back(F);
        for (int i = 0; i < 4; ++i) { F = F > 0 ? conv_out_f(F, c.k1_f) : -1; m->Fchain.push_back(F); }
        if (F < 1) return fail("fft_num too small for the five stride-2 encoder stages");
        m->Fb = F;
    }
    if (c.d_feat != 64 * m->Fb) return fail("d_feat must equal 64 * bottleneck_F (the reference fails at run time otherwise)");
    if (c.d_feat / 64 > 8) return fail("GaGNet: d_feat above 512 is not supported");
    if (c.is_u2) {
        m->en_mod.push_back(bd.module("en.meta_unet_list.0", 2 * c.M, 2, 5, 4, false));
        for (int i = 1; i < 4; ++i)
            m->en_mod.push_back(bd.module("en.meta_unet_list." + std::to_string(i), c.c, c.k1_t, c.k1_f, 4 - i, false));
        m->en_mod[0].in_conv.perm_ri = true;
        m->en_mod[0].in_conv.M = c.M;
        m->en_last = bd.gated("en.last_conv", c.c, 64, c.k1_t, c.k1_f, false, true);
        for (auto& U : m->en_mod) { for (auto& L : U.enco) L.zone = 2; for (auto& L : U.deco) L.zone = 2; }
    } else {
        // UNet_Encoder of GaGNet.py:368-413: every layer carries its norm (EaBNet's variant drops two of them)
        m->en_plain.push_back(bd.gated("en.unet_list.0", 2 * c.M, c.c, 2, 5, false, true));
        m->en_plain[0].perm_ri = true;
        m->en_plain[0].M = c.M;
        for (int i = 1; i < 4; ++i)
            m->en_plain.push_back(bd.gated("en.unet_list." + std::to_string(i), c.c, c.c, c.k1_t, c.k1_f, false, true));
        m->en_plain.push_back(bd.gated("en.unet_list.4", c.c, 64, c.k1_t, c.k1_f, false, true));
    }
    const int Fq = c.n_freq, ci = 2 * Fq + c.d_feat;
    auto tcm_groups = [&](const std::string& pfx, std::vector<TcmLayer>& dst) -> int {
        for (int gi = 0; gi < c.p; ++gi)
            for (int i = 0; i < g.n_dilas; ++i) {
                TcmLayer t;
                t.single = true;
                t.perm = false;
                t.dilation = g.dilas[i];
                if (t.dilation < 1) return fail("GaGNet: dilation rates must be positive");
                const std::string p = pfx + "." + std::to_string(gi) + ".tcns." + std::to_string(i);
                t.w_in = bd.add(p + ".in_conv.weight", {c.cd1, c.d_feat, 1}, EAB_P_CONV_W, c.d_feat);
                t.na_left.C = t.na_out.C = c.cd1;
                t.na_left.alpha = bd.add(p + ".d_conv.0.weight", {c.cd1}, EAB_P_PRELU, c.cd1);
                bd.norm(p + ".d_conv.1", c.cd1, t.na_left);
                t.w_left = bd.add(p + ".d_conv.3.weight", {c.cd1, c.cd1, c.kd1}, EAB_P_CONV_W, c.cd1 * c.kd1);
                t.na_out.alpha = bd.add(p + ".out_conv.0.weight", {c.cd1}, EAB_P_PRELU, c.cd1);
                bd.norm(p + ".out_conv.1", c.cd1, t.na_out);
                t.w_out = bd.add(p + ".out_conv.2.weight", {c.d_feat, c.cd1, 1}, EAB_P_CONV_W, c.cd1);
                const int span = (c.kd1 - 1) * t.dilation;
                if (!c.is_causal && (span & 1)) return fail("non-causal TCM needs an even (kd1-1)*dilation");
                const int pad_left = c.is_causal ? span : span / 2;
                for (int k = 0; k < c.kd1; ++k) t.dt[k] = pad_left - k * t.dilation;
                dst.push_back(t);
            }
        return 0;
    };
    auto in_convs = [&](const std::string& p, GagIn& in) {
        in.w_main = bd.add(p + ".in_conv_main.weight", {c.d_feat, ci, 1}, EAB_P_CONV_W, ci);
        in.b_main = bd.add(p + ".in_conv_main.bias", {c.d_feat}, EAB_P_CONV_B, ci);
        in.w_gate = bd.add(p + ".in_conv_gate.0.weight", {c.d_feat, ci, 1}, EAB_P_CONV_W, ci);
        in.b_gate = bd.add(p + ".in_conv_gate.0.bias", {c.d_feat}, EAB_P_CONV_B, ci);
    };
    auto lin = [&](const std::string& p, GagLin& l) {
        l.w = bd.add(p + ".weight", {Fq, c.d_feat, 1}, EAB_P_CONV_W, c.d_feat);
        l.b = bd.add(p + ".bias", {Fq}, EAB_P_CONV_B, c.d_feat);
    };
    m->gags.resize(c.q);
    for (int i = 0; i < c.q; ++i) {
        GagModule& G = m->gags[i];
        std::string p = "gags." + std::to_string(i) + ".glance_block";
        in_convs(p, G.in_g);
        EAB_TRY(tcm_groups(p + ".tcn_g", G.tcn_g));
        lin(p + ".linear_g.0", G.lin_g);
        p = "gags." + std::to_string(i) + ".gaze_block";
        in_convs(p, G.in_z);
        if (g.is_squeezed) {
            EAB_TRY(tcm_groups(p + ".tcm_ri", G.tcm_r));
        } else {
            EAB_TRY(tcm_groups(p + ".tcm_r", G.tcm_r));
            EAB_TRY(tcm_groups(p + ".tcm_i", G.tcm_i));
        }
        lin(p + ".linear_r", G.lin_r);
        lin(p + ".linear_i", G.lin_i);
    }
    return 0;
}

// ================================================================================================ pack
struct Packer {
    eab_model* m;
    std::vector<float> blob;
    size_t alloc(size_t n) {
        size_t off = (blob.size() + 63) / 64 * 64;
        blob.resize(off + n, 0.f);
        return off;
    }
    const std::vector<float>& P(int i) const { return m->params[i].host; }

    void normact(NormAct& na) {
        const int C = na.C;
        if (na.has_norm) {
            na.off_scale = alloc(C);
            na.off_shift = alloc(C);
            for (int c = 0; c < C; ++c) {
                if (m->cfg.norm_type == 1) {          // BatchNorm eval: fold running statistics
                    const double s = (double)P(na.gamma)[c] / sqrt((double)P(na.var)[c] + 1e-5);
                    blob[na.off_scale + c] = (float)s;
                    blob[na.off_shift + c] = (float)((double)P(na.beta)[c] - (double)P(na.mean)[c] * s);
                } else {
                    blob[na.off_scale + c] = P(na.gamma)[c];
                    blob[na.off_shift + c] = P(na.beta)[c];
                }
            }
        }
        na.off_alpha = alloc(C);
        na.alpha01 = true;
        for (int c = 0; c < C; ++c) {
            const float al = P(na.alpha)[c];
            blob[na.off_alpha + c] = al;
            if (!(al >= 0.f && al <= 1.f)) na.alpha01 = false;
        }
    }

    void conv(ConvLayer& L) {
        const int cout_t = L.gated ? 2 * L.cout : L.cout;
        L.N = L.gated ? 2 * ceil64(L.cout) : pad_n(L.cout);
        L.gate_off = L.gated ? ceil64(L.cout) : 0;
        auto col_of = [&](int n_orig) { return (L.gated && n_orig >= L.cout) ? L.gate_off + (n_orig - L.cout) : n_orig; };
        auto cin_of = [&](int cin_ref) {          // reference input channel -> memory channel
            if (!L.perm_ri) return cin_ref;
            const int ri = cin_ref / L.M, mic = cin_ref - ri * L.M;
            return mic * 2 + ri;
        };
        const std::vector<float>& W = P(L.w);
        L.nvar = L.deconv ? 2 : 1;
        for (int v = 0; v < L.nvar; ++v) {
            int nt = 0;
            std::vector<int> kj, kk;
            for (int j = 0; j < L.kt; ++j)
                for (int k = 0; k < L.kf; ++k) {
                    if (L.deconv) {
                        if ((k & 1) != v) continue;
                        L.dt[v][nt] = j;                  // transposed conv + chomp: tap j reads frame t - j
                        L.df[v][nt] = -(k / 2);           // fo = 2e + v, fi = e - (k - v)/2
                    } else {
                        L.dt[v][nt] = L.kt - 1 - j;       // top padding kt-1: tap j reads frame t - (kt-1-j)
                        L.df[v][nt] = k;                  // fi = 2 fo + k
                    }
                    kj.push_back(j); kk.push_back(k);
                    ++nt;
                }
            L.ntaps[v] = nt;
            L.off_w[v] = alloc((size_t)(nt > 0 ? nt : 1) * L.cin * L.N);
            for (int tp = 0; tp < nt; ++tp)
                for (int ci = 0; ci < L.cin; ++ci)
                    for (int n = 0; n < cout_t; ++n) {
                        const size_t src = L.deconv
                            ? (((size_t)ci * cout_t + n) * L.kt + kj[tp]) * L.kf + kk[tp]
                            : (((size_t)n * L.cin + ci) * L.kt + kj[tp]) * L.kf + kk[tp];
                        blob[L.off_w[v] + ((size_t)tp * L.cin + cin_of(ci)) * L.N + col_of(n)] = W[src];
                    }
        }
        L.off_b = alloc(L.N);
        for (int n = 0; n < cout_t; ++n) blob[L.off_b + col_of(n)] = P(L.b)[n];
        pack_umma(L, W);
        normact(L.na);
    }

    // tcgen05 weight images: [variant][tap][slab][N rows][32 k], 128B-swizzled, hi = tf32(w), lo = tf32(w - hi)
    void pack_umma(ConvLayer& L, const std::vector<float>& W) {
        const int co = L.cout;
        L.umma_ok = false;
        if (co != 16 && co != 32 && co != 64 && co != 128) return;
        if (L.gated && co > 128) return;
        L.wide = L.perm_ri;
        if (!L.wide && (L.cin % 64 != 0)) return;
        if (L.wide && L.deconv) return;
        const int cout_t = L.gated ? 2 * co : co;
        L.u_N = cout_t;
        L.u_gate_off = L.gated ? co : 0;
        L.u_kwidth = L.wide ? L.kf * L.cin : 0;
        L.u_nslab = L.wide ? (L.u_kwidth + 63) / 64 : L.cin / 64;
        auto cin_mem = [&](int cin_ref) {
            if (!L.perm_ri) return cin_ref;
            const int ri = cin_ref / L.M, mic = cin_ref - ri * L.M;
            return mic * 2 + ri;
        };
        for (int v = 0; v < L.nvar; ++v) {
            // taps: wide mode has one tap per temporal tap (window over kf positions); otherwise as the generic path
            std::vector<int> tj, tk;
            int nt = 0;
            if (L.wide) {
                for (int j = 0; j < L.kt; ++j) { L.u_dt[v][nt] = L.kt - 1 - j; L.u_df[v][nt] = 0; tj.push_back(j); tk.push_back(0); ++nt; }
            } else {
                nt = L.ntaps[v];
                int q = 0;
                for (int j = 0; j < L.kt; ++j)
                    for (int k = 0; k < L.kf; ++k) {
                        if (L.deconv && (k & 1) != v) continue;
                        L.u_dt[v][q] = L.dt[v][q]; L.u_df[v][q] = L.df[v][q];
                        tj.push_back(j); tk.push_back(k);
                        ++q;
                    }
            }
            L.u_ntaps[v] = nt;
            if (nt == 0) return;
            const size_t img = (size_t)nt * L.u_nslab * cout_t * 32;      // floats: N rows x 128 B per (tap, slab)
            L.off_whi[v] = alloc(img);
            L.off_wlo[v] = alloc(img);
            // dense [tap][kk][n] first (kk = K index inside the tap in MEMORY order), then swizzle per slab
            const int kper = L.u_nslab * 64;
            std::vector<float> dense((size_t)nt * kper * cout_t, 0.f);
            for (int tp = 0; tp < nt; ++tp)
                for (int ci = 0; ci < L.cin; ++ci)
                    for (int n = 0; n < cout_t; ++n) {
                        if (L.wide) {
                            for (int k = 0; k < L.kf; ++k) {
                                const size_t src = (((size_t)n * L.cin + ci) * L.kt + tj[tp]) * L.kf + k;
                                dense[((size_t)tp * kper + (size_t)k * L.cin + cin_mem(ci)) * cout_t + n] = W[src];
                            }
                        } else {
                            const size_t src = L.deconv ? (((size_t)ci * cout_t + n) * L.kt + tj[tp]) * L.kf + tk[tp]
                                                        : (((size_t)n * L.cin + ci) * L.kt + tj[tp]) * L.kf + tk[tp];
                            dense[((size_t)tp * kper + ci) * cout_t + n] = W[src];
                        }
                    }
            __half* img_hi = reinterpret_cast<__half*>(blob.data() + L.off_whi[v]);
            __half* img_lo = reinterpret_cast<__half*>(blob.data() + L.off_wlo[v]);
            for (int tp = 0; tp < nt; ++tp)
                for (int sl = 0; sl < L.u_nslab; ++sl) {
                    const size_t base = ((size_t)tp * L.u_nslab + sl) * cout_t * 64;       // in halves
                    for (int n = 0; n < cout_t; ++n)
                        for (int k = 0; k < 64; ++k) {
                            const float w = dense[((size_t)tp * kper + sl * 64 + k) * cout_t + n];
                            const __half hi = __float2half_rn(w);
                            img_hi[base + sw128_index_h(n, k)] = hi;
                            img_lo[base + sw128_index_h(n, k)] = __float2half_rn(w - __half2float(hi));
                        }
                }
        }
        L.off_ub = alloc(cout_t);
        for (int n = 0; n < cout_t; ++n) blob[L.off_ub + n] = P(L.b)[n];
        L.umma_ok = true;
        if (L.wide && !L.deconv && 2 * L.cin <= 64 && L.kt * ((L.kf + 1) / 2) <= kMaxTaps) {
            const int ns = (L.kf + 1) / 2;
            L.p_ntaps = L.kt * ns;
            const size_t img = (size_t)L.p_ntaps * cout_t * 32;            // floats: one 64-wide slab per tap
            L.off_phi = alloc(img);
            L.off_plo = alloc(img);
            __half* phi = reinterpret_cast<__half*>(blob.data() + L.off_phi);
            __half* plo = reinterpret_cast<__half*>(blob.data() + L.off_plo);
            int q = 0;
            for (int j = 0; j < L.kt; ++j)
                for (int sft = 0; sft < ns; ++sft, ++q) {
                    L.p_dt[q] = L.kt - 1 - j;
                    L.p_ds[q] = sft;
                    const size_t base = (size_t)q * cout_t * 64;
                    for (int n = 0; n < cout_t; ++n)
                        for (int k = 0; k < 64; ++k) {
                            const int pos = k / L.cin, cm = k - pos * L.cin;       // position inside the pair, memory channel
                            float w = 0.f;
                            if (pos < 2 && 2 * sft + pos < L.kf) {
                                int ci = cm;                                      // memory channel -> reference channel
                                if (L.perm_ri) { const int mic = cm / 2, ri = cm - 2 * mic; ci = ri * L.M + mic; }
                                w = W[(((size_t)n * L.cin + ci) * L.kt + j) * L.kf + 2 * sft + pos];
                            }
                            const __half hi = __float2half_rn(w);
                            phi[base + sw128_index_h(n, k)] = hi;
                            plo[base + sw128_index_h(n, k)] = __float2half_rn(w - __half2float(hi));
                        }
                }
            L.pair_ok = true;
        }
    }

    void tcm(TcmLayer& t) {
        const eab_config& c = m->cfg;
        const int Fb = m->Fb, cd = c.cd1, df = c.d_feat, kd = c.kd1;
        const int Nin = pad_n(cd);
        // residual-stream channel of reference channel cr: bottleneck order f*64 + cc for cr = cc*Fb + f (EaBNet), or cr itself
        auto mem_ch = [&](int cr) { if (!t.perm) return cr; const int cc = cr / Fb, f = cr - cc * Fb; return f * 64 + cc; };
        // 1x1 squeeze
        t.off_in = alloc((size_t)df * Nin);
        for (int n = 0; n < cd; ++n)
            for (int cr = 0; cr < df; ++cr)
                blob[t.off_in + (size_t)mem_ch(cr) * Nin + n] = P(t.w_in)[(size_t)n * df + cr];
        if (t.single) {
            // one dilated branch, no gate (GaGNet.py:310-315)
            t.off_dil = alloc((size_t)kd * cd * Nin);
            for (int k = 0; k < kd; ++k)
                for (int ci = 0; ci < cd; ++ci)
                    for (int n = 0; n < cd; ++n)
                        blob[t.off_dil + ((size_t)k * cd + ci) * Nin + n] = P(t.w_left)[((size_t)n * cd + ci) * kd + k];
        } else {
            // dilated pair as one gated conv over K = [left-branch channels | right-branch channels]
            const int Nd = 2 * ceil64(cd), goff = ceil64(cd);
            t.off_dil = alloc((size_t)kd * 2 * cd * Nd);
            for (int k = 0; k < kd; ++k)
                for (int ci = 0; ci < cd; ++ci)
                    for (int n = 0; n < cd; ++n) {
                        blob[t.off_dil + ((size_t)k * 2 * cd + ci) * Nd + n] = P(t.w_left)[((size_t)n * cd + ci) * kd + k];
                        blob[t.off_dil + ((size_t)k * 2 * cd + cd + ci) * Nd + goff + n] = P(t.w_right)[((size_t)n * cd + ci) * kd + k];
                    }
        }
        // 1x1 expand
        t.off_out = alloc((size_t)cd * df);
        for (int nr = 0; nr < df; ++nr)
            for (int ci = 0; ci < cd; ++ci) blob[t.off_out + (size_t)ci * df + mem_ch(nr)] = P(t.w_out)[(size_t)nr * cd + ci];
        t.u_in = umma_images(t.off_in, 1, df, Nin, cd, false, 0, nullptr);
        if (t.single) t.u_dil = umma_images(t.off_dil, kd, cd, Nin, cd, false, 0, nullptr);
        else {
            t.u_dil = umma_images(t.off_dil, kd, 2 * cd, 2 * ceil64(cd), 2 * cd, true, ceil64(cd), nullptr);
            if (cd == 64) {
                // the branches separately (the merged form above is block-diagonal: half of its MMAs multiply zeros)
                for (int br = 0; br < 2; ++br) {
                    const size_t off = alloc((size_t)kd * cd * Nin);
                    const std::vector<float>& W = P(br ? t.w_right : t.w_left);
                    for (int k = 0; k < kd; ++k)
                        for (int ci = 0; ci < cd; ++ci)
                            for (int n = 0; n < cd; ++n) blob[off + ((size_t)k * cd + ci) * Nin + n] = W[((size_t)n * cd + ci) * kd + k];
                    (br ? t.u_dr : t.u_dl) = umma_images(off, kd, cd, Nin, cd, false, 0, nullptr);
                }
            }
        }
        t.u_out = umma_images(t.off_out, 1, cd, df, df, false, 0, nullptr);
        normact(t.na_left);
        if (!t.single) normact(t.na_right);
        normact(t.na_out);
    }

    // GaGNet glance / gaze input convs (GaGNet.py:161-165, 190): K = [encoder feature, bottleneck order f*64+c | pre_x row
    // ri*F+f, zero-padded to KP], columns of split sp = value channels 64 sp .. | gate channels 64 sp ..
    void gag_in(GagIn& in) {
        const eab_config& c = m->cfg;
        const int Fb = m->Fb, df = c.d_feat, Fq = c.n_freq, KP = ceil64(2 * Fq), ci = 2 * Fq + df;
        // widest gated split the tensor-core kernel takes: 128 value + 128 gate columns (each split re-reads the whole input)
        const int SW = df % 128 == 0 ? 128 : 64;
        in.K = df + KP;
        in.SW = SW;
        in.nsplit = df / SW;
        for (int sp = 0; sp < in.nsplit; ++sp) {
            in.off_dense[sp] = alloc((size_t)in.K * 2 * SW);
            std::vector<float> bias(2 * SW);
            for (int n = 0; n < 2 * SW; ++n) {
                const bool gate = n >= SW;
                const int co = sp * SW + (n % SW);
                const std::vector<float>& W = P(gate ? in.w_gate : in.w_main);
                bias[n] = P(gate ? in.b_gate : in.b_main)[co];
                for (int cr = 0; cr < df; ++cr) {
                    const int cc = cr / Fb, f = cr - cc * Fb;
                    blob[in.off_dense[sp] + (size_t)(f * 64 + cc) * 2 * SW + n] = W[(size_t)co * ci + cr];
                }
                for (int k = 0; k < 2 * Fq; ++k) blob[in.off_dense[sp] + (size_t)(df + k) * 2 * SW + n] = W[(size_t)co * ci + df + k];
            }
            in.u[sp] = umma_images(in.off_dense[sp], 1, in.K, 2 * SW, 2 * SW, true, SW, bias.data());
        }
    }

    void gag_lin(GagLin& l) {
        const eab_config& c = m->cfg;
        l.off_w = linear(l.w, l.b, c.n_freq, c.d_feat, &l.N, &l.off_b);
        std::vector<float> b0(blob.begin() + l.off_b, blob.begin() + l.off_b + l.N);
        l.u = umma_images(l.off_w, 1, c.d_feat, l.N, c.n_freq, false, 0, b0.data());
    }

    // Build fp16 hi/lo images from a dense [ntaps][K][ldn] fp32 matrix that already sits in the blob at `off`
    // (the layout of the CUDA-core path).  Columns [0, ncols) are used; gated => value|gate halves of `cout` each
    // located at columns [0,cout) and [gate_col, gate_col+cout) of the dense matrix.
    UmmaW umma_images(size_t off, int ntaps, int K, int ldn, int ncols, bool gated, int gate_col, const float* bias) {
        UmmaW u;
        if (K % 64 != 0) return u;
        const int cout = gated ? ncols / 2 : ncols;
        int padded = cout <= 16 ? 16 : cout <= 32 ? 32 : cout <= 64 ? 64 : (cout + 127) / 128 * 128;
        if (gated && padded != cout) return u;
        if (gated && cout > 128) return u;
        u.ntaps = ntaps;
        u.nslab = K / 64;
        u.gate_off = gated ? cout : 0;
        u.nsplit = (!gated && padded > 128) ? padded / 128 : 1;
        if (u.nsplit > 4) return u;
        u.cout = padded / u.nsplit;
        u.ncol = gated ? 2 * cout : u.cout;
        u.ld = padded;
        u.has_bias = bias != nullptr;
        for (int sp = 0; sp < u.nsplit; ++sp) {
            const size_t img = (size_t)ntaps * u.nslab * u.ncol * 32;
            u.off_hi[sp] = alloc(img);
            u.off_lo[sp] = alloc(img);
            u.off_bias[sp] = alloc(u.ncol);
            __half* hi = reinterpret_cast<__half*>(blob.data() + u.off_hi[sp]);
            __half* lo = reinterpret_cast<__half*>(blob.data() + u.off_lo[sp]);
            for (int n = 0; n < u.ncol; ++n) {
                // column of the dense matrix feeding image row n
                int col;
                if (gated) col = n < cout ? n : gate_col + (n - cout);
                else col = sp * u.cout + n;
                const bool real = gated ? true : col < ncols;
                blob[u.off_bias[sp] + n] = (bias && real) ? bias[col] : 0.f;
                for (int tp = 0; tp < ntaps; ++tp)
                    for (int sl = 0; sl < u.nslab; ++sl) {
                        const size_t base = ((size_t)tp * u.nslab + sl) * u.ncol * 64;
                        for (int k = 0; k < 64; ++k) {
                            const float w = real ? blob[off + ((size_t)tp * K + sl * 64 + k) * ldn + col] : 0.f;
                            const __half h = __float2half_rn(w);
                            hi[base + sw128_index_h(n, k)] = h;
                            lo[base + sw128_index_h(n, k)] = __float2half_rn(w - __half2float(h));
                        }
                    }
            }
        }
        u.ok = true;
        return u;
    }

    size_t linear(int w, int b, int nout, int nin, int* N, size_t* off_b) {
        *N = pad_n(nout);
        const size_t off = alloc((size_t)nin * *N);
        for (int n = 0; n < nout; ++n)
            for (int k = 0; k < nin; ++k) blob[off + (size_t)k * *N + n] = P(w)[(size_t)n * nin + k];
        *off_b = alloc(*N);
        for (int n = 0; n < nout; ++n) blob[*off_b + n] = P(b)[n];
        return off;
    }

    void head() {
        const eab_config& c = m->cfg;
        if (m->rnn[0][0] >= 0) {
            const int H = 64;
            for (int r = 0; r < 2; ++r) {
                const int E = r ? H : c.embed_dim;
                m->off_rnn[r][0] = alloc((size_t)E * H * 4);
                m->off_rnn[r][1] = alloc((size_t)H * H * 4);
                m->off_rnn[r][2] = alloc((size_t)H * 4);
                for (int g = 0; g < 4; ++g)
                    for (int j = 0; j < H; ++j) {
                        for (int k = 0; k < E; ++k)
                            blob[m->off_rnn[r][0] + ((size_t)k * H + j) * 4 + g] = P(m->rnn[r][0])[(size_t)(g * H + j) * E + k];
                        for (int k = 0; k < H; ++k)
                            blob[m->off_rnn[r][1] + ((size_t)k * H + j) * 4 + g] = P(m->rnn[r][1])[(size_t)(g * H + j) * H + k];
                        blob[m->off_rnn[r][2] + (size_t)j * 4 + g] = P(m->rnn[r][2])[g * H + j] + P(m->rnn[r][3])[g * H + j];
                    }
            }
            // tcgen05 LSTM: image row n = half*128 + quarter*32 + gate*8 + j  <->  torch row gate*64 + (quarter*16 + half*8 + j);
            // K slab 0 = W_ih (input channels), slab 1 = W_hh
            m->rnn_umma_ok = c.embed_dim == 64;
            if (m->rnn_umma_ok) {
                for (int r = 0; r < 2; ++r) {
                    m->off_rnn_img[r] = alloc((size_t)4 * 256 * 32);
                    m->off_rnn_ubias[r] = alloc(256);
                    __half* img = reinterpret_cast<__half*>(blob.data() + m->off_rnn_img[r]);
                    for (int n = 0; n < 256; ++n) {
                        const int hf_ = n >> 7, qtr_ = (n >> 5) & 3, g = (n >> 3) & 3, jj = n & 7;
                        const int row = g * H + qtr_ * 16 + hf_ * 8 + jj;
                        blob[m->off_rnn_ubias[r] + n] = P(m->rnn[r][2])[row] + P(m->rnn[r][3])[row];
                        for (int slab = 0; slab < 2; ++slab)
                            for (int k = 0; k < 64; ++k) {
                                const float w = P(m->rnn[r][slab])[(size_t)row * 64 + k];
                                const __half hi = __float2half_rn(w);
                                const __half lo = __float2half_rn(w - __half2float(hi));
                                img[((size_t)(0 * 2 + slab) * 256) * 64 + sw128_index_h(n, k)] = hi;
                                img[((size_t)(1 * 2 + slab) * 256) * 64 + sw128_index_h(n, k)] = lo;
                            }
                    }
                }
            }
            m->off_dnn_w[0] = linear(m->dnn_w[0], m->dnn_b[0], H, H, &m->dnn_N[0], &m->off_dnn_b[0]);
            m->off_dnn_w[1] = linear(m->dnn_w[1], m->dnn_b[1], 2 * c.M, H, &m->dnn_N[1], &m->off_dnn_b[1]);
            {
                std::vector<float> b0(blob.begin() + m->off_dnn_b[0], blob.begin() + m->off_dnn_b[0] + m->dnn_N[0]);
                std::vector<float> b1(blob.begin() + m->off_dnn_b[1], blob.begin() + m->off_dnn_b[1] + m->dnn_N[1]);
                m->u_dnn[0] = umma_images(m->off_dnn_w[0], 1, H, m->dnn_N[0], H, false, 0, b0.data());
                m->u_dnn[1] = umma_images(m->off_dnn_w[1], 1, H, m->dnn_N[1], 2 * c.M, false, 0, b1.data());
            }
            m->off_ln_g = alloc(c.embed_dim);
            m->off_ln_b = alloc(c.embed_dim);
            for (int i = 0; i < c.embed_dim; ++i) {
                blob[m->off_ln_g + i] = P(m->ln_g)[i];
                blob[m->off_ln_b + i] = P(m->ln_b)[i];
            }
        } else {
            const int n = c.topo_type == 0 ? 2 * c.M : 2;
            m->off_cnn_w = linear(m->cnn_w, m->cnn_b, n, c.embed_dim, &m->cnn_N, &m->off_cnn_b);
            {
                std::vector<float> b0(blob.begin() + m->off_cnn_b, blob.begin() + m->off_cnn_b + m->cnn_N);
                m->u_cnn = umma_images(m->off_cnn_w, 1, c.embed_dim, m->cnn_N, n, false, 0, b0.data());
            }
        }
    }
};

int commit(eab_model* m, cudaStream_t st) {
    for (const Param& p : m->params)
        if (!p.set && p.kind != EAB_P_BN_COUNT) return fail("parameter not set: " + p.name);
    Packer pk{m};
    for (auto& U : m->en_mod) { pk.conv(U.in_conv); for (auto& L : U.enco) pk.conv(L); for (auto& L : U.deco) pk.conv(L); }
    for (auto& U : m->de_mod) { pk.conv(U.in_conv); for (auto& L : U.enco) pk.conv(L); for (auto& L : U.deco) pk.conv(L); }
    if (m->cfg.is_u2) { pk.conv(m->en_last); if (m->kind == 0) pk.conv(m->de_last); }
    for (auto& L : m->en_plain) pk.conv(L);
    for (auto& L : m->de_plain) pk.conv(L);
    for (auto& t : m->tcms) pk.tcm(t);
    if (m->kind == 0) pk.head();
    for (auto& G : m->gags) {
        pk.gag_in(G.in_g);
        pk.gag_in(G.in_z);
        for (auto& t : G.tcn_g) pk.tcm(t);
        for (auto& t : G.tcm_r) pk.tcm(t);
        for (auto& t : G.tcm_i) pk.tcm(t);
        pk.gag_lin(G.lin_g);
        pk.gag_lin(G.lin_r);
        pk.gag_lin(G.lin_i);
    }
    if (m->blob && m->blob_floats < pk.blob.size()) { cudaFree(m->blob); m->blob = nullptr; }
    if (!m->blob) {
        EAB_CUDA(cudaMalloc(&m->blob, pk.blob.size() * sizeof(float)));
        m->blob_floats = pk.blob.size();
    }
    EAB_CUDA(cudaMemcpyAsync(m->blob, pk.blob.data(), pk.blob.size() * sizeof(float), cudaMemcpyHostToDevice, st));
    EAB_CUDA(cudaStreamSynchronize(st));        // the staging vector dies with this scope
    m->dirty = false;
    ++m->param_version;
    return 0;
}

// ================================================================================================ run
struct Ctx {
    eab_model* m;
    bool dry;                 // size planning only: no launches, no dereference
    char* base;
    size_t stats_off = 0, stats_cap = 0;      // [0, stats_cap): zeroed once per forward
    size_t act_off = 0;
    int B, T;
    cudaStream_t st;

    // streaming (eab_stream_step): T == 1, every activation is a persistent ring of `last_RT` frames in the caller's
    // state blob (same allocation order every step => same addresses), nothing is reused, tensor-core kernels are off
    bool streaming = false;
    const int* step = nullptr;
    const int* start = nullptr;   // [streams] first absolute frame of each stream (eab_stream_reset_one)
    int gag_in_RT = 1; long long gag_in_slot = 0;      // streaming GaGNet: ring geometry of the `inpt` frame source
    int next_RT = 0;          // ring size of the next allocation (0 = the default of 2: current + previous frame)
    int last_RT = 0;          // ring size of the last allocation (0 offline)
    bool tensor_ok() const { return m->opt_umma && !streaming; }
    bool stream_umma() const { return m->opt_umma && streaming && m->opt_stream_umma; }      // conv_umma with ring addressing
    std::vector<TcmStreamDesc>* tcm_desc = nullptr;     // planning pass of eab_stream_reset: receives the descriptors
    std::vector<std::pair<size_t, size_t>>* per_stream = nullptr;      // planning pass: carried per-stream state to zero on a restart
    const TcmStreamDesc* tcm_desc_dev = nullptr;        // step: the table inside the state blob

    size_t act_peak = 0;
    float* alloc_act(size_t floats) {
        last_RT = streaming ? (next_RT ? next_RT : 2) : 0;
        next_RT = 0;
        if (streaming) floats *= last_RT;
        const size_t bytes = (floats * sizeof(float) + 255) / 256 * 256;
        float* p = reinterpret_cast<float*>(base + act_off);
        act_off += bytes;
        if (act_off > act_peak) act_peak = act_off;
        return p;
    }
    // scoped reuse: everything allocated after mark() is dead at release() (single stream => later kernels that
    // overwrite the region are ordered after the kernels that read it)
    size_t mark() const { return act_off; }
    void release(size_t m) { if (!streaming) act_off = m; }
    double* alloc_stats(int C) {
        const size_t bytes = ((size_t)B * C * 2 * sizeof(double) + 255) / 256 * 256;
        double* p = reinterpret_cast<double*>(base + stats_off);
        stats_off += bytes;
        return p;
    }
    const float* W(size_t off) const { return m->blob + off; }
};

// the Xform a consumer uses for the output of a conv followed by NormAct (2-D convention: norm -> PReLU)
Xform xf_after(Ctx& cx, const NormAct& na, double* stats, int count, int prelu_pos) {
    Xform x = xform_identity();
    if (na.has_norm) {
        if (cx.m->cfg.norm_type == 0) {
            x.affine = 1; x.stats = stats; x.inv_count = 1.f / (float)count;
            if (cx.m->opt_norm_log && !cx.dry && stats) cx.m->norm_log.push_back({na.gamma, stats, na.C, count, cx.B});
        }
        else x.affine = 2;
        x.scale = cx.W(na.off_scale);
        x.shift = cx.W(na.off_shift);
    }
    x.alpha = cx.W(na.off_alpha);
    x.alpha01 = na.alpha01 ? 1 : 0;
    x.prelu = prelu_pos;
    return x;
}

// Re-express a per-tap gather launch in the padded-pitch row space of conv_raw / the staged pair; false if the shape does not qualify.
bool to_plane_args(const UmmaConvArgs& u, PlaneConvArgs* p, int force_P = 0) {
    if (u.wide) return false;
    memset(p, 0, sizeof(*p));
    p->nsrc = u.nsrc;
    for (int i = 0; i < u.nsrc; ++i) p->src[i] = u.src[i];
    p->B = u.B; p->T = u.T; p->Fin = u.Fin; p->E = u.E;
    int min_df = 0, max_df = 0;
    for (int i = 0; i < u.ntaps; ++i) { min_df = std::min(min_df, u.df[i]); max_df = std::max(max_df, u.df[i]); }
    if (u.in_stride == 2) {
        if (min_df < 0) return false;
        p->nplanes = 2;
        p->plane_cols[0] = (u.Fin + 1) / 2; p->plane_cols[1] = u.Fin / 2;
        p->col_stride = 2; p->col_off[0] = 0; p->col_off[1] = 1;
        p->P = std::max(std::max(u.E + max_df / 2, p->plane_cols[0]), force_P);
    } else if (u.in_stride == 1) {
        if (max_df > 0) return false;
        p->nplanes = 1;
        p->plane_cols[0] = u.Fin; p->plane_cols[1] = 0;
        p->col_stride = 1; p->col_off[0] = 0; p->col_off[1] = 0;
        p->P = std::max(std::max(u.E, u.Fin - min_df), force_P);      // the pad columns [Fin, P) absorb the negative column offsets
    } else {
        return false;
    }
    p->ntaps = u.ntaps;
    int back = 0, fwd = 0;
    for (int i = 0; i < u.ntaps; ++i) {
        p->tap_plane[i] = u.in_stride == 2 ? (u.df[i] & 1) : 0;
        p->tap_shift[i] = -u.dt[i] * p->P + (u.in_stride == 2 ? (u.df[i] >> 1) : u.df[i]);
        back = std::max(back, -p->tap_shift[i]);
        fwd = std::max(fwd, p->tap_shift[i]);
    }
    p->back = back; p->fwd = fwd;
    p->out_stride = u.out_stride; p->out_off = u.out_off; p->Fout = u.Fout;
    p->nslab = u.nslab; p->ncoef = u.ncoef; p->npass = u.npass;
    p->Whi = u.Whi; p->Wlo = u.Wlo; p->bias = u.bias;
    p->Cout = u.Cout; p->N = u.N; p->gate_off = u.gate_off; p->relu = u.relu; p->algo_frac = u.algo_frac;
    p->out = u.out; p->out_ld = u.out_ld; p->out_coff = u.out_coff; p->resid = u.resid;
    p->nstats = u.nstats;
    p->stats_ld = u.stats_ld; p->stats_coff = u.stats_coff;
    for (int i = 0; i < 2; ++i) { p->stats[i] = u.stats[i]; p->stat_alpha[i] = u.stat_alpha[i]; }
    p->tiles_per_b = (int)(((long long)u.T * p->P + 127) / 128);
    p->nbuf = 1;
    return plane_conv_supported(*p);
}

// streaming: the variants of one layer on the tcgen05 gather kernel - all streams share one row space (B = 1, "frames" = streams)
int run_umma_stream(Ctx& cx, UmmaConvArgs* us, int n, int out_RT, int resid_RT) {
    if (cx.dry) return 0;
    for (int i = 0; i < n; ++i) {
        UmmaConvArgs& u = us[i];
        for (int k = 0; k < u.nsrc; ++k)
            if (u.src[k].x2) return fail("internal: lazy residual sum in a streaming step");
        u.B = 1; u.T = cx.B;
        u.tiles_per_b = (int)(((long long)cx.B * u.E + 127) / 128);
        u.step = cx.step; u.start = cx.start; u.out_RT = out_RT; u.resid_RT = resid_RT;
        u.nstats = 0;
        EAB_TRY(launch_conv_umma(u, cx.st));
    }
    return 0;
}

int launch_tensor_conv(eab_model* m, const UmmaConvArgs& u, cudaStream_t st) {
    (void)m;
    return launch_conv_umma(u, st);      // per-tap gather ring: what neither conv_raw nor the staged pair takes (K = 578, `staged` = 0)
}

// Launch the 1-4 tensor-core variants of one layer (output parities of a transposed conv, column splits of a wide
// 1x1) that read the same inputs.  Preferred path: stage the normalised fp16 planes ONCE (stage_kernel) and run the
// TMA-fed GEMM per variant; otherwise the fused-producer kernels.  Also runs in planning mode (allocations only).
// Can the 1-4 variants of a layer run as one stage launch + conv_tma launches?  Fills their plane arguments.
bool plan_planes(const UmmaConvArgs* us, int n, PlaneConvArgs* p, PlaneConvArgs* ps_out) {
    if (us[0].wide || n > 4) return false;
    bool ok = true;
    int P = 0;
    for (int i = 0; i < n; ++i) { ok = ok && to_plane_args(us[i], &p[i]); if (ok) P = std::max(P, p[i].P); }
    if (ok)
        for (int i = 0; i < n; ++i)
            if (p[i].P != P) ok = ok && to_plane_args(us[i], &p[i], P);
    if (ok)
        for (int i = 1; i < n; ++i)
            ok = ok && p[i].nplanes == p[0].nplanes && p[i].plane_cols[0] == p[0].plane_cols[0] &&
                 p[i].plane_cols[1] == p[0].plane_cols[1] && p[i].nslab == p[0].nslab && p[i].npass == p[0].npass &&
                 p[i].tiles_per_b == p[0].tiles_per_b;
    if (!ok) return false;
    PlaneConvArgs ps = p[0];
    for (int i = 1; i < n; ++i) { ps.back = std::max(ps.back, p[i].back); ps.fwd = std::max(ps.fwd, p[i].fwd); }
    *ps_out = ps;
    return true;
}

int run_tensor_convs(Ctx& cx, UmmaConvArgs* us, int n) {
    eab_model* m = cx.m;
    PlaneConvArgs p[4];
    PlaneConvArgs ps;
    const bool planes_ok = plan_planes(us, n, p, &ps);
    if (planes_ok && m->opt_raw && raw_conv_supported(p, n)) {
        {
            // one launch per layer: raw tiles -> norm + PReLU -> fp16 operand in shared memory -> GEMM (both parities)
            if (cx.dry) return 0;
            unsigned long long* dbg = nullptr;
            if (m->umma_launch_idx++ == m->opt_dbg_launch && m->dbg_buf) dbg = m->dbg_buf;
            return launch_conv_raw(p, n, cx.st, dbg, m->opt_raw_grid);
        }
    }
    bool staged_ok = planes_ok && m->opt_staged && staged_conv_supported(ps);
    for (int i = 0; staged_ok && i < n; ++i) staged_ok = staged_conv_supported(p[i]);
    if (staged_ok) {
        int front = 0;
        const int rows = staged_rows(ps, &front);
        const int nimg = ps.nplanes * ps.nslab * (ps.npass == 3 ? 2 : 1);
        ps.np_rows = rows; ps.np_front = front;
        const size_t scratch = cx.mark();           // the staged planes die with this layer
        for (int k = 0; k < nimg; ++k) ps.np[k] = cx.alloc_act((size_t)cx.B * rows * 32);     // 128 B per row
        cx.release(scratch);
        if (cx.dry) return 0;
        EAB_TRY(launch_stage(ps, cx.st));
        for (int i = 0; i < n; ++i) {
            if (m->umma_launch_idx++ == m->opt_dbg_launch && m->dbg_buf) p[i].dbg = m->dbg_buf;
            p[i].np_rows = rows; p[i].np_front = front;
            for (int k = 0; k < nimg; ++k) p[i].np[k] = ps.np[k];
            p[i].algo_in_share = 1.f / (float)n;
            EAB_TRY(launch_conv_staged(p[i], cx.st));
        }
        return 0;
    }
    for (int i = 0; i < n; ++i)
        for (int k = 0; k < us[i].nsrc; ++k)
            if (us[i].src[k].x2) return fail("internal: lazy residual sum reached a kernel that cannot read it");
    if (cx.dry) return 0;
    for (int i = 0; i < n; ++i) EAB_TRY(launch_tensor_conv(m, us[i], cx.st));
    return 0;
}

inline int zone_passes(const eab_model* m, int zone) {
    return zone == 3 ? m->opt_first_passes : zone == 0 ? m->opt_enc_passes : zone == 1 ? m->opt_dec_passes : m->opt_inner_passes;
}

// one 2-D layer: conv/deconv (+gate) -> raw output + statistics; returns the Act a consumer should read
int materialize(Ctx& cx, Act* a);
int run_conv2d(Ctx& cx, const ConvLayer& L, const Act* srcs_in, int nsrc, Act* out, float* prealloc = nullptr) {
    Act srcs[2];
    for (int i = 0; i < nsrc; ++i) srcs[i] = srcs_in[i];
    if (!(cx.tensor_ok() && L.umma_ok && cx.m->opt_staged && !L.wide))
        for (int i = 0; i < nsrc; ++i) EAB_TRY(materialize(cx, &srcs[i]));       // only conv_raw / the staged pair read lazy sums
    const int Fin = srcs[0].F;
    int cin = 0;
    for (int i = 0; i < nsrc; ++i) {
        if (srcs[i].F != Fin) return fail("skip connection width mismatch (the reference's torch.cat would raise too)");
        cin += srcs[i].C;
    }
    if (cin != L.cin) return fail("internal: channel mismatch in conv layer");
    const int Fout = L.deconv ? deconv_out_f(Fin, L.kf) : conv_out_f(Fin, L.kf);
    if (Fout < 1) return fail("frequency axis too short for this layer");
    out->F = Fout;
    out->C = L.cout;
    out->data2 = nullptr;
    const size_t out_elems = (size_t)cx.B * cx.T * Fout * L.cout;
    const bool in_stats = L.na.has_norm && cx.m->cfg.norm_type == 0;
    double* stats = in_stats ? cx.alloc_stats(L.cout) : nullptr;
    out->xf = xf_after(cx, L.na, stats, cx.T * Fout, 2);
    auto allocate_out = [&]() {
        out->data = prealloc ? prealloc : cx.alloc_act(out_elems);
        out->RT = prealloc ? 0 : cx.last_RT;
    };
    if (cx.tensor_ok() && L.umma_ok && L.wide && cx.m->opt_staged && nsrc == 1 && !L.deconv &&
        srcs[0].xf.affine == 0 && srcs[0].xf.prelu == 0 && !srcs[0].data2 && (srcs[0].C * 2) % 2 == 0) {
        // first layer (2M input channels) on the staged path: a plane row is a frequency PAIR (or the whole kf x C tap window)
        allocate_out();
        PlaneConvArgs p;
        memset(&p, 0, sizeof(p));
        p.nsrc = 1;
        set_src(p.src[0], srcs[0]);
        const bool pair = L.pair_ok;
        p.B = cx.B; p.T = cx.T; p.Fin = Fin; p.E = Fout;
        p.nplanes = 1; p.col_stride = 2; p.col_off[0] = 0;
        int back = 0, fwd = 0;
        if (pair) {
            // rows = frequency PAIRS (2 x cin values each): 1.8x the input bytes per pass-plane instead of 3.6x for window rows
            p.P = (Fin + 1) / 2;
            p.plane_cols[0] = p.P;
            p.ntaps = L.p_ntaps;
            for (int i = 0; i < p.ntaps; ++i) {
                p.tap_plane[i] = 0; p.tap_shift[i] = -L.p_dt[i] * p.P + L.p_ds[i];
                back = std::max(back, -p.tap_shift[i]); fwd = std::max(fwd, p.tap_shift[i]);
            }
            p.nslab = 1;
            p.Whi = cx.W(L.off_phi); p.Wlo = cx.W(L.off_plo);
            p.algo_frac = (float)(L.kf * cin) / (float)(((L.kf + 1) / 2) * 64);
            p.wide_k = 2 * cin;
        } else {
            p.P = Fout;
            p.plane_cols[0] = Fout;
            p.ntaps = L.u_ntaps[0];
            for (int i = 0; i < p.ntaps; ++i) { p.tap_plane[i] = 0; p.tap_shift[i] = -L.u_dt[0][i] * p.P; back = std::max(back, -p.tap_shift[i]); }
            p.nslab = L.u_nslab;
            p.Whi = cx.W(L.off_whi[0]); p.Wlo = cx.W(L.off_wlo[0]);
            p.algo_frac = (float)L.u_kwidth / (float)(L.u_nslab * 64);
            p.wide_k = L.u_kwidth;
        }
        p.back = back; p.fwd = fwd;
        p.out_stride = 1; p.out_off = 0; p.Fout = Fout;
        p.ncoef = cin; p.npass = zone_passes(cx.m, L.zone);
        p.bias = cx.W(L.off_ub);
        p.Cout = L.cout; p.N = L.u_N; p.gate_off = L.u_gate_off;
        p.out = out->data; p.out_ld = L.cout; p.out_coff = 0;
        if (stats) { p.nstats = 1; p.stats[0] = stats; }
        p.tiles_per_b = (int)(((long long)cx.T * p.P + 127) / 128);
        const bool geom_ok = (srcs[0].C * p.col_stride) % 2 == 0 && (Fin * srcs[0].C) % 2 == 0 &&      // 8-byte aligned windows
                             (Fout - 1) * 2 + L.kf <= Fin && p.P >= 1 &&
                             ((long long)cx.T * p.P + p.back + 4 * 128 + 2ll * p.P) * p.P < (1ll << 31);
        if (geom_ok && staged_conv_fits(p)) {
            int front = 0;
            const int rows = staged_rows(p, &front);
            const int nimg = p.nslab * (p.npass == 3 ? 2 : 1);
            p.np_rows = rows; p.np_front = front;
            const size_t scratch = cx.mark();
            for (int k = 0; k < nimg; ++k) p.np[k] = cx.alloc_act((size_t)cx.B * rows * 32);
            cx.release(scratch);
            if (cx.dry) return 0;
            EAB_TRY(launch_stage(p, cx.st));
            if (cx.m->umma_launch_idx++ == cx.m->opt_dbg_launch && cx.m->dbg_buf) p.dbg = cx.m->dbg_buf;
            return launch_conv_staged(p, cx.st);
        }
    }
    if ((cx.tensor_ok() || cx.stream_umma()) && L.umma_ok) {
        UmmaConvArgs us[4];
        bool all_ok = true;
        int nus = 0;
        for (int v = 0; v < L.nvar; ++v) {
            const int npass = zone_passes(cx.m, L.zone);
            {
                UmmaConvArgs& u = us[nus++];
                memset(&u, 0, sizeof(u));
                u.nsrc = nsrc;
                for (int i = 0; i < nsrc; ++i) set_src(u.src[i], srcs[i]);
                u.B = cx.B; u.T = cx.T; u.Fin = Fin; u.Fout = Fout;
                if (L.deconv) { u.in_stride = 1; u.out_stride = 2; u.out_off = v; u.E = (Fout - v + 1) / 2; }
                else          { u.in_stride = 2; u.out_stride = 1; u.out_off = 0; u.E = Fout; }
                u.ntaps = L.u_ntaps[v];
                for (int i = 0; i < u.ntaps; ++i) { u.dt[i] = L.u_dt[v][i]; u.df[i] = L.u_df[v][i]; }
                u.wide = L.wide; u.kwidth = L.u_kwidth; u.nslab = L.u_nslab; u.ncoef = cin;
                u.npass = npass;
                u.algo_frac = 1.f;
                u.out = out->data; u.out_ld = L.cout;
                u.Whi = cx.W(L.off_whi[v]); u.Wlo = cx.W(L.off_wlo[v]); u.bias = cx.W(L.off_ub);
                u.Cout = L.cout; u.N = L.u_N; u.gate_off = L.u_gate_off; u.out_coff = 0;
                if (stats) { u.nstats = 1; u.stats[0] = stats; }
                u.tiles_per_b = (cx.T * u.E + 127) / 128;
                all_ok = all_ok && umma_conv_supported(u);
            }
        }
        if (all_ok) {
            allocate_out();
            for (int i = 0; i < nus; ++i) us[i].out = out->data;
            if (cx.streaming) return run_umma_stream(cx, us, nus, out->RT, 0);
            return run_tensor_convs(cx, us, nus);
        }
    }
    allocate_out();
    for (int i = 0; i < nsrc; ++i) EAB_TRY(materialize(cx, &srcs[i]));
    if (cx.dry) return 0;
    for (int v = 0; v < L.nvar; ++v) {
        ConvArgs a;
        memset(&a, 0, sizeof(a));
        a.nsrc = nsrc;
        for (int i = 0; i < nsrc; ++i) { a.src[i].x = srcs[i].data; a.src[i].C = srcs[i].C; a.src[i].xf = srcs[i].xf; a.src[i].RT = srcs[i].RT; }
        a.step = cx.step; a.start = cx.start; a.out_RT = out->RT;
        a.B = cx.B; a.T = cx.T; a.Fin = Fin; a.Fout = Fout;
        if (L.deconv) { a.in_stride = 1; a.out_stride = 2; a.out_off = v; a.E = (Fout - v + 1) / 2; }
        else          { a.in_stride = 2; a.out_stride = 1; a.out_off = 0; a.E = Fout; }
        a.ntaps = L.ntaps[v];
        for (int i = 0; i < a.ntaps; ++i) { a.dt[i] = L.dt[v][i]; a.df[i] = L.df[v][i]; }
        if (a.ntaps == 0) return fail("transposed conv with kf == 1 is not supported");
        a.W = cx.W(L.off_w[v]);
        a.bias = cx.W(L.off_b);
        a.Cout = L.cout; a.N = L.N; a.gate_off = L.gate_off;
        a.algo_frac = 1.f;
        a.out = out->data;
        if (stats) { a.nstats = 1; a.stats[0] = stats; }
        EAB_TRY(launch_conv(a, cx.st));
    }
    return 0;
}

int run_combine_into(Ctx& cx, const Act* srcs, int nsrc, Act* out);
int run_combine(Ctx& cx, const Act* srcs, int nsrc, Act* out) {
    out->F = srcs[0].F;
    out->C = srcs[0].C;
    out->xf = xform_identity();
    out->data = cx.alloc_act((size_t)cx.B * cx.T * out->F * out->C);
    out->RT = cx.last_RT;
    return run_combine_into(cx, srcs, nsrc, out);
}

// same, into an Act whose buffer (F, C, data) the caller has already allocated
int run_combine_into(Ctx& cx, const Act* srcs, int nsrc, Act* out) {
    if (cx.dry) return 0;
    CombineArgs a;
    memset(&a, 0, sizeof(a));
    int n = 0;
    for (int i = 0; i < nsrc; ++i) {
        if (srcs[i].F != out->F || srcs[i].C != out->C) return fail("internal: combine shape mismatch");
        if (n + (srcs[i].data2 ? 2 : 1) > 3) return fail("internal: too many addends in combine");
        a.src[n].x = srcs[i].data; a.src[n].C = srcs[i].C; a.src[n].xf = srcs[i].xf; a.src[n].RT = srcs[i].RT; ++n;
        if (srcs[i].data2) { a.src[n].x = srcs[i].data2; a.src[n].C = srcs[i].C; a.src[n].xf = srcs[i].xf2; a.src[n].RT = srcs[i].RT; ++n; }
    }
    a.nsrc = n;
    a.B = cx.B; a.P = cx.T * out->F; a.C = out->C; a.out = out->data;
    a.step = cx.step; a.out_RT = out->RT;
    return launch_combine(a, cx.st);
}

// materialise a lazy residual sum (only needed in front of kernels that cannot read one)
int materialize(Ctx& cx, Act* a) {
    if (!a->data2) return 0;
    Act src = *a;
    Act dst;
    EAB_TRY(run_combine(cx, &src, 1, &dst));
    *a = dst;
    return 0;
}

// En_unet_module.forward (EaBNet.py:372-388)
int run_combine_into(Ctx& cx, const Act* srcs, int nsrc, Act* out);
int run_module(Ctx& cx, const UnetModule& U, const Act* srcs, int nsrc, Act* out) {
    // the module output is allocated first; everything else (in_conv output, inner U-Net maps) is scoped scratch
    const int Fin = srcs[0].F;
    const int Fw = U.in_conv.deconv ? deconv_out_f(Fin, U.in_conv.kf) : conv_out_f(Fin, U.in_conv.kf);
    if (Fw < 1) return fail("frequency axis too short for this layer");
    out->F = Fw;
    out->C = U.in_conv.cout;
    out->xf = xform_identity();
    // lazy mode: the module result x0 + y is never written; its two addends (in_conv output, last inner deconv output)
    // outlive the module instead and every consumer's stage kernel sums them while staging
    const bool lazy = cx.m->opt_lazy && cx.tensor_ok() && cx.m->opt_staged && U.in_conv.umma_ok && !U.deco.empty() &&
                      U.deco.back().umma_ok;
    const size_t nel = (size_t)cx.B * cx.T * Fw * out->C;
    float* buf_x0 = nullptr;
    float* buf_y = nullptr;
    if (lazy) { buf_x0 = cx.alloc_act(nel); buf_y = cx.alloc_act(nel); out->data = nullptr; }
    else { out->data = cx.alloc_act(nel); out->RT = cx.last_RT; }
    const size_t scope = cx.mark();
    Act x0;
    EAB_TRY(run_conv2d(cx, U.in_conv, srcs, nsrc, &x0, buf_x0));
    Act y = x0;
    std::vector<Act> keep;
    for (size_t i = 0; i < U.enco.size(); ++i) {
        Act z;
        EAB_TRY(run_conv2d(cx, U.enco[i], &y, 1, &z, nullptr));
        keep.push_back(z);
        y = z;
    }
    for (size_t i = 0; i < U.deco.size(); ++i) {
        Act z;
        float* pre = (lazy && i + 1 == U.deco.size()) ? buf_y : nullptr;
        if (i == 0) {
            EAB_TRY(run_conv2d(cx, U.deco[i], &y, 1, &z, pre));
        } else {
            Act pair[2] = {y, keep[keep.size() - 1 - i]};
            if (cx.m->cfg.intra_connect == 0) {
                EAB_TRY(run_conv2d(cx, U.deco[i], pair, 2, &z, pre));
            } else {
                Act sum;
                EAB_TRY(run_combine(cx, pair, 2, &sum));
                EAB_TRY(run_conv2d(cx, U.deco[i], &sum, 1, &z, pre));
            }
        }
        y = z;
    }
    if (lazy) {
        out->data = x0.data; out->xf = x0.xf;
        out->data2 = y.data; out->xf2 = y.xf;
    } else {
        Act pair[2] = {x0, y};
        EAB_TRY(run_combine_into(cx, pair, 2, out));
    }
    cx.release(scope);
    return 0;
}

// 1x1 "conv" over positions with optional bias / relu / residual / statistics (used by TCMs and the head)
int run_pointwise(Ctx& cx, const Act* srcs, int nsrc, const float* W, const float* bias, int Cout, int N, int gate_off,
                  int ntaps, const int* dt, int relu, const float* resid, int nstats, double** stats,
                  const float** stat_alpha, Act* out, const UmmaW* uw = nullptr, bool preallocated = false, int resid_RT = 0) {
    const bool use_umma = uw && uw->ok && (cx.tensor_ok() || cx.stream_umma()) && (resid == nullptr || uw->ld == Cout);
    if (!preallocated) {
        out->F = srcs[0].F;
        out->C = use_umma ? uw->ld : Cout;      // the tcgen05 path may pad the channel count (e.g. 18 -> 32, zeros)
        out->xf = xform_identity();
        out->data = cx.alloc_act((size_t)cx.B * cx.T * out->F * out->C);
        out->RT = cx.last_RT;
    }
    if (use_umma) {
        UmmaConvArgs us[4];
        for (int sp = 0; sp < uw->nsplit; ++sp) {
            UmmaConvArgs& u = us[sp];
            memset(&u, 0, sizeof(u));
            u.nsrc = nsrc;
            int cin = 0;
            for (int i = 0; i < nsrc; ++i) { set_src(u.src[i], srcs[i]); cin += srcs[i].C; }
            u.B = cx.B; u.T = cx.T; u.Fin = srcs[0].F; u.Fout = srcs[0].F; u.E = srcs[0].F;
            u.in_stride = 1; u.out_stride = 1; u.out_off = 0;
            u.ntaps = uw->ntaps;
            for (int i = 0; i < u.ntaps; ++i) { u.dt[i] = dt ? dt[i] : 0; u.df[i] = 0; }
            u.wide = 0; u.kwidth = 0; u.nslab = uw->nslab; u.ncoef = cin;
            u.npass = 3;                        // these layers are < 5 % of the FLOPs: keep them fp32-grade
            u.Whi = cx.W(uw->off_hi[sp]); u.Wlo = cx.W(uw->off_lo[sp]);
            u.bias = uw->has_bias ? cx.W(uw->off_bias[sp]) : nullptr;
            u.Cout = uw->cout; u.N = uw->ncol; u.gate_off = uw->gate_off; u.relu = relu;
            u.algo_frac = (nsrc == 2 && gate_off > 0) ? 0.5f : 1.f;
            u.out = out->data; u.out_ld = uw->ld; u.out_coff = sp * uw->cout;
            u.resid = resid;
            u.nstats = nstats;
            for (int i = 0; i < nstats; ++i) { u.stats[i] = stats[i]; u.stat_alpha[i] = stat_alpha[i]; }
            u.tiles_per_b = (cx.T * u.E + 127) / 128;
            if (!umma_conv_supported(u)) return fail("internal: pointwise layer rejected by the tcgen05 path");
        }
        if (cx.streaming) return run_umma_stream(cx, us, uw->nsplit, out->RT, resid_RT);
        return run_tensor_convs(cx, us, uw->nsplit);
    }
    if (cx.dry) return 0;
    for (int i = 0; i < nsrc; ++i)
        if (srcs[i].data2) return fail("internal: lazy activation reached the CUDA-core pointwise kernel");
    ConvArgs a;
    memset(&a, 0, sizeof(a));
    a.nsrc = nsrc;
    for (int i = 0; i < nsrc; ++i) { a.src[i].x = srcs[i].data; a.src[i].C = srcs[i].C; a.src[i].xf = srcs[i].xf; a.src[i].RT = srcs[i].RT; }
    a.step = cx.step; a.start = cx.start; a.out_RT = out->RT; a.resid_RT = resid_RT;
    a.B = cx.B; a.T = cx.T; a.Fin = srcs[0].F; a.Fout = srcs[0].F; a.E = srcs[0].F;
    a.in_stride = 1; a.out_stride = 1; a.out_off = 0;
    a.ntaps = ntaps;
    for (int i = 0; i < ntaps; ++i) { a.dt[i] = dt ? dt[i] : 0; a.df[i] = 0; }
    a.W = W; a.bias = bias; a.Cout = Cout; a.N = N; a.gate_off = gate_off; a.relu = relu;
    a.algo_frac = (nsrc == 2 && gate_off > 0) ? 0.5f : 1.f;      // merged TCM branches: block-diagonal weights
    a.out = out->data; a.resid = resid;
    a.nstats = nstats;
    for (int i = 0; i < nstats; ++i) { a.stats[i] = stats[i]; a.stat_alpha[i] = stat_alpha[i]; }
    return launch_conv(a, cx.st);
}

// SqueezedTCM.forward (EaBNet.py:572-578) on the channels-last residual stream x [B,T,1,d_feat]
int run_tcm(Ctx& cx, const TcmLayer& t, const Act& x, Act* out) {
    const eab_config& c = cx.m->cfg;
    const bool in_stats = c.norm_type == 0;
    // the residual-stream output first, the squeezed intermediates are scoped scratch
    out->F = x.F; out->C = c.d_feat; out->xf = xform_identity();
    out->data = cx.alloc_act((size_t)cx.B * cx.T * x.F * c.d_feat);
    out->RT = cx.last_RT;
    const size_t scope = cx.mark();
    if (cx.streaming) {                        // the dilated taps reach (kd1-1)*d frames back into the squeezed tensor
        int back = 0;
        for (int i = 0; i < c.kd1; ++i) back = std::max(back, t.dt[i]);
        cx.next_RT = back + 1;
    }
    Act z;
    double* st_o[1] = {nullptr};
    if (t.single) {
        // GaGNet's SqueezedTCM: squeeze 1x1 (statistics of PReLU(y)), one plain dilated conv
        double* st_d[1] = {in_stats ? cx.alloc_stats(c.cd1) : nullptr};
        const float* al_d[1] = {cx.W(t.na_left.off_alpha)};
        Act y;
        EAB_TRY(run_pointwise(cx, &x, 1, cx.W(t.off_in), nullptr, c.cd1, pad_n(c.cd1), 0, 1, nullptr, 0, nullptr,
                              in_stats ? 1 : 0, st_d, al_d, &y, &t.u_in));
        y.xf = xf_after(cx, t.na_left, st_d[0], cx.T, 1);
        st_o[0] = in_stats ? cx.alloc_stats(c.cd1) : nullptr;
        const float* al_o[1] = {cx.W(t.na_out.off_alpha)};
        EAB_TRY(run_pointwise(cx, &y, 1, cx.W(t.off_dil), nullptr, c.cd1, pad_n(c.cd1), 0, c.kd1, t.dt, 0, nullptr,
                              in_stats ? 1 : 0, st_o, al_o, &z, &t.u_dil));
    } else {
    // squeeze 1x1; statistics of PReLU_left(y) and PReLU_right(y) for the two branch norms
    double* st_lr[2] = {in_stats ? cx.alloc_stats(c.cd1) : nullptr, in_stats ? cx.alloc_stats(c.cd1) : nullptr};
    const float* al_lr[2] = {cx.W(t.na_left.off_alpha), cx.W(t.na_right.off_alpha)};
    Act y;
    EAB_TRY(run_pointwise(cx, &x, 1, cx.W(t.off_in), nullptr, c.cd1, pad_n(c.cd1), 0, 1, nullptr, 0, nullptr,
                          in_stats ? 2 : 0, st_lr, al_lr, &y, &t.u_in));
    // both dilated branches as one gated conv: value = left branch, gate = right branch (sigmoid)
    Act br[2] = {y, y};
    br[0].xf = xf_after(cx, t.na_left, st_lr[0], cx.T, 1);
    br[1].xf = xf_after(cx, t.na_right, st_lr[1], cx.T, 1);
    st_o[0] = in_stats ? cx.alloc_stats(c.cd1) : nullptr;
    const float* al_o[1] = {cx.W(t.na_out.off_alpha)};
    EAB_TRY(run_pointwise(cx, br, 2, cx.W(t.off_dil), nullptr, c.cd1, 2 * ceil64(c.cd1), ceil64(c.cd1), c.kd1, t.dt, 0,
                          nullptr, in_stats ? 1 : 0, st_o, al_o, &z, &t.u_dil));
    }
    // expand 1x1 + residual
    z.xf = xf_after(cx, t.na_out, st_o[0], cx.T, 1);
    EAB_TRY(run_pointwise(cx, &z, 1, cx.W(t.off_out), nullptr, c.d_feat, c.d_feat, 0, 1, nullptr, 0, x.data, 0, nullptr,
                          nullptr, out, &t.u_out, /*preallocated=*/true, x.RT));
    cx.release(scope);
    return 0;
}

void tap(Ctx& cx, const char* name, const Act& a) {
    if (cx.dry || cx.streaming) return;
    Tap t;
    t.act = a; t.B = cx.B; t.T = cx.T;
    cx.m->taps[name] = t;
}

struct ChainRef { const TcmLayer* l; int n; };
bool tcm_chain_ok(Ctx& cx, const ChainRef* chains, int nch);
int run_tcm_chains(Ctx& cx, const ChainRef* chains, int nch, const Act* ins, Act* outs);

int run_forward(Ctx& cx, const float* inpt, float* out_dev) {
    eab_model* m = cx.m;
    const eab_config& c = m->cfg;
    Act x;
    x.data = const_cast<float*>(inpt);
    x.F = c.n_freq;
    x.C = 2 * c.M;
    x.RT = cx.streaming ? 2 : 0;          // streaming: inpt is the [S][2][F][M][2] spectrum ring of the stream state
    const int inpt_RT = x.RT;
    std::vector<Act> skips;
    // ---------------- encoder (EaBNet.py:190-197 / :234-239)
    if (c.is_u2) {
        for (size_t i = 0; i < m->en_mod.size(); ++i) {
            Act y;
            EAB_TRY(run_module(cx, m->en_mod[i], &x, 1, &y));
            skips.push_back(y);
            tap(cx, ("en." + std::to_string(i)).c_str(), y);
            x = y;
        }
        Act y;
        EAB_TRY(run_conv2d(cx, m->en_last, &x, 1, &y));
        skips.push_back(y);
        tap(cx, "en.4", y);
        x = y;
    } else {
        for (size_t i = 0; i < m->en_plain.size(); ++i) {
            Act y;
            EAB_TRY(run_conv2d(cx, m->en_plain[i], &x, 1, &y));
            skips.push_back(y);
            tap(cx, ("en." + std::to_string(i)).c_str(), y);
            x = y;
        }
    }
    if (x.F != m->Fb || x.C != 64) return fail("internal: bottleneck shape");
    // ---------------- squeezed TCM stack (EaBNet.py:99-106); channel order f*64+c is kept both ways
    Act r;                                            // residual stream [B,T,1,d_feat], finalised
    EAB_TRY(run_combine(cx, &x, 1, &r));
    r.F = 1;
    r.C = c.d_feat;
    r.xf = xform_identity();
    Act acc;
    if (cx.streaming && tcm_stream_supported(c.cd1, c.d_feat, c.kd1) && c.norm_type == 1 && m->opt_stream_tcm) {
        // one launch for the whole stack (tcm_stream.cu): per-TCM history rings of the squeezed tensor + the group sum
        std::vector<TcmStreamDesc> descs;
        for (const TcmLayer& t : m->tcms) {
            TcmStreamDesc d;
            memset(&d, 0, sizeof(d));
            int back = 0;
            for (int i = 0; i < c.kd1; ++i) { d.dt[i] = t.dt[i]; back = std::max(back, t.dt[i]); }
            cx.next_RT = back + 1;
            float* ring = cx.alloc_act((size_t)cx.B * c.cd1);
            d.RT = cx.last_RT;
            d.ring_off = (long long)((reinterpret_cast<char*>(ring) - cx.base) / (ptrdiff_t)sizeof(float));
            d.W_in = (long long)t.off_in; d.W_dil = (long long)t.off_dil; d.W_out = (long long)t.off_out;
            d.sL = (long long)t.na_left.off_scale; d.hL = (long long)t.na_left.off_shift; d.aL = (long long)t.na_left.off_alpha;
            d.sR = (long long)t.na_right.off_scale; d.hR = (long long)t.na_right.off_shift; d.aR = (long long)t.na_right.off_alpha;
            d.sO = (long long)t.na_out.off_scale; d.hO = (long long)t.na_out.off_shift; d.aO = (long long)t.na_out.off_alpha;
            descs.push_back(d);
        }
        acc.F = 1; acc.C = c.d_feat; acc.xf = xform_identity();
        acc.data = cx.alloc_act((size_t)cx.B * c.d_feat);
        acc.RT = cx.last_RT;
        if (cx.tcm_desc) *cx.tcm_desc = descs;
        if (!cx.dry) {
            TcmStreamArgs a;
            memset(&a, 0, sizeof(a));
            a.desc = cx.tcm_desc_dev; a.blob = m->blob;
            a.ntcm = (int)m->tcms.size(); a.p = c.p; a.kd = c.kd1; a.S = cx.B;
            a.step = cx.step; a.start = cx.start; a.act_base = reinterpret_cast<float*>(cx.base);
            a.x = r.data; a.x_RT = r.RT; a.out = acc.data; a.out_RT = acc.RT;
            EAB_TRY(launch_tcm_stream(a, cx.st));
        }
    } else {
    std::vector<Act> group_out;
    size_t ti = 0;
    for (int g = 0; g < c.q; ++g) {
        const ChainRef chain = {m->tcms.data() + (size_t)g * c.p, c.p};
        if (!cx.streaming && tcm_chain_ok(cx, &chain, 1)) {
            // a whole group (p gated TCMs) as one persistent cooperative launch (tcm_chain.cu)
            Act nx;
            EAB_TRY(run_tcm_chains(cx, &chain, 1, &r, &nx));
            r = nx;
            ti += c.p;
        } else {
            for (int i = 0; i < c.p; ++i) {
                Act nx;
                EAB_TRY(run_tcm(cx, m->tcms[ti++], r, &nx));
                r = nx;
            }
        }
        group_out.push_back(r);
    }
    if (c.q == 1) acc = group_out[0];
    else EAB_TRY(run_combine(cx, group_out.data(), c.q, &acc));
    }
    acc.F = m->Fb;
    acc.C = 64;
    tap(cx, "tcm", acc);
    // ---------------- decoder (EaBNet.py:273-279 / :324-328)
    x = acc;
    Act emb;
    if (c.is_u2) {
        for (size_t i = 0; i < m->de_mod.size(); ++i) {
            Act pair[2] = {x, skips[skips.size() - 1 - i]};
            Act y;
            EAB_TRY(run_module(cx, m->de_mod[i], pair, 2, &y));
            tap(cx, ("de." + std::to_string(i)).c_str(), y);
            x = y;
        }
        Act pair[2] = {x, skips[0]};
        EAB_TRY(run_conv2d(cx, m->de_last, pair, 2, &emb));
    } else {
        for (size_t i = 0; i < m->de_plain.size(); ++i) {
            Act pair[2] = {x, skips[skips.size() - 1 - i]};
            Act y;
            EAB_TRY(run_conv2d(cx, m->de_plain[i], pair, 2, &y));
            if (i + 1 < m->de_plain.size()) tap(cx, ("de." + std::to_string(i)).c_str(), y);
            x = y;
        }
        emb = x;
    }
    if (emb.F != c.n_freq) return fail("decoder output width differs from the input width (the reference would fail in filter-and-sum)");
    tap(cx, "embed", emb);
    // ---------------- beam-weight head + filter-and-sum (EaBNet.py:108-125, 600-614)
    Act w;
    if (c.topo_type == 0 && c.bf_type == 0) {
        Act h[2];
        for (int l = 0; l < 2; ++l) {
            h[l].F = c.n_freq; h[l].C = 64; h[l].xf = xform_identity();
            h[l].data = cx.alloc_act((size_t)cx.B * cx.T * c.n_freq * 64);
            h[l].RT = cx.last_RT;
            float* hc_state[2] = {nullptr, nullptr};
            if (cx.streaming)
                for (int k = 0; k < 2; ++k) {
                    cx.next_RT = 1; hc_state[k] = cx.alloc_act((size_t)cx.B * c.n_freq * 64);
                    if (cx.per_stream) cx.per_stream->push_back({(size_t)(reinterpret_cast<char*>(hc_state[k]) - cx.base), (size_t)c.n_freq * 64 * sizeof(float)});
                }
            if (!cx.dry) {
                LstmArgs a;
                memset(&a, 0, sizeof(a));
                const Act& src = l ? h[0] : emb;
                a.src.x = src.data; a.src.C = src.C; a.src.xf = src.xf; a.src.RT = src.RT;
                a.step = cx.step; a.h_state = hc_state[0]; a.c_state = hc_state[1]; a.out_RT = h[l].RT;
                a.layer_norm = l == 0;
                a.ln_g = cx.W(m->off_ln_g); a.ln_b = cx.W(m->off_ln_b);
                a.Wx = cx.W(m->off_rnn[l][0]); a.Wh = cx.W(m->off_rnn[l][1]); a.bias = cx.W(m->off_rnn[l][2]);
                a.B = cx.B; a.T = cx.T; a.F = c.n_freq; a.E = src.C;
                a.out = h[l].data;
                if (cx.tensor_ok() && m->rnn_umma_ok) {
                    LstmArgs u = a;
                    if (m->opt_dbg_launch == -100 - l && m->dbg_buf) u.dbg = m->dbg_buf;
                    u.exp_flags = m->opt_lstm_exp;
                    u.Wimg = cx.W(m->off_rnn_img[l]);
                    u.bias = cx.W(m->off_rnn_ubias[l]);
                    if (lstm_umma_supported(u)) {
                        EAB_TRY(launch_lstm_umma(u, cx.st));
                        tap(cx, l ? "h2" : "h1", h[l]);
                        continue;
                    }
                }
                EAB_TRY(launch_lstm(a, cx.st));
            }
            tap(cx, l ? "h2" : "h1", h[l]);
        }
        // fused w_dnn + filter-and-sum (head_fused.cu): one read of h2, nothing else touches HBM
        if (cx.tensor_ok() && m->opt_fused_head && m->u_dnn[0].ok && m->u_dnn[1].ok && m->u_dnn[0].nsplit == 1 &&
            m->u_dnn[1].nsplit == 1 && m->u_dnn[0].ncol == 64 && m->u_dnn[1].ncol == 32 && h[1].xf.affine == 0 &&
            h[1].xf.prelu == 0 && c.M <= 16) {
            float* wtap = nullptr;
            if (m->opt_head_w_tap) wtap = cx.alloc_act((size_t)cx.B * cx.T * c.n_freq * 32);
            if (!cx.dry) {
                HeadArgs a;
                memset(&a, 0, sizeof(a));
                a.h = h[1].data; a.inpt = inpt; a.out = out_dev; a.w_out = wtap; a.w_ld = 32;
                a.rows = (long long)cx.B * cx.T * c.n_freq; a.T = cx.T; a.F = c.n_freq; a.M = c.M;
                a.W1hi = cx.W(m->u_dnn[0].off_hi[0]); a.W1lo = cx.W(m->u_dnn[0].off_lo[0]);
                a.W2hi = cx.W(m->u_dnn[1].off_hi[0]); a.W2lo = cx.W(m->u_dnn[1].off_lo[0]);
                a.b1 = cx.W(m->u_dnn[0].off_bias[0]); a.b2 = cx.W(m->u_dnn[1].off_bias[0]);
                EAB_TRY(launch_head_fused(a, cx.st));
                if (wtap) { Act wa; wa.data = wtap; wa.F = c.n_freq; wa.C = 32; tap(cx, "w", wa); }
            }
            return 0;
        }
        Act u;
        EAB_TRY(run_pointwise(cx, &h[1], 1, cx.W(m->off_dnn_w[0]), cx.W(m->off_dnn_b[0]), 64, m->dnn_N[0], 0, 1, nullptr, 1,
                              nullptr, 0, nullptr, nullptr, &u, &m->u_dnn[0]));
        EAB_TRY(run_pointwise(cx, &u, 1, cx.W(m->off_dnn_w[1]), cx.W(m->off_dnn_b[1]), 2 * c.M, m->dnn_N[1], 0, 1, nullptr, 0,
                              nullptr, 0, nullptr, nullptr, &w, &m->u_dnn[1]));
    } else {
        const int n = c.topo_type == 0 ? 2 * c.M : 2;
        EAB_TRY(run_pointwise(cx, &emb, 1, cx.W(m->off_cnn_w), cx.W(m->off_cnn_b), n, m->cnn_N, 0, 1, nullptr, 0, nullptr, 0,
                              nullptr, nullptr, &w, &m->u_cnn));
    }
    tap(cx, "w", w);
    if (!cx.dry) {
        BeamArgs a;
        memset(&a, 0, sizeof(a));
        a.step = cx.step; a.w_RT = w.RT; a.inpt_RT = inpt_RT;
        a.w = w.data; a.w_ld = w.C; a.inpt = inpt; a.B = cx.B; a.T = cx.T; a.F = c.n_freq; a.M = c.M; a.miso = c.topo_type == 1;
        a.out = out_dev;
        EAB_TRY(launch_beam(a, cx.st));
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------- GaGNet post-filter
// in_conv_main(cat(feat, pre)) * sigmoid(in_conv_gate(cat(feat, pre)))  (GaGNet.py:189-191, 249-251): d_feat/64 gated
// tcgen05 launches over K = d_feat + KP (two sources, the concat is never materialised)
int run_gag_in(Ctx& cx, const GagIn& in, const Act& feat, const Act& pre, Act* out) {
    const eab_config& c = cx.m->cfg;
    out->F = 1; out->C = c.d_feat; out->xf = xform_identity();
    out->data = cx.alloc_act((size_t)cx.B * cx.T * c.d_feat);
    out->RT = cx.last_RT;
    if (!cx.tensor_ok() && !cx.stream_umma()) {
        // CUDA-core path (streaming with stream_umma = 0, option umma = 0): the dense [K][value | gate] matrix of every column split; the split's
        // SW output channels are position `sp` of a [.., nsplit, SW] view of the d_feat-wide row
        if (cx.dry) return 0;
        for (int sp = 0; sp < in.nsplit; ++sp) {
            if (!in.u[sp].ok) return fail("internal: GaGNet input conv bias images missing");
            ConvArgs a;
            memset(&a, 0, sizeof(a));
            a.nsrc = 2;
            const Act* srcs[2] = {&feat, &pre};
            for (int i = 0; i < 2; ++i) { a.src[i].x = srcs[i]->data; a.src[i].C = srcs[i]->C; a.src[i].xf = srcs[i]->xf; a.src[i].RT = srcs[i]->RT; }
            a.step = cx.step; a.start = cx.start; a.out_RT = out->RT;
            a.B = cx.B; a.T = cx.T; a.Fin = 1; a.E = 1; a.Fout = in.nsplit;
            a.in_stride = 1; a.out_stride = 1; a.out_off = sp;
            a.ntaps = 1; a.dt[0] = 0; a.df[0] = 0;
            a.W = cx.W(in.off_dense[sp]); a.bias = cx.W(in.u[sp].off_bias[0]);
            a.Cout = in.SW; a.N = 2 * in.SW; a.gate_off = in.SW;
            a.algo_frac = (float)(c.d_feat + 2 * c.n_freq) / (float)in.K;
            a.out = out->data;
            EAB_TRY(launch_conv(a, cx.st));
        }
        return 0;
    }
    for (int sp = 0; sp < in.nsplit; ++sp) {
        if (!cx.dry && !in.u[sp].ok) return fail("internal: GaGNet input conv images missing");
        UmmaConvArgs u;
        memset(&u, 0, sizeof(u));
        u.nsrc = 2;
        set_src(u.src[0], feat);
        set_src(u.src[1], pre);
        u.B = cx.B; u.T = cx.T; u.Fin = 1; u.Fout = 1; u.E = 1;
        u.in_stride = 1; u.out_stride = 1; u.out_off = 0;
        u.ntaps = 1; u.dt[0] = 0; u.df[0] = 0;
        u.nslab = in.K / 64; u.ncoef = in.K; u.npass = 3;
        u.Whi = cx.W(in.u[sp].off_hi[0]); u.Wlo = cx.W(in.u[sp].off_lo[0]); u.bias = cx.W(in.u[sp].off_bias[0]);
        u.Cout = in.SW; u.N = 2 * in.SW; u.gate_off = in.SW;
        u.algo_frac = (float)(c.d_feat + 2 * c.n_freq) / (float)in.K;
        u.out = out->data; u.out_ld = c.d_feat; u.out_coff = sp * in.SW;
        u.tiles_per_b = (cx.T + 127) / 128;
        if (feat.C + pre.C != in.K || !umma_conv_supported(u)) return fail("internal: GaGNet input conv rejected by the tcgen05 path");
        if (cx.streaming) EAB_TRY(run_umma_stream(cx, &u, 1, out->RT, 0));
        else EAB_TRY(run_tensor_convs(cx, &u, 1));
    }
    return 0;
}

int run_gag_lin(Ctx& cx, const GagLin& l, const Act& x, Act* out) {
    const eab_config& c = cx.m->cfg;
    return run_pointwise(cx, &x, 1, cx.W(l.off_w), cx.W(l.off_b), c.n_freq, l.N, 0, 1, nullptr, 0, nullptr, 0, nullptr, nullptr,
                         out, &l.u);
}

// 1-3 equally long chains of single-branch TCMs as one cooperative launch (tcm_chain.cu)
bool tcm_chain_ok(Ctx& cx, const ChainRef* chains, int nch) {
    const eab_config& c = cx.m->cfg;
    if (!cx.m->opt_tcm_chain || !cx.tensor_ok() || c.cd1 != 64 || c.d_feat != 256 || c.kd1 > 8) return false;
    const bool gated = !chains[0].l[0].single;
    if (gated ? (c.kd1 != 3 && c.kd1 != 5) : c.kd1 > 4) return false;
    const size_t nl = (size_t)chains[0].n;
    if (nl < 1 || nch < 1 || nch > 3 || nch * nl > (size_t)kMaxChainLayers) return false;
    for (int i = 0; i < nch; ++i) {
        if ((size_t)chains[i].n != nl) return false;
        for (int li = 0; li < chains[i].n; ++li) {
            const TcmLayer& t = chains[i].l[li];
            // (planning runs before the weights are packed: the image flags are only known once committed)
            if (t.single == gated) return false;
            if (!cx.m->dirty && !(t.u_in.ok && t.u_out.ok && t.u_out.nsplit == 2 && t.u_in.nslab == 4)) return false;
            if (!cx.m->dirty && !(gated ? (t.u_dl.ok && t.u_dr.ok) : t.u_dil.ok)) return false;
            for (int k = 0; k < c.kd1; ++k) if (t.dt[k] > 30000 || t.dt[k] < -30000) return false;
        }
    }
    return true;
}

int run_tcm_chains(Ctx& cx, const ChainRef* chains, int nch, const Act* ins, Act* outs) {
    const eab_config& c = cx.m->cfg;
    const int nl = chains[0].n;
    const size_t rows = (size_t)cx.B * cx.T;
    TcmChainArgs a;
    memset(&a, 0, sizeof(a));
    for (int i = 0; i < nch; ++i) {
        if (ins[i].xf.affine != 0 || ins[i].xf.prelu != 0 || ins[i].data2 || ins[i].C != 256)
            return fail("internal: TCM chain input must be a plain fp32 [B,T,256] tensor");
        outs[i].F = 1; outs[i].C = c.d_feat; outs[i].xf = xform_identity();
        outs[i].data = cx.alloc_act(rows * 256);
        outs[i].RT = cx.last_RT;
        a.x_in[i] = ins[i].data;
        a.x_buf[i] = outs[i].data;
    }
    const size_t scope = cx.mark();
    for (int i = 0; i < nch; ++i) { a.y[i] = cx.alloc_act(rows * 64); a.z[i] = cx.alloc_act(rows * 64); }
    cx.release(scope);                                   // y / z die with the launch
    const bool in_stats = c.norm_type == 0;
    double* sbase = reinterpret_cast<double*>(cx.base);
    a.barrier = reinterpret_cast<unsigned*>(cx.alloc_stats(1));
    for (int i = 0; i < nch; ++i)
        for (int l = 0; l < nl; ++l) {
            const TcmLayer& t = chains[i].l[l];
            TcmChainLayer& L = a.L[i * nl + l];
            L.win_hi = (unsigned)t.u_in.off_hi[0]; L.win_lo = (unsigned)t.u_in.off_lo[0];
            const bool gated = !t.single;
            L.wd_hi = (unsigned)(gated ? t.u_dl : t.u_dil).off_hi[0]; L.wd_lo = (unsigned)(gated ? t.u_dl : t.u_dil).off_lo[0];
            if (gated) {
                L.wr_hi = (unsigned)t.u_dr.off_hi[0]; L.wr_lo = (unsigned)t.u_dr.off_lo[0];
                L.sc_r = (unsigned)t.na_right.off_scale; L.sh_r = (unsigned)t.na_right.off_shift; L.al_r = (unsigned)t.na_right.off_alpha;
                if (in_stats) L.st_r = (unsigned)(cx.alloc_stats(64) - sbase);
            }
            for (int sp = 0; sp < 2; ++sp) { L.wo_hi[sp] = (unsigned)t.u_out.off_hi[sp]; L.wo_lo[sp] = (unsigned)t.u_out.off_lo[sp]; }
            L.sc_d = (unsigned)t.na_left.off_scale; L.sh_d = (unsigned)t.na_left.off_shift; L.al_d = (unsigned)t.na_left.off_alpha;
            L.sc_o = (unsigned)t.na_out.off_scale; L.sh_o = (unsigned)t.na_out.off_shift; L.al_o = (unsigned)t.na_out.off_alpha;
            if (in_stats) {
                L.st_d = (unsigned)(cx.alloc_stats(64) - sbase);
                L.st_o = (unsigned)(cx.alloc_stats(64) - sbase);
            }
            for (int k = 0; k < c.kd1; ++k) L.dt[k] = (short)t.dt[k];
        }
    if (cx.dry) return 0;
    a.blob = cx.m->blob; a.stats = sbase;
    a.nchains = nch; a.nlayers = nl; a.kd = c.kd1; a.B = cx.B; a.T = cx.T;
    a.gated = chains[0].l[0].single ? 0 : 1;
    a.instance_norm = in_stats ? 1 : 0; a.inv_count = 1.f / (float)cx.T;
    if (cx.m->opt_dbg_launch == -200 && cx.m->dbg_buf) a.dbg = cx.m->dbg_buf;
    a.no_cluster = cx.m->opt_tcm_chain == 3 ? 1 : 0;
    if (cx.m->opt_tcm_chain == 2 && nch > 1) {
        // one chain per launch: a single chain's residual stream + scratch (59 MB at 64 x 6 s) stays in the 126 MB L2
        for (int i = 0; i < nch; ++i) {
            TcmChainArgs s = a;
            s.nchains = 1;
            s.x_in[0] = a.x_in[i]; s.x_buf[0] = a.x_buf[i]; s.y[0] = a.y[i]; s.z[0] = a.z[i];
            for (int l = 0; l < nl; ++l) s.L[l] = a.L[i * nl + l];
            s.barrier = a.barrier + 4 * i;
            EAB_TRY(launch_tcm_chain(s, cx.st));
        }
        return 0;
    }
    return launch_tcm_chain(a, cx.st);
}

// GaGNet.forward (GaGNet.py:75-89).  inpt through strides sb, sc, st, sf (floats); pre [B,2,T,F]; out [q][B,2,T,F].
int run_gag_forward(Ctx& cx, const float* inpt, const long long* strides, const float* pre_in, float* out_dev) {
    eab_model* m = cx.m;
    const eab_config& c = m->cfg;
    const int F = c.n_freq, KP = ceil64(2 * F);
    Act x;
    x.F = F; x.C = 4; x.xf = xform_identity();
    x.data = cx.alloc_act((size_t)cx.B * cx.T * F * 4);          // streaming: a ring of 2 frames (the kt = 2 first conv)
    x.RT = cx.last_RT;
    Act pre;
    pre.F = 1; pre.C = KP; pre.xf = xform_identity();
    if (cx.streaming) cx.next_RT = 1;                            // read by 1x1 convs and the elementwise kernels only
    pre.data = cx.alloc_act((size_t)cx.B * cx.T * KP);
    pre.RT = cx.last_RT;
    if (!cx.dry) {
        GagPackArgs a;
        memset(&a, 0, sizeof(a));
        a.inpt = inpt; a.sb = strides[0]; a.sc = strides[1]; a.st = strides[2]; a.sf = strides[3];
        a.pre = pre_in; a.x4 = x.data; a.pre_row = pre.data;
        a.B = cx.B; a.T = cx.T; a.F = F; a.KP = KP; a.KP2 = KP / 2;
        if (cx.streaming) { a.step = cx.step; a.in_RT = cx.gag_in_RT; a.in_slot = cx.gag_in_slot; a.x_RT = x.RT; }
        EAB_TRY(launch_gag_pack(a, cx.st));
    }
    // ---------------- encoder (GaGNet.py:361-365 / :408-412): only the bottleneck is used
    if (c.is_u2) {
        for (size_t i = 0; i < m->en_mod.size(); ++i) {
            Act y;
            EAB_TRY(run_module(cx, m->en_mod[i], &x, 1, &y));
            tap(cx, ("en." + std::to_string(i)).c_str(), y);
            x = y;
        }
        Act y;
        EAB_TRY(run_conv2d(cx, m->en_last, &x, 1, &y));
        x = y;
    } else {
        for (size_t i = 0; i < m->en_plain.size(); ++i) {
            Act y;
            EAB_TRY(run_conv2d(cx, m->en_plain[i], &x, 1, &y));
            x = y;
        }
    }
    tap(cx, "en.4", x);
    if (x.F != m->Fb || x.C != 64) return fail("internal: bottleneck shape");
    Act feat;                                             // [B,T,1,d_feat] finalised, channel f*64+c
    EAB_TRY(run_combine(cx, &x, 1, &feat));
    feat.F = 1; feat.C = c.d_feat; feat.xf = xform_identity();
    // ---------------- glance-gaze modules (GaGNet.py:84-88, 120-134)
    const size_t stage_elems = (size_t)cx.B * 2 * cx.T * F;
    for (size_t gi = 0; gi < m->gags.size(); ++gi) {
        const GagModule& G = m->gags[gi];
        const bool last = gi + 1 == m->gags.size();
        Act next;
        next.F = 1; next.C = KP; next.xf = xform_identity();
        if (cx.streaming) cx.next_RT = 1;
        next.data = last ? nullptr : cx.alloc_act((size_t)cx.B * cx.T * KP);      // outlives the module's scratch
        next.RT = cx.last_RT;
        const size_t scope = cx.mark();
        Act xg, xz;
        EAB_TRY(run_gag_in(cx, G.in_g, feat, pre, &xg));
        EAB_TRY(run_gag_in(cx, G.in_z, feat, pre, &xz));
        tap(cx, ("g.in_g." + std::to_string(gi)).c_str(), xg);
        Act xr = xz, xi = xz;
        {
            const ChainRef chains[3] = {{G.tcn_g.data(), (int)G.tcn_g.size()}, {G.tcm_r.data(), (int)G.tcm_r.size()},
                                        {G.tcm_i.data(), (int)G.tcm_i.size()}};
            const int nch = m->gcfg.is_squeezed ? 2 : 3;
            if (tcm_chain_ok(cx, chains, nch)) {
                Act ins[3] = {xg, xz, xz}, outs[3];
                EAB_TRY(run_tcm_chains(cx, chains, nch, ins, outs));
                xg = outs[0]; xr = outs[1]; xi = nch == 3 ? outs[2] : outs[1];
            } else {
                for (const TcmLayer& t : G.tcn_g) { Act nx; EAB_TRY(run_tcm(cx, t, xg, &nx)); xg = nx; }
                for (const TcmLayer& t : G.tcm_r) { Act nx; EAB_TRY(run_tcm(cx, t, xr, &nx)); xr = nx; }
                if (m->gcfg.is_squeezed) xi = xr;
                else for (const TcmLayer& t : G.tcm_i) { Act nx; EAB_TRY(run_tcm(cx, t, xi, &nx)); xi = nx; }
            }
        }
        tap(cx, ("g.tcn_g." + std::to_string(gi)).c_str(), xg);
        Act gain;
        if (cx.streaming) cx.next_RT = 1;                        // (gag_crm_kernel reads plain rows)
        EAB_TRY(run_gag_lin(cx, G.lin_g, xg, &gain));
        Act rr, ri;
        if (cx.streaming) cx.next_RT = 1;
        EAB_TRY(run_gag_lin(cx, G.lin_r, xr, &rr));
        if (cx.streaming) cx.next_RT = 1;
        EAB_TRY(run_gag_lin(cx, G.lin_i, xi, &ri));
        tap(cx, ("g.gain." + std::to_string(gi)).c_str(), gain);
        tap(cx, ("g.res_r." + std::to_string(gi)).c_str(), rr);
        if (rr.C != ri.C) return fail("internal: GaGNet residual widths differ");
        if (!cx.dry) {
            GagCrmArgs a;
            memset(&a, 0, sizeof(a));
            a.pre_row = pre.data; a.gain = gain.data; a.res_r = rr.data; a.res_i = ri.data;
            a.ld_g = gain.C; a.ld_r = rr.C; a.acti = m->gcfg.acti_type;
            a.next_row = next.data; a.out = out_dev + gi * stage_elems;
            a.B = cx.B; a.T = cx.T; a.F = F; a.KP = KP; a.KP2 = KP / 2;
            EAB_TRY(launch_gag_crm(a, cx.st));
        }
        cx.release(scope);
        pre = next;
    }
    return 0;
}

int plan(eab_model* m, int B, int T, size_t* stats_bytes, size_t* total_bytes) {
    Ctx cx;
    cx.m = m; cx.dry = true; cx.base = nullptr; cx.B = B; cx.T = T; cx.st = nullptr;
    if (m->kind == 1) {
        const long long zero[4] = {0, 0, 0, 0};
        EAB_TRY(run_gag_forward(cx, nullptr, zero, nullptr, nullptr));
    } else
    EAB_TRY(run_forward(cx, nullptr, nullptr));
    *stats_bytes = cx.stats_off;
    *total_bytes = cx.stats_off + cx.act_peak;
    return 0;
}

// ---------------------------------------------------------------------------------------------- streaming
// State blob layout (device, caller-owned): [0,256) absolute frame counter | carried hop [S][M][160] | iSTFT tail
// [S][160] | spectrum ring [S][2][F][M][2] | output frame [S][2][F] | activation rings + LSTM state (run_forward order)
struct StreamLayout {
    size_t off_start, off_desc, off_prev, off_tail, off_spec, off_out, off_act, total;
    std::vector<TcmStreamDesc> descs;
    std::vector<std::pair<size_t, size_t>> per_stream;      // (offset from off_act, bytes per stream) of the carried LSTM (h, c)
};
inline size_t up256(size_t x) { return (x + 255) / 256 * 256; }

int stream_layout(eab_model* m, int S, StreamLayout* L) {
    const eab_config& c = m->cfg;
    if (S < 1) return fail("stream: need at least one stream");
    if (c.norm_type != 1) return fail("streaming needs norm_type='BN': InstanceNorm statistics span the whole utterance (EaBNet.py:684-686)");
    if (!c.is_causal) return fail("streaming needs is_causal=True");
    size_t o = 256;
    L->off_start = o; o += up256((size_t)S * sizeof(int));
    if (m->kind == 1) {
        // GaGNet post-filter: [0,256) frame counter | start | estimates of the q modules [q][S][2][F] | activation rings
        L->off_desc = L->off_prev = L->off_tail = L->off_spec = o;
        L->off_out = o; o += up256(m->gags.size() * (size_t)S * 2 * c.n_freq * sizeof(float));
        L->off_act = o;
        Ctx cx;
        cx.m = m; cx.dry = true; cx.base = nullptr; cx.B = S; cx.T = 1; cx.st = nullptr; cx.streaming = true;
        const long long zero[4] = {0, 0, 0, 0};
        EAB_TRY(run_gag_forward(cx, nullptr, zero, nullptr, nullptr));
        if (cx.stats_off != 0) return fail("internal: streaming plan allocated statistics");
        L->total = o + cx.act_peak;
        return 0;
    }
    L->off_desc = o; o += up256(m->tcms.size() * sizeof(TcmStreamDesc));
    L->off_prev = o; o += up256((size_t)S * c.M * 160 * sizeof(float));
    L->off_tail = o; o += up256((size_t)S * 160 * sizeof(float));
    L->off_spec = o; o += up256((size_t)S * 2 * c.n_freq * c.M * 2 * sizeof(float));
    L->off_out = o;  o += up256((size_t)S * 2 * c.n_freq * sizeof(float));
    L->off_act = o;
    Ctx cx;
    cx.m = m; cx.dry = true; cx.base = nullptr; cx.B = S; cx.T = 1; cx.st = nullptr; cx.streaming = true;
    cx.tcm_desc = &L->descs;
    cx.per_stream = &L->per_stream;
    EAB_TRY(run_forward(cx, nullptr, nullptr));
    if (cx.stats_off != 0) return fail("internal: streaming plan allocated statistics");
    L->total = o + cx.act_peak;
    return 0;
}

int stream_forward(eab_model* m, char* state, const StreamLayout& L, int S, cudaStream_t st) {
    if (m->dirty) return fail("parameters not committed: call eab_commit_params first");
    Ctx cx;
    cx.m = m; cx.dry = false; cx.base = state + L.off_act; cx.B = S; cx.T = 1; cx.st = st;
    cx.streaming = true; cx.step = reinterpret_cast<const int*>(state);
    cx.start = reinterpret_cast<const int*>(state + L.off_start);
    cx.tcm_desc_dev = reinterpret_cast<const TcmStreamDesc*>(state + L.off_desc);
    return run_forward(cx, reinterpret_cast<const float*>(state + L.off_spec), reinterpret_cast<float*>(state + L.off_out));
}

// one frame of every stream through the post-filter: inpt element (s, ri, f) at inpt[s*strides[0] + ri*strides[1] + f*strides[3]
// + slot*in_slot] (slot = *step % in_RT: the spectrum ring of an EaBNet stream state, or in_RT = 1 for a plain frame);
// pre [S][2][F]; the q estimates go to the state's [q][S][2][F] block
int gag_stream_forward(eab_model* m, char* state, const StreamLayout& L, const float* inpt, const long long* strides, int in_RT,
                       long long in_slot, const float* pre, int S, cudaStream_t st) {
    Ctx cx;
    cx.m = m; cx.dry = false; cx.base = state + L.off_act; cx.B = S; cx.T = 1; cx.st = st;
    cx.streaming = true; cx.step = reinterpret_cast<const int*>(state);
    cx.start = reinterpret_cast<const int*>(state + L.off_start);
    cx.gag_in_RT = in_RT; cx.gag_in_slot = in_slot;
    return run_gag_forward(cx, inpt, strides, pre, reinterpret_cast<float*>(state + L.off_out));
}

int gag_forward(eab_model* m, const float* inpt, const long long* strides, const float* pre, float* out, int B, int T, void* ws,
                size_t ws_bytes, cudaStream_t st) {
    if (m->kind != 1) return fail("eab_gag_forward: the handle is not a GaGNet (create it with eab_gag_create)");
    if (B < 1 || T < 1) return fail("forward: B and T must be positive");
    if (m->cfg.norm_type == 0 && T < 2) return fail("InstanceNorm1d needs more than one frame (the reference raises too)");
    if (m->dirty) return fail("parameters not committed: call eab_commit_params first");
    size_t sb = 0, tb = 0;
    EAB_TRY(plan(m, B, T, &sb, &tb));
    if (ws_bytes < tb) return fail("workspace too small: need " + std::to_string(tb) + " bytes");
    if ((reinterpret_cast<uintptr_t>(ws) & 255) != 0) return fail("workspace must be 256-byte aligned");
    m->taps.clear();
    m->norm_log.clear();
    m->umma_launch_idx = 0;
    if (sb) EAB_CUDA(cudaMemsetAsync(ws, 0, sb, st));
    Ctx cx;
    cx.m = m; cx.dry = false; cx.base = static_cast<char*>(ws); cx.B = B; cx.T = T; cx.st = st;
    cx.stats_off = 0; cx.stats_cap = sb; cx.act_off = sb; cx.act_peak = sb;
    return run_gag_forward(cx, inpt, strides, pre, out);
}

int forward(eab_model* m, const float* inpt, float* out, int B, int T, void* ws, size_t ws_bytes, cudaStream_t st) {
    if (m->kind != 0) return fail("this entry point needs an EaBNet handle (eab_create); a GaGNet runs through eab_gag_forward");
    if (B < 1 || T < 1) return fail("forward: B and T must be positive");
    if (m->cfg.norm_type == 0 && T < 2)
        return fail("InstanceNorm1d needs more than one frame (the reference raises too)");
    if (m->dirty) return fail("parameters not committed: call eab_commit_params first");
    size_t sb = 0, tb = 0;
    EAB_TRY(plan(m, B, T, &sb, &tb));
    if (ws_bytes < tb) return fail("workspace too small: need " + std::to_string(tb) + " bytes");
    if ((reinterpret_cast<uintptr_t>(ws) & 255) != 0) return fail("workspace must be 256-byte aligned");
    m->taps.clear();
    m->umma_launch_idx = 0;
    if (sb) EAB_CUDA(cudaMemsetAsync(ws, 0, sb, st));
    Ctx cx;
    cx.m = m; cx.dry = false; cx.base = static_cast<char*>(ws); cx.B = B; cx.T = T; cx.st = st;
    cx.stats_off = 0; cx.stats_cap = sb; cx.act_off = sb; cx.act_peak = sb;
    return run_forward(cx, inpt, out);
}

}  // namespace
}  // namespace eab

// =================================================================================================== C ABI
extern "C" {

size_t eab_workspace_bytes(const eab_model* m, int B, int T);

int eab_create(const eab_config* cfg, eab_model** out) {
    if (!cfg || !out) return fail("eab_create: null argument");
    std::unique_ptr<eab_model> m(new eab_model());
    m->cfg = *cfg;
    if (m->cfg.n_freq <= 0) m->cfg.n_freq = 161;
    if (build(m.get())) return 1;
    // the 1x1-conv heads feed decoder round-off straight into the beam weights (no recurrent smoothing, and "miso"
    // sums 161 bins): keep their decoder fp32-grade as well
    if (!(m->cfg.topo_type == 0 && m->cfg.bf_type == 0)) m->opt_dec_passes = 3;
    *out = m.release();
    return 0;
}

int eab_gag_create(const eab_gag_config* cfg, eab_model** out) {
    if (!cfg || !out) return fail("eab_gag_create: null argument");
    std::unique_ptr<eab_model> m(new eab_model());
    m->kind = 1;
    m->gcfg = *cfg;
    if (build_gag(m.get())) return 1;
    // the post-filter refines an estimate that is already within tolerance: keep its whole encoder fp32-grade
    m->opt_enc_passes = 3;
    m->opt_inner_passes = 3;
    *out = m.release();
    return 0;
}

size_t eab_gag_workspace_bytes(const eab_model* m, int B, int T) {
    if (!m || m->kind != 1) return 0;
    return eab_workspace_bytes(m, B, T);
}

int eab_gag_forward(eab_model* m, const float* inpt, const int64_t inpt_strides[4], const float* pre, float* out, int B, int T,
                    void* ws, size_t ws_bytes, void* stream) {
    if (!m || !inpt || !inpt_strides || !pre || !out || !ws) return fail("eab_gag_forward: null argument");
    const long long st4[4] = {(long long)inpt_strides[0], (long long)inpt_strides[1], (long long)inpt_strides[2], (long long)inpt_strides[3]};
    reset_launch_count();
    const int rc = gag_forward(m, inpt, st4, pre, out, B, T, ws, ws_bytes, static_cast<cudaStream_t>(stream));
    m->last_launches = launch_count();
    return rc;
}

void eab_destroy(eab_model* m) {
    if (!m) return;
    if (m->blob) cudaFree(m->blob);
    if (m->scratch) cudaFree(m->scratch);
    for (auto& sg : m->slot_graph) if (sg.exec) cudaGraphExecDestroy(sg.exec);
    if (m->s_in) {
        cudaStreamDestroy(m->s_in);
        cudaStreamDestroy(m->s_out);
        cudaStreamDestroy(m->s_comp);
        cudaStreamDestroy(m->s_comp2);
        for (int i = 0; i < 2; ++i) { cudaEventDestroy(m->ev_in[i]); cudaEventDestroy(m->ev_comp[i]); cudaEventDestroy(m->ev_out[i]); }
    }
    delete m;
}

int eab_param_count(const eab_model* m) { return m ? (int)m->params.size() : 0; }

int eab_param_info(const eab_model* m, int i, const char** name, int* ndim, int64_t shape[4], int* kind, int* fan_in) {
    if (!m || i < 0 || i >= (int)m->params.size()) return fail("eab_param_info: index out of range");
    const Param& p = m->params[i];
    if (name) *name = p.name.c_str();
    if (ndim) *ndim = p.ndim;
    if (shape) for (int k = 0; k < 4; ++k) shape[k] = p.shape[k];
    if (kind) *kind = p.kind;
    if (fan_in) *fan_in = p.fan_in;
    return 0;
}

int eab_set_param(eab_model* m, const char* name, const float* host, int64_t numel) {
    if (!m || !name) return fail("eab_set_param: null argument");
    auto it = m->index.find(name);
    if (it == m->index.end()) return fail(std::string("unexpected key in state_dict: ") + name);
    Param& p = m->params[it->second];
    if (p.kind == EAB_P_BN_COUNT) { p.set = true; return 0; }
    if (numel != p.numel()) return fail(std::string("size mismatch for ") + name);
    if (!host) return fail("eab_set_param: null data");
    p.host.assign(host, host + numel);
    p.set = true;
    m->dirty = true;
    return 0;
}

int eab_commit_params(eab_model* m, void* stream) {
    if (!m) return fail("null model");
    return commit(m, static_cast<cudaStream_t>(stream));
}

size_t eab_workspace_bytes(const eab_model* m, int B, int T) {
    size_t sb = 0, tb = 0;
    if (!m || B < 1 || T < 1) return 0;
    if (plan(const_cast<eab_model*>(m), B, T, &sb, &tb)) return 0;
    return tb;
}

int eab_forward(eab_model* m, const float* inpt, float* out, int B, int T, void* ws, size_t ws_bytes, void* stream) {
    if (!m || !inpt || !out || !ws) return fail("eab_forward: null argument");
    reset_launch_count();
    const int rc = forward(m, inpt, out, B, T, ws, ws_bytes, static_cast<cudaStream_t>(stream));
    m->last_launches = launch_count();
    return rc;
}

int eab_stft(const float* wave, float* spec, int B, int M, int L, void* stream) {
    if (!wave || !spec) return fail("eab_stft: null argument");
    return launch_stft(wave, spec, B, M, L, static_cast<cudaStream_t>(stream));
}

int eab_istft(const float* spec, float* wave, int B, int T, void* stream) {
    if (!spec || !wave) return fail("eab_istft: null argument");
    return launch_istft(spec, wave, B, T, static_cast<cudaStream_t>(stream));
}

static size_t align256(size_t x) { return (x + 255) / 256 * 256; }

size_t eab_enhance_workspace_bytes(const eab_model* m, int B, int L) {
    if (!m || B < 1 || L < 161) return 0;
    const int T = 1 + L / 160;
    const size_t fw = eab_workspace_bytes(m, B, T);
    if (!fw) return 0;
    const size_t spec = align256((size_t)B * T * m->cfg.n_freq * m->cfg.M * 2 * sizeof(float));
    const size_t outp = align256((size_t)B * 2 * T * m->cfg.n_freq * sizeof(float));
    return spec + outp + fw;
}

// wave -> wave on device buffers.  Input: fp32 [B][M][L], or (pcm != null) int16 PCM in file channel order with microphone
// m = file channel order[m]; output: fp32 samples, or (enhanced16 != null) the reference's int16 writer.
static int enhance_dev(eab_model* m, const float* wave, const int16_t* pcm, const int* order, float* enhanced, int16_t* enhanced16,
                       int B, int L, void* ws, size_t ws_bytes, void* stream) {
    if (!m || (!wave && !pcm) || (!enhanced && !enhanced16) || !ws) return fail("eab_enhance: null argument");
    if (m->kind != 0) return fail("eab_enhance needs an EaBNet handle");
    if (m->cfg.topo_type == 1) return fail("eab_enhance: the 'miso' topology returns [B,2,T], which has no iSTFT");
    if (m->cfg.n_freq != 161) return fail("eab_enhance: the 320-point STFT gives 161 bins");
    if (L < 161) return fail("eab_enhance: need at least 161 samples");
    const size_t need = eab_enhance_workspace_bytes(m, B, L);
    if (!need) return 1;
    if (ws_bytes < need) return fail("workspace too small: need " + std::to_string(need) + " bytes");
    const int T = 1 + L / 160;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    char* p = static_cast<char*>(ws);
    float* spec = reinterpret_cast<float*>(p);
    p += align256((size_t)B * T * 161 * m->cfg.M * 2 * sizeof(float));
    float* outp = reinterpret_cast<float*>(p);
    p += align256((size_t)B * 2 * T * 161 * sizeof(float));
    reset_launch_count();
    // the forward workspace is idle until the STFT has finished: it doubles as the STFT's plane scratch
    const size_t rest = ws_bytes - (size_t)(p - static_cast<char*>(ws));
    int rc = pcm ? launch_stft_pcm16(pcm, order, spec, B, m->cfg.M, L, st, p, rest) : launch_stft(wave, spec, B, m->cfg.M, L, st, p, rest);
    if (!rc) rc = forward(m, spec, outp, B, T, p, rest, st);
    if (!rc) rc = launch_istft(outp, enhanced, B, T, st, enhanced16);
    m->last_launches = launch_count();
    return rc;
}

int eab_enhance(eab_model* m, const float* wave, float* enhanced, int B, int L, void* ws, size_t ws_bytes, void* stream) {
    if (!wave || !enhanced) return fail("eab_enhance: null argument");
    return enhance_dev(m, wave, nullptr, nullptr, enhanced, nullptr, B, L, ws, ws_bytes, stream);
}

// enhance.py:49-62 on device buffers: STFT + compression, EaBNet, GaGNet on (reference microphone, EaBNet estimate), iSTFT
// of the last glance-gaze module's estimate.
size_t eab_enhance_postnet_workspace_bytes(const eab_model* eab, const eab_model* gag, int B, int L) {
    if (!eab || !gag || eab->kind != 0 || gag->kind != 1 || B < 1 || L < 161) return 0;
    const int T = 1 + L / 160;
    const size_t f1 = eab_workspace_bytes(eab, B, T), f2 = eab_workspace_bytes(gag, B, T);
    if (!f1 || !f2) return 0;
    const size_t spec = align256((size_t)B * T * eab->cfg.n_freq * eab->cfg.M * 2 * sizeof(float));
    const size_t est = align256((size_t)B * 2 * T * eab->cfg.n_freq * sizeof(float));
    return spec + est * (1 + gag->cfg.q) + std::max(f1, f2);
}

static int enhance_postnet_dev(eab_model* eab, eab_model* gag, int ref_mic, const float* wave, const int16_t* pcm, const int* order,
                               float* enhanced, int16_t* enhanced16, int B, int L, void* ws, size_t ws_bytes, void* stream) {
    if (!eab || !gag || (!wave && !pcm) || (!enhanced && !enhanced16) || !ws) return fail("eab_enhance_postnet: null argument");
    if (eab->kind != 0 || gag->kind != 1) return fail("eab_enhance_postnet: needs an EaBNet handle and a GaGNet handle");
    if (eab->cfg.topo_type == 1) return fail("eab_enhance_postnet: the 'miso' topology returns [B,2,T]");
    if (eab->cfg.n_freq != 161 || gag->cfg.n_freq != 161) return fail("eab_enhance_postnet: the 320-point STFT gives 161 bins");
    if (ref_mic < 0 || ref_mic >= eab->cfg.M) return fail("eab_enhance_postnet: ref_mic out of range");
    const size_t need = eab_enhance_postnet_workspace_bytes(eab, gag, B, L);
    if (!need) return fail("eab_enhance_postnet: bad shape");
    if (ws_bytes < need) return fail("workspace too small: need " + std::to_string(need) + " bytes");
    const int T = 1 + L / 160, F = 161, M = eab->cfg.M;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    char* p = static_cast<char*>(ws);
    float* spec = reinterpret_cast<float*>(p);
    p += align256((size_t)B * T * F * M * 2 * sizeof(float));
    const size_t est = align256((size_t)B * 2 * T * F * sizeof(float));
    float* est0 = reinterpret_cast<float*>(p);
    p += est;
    float* stages = reinterpret_cast<float*>(p);
    p += est * gag->cfg.q;
    // (the q estimates are written back to back, B*2*T*F floats each; only the last one is read here)
    const size_t rest = ws_bytes - (size_t)(p - static_cast<char*>(ws));
    reset_launch_count();
    // (the forward workspace is idle until the STFT has finished: it doubles as the STFT's plane scratch)
    int rc = pcm ? launch_stft_pcm16(pcm, order, spec, B, M, L, st, p, rest) : launch_stft(wave, spec, B, M, L, st, p, rest);
    if (!rc) rc = forward(eab, spec, est0, B, T, p, rest, st);
    const long long strides[4] = {(long long)T * F * M * 2, 1, (long long)F * M * 2, (long long)M * 2};
    if (!rc) rc = gag_forward(gag, spec + (size_t)ref_mic * 2, strides, est0, stages, B, T, p, rest, st);
    if (!rc) rc = launch_istft(stages + (size_t)(gag->cfg.q - 1) * B * 2 * T * F, enhanced, B, T, st, enhanced16);
    const int n = launch_count();
    eab->last_launches = n;
    gag->last_launches = n;
    return rc;
}

int eab_enhance_postnet(eab_model* eab, eab_model* gag, int ref_mic, const float* wave, float* enhanced, int B, int L, void* ws,
                        size_t ws_bytes, void* stream) {
    if (!wave || !enhanced) return fail("eab_enhance_postnet: null argument");
    return enhance_postnet_dev(eab, gag, ref_mic, wave, nullptr, nullptr, enhanced, nullptr, B, L, ws, ws_bytes, stream);
}

// Host front door.  Batch i+1 is uploaded (copy stream) and batch i-1 downloaded (second copy stream) while batch i
// computes on the caller's stream: two device input slots, two output slots, one workspace.
// pcm: the batches are int16 PCM in file channel order (microphone m = file channel order[m]) and the results int16 PCM
static int host_batches(eab_model* m, const void* const* waves_host, void* const* enhanced_host, bool pcm, const int* order,
                        int n_batches, int B, int L, void* stream) {
    if (!m || !waves_host || !enhanced_host) return fail("eab_enhance_host_batches: null argument");
    if (m->kind != 0) return fail("eab_enhance_host_batches needs an EaBNet handle");
    int ord[64];
    unsigned long long ord_hash = pcm ? 1469598103934665603ull : 0ull;
    if (pcm) {
        if (m->cfg.M > 64) return fail("pcm16: at most 64 microphones");
        for (int i = 0; i < m->cfg.M; ++i) {
            ord[i] = order ? order[i] : i;
            if (ord[i] < 0 || ord[i] >= m->cfg.M) return fail("pcm16: mic_order entries must be in [0, M)");
            ord_hash = (ord_hash ^ (unsigned long long)(ord[i] + 1)) * 1099511628211ull;
        }
    }
    const size_t esz = pcm ? sizeof(int16_t) : sizeof(float);
    if (n_batches < 1) return 0;
    for (int i = 0; i < n_batches; ++i)
        if (!waves_host[i] || !enhanced_host[i]) return fail("eab_enhance_host_batches: null batch pointer");
    const size_t need = eab_enhance_workspace_bytes(m, B, L);
    if (!need) return fail("eab_enhance_host_batches: bad shape");
    const int nslot = n_batches > 1 ? 2 : 1;
    const size_t in_bytes = (size_t)B * m->cfg.M * L * esz;
    const size_t out_bytes = (size_t)B * 160 * (L / 160) * esz;
    const size_t in_b = align256(in_bytes), out_b = align256(out_bytes);
    // two compute streams (even / odd batches, a workspace each): consecutive batches overlap wherever one leaves SMs idle
    // (the LSTM runs on 108 of the 148 SMs for 4 ms of a 22 ms step)
    const bool dual = m->opt_dual_stream && nslot == 2;
    const size_t total = nslot * (in_b + out_b) + (dual ? 2 : 1) * align256(need);
    if (m->scratch_bytes < total) {
        if (m->scratch) cudaFree(m->scratch);
        m->scratch = nullptr;
        m->scratch_bytes = 0;
        EAB_CUDA(cudaMalloc(&m->scratch, total));
        m->scratch_bytes = total;
    }
    if (!m->s_in) {
        EAB_CUDA(cudaStreamCreateWithFlags(&m->s_in, cudaStreamNonBlocking));
        EAB_CUDA(cudaStreamCreateWithFlags(&m->s_out, cudaStreamNonBlocking));
        EAB_CUDA(cudaStreamCreateWithFlags(&m->s_comp, cudaStreamNonBlocking));
        EAB_CUDA(cudaStreamCreateWithFlags(&m->s_comp2, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) {
            EAB_CUDA(cudaEventCreateWithFlags(&m->ev_in[i], cudaEventDisableTiming));
            EAB_CUDA(cudaEventCreateWithFlags(&m->ev_comp[i], cudaEventDisableTiming));
            EAB_CUDA(cudaEventCreateWithFlags(&m->ev_out[i], cudaEventDisableTiming));
        }
    }
    cudaStream_t caller = static_cast<cudaStream_t>(stream);
    // the legacy default stream cannot be captured into a graph: compute on the library's own stream then, ordered after the
    // caller's queued work (the call synchronises before it returns, so the caller's stream semantics are unchanged)
    cudaStream_t st = caller ? caller : m->s_comp;
    if (!caller) {
        EAB_CUDA(cudaEventRecord(m->ev_out[0], caller));
        EAB_CUDA(cudaStreamWaitEvent(st, m->ev_out[0], 0));
        stream = st;
    }
    char* p = static_cast<char*>(m->scratch);
    char* din[2] = {p, p + (nslot - 1) * in_b};
    char* dout[2] = {p + nslot * in_b, p + nslot * in_b + (nslot - 1) * out_b};
    char* ws_slot[2] = {p + nslot * (in_b + out_b), p + nslot * (in_b + out_b) + (dual ? align256(need) : 0)};
    cudaStream_t cs[2] = {st, dual ? m->s_comp2 : st};
    // the copy streams start after everything already queued on the caller's stream (scratch may still be in use)
    EAB_CUDA(cudaEventRecord(m->ev_comp[0], st));
    EAB_CUDA(cudaStreamWaitEvent(m->s_in, m->ev_comp[0], 0));
    if (dual) EAB_CUDA(cudaStreamWaitEvent(m->s_comp2, m->ev_comp[0], 0));
    int launches = 0;
    for (int i = 0; i < n_batches; ++i) {
        const int s = i & 1;
        cudaStream_t cst = cs[s];
        char* ws = ws_slot[s];
        if (i >= 2) EAB_CUDA(cudaStreamWaitEvent(m->s_in, m->ev_comp[s], 0));          // batch i-2 has consumed this input slot
        EAB_CUDA(cudaMemcpyAsync(din[s], waves_host[i], in_bytes, cudaMemcpyHostToDevice, m->s_in));
        EAB_CUDA(cudaEventRecord(m->ev_in[s], m->s_in));
        EAB_CUDA(cudaStreamWaitEvent(cst, m->ev_in[s], 0));
        if (i >= 2) EAB_CUDA(cudaStreamWaitEvent(cst, m->ev_out[s], 0));               // batch i-2 has left this output slot
        // the slot's step: replayed from a CUDA graph once the slot has run it directly (same buffers, same shape, same weights)
        {
            eab_model::SlotGraph& sg = m->slot_graph[s];
            const bool match = sg.exec && sg.in == din[s] && sg.out == dout[s] && sg.ws == ws && sg.B == B && sg.L == L &&
                               sg.version == m->param_version && sg.mode == ord_hash;
            auto step = [&]() {
                return pcm ? enhance_dev(m, nullptr, reinterpret_cast<const int16_t*>(din[s]), ord, nullptr, reinterpret_cast<int16_t*>(dout[s]),
                                         B, L, ws, need, cst)
                           : enhance_dev(m, reinterpret_cast<const float*>(din[s]), nullptr, nullptr, reinterpret_cast<float*>(dout[s]), nullptr,
                                         B, L, ws, need, cst);
            };
            if (match) {
                EAB_CUDA(cudaGraphLaunch(sg.exec, cst));
                m->last_launches = sg.launches;
            } else if (m->opt_host_graph && i >= 2 && !g_prof_on()) {
                if (sg.exec) { cudaGraphExecDestroy(sg.exec); sg.exec = nullptr; }
                cudaGraph_t graph = nullptr;
                bool ok = cudaStreamBeginCapture(cst, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
                int rc = 1;
                if (ok) {
                    rc = step();
                    ok = cudaStreamEndCapture(cst, &graph) == cudaSuccess && rc == 0 && graph != nullptr;
                }
                if (ok) ok = cudaGraphInstantiate(&sg.exec, graph, 0) == cudaSuccess;
                if (graph) cudaGraphDestroy(graph);
                if (ok) {
                    sg.in = din[s]; sg.out = dout[s]; sg.ws = ws; sg.B = B; sg.L = L; sg.version = m->param_version; sg.mode = ord_hash;
                    sg.launches = m->last_launches;
                    EAB_CUDA(cudaGraphLaunch(sg.exec, cst));
                } else {
                    sg.exec = nullptr;
                    cudaGetLastError();                        // capture refused (e.g. a legacy-stream caller): run the step directly
                    EAB_TRY(step());
                }
            } else {
                EAB_TRY(step());
            }
        }
        launches += m->last_launches;
        EAB_CUDA(cudaEventRecord(m->ev_comp[s], cst));
        EAB_CUDA(cudaStreamWaitEvent(m->s_out, m->ev_comp[s], 0));
        EAB_CUDA(cudaMemcpyAsync(enhanced_host[i], dout[s], out_bytes, cudaMemcpyDeviceToHost, m->s_out));
        EAB_CUDA(cudaEventRecord(m->ev_out[s], m->s_out));
    }
    m->last_launches = launches;
    EAB_CUDA(cudaStreamSynchronize(m->s_out));
    EAB_CUDA(cudaStreamSynchronize(st));
    if (dual) EAB_CUDA(cudaStreamSynchronize(m->s_comp2));
    return 0;
}

int eab_enhance_host_batches(eab_model* m, const float* const* waves_host, float* const* enhanced_host, int n_batches,
                             int B, int L, void* stream) {
    return host_batches(m, reinterpret_cast<const void* const*>(waves_host), reinterpret_cast<void* const*>(enhanced_host), false, nullptr,
                        n_batches, B, L, stream);
}

// The same pipeline on the 16-bit PCM wire format (half the H2D / D2H bytes): the int16 -> float conversion and the microphone
// permutation happen in the STFT's operand staging, the int16 writer in the iSTFT's store.
int eab_enhance_host_batches_pcm16(eab_model* m, const int16_t* const* pcm_host, const int* mic_order, int16_t* const* enhanced_pcm_host,
                                   int n_batches, int B, int L, void* stream) {
    return host_batches(m, reinterpret_cast<const void* const*>(pcm_host), reinterpret_cast<void* const*>(enhanced_pcm_host), true, mic_order,
                        n_batches, B, L, stream);
}

// enhance.py:35-43 + the int16 writer of the dataset tools, end to end on HOST 16-bit PCM buffers: H2D of the PCM (half the
// bytes of the fp32 front door), int16 -> float / 32768 with the microphone permutation, EaBNet (+ GaGNet), float -> int16, D2H.
int eab_enhance_host_pcm16(eab_model* m, eab_model* gag, int ref_mic, const int16_t* pcm_host, const int* mic_order,
                           int16_t* enhanced_pcm_host, int B, int L, void* stream) {
    if (!m || !pcm_host || !enhanced_pcm_host) return fail("eab_enhance_host_pcm16: null argument");
    if (m->kind != 0 || (gag && gag->kind != 1)) return fail("eab_enhance_host_pcm16: needs an EaBNet handle (and optionally a GaGNet handle)");
    if (!gag) return eab_enhance_host_batches_pcm16(m, &pcm_host, mic_order, &enhanced_pcm_host, 1, B, L, stream);
    const int M = m->cfg.M;
    if (M > 64) return fail("eab_enhance_host_pcm16: at most 64 microphones");
    int order[64];
    for (int i = 0; i < M; ++i) {
        order[i] = mic_order ? mic_order[i] : i;
        if (order[i] < 0 || order[i] >= M) return fail("eab_enhance_host_pcm16: mic_order entries must be in [0, M)");
    }
    const size_t need = eab_enhance_postnet_workspace_bytes(m, gag, B, L);
    if (!need) return fail("eab_enhance_host_pcm16: bad shape");
    const size_t n_in = (size_t)B * M * L, n_out = (size_t)B * 160 * (L / 160);
    const size_t o_pcm = 0, o_epcm = align256(n_in * 2), o_ws = o_epcm + align256(n_out * 2), total = o_ws + need;
    if (m->scratch_bytes < total) {
        if (m->scratch) cudaFree(m->scratch);
        m->scratch = nullptr;
        m->scratch_bytes = 0;
        EAB_CUDA(cudaMalloc(&m->scratch, total));
        m->scratch_bytes = total;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    char* p = static_cast<char*>(m->scratch);
    int16_t* d_pcm = reinterpret_cast<int16_t*>(p + o_pcm);
    int16_t* d_epcm = reinterpret_cast<int16_t*>(p + o_epcm);
    EAB_CUDA(cudaMemcpyAsync(d_pcm, pcm_host, n_in * 2, cudaMemcpyHostToDevice, st));
    EAB_TRY(enhance_postnet_dev(m, gag, ref_mic, nullptr, d_pcm, order, nullptr, d_epcm, B, L, p + o_ws, need, stream));
    EAB_CUDA(cudaMemcpyAsync(enhanced_pcm_host, d_epcm, n_out * 2, cudaMemcpyDeviceToHost, st));
    EAB_CUDA(cudaStreamSynchronize(st));
    return 0;
}

int eab_enhance_host(eab_model* m, const float* wave_host, float* enhanced_host, int B, int L, void* stream) {
    if (!m || !wave_host || !enhanced_host) return fail("eab_enhance_host: null argument");
    return eab_enhance_host_batches(m, &wave_host, &enhanced_host, 1, B, L, stream);
}

size_t eab_stream_state_bytes(const eab_model* m, int n_streams) {
    StreamLayout L;
    if (!m || stream_layout(const_cast<eab_model*>(m), n_streams, &L)) return 0;
    return L.total;
}

static int stream_check(eab_model* m, void* state, size_t state_bytes, int S, StreamLayout* L) {
    if (!m || !state) return fail("eab_stream: null argument");
    EAB_TRY(stream_layout(m, S, L));
    if (state_bytes < L->total) return fail("stream state too small: need " + std::to_string(L->total) + " bytes");
    if ((reinterpret_cast<uintptr_t>(state) & 255) != 0) return fail("stream state must be 256-byte aligned");
    return 0;
}

int eab_stream_reset(eab_model* m, void* state, size_t state_bytes, int n_streams, void* stream) {
    StreamLayout L;
    EAB_TRY(stream_check(m, state, state_bytes, n_streams, &L));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    EAB_CUDA(cudaMemsetAsync(state, 0, L.total, st));
    if (!L.descs.empty()) {
        EAB_CUDA(cudaMemcpyAsync(static_cast<char*>(state) + L.off_desc, L.descs.data(), L.descs.size() * sizeof(TcmStreamDesc),
                                 cudaMemcpyHostToDevice, st));
        EAB_CUDA(cudaStreamSynchronize(st));       // the descriptor vector dies with this scope
    }
    return 0;
}

int eab_stream_step_spec(eab_model* m, void* state, size_t state_bytes, const float* frame, float* out_frame, int n_streams,
                         void* stream) {
    StreamLayout L;
    if (m && m->kind != 0) return fail("eab_stream_step_spec needs an EaBNet handle (a GaGNet steps through eab_gag_stream_step_spec)");
    EAB_TRY(stream_check(m, state, state_bytes, n_streams, &L));
    if (!frame || !out_frame) return fail("eab_stream_step_spec: null argument");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    char* p = static_cast<char*>(state);
    const eab_config& c = m->cfg;
    reset_launch_count();
    {   // new spectrum frame -> slot (*step % 2) of the spectrum ring
        CombineArgs a;
        memset(&a, 0, sizeof(a));
        a.nsrc = 1;
        a.src[0].x = frame; a.src[0].C = 2 * c.M; a.src[0].xf = xform_identity(); a.src[0].RT = 1;
        a.B = n_streams; a.P = c.n_freq; a.C = 2 * c.M;
        a.out = reinterpret_cast<float*>(p + L.off_spec); a.out_RT = 2;
        a.step = reinterpret_cast<const int*>(p);
        EAB_TRY(launch_combine(a, st));
    }
    EAB_TRY(stream_forward(m, p, L, n_streams, st));
    const size_t out_bytes = (size_t)n_streams * (c.topo_type == 1 ? 2 : 2 * c.n_freq) * sizeof(float);
    EAB_CUDA(cudaMemcpyAsync(out_frame, p + L.off_out, out_bytes, cudaMemcpyDeviceToDevice, st));
    EAB_TRY(launch_step_advance(reinterpret_cast<int*>(p), st));
    m->last_launches = launch_count();
    return 0;
}

static int stream_step_any(eab_model* m, void* state, size_t state_bytes, eab_model* gag, void* gstate, size_t gstate_bytes, int ref_mic,
                           const float* hop, const int16_t* hop16, float* enhanced_hop, int16_t* enhanced16, int n_streams, void* stream) {
    StreamLayout L, GL;
    if (m && m->kind != 0) return fail("eab_stream_step needs an EaBNet handle");
    EAB_TRY(stream_check(m, state, state_bytes, n_streams, &L));
    if ((!hop && !hop16) || (!enhanced_hop && !enhanced16)) return fail("eab_stream_step: null argument");
    const eab_config& c = m->cfg;
    if (c.topo_type == 1) return fail("eab_stream_step: the 'miso' topology returns [B,2,T], which has no iSTFT");
    if (c.n_freq != 161) return fail("eab_stream_step: the 320-point STFT gives 161 bins");
    if (gag) {
        if (gag->kind != 1) return fail("eab_stream_step_postnet: the post-filter handle is not a GaGNet");
        if (gag->cfg.n_freq != c.n_freq) return fail("eab_stream_step_postnet: EaBNet and GaGNet disagree on the number of bins");
        if (ref_mic < 0 || ref_mic >= c.M) return fail("eab_stream_step_postnet: reference microphone out of range");
        EAB_TRY(stream_check(gag, gstate, gstate_bytes, n_streams, &GL));
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    char* p = static_cast<char*>(state);
    int* step = reinterpret_cast<int*>(p);
    const int* start = reinterpret_cast<const int*>(p + L.off_start);
    reset_launch_count();
    EAB_TRY(launch_stft_frame(hop, hop16, reinterpret_cast<float*>(p + L.off_prev), reinterpret_cast<float*>(p + L.off_spec), 2, step,
                              start, n_streams, c.M, st));
    EAB_TRY(stream_forward(m, p, L, n_streams, st));
    const float* est = reinterpret_cast<const float*>(p + L.off_out);
    if (gag) {
        // enhance.py:49-62 frame by frame: GaGNet on (reference microphone of the compressed spectrum - read in place from the
        // spectrum ring's current slot -, EaBNet's estimate); the last module's estimate goes to the iSTFT
        char* g = static_cast<char*>(gstate);
        const long long FM2 = (long long)c.n_freq * c.M * 2;
        const long long strides[4] = {2 * FM2, 1, 0, (long long)c.M * 2};
        EAB_TRY(gag_stream_forward(gag, g, GL, reinterpret_cast<const float*>(p + L.off_spec) + ref_mic * 2, strides, 2, FM2, est, n_streams, st));
        est = reinterpret_cast<const float*>(g + GL.off_out) + (gag->gags.size() - 1) * (size_t)n_streams * 2 * c.n_freq;
        EAB_TRY(launch_step_advance(reinterpret_cast<int*>(g), st));
    }
    EAB_TRY(launch_istft_frame(est, reinterpret_cast<float*>(p + L.off_tail), enhanced_hop, enhanced16, step, start, n_streams, st));
    EAB_TRY(launch_step_advance(step, st));
    m->last_launches = launch_count();
    return 0;
}

int eab_stream_step(eab_model* m, void* state, size_t state_bytes, const float* hop, float* enhanced_hop, int n_streams,
                    void* stream) {
    if (!hop || !enhanced_hop) return fail("eab_stream_step: null argument");
    return stream_step_any(m, state, state_bytes, nullptr, nullptr, 0, 0, hop, nullptr, enhanced_hop, nullptr, n_streams, stream);
}

// the same step on the 16-bit PCM wire format: hop [S][M][160] int16 (sample / 32768), enhanced hop [S][160] int16
int eab_stream_step_pcm16(eab_model* m, void* state, size_t state_bytes, const int16_t* hop, int16_t* enhanced_hop, int n_streams,
                          void* stream) {
    if (!hop || !enhanced_hop) return fail("eab_stream_step_pcm16: null argument");
    return stream_step_any(m, state, state_bytes, nullptr, nullptr, 0, 0, nullptr, hop, nullptr, enhanced_hop, n_streams, stream);
}

// EaBNet + GaGNet post-filter, one hop per stream (enhance.py:49-62 as a causal stream).  Both states carry their own frame
// counter: reset them together (eab_stream_reset / eab_stream_reset_one on each).
int eab_stream_step_postnet(eab_model* eabnet, void* state, size_t state_bytes, eab_model* gagnet, void* gag_state, size_t gag_state_bytes,
                            int ref_mic, const float* hop, float* enhanced_hop, int n_streams, void* stream) {
    if (!hop || !enhanced_hop || !gagnet) return fail("eab_stream_step_postnet: null argument");
    return stream_step_any(eabnet, state, state_bytes, gagnet, gag_state, gag_state_bytes, ref_mic, hop, nullptr, enhanced_hop, nullptr,
                           n_streams, stream);
}
int eab_stream_step_postnet_pcm16(eab_model* eabnet, void* state, size_t state_bytes, eab_model* gagnet, void* gag_state,
                                  size_t gag_state_bytes, int ref_mic, const int16_t* hop, int16_t* enhanced_hop, int n_streams, void* stream) {
    if (!hop || !enhanced_hop || !gagnet) return fail("eab_stream_step_postnet_pcm16: null argument");
    return stream_step_any(eabnet, state, state_bytes, gagnet, gag_state, gag_state_bytes, ref_mic, nullptr, hop, nullptr, enhanced_hop,
                           n_streams, stream);
}

// GaGNet.forward one frame at a time: inpt_frame / pre_frame [S][2][F] -> the q modules' estimates [q][S][2][F]
int eab_gag_stream_step_spec(eab_model* gagnet, void* gag_state, size_t gag_state_bytes, const float* inpt_frame, const float* pre_frame,
                             float* out_frames, int n_streams, void* stream) {
    if (!gagnet || gagnet->kind != 1) return fail("eab_gag_stream_step_spec: the handle is not a GaGNet");
    if (!inpt_frame || !pre_frame || !out_frames) return fail("eab_gag_stream_step_spec: null argument");
    StreamLayout GL;
    EAB_TRY(stream_check(gagnet, gag_state, gag_state_bytes, n_streams, &GL));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    char* g = static_cast<char*>(gag_state);
    const int F = gagnet->cfg.n_freq;
    const long long strides[4] = {2LL * F, (long long)F, 0, 1};
    reset_launch_count();
    EAB_TRY(gag_stream_forward(gagnet, g, GL, inpt_frame, strides, 1, 0, pre_frame, n_streams, st));
    EAB_CUDA(cudaMemcpyAsync(out_frames, g + GL.off_out, gagnet->gags.size() * (size_t)n_streams * 2 * F * sizeof(float),
                             cudaMemcpyDeviceToDevice, st));
    EAB_TRY(launch_step_advance(reinterpret_cast<int*>(g), st));
    gagnet->last_launches = launch_count();
    return 0;
}

// One stream leaves and a new one joins in its slot: from the next step on stream `idx` starts over (its frame 0), the other
// streams carry on untouched.  Its history rings need no clearing: every kernel treats frames before the stream's start index
// as the literal zeros of the causal padding; the carried LSTM state is zeroed.  Stream-ordered, no synchronisation.
int eab_stream_reset_one(eab_model* m, void* state, size_t state_bytes, int n_streams, int idx, void* stream) {
    StreamLayout L;
    EAB_TRY(stream_check(m, state, state_bytes, n_streams, &L));
    if (idx < 0 || idx >= n_streams) return fail("eab_stream_reset_one: stream index out of range");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    char* p = static_cast<char*>(state);
    EAB_TRY(launch_stream_restart(reinterpret_cast<int*>(p + L.off_start), reinterpret_cast<const int*>(p), idx, st));
    for (const auto& ps : L.per_stream)
        EAB_CUDA(cudaMemsetAsync(p + L.off_act + ps.first + (size_t)idx * ps.second, 0, ps.second, st));
    return 0;
}

int eab_last_launch_count(const eab_model* m) { return m ? m->last_launches : 0; }

// InstanceNorm statistics of the last eab_forward (option "norm_log" = 1 before it): the i-th normalisation layer the
// forward applied, in execution order.  sums_host [C][2] = (sum, sum of squares) of the layer's input over `count`
// positions (all batch items pooled).  Synchronises the stream; the workspace of that forward must still be intact.
int eab_norm_stats_count(const eab_model* m) { return m ? (int)m->norm_log.size() : 0; }
int eab_norm_stats(eab_model* m, int i, const char** weight_name, int* C, int64_t* count, double* sums_host, void* stream) {
    if (!m) return fail("eab_norm_stats: null handle");
    if (i < 0 || i >= (int)m->norm_log.size()) return fail("eab_norm_stats: index out of range (set option norm_log = 1 before the forward)");
    const auto& r = m->norm_log[i];
    if (weight_name) *weight_name = m->params[r.gamma].name.c_str();
    if (C) *C = r.C;
    if (count) *count = (int64_t)r.count * r.B;
    if (sums_host) {
        std::vector<double> h((size_t)r.B * r.C * 2);
        EAB_CUDA(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
        EAB_CUDA(cudaMemcpy(h.data(), r.stats, h.size() * sizeof(double), cudaMemcpyDeviceToHost));
        for (int c = 0; c < r.C * 2; ++c) sums_host[c] = 0.0;
        for (int b = 0; b < r.B; ++b)
            for (int c = 0; c < r.C * 2; ++c) sums_host[c] += h[(size_t)b * r.C * 2 + c];
    }
    return 0;
}

int64_t eab_debug_tap(eab_model* m, const char* name, float* dst, int64_t capacity, void* stream) {
    if (!m || !name || !dst) { fail("eab_debug_tap: null argument"); return -1; }
    auto it = m->taps.find(name);
    if (it == m->taps.end()) { fail(std::string("no such tap: ") + name); return -1; }
    const Tap& t = it->second;
    const int64_t n = (int64_t)t.B * t.T * t.act.F * t.act.C;
    if (n > capacity) { fail("eab_debug_tap: destination too small"); return -1; }
    CombineArgs a;
    memset(&a, 0, sizeof(a));
    a.nsrc = 1;
    a.src[0].x = t.act.data; a.src[0].C = t.act.C; a.src[0].xf = t.act.xf;
    if (t.act.data2) { a.nsrc = 2; a.src[1].x = t.act.data2; a.src[1].C = t.act.C; a.src[1].xf = t.act.xf2; }
    a.B = t.B; a.P = t.T * t.act.F; a.C = t.act.C; a.out = dst;
    if (launch_combine(a, static_cast<cudaStream_t>(stream))) return -1;
    return n;
}

int eab_set_option(eab_model* m, const char* name, int value) {
    if (!m || !name) return fail("eab_set_option: null argument");
    const std::string n(name);
    if (n == "umma") m->opt_umma = value != 0;
    else if (n == "staged") m->opt_staged = value != 0;
    else if (n == "raw") m->opt_raw = value != 0;
    else if (n == "raw_grid") m->opt_raw_grid = value;
    else if (n == "lazy") m->opt_lazy = value != 0;
    else if (n == "norm_log") { m->opt_norm_log = value != 0; m->norm_log.clear(); }
    else if (n == "tcm_chain") m->opt_tcm_chain = value;
    else if (n == "host_graph") m->opt_host_graph = value != 0;
    else if (n == "dual_stream") m->opt_dual_stream = value != 0;
    else if (n == "stream_tcm") m->opt_stream_tcm = value != 0;
    else if (n == "stream_umma") m->opt_stream_umma = value != 0;
    else if (n == "lstm_exp") m->opt_lstm_exp = value;
    else if (n == "stft_tc") g_stft_tc = value != 0;
    else if (n == "fused_head") m->opt_fused_head = value != 0;
    else if (n == "head_w_tap") m->opt_head_w_tap = value != 0;
    else if (n == "enc_passes" && (value == 1 || value == 3)) { m->opt_enc_passes = value; m->opt_inner_passes = value; }
    else if (n == "dec_passes" && (value == 1 || value == 3)) m->opt_dec_passes = value;
    else if (n == "inner_passes" && (value == 1 || value == 3)) m->opt_inner_passes = value;
    else if (n == "first_passes" && (value == 1 || value == 3)) m->opt_first_passes = value;
    else if (n == "dbg_launch") {
        m->opt_dbg_launch = value;
        if (!m->dbg_buf) { if (check_cuda(cudaMalloc(&m->dbg_buf, 16 * sizeof(unsigned long long)), "dbg alloc")) return 1; }
        cudaMemset(m->dbg_buf, 0, 16 * sizeof(unsigned long long));
    }
    else return fail("eab_set_option: unknown option or bad value: " + n);
    for (auto& sg : m->slot_graph)               // graphs captured under the old options must not be replayed
        if (sg.exec) { cudaGraphExecDestroy(sg.exec); sg.exec = nullptr; }
    return 0;
}

int eab_debug_counters(eab_model* m, unsigned long long* out16) {
    if (!m || !out16 || !m->dbg_buf) return fail("eab_debug_counters: not enabled");
    EAB_CUDA(cudaDeviceSynchronize());
    EAB_CUDA(cudaMemcpy(out16, m->dbg_buf, 16 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    return 0;
}

int eab_profile_enable(eab_model* m, int on) {
    (void)m;
    g_prof.clear();
    g_prof.on = on != 0;
    g_prof.detail = on == 2;          // 2: one summary entry per launch instead of per kernel family
    return 0;
}

int64_t eab_profile_summary(eab_model* m, char* buf, int64_t cap) {
    (void)m;
    if (!buf || cap < 2) { fail("eab_profile_summary: bad buffer"); return -1; }
    if (check_cuda(cudaDeviceSynchronize(), "profile sync")) return -1;
    struct Agg { int n = 0; double ms = 0, flops = 0, bytes = 0, moved = 0; };
    std::map<std::string, Agg> agg;
    int seq = 0;
    for (auto& r : g_prof.recs) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, r.a, r.b) != cudaSuccess) ms = 0.f;
        char key[64];
        if (g_prof.detail) snprintf(key, sizeof(key), "%04d:%s", seq++, r.cat); else snprintf(key, sizeof(key), "%s", r.cat);
        Agg& a = agg[key];
        a.n += 1; a.ms += ms; a.flops += r.flops; a.bytes += r.bytes; a.moved += r.moved;
    }
    std::string js = "[";
    bool first = true;
    for (auto& kv : agg) {
        char tmp[256];
        snprintf(tmp, sizeof(tmp), "%s{\"kernel\":\"%s\",\"launches\":%d,\"ms\":%.6f,\"flops\":%.6e,\"bytes\":%.6e,\"moved_bytes\":%.6e}",
                 first ? "" : ",", kv.first.c_str(), kv.second.n, kv.second.ms, kv.second.flops, kv.second.bytes, kv.second.moved);
        js += tmp;
        first = false;
    }
    js += "]";
    g_prof.clear();
    if ((int64_t)js.size() + 1 > cap) { fail("eab_profile_summary: buffer too small"); return -1; }
    memcpy(buf, js.c_str(), js.size() + 1);
    return (int64_t)js.size();
}

const char* eab_last_error(void) { return g_err.c_str(); }

const char* eab_build_info(void) {
#define EAB_STR2(x) #x
#define EAB_STR(x) EAB_STR2(x)
    return "sm_100a;" __DATE__ ";nvcc " EAB_STR(__CUDACC_VER_MAJOR__) "." EAB_STR(__CUDACC_VER_MINOR__);
}

}  // extern "C"
