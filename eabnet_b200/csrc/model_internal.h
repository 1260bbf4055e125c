// Private header of the model translation units (model.cu: runtime + C ABI; model_build.cu: architecture walk / parameter
// table; model_pack.cu: weight packing; model_run.cu: forward orchestration, planning, streaming).  Not part of the C ABI.
#pragma once
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
namespace detail {


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
    int p_stack = 1;                     // kt when the time taps are stacked along K (kt * 2 cin <= 64): taps = frequency shifts only
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


// element (n, k) of a [N][64] fp16 K-major tile with the 128-byte swizzle the tensor core expects (index in halves)
inline size_t sw128_index_h(int n, int k) { return (size_t)n * 64 + (size_t)((((k >> 3) ^ (n & 7)) << 3) | (k & 7)); }

inline int conv_out_f(int Fin, int kf) { return Fin < kf ? -1 : (Fin - kf) / 2 + 1; }
inline int deconv_out_f(int Fin, int kf) { return 2 * (Fin - 1) + kf; }

}  // namespace detail
}  // namespace eab

using namespace eab;
using namespace eab::detail;

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
    size_t off_rnn_step[2] = {0, 0}, off_rnn_step_b[2] = {0, 0};      // streaming gate GEMM: dense [x | h][256] + bias
    UmmaW u_rnn_step[2];

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
    int opt_lstm_pp = 0;          // LSTM layers on lstm_pp.cu (two interleaved sub-batches of 64 per CTA: measured slower, tensor-pipe bound); 0 = lstm_umma.cu
    int opt_stream_fuse = 1;      // streaming: a U-Net module's residual sum in the last inner deconv's epilogue (no combine launch); changes the state layout
    int opt_stream_pair = 1;      // streaming: both output parities of a transposed conv as one launch (conv_umma second variant)
    int opt_stream_lstm = 1;      // streaming: the LSTM step as a tensor-core gate GEMM + cell kernel (needs stream_umma); 0 = CUDA-core lstm kernel
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
namespace detail {

// model_build.cu
int build(eab_model* m);
int build_gag(eab_model* m);
// model_pack.cu
int commit(eab_model* m, cudaStream_t st);
// model_run.cu
int plan(eab_model* m, int B, int T, size_t* stats_bytes, size_t* total_bytes);
int forward(eab_model* m, const float* inpt, float* out, int B, int T, void* ws, size_t ws_bytes, cudaStream_t st);
int gag_forward(eab_model* m, const float* inpt, const long long* strides, const float* pre, float* out, int B, int T, void* ws,
                size_t ws_bytes, cudaStream_t st);
// State blob layout of a streaming session (device, caller-owned): [0,256) absolute frame counter | per-stream start frames |
// TCM descriptors | carried hop [S][M][160] | iSTFT tail [S][160] | spectrum ring [S][2][F][M][2] | output frame(s) |
// activation rings + LSTM state (run_forward order)
struct StreamLayout {
    size_t off_start, off_desc, off_prev, off_tail, off_spec, off_out, off_act, total;
    std::vector<TcmStreamDesc> descs;
    std::vector<std::pair<size_t, size_t>> per_stream;      // (offset from off_act, bytes per stream) of the carried LSTM (h, c)
};
inline size_t up256(size_t x) { return (x + 255) / 256 * 256; }
int stream_layout(eab_model* m, int S, StreamLayout* L);
int stream_forward(eab_model* m, char* state, const StreamLayout& L, int S, cudaStream_t st);
int gag_stream_forward(eab_model* m, char* state, const StreamLayout& L, const float* inpt, const long long* strides, int in_RT,
                       long long in_slot, const float* pre, int S, cudaStream_t st);

}  // namespace detail
}  // namespace eab
