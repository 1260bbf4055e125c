// Shared device/host helpers for the eabnet_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>

namespace eab {

// ---------------------------------------------------------------------------------------------------
// error plumbing: every launcher returns 0/1 and leaves a message for eab_last_error()
// ---------------------------------------------------------------------------------------------------
void set_error(const std::string& msg);
int  fail(const std::string& msg);                       // sets the message, returns 1
int  check_cuda(cudaError_t e, const char* what);        // 0 if ok
// per-device launch state, safe for several devices / host threads in one process: SM count, and the largest dynamic
// shared-memory size a kernel has been configured for on the CURRENT device (cudaFuncSetAttribute is per device)
int device_sm_count(int* sms);
int ensure_dynamic_smem(const void* kernel, int bytes);
void count_launch(int n = 1);                            // per-thread launch counter
int  launch_count();
void reset_launch_count();

// Optional per-launch timing (CUDA events on the launching stream), switched on by eab_profile_enable; used by
// bench.py for the live roofline.  `flops` / `bytes` are ALGORITHMIC figures of the launch in SURVEY.md section 8(d)'s
// sense (a fused layer reads each input tensor once and writes its output once, fp32; a lazy residual pair counts as one
// tensor; weights and design-specific copies are not counted); `moved` is what THIS design asks the memory system for
// (both addends of a lazy sum, staged fp16 planes written and read back, weights) - default: the same figure.
struct ProfScope {
    ProfScope(const char* category, double flops, double bytes, cudaStream_t st, double moved = -1.0);
    ~ProfScope();
    void* rec;
    cudaStream_t st;
};

#define EAB_CUDA(x)                                                   \
    do {                                                              \
        if (eab::check_cuda((x), #x)) return 1;                       \
    } while (0)
#define EAB_TRY(x)                                                    \
    do {                                                              \
        if ((x) != 0) return 1;                                       \
    } while (0)
#define EAB_LAUNCH_CHECK(name)                                        \
    do {                                                              \
        eab::count_launch();                                          \
        if (eab::check_cuda(cudaGetLastError(), name)) return 1;      \
    } while (0)

// ---------------------------------------------------------------------------------------------------
// On-load activation transform.  Every activation tensor is stored RAW (exactly what its producing
// conv wrote) next to per-(b,c) double sums {sum, sumsq}; whoever reads it applies
//     v = x ; [PReLU if prelu==1] ; v = v*s + h ; [PReLU if prelu==2]
// with (s,h) from the statistics (InstanceNorm: biased variance, eps inside the sqrt, affine;
// EaBNet.py:684-686), from precomputed per-channel arrays (BatchNorm eval, EaBNet.py:678-681), or identity.
// 2-D blocks are conv -> norm -> PReLU (prelu==2); TCM branches are PReLU -> norm (prelu==1).
// ---------------------------------------------------------------------------------------------------
struct Xform {
    const double* stats;   // affine==1: [B][C][2] running sums of the (pre-PReLU'd if prelu==1) values
    const float* scale;    // affine==1: gamma[C];  affine==2: precomputed scale[C]
    const float* shift;    // affine==1: beta[C];   affine==2: precomputed shift[C]
    const float* alpha;    // PReLU slopes [C] (prelu != 0)
    float inv_count;       // 1 / (#elements per (b,c)) for affine==1
    int affine;            // 0 none, 1 instance statistics, 2 precomputed
    int prelu;             // 0 none, 1 before the affine, 2 after it
    int alpha01;           // all slopes in [0, 1] (checked on the host at commit): PReLU(z) == max(z, alpha z), bit for bit
};

static inline Xform xform_identity() {
    Xform x;
    x.stats = nullptr; x.scale = nullptr; x.shift = nullptr; x.alpha = nullptr;
    x.inv_count = 0.f; x.affine = 0; x.prelu = 0; x.alpha01 = 0;
    return x;
}

// Plain launch helper.  (Programmatic dependent launch was measured in round 1: +6 % step time - dependents crowd the
// multi-wave kernels - and removed.)
#ifdef __CUDACC__
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
#endif

#ifdef __CUDACC__
// Programmatic dependent launch, for the streaming step only (chains of ~90 launches of 5-20 us: the next kernel's launch latency
// and prologue - barrier init, TMEM allocation, static weights - run under the tail of its predecessor).  The kernel must
// execute pdl_wait() before it reads anything an earlier kernel of the stream wrote, and may call pdl_trigger() at any point.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#endif

#ifdef __CUDACC__
// coefficients for channel c of batch b (called once per CTA per channel, not per element)
__device__ __forceinline__ void xform_coeffs(const Xform& xf, int b, int C, int c, float& s, float& h, float& a) {
    s = 1.f; h = 0.f; a = 1.f;
    if (xf.affine == 1) {
        const double* st = xf.stats + ((size_t)b * C + c) * 2;
        double mean = st[0] * (double)xf.inv_count;
        double var = st[1] * (double)xf.inv_count - mean * mean;
        if (var < 0.0) var = 0.0;
        double rstd = rsqrt(var + 1e-5);
        double g = (double)xf.scale[c];
        s = (float)(g * rstd);
        h = (float)((double)xf.shift[c] - mean * g * rstd);
    } else if (xf.affine == 2) {
        s = xf.scale[c];
        h = xf.shift[c];
    }
    if (xf.prelu) a = xf.alpha[c];
}

__device__ __forceinline__ float prelu_f(float v, float a) { return v > 0.f ? v : a * v; }

__device__ __forceinline__ float xform_apply(float x, float s, float h, float a, int prelu) {
    if (prelu == 1) x = prelu_f(x, a);
    x = fmaf(x, s, h);
    if (prelu == 2) x = prelu_f(x, a);
    return x;
}

// accurate-enough logistic / tanh on the SFU path (abs error ~1e-7; ex2.approx + rcp)
__device__ __forceinline__ float sigmoid_f(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float tanh_f(float x) {
    // 1 - 2/(1+e^{2x}); saturates cleanly for |x| large (e^{2x} -> inf or 0)
    return 1.f - __fdividef(2.f, 1.f + __expf(2.f * x));
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src, bool valid) {
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    int n = valid ? 16 : 0;                                  // src-size 0 => the 16 bytes are zero-filled
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(gmem_src), "r"(n));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }
#endif

// ---------------------------------------------------------------------------------------------------
// launch descriptors shared between model.cu and the kernel translation units
// ---------------------------------------------------------------------------------------------------
constexpr int kMaxTaps = 16;
constexpr int kMaxConvUnits = 96;       // taps x 64-channel slabs x passes of one conv_tma tile
constexpr int kMaxCin = 1024;

struct ConvSrc {
    const float* x;      // [B][T][Fin][C] raw
    int C;
    Xform xf;
    // optional second addend: value = xf(x) + xf2(x2)  (the residual sum that closes every En_unet_module,
    // EaBNet.py:386).  Honoured by stage_kernel and combine_kernel only; every other kernel requires x2 == null.
    const float* x2;
    Xform xf2;
    int RT;              // streaming only: frames in this tensor's ring (see StreamPos); 0 = offline [B][T][...]
};

// Frame-by-frame ("streaming", BASELINE configs[2]) addressing.  Offline, a tensor holds frames 0..T-1 of every batch
// item.  In a streaming launch (step != null) exactly ONE new frame is computed per stream (T == 1); its absolute index
// n = *step is read on the device (so a whole step is CUDA-graph capturable), every activation is a ring of RT frames
// per stream with frame n in slot n % RT, and a causal tap that reaches d frames back reads slot (n - d) % RT, or
// zeros when n - d < 0 (the literal zero padding of the offline path: EaBNet.py:449, :480, :557).
#ifdef __CUDACC__
__device__ __forceinline__ int ring_slot(int n, int RT) { return RT <= 1 ? 0 : n % RT; }
#endif

// One "virtual" stride-1-output convolution over rows (t, e):  fi = e*in_stride + df[tap],
// frame = t - dt[tap], fo = e*out_stride + out_off.  A Conv2d with stride (1,2) is one launch
// (in_stride 2, out_stride 1); a ConvTranspose2d with stride (1,2) is two launches, one per output parity
// (in_stride 1, out_stride 2, out_off = parity, df = -i) -- no zero stuffing.
struct ConvArgs {
    ConvSrc src[2];
    int nsrc;
    int B, T, Fin;
    int E;                       // rows per frame in this launch
    int in_stride, out_stride, out_off, Fout;
    int ntaps;
    int dt[kMaxTaps], df[kMaxTaps];
    const float* W;              // packed [ntaps][Cin_total][N]
    const float* bias;           // [N] or null
    int Cout, N, gate_off;       // N = padded column count (multiple of 64); gate_off>0 => gated
    int relu;                    // ReLU on the output (w_dnn hidden layer)
    float algo_frac;             // share of the dense K x N product that is algorithmic work (profiling only)
    float* out;                  // [B][T][Fout][Cout]
    const float* resid;          // optional, same layout as out (TCM residual)
    double* stats[2];            // optional [B][Cout][2] accumulators
    const float* stat_alpha[2];  // if set, statistics are taken of PReLU(out, alpha) (TCM convention)
    int nstats;
    const int* step;             // streaming: device pointer to the absolute frame index (null = offline)
    const int* start;            // streaming: [streams] absolute frame index at which each stream (re)started: frames before it
                                 // read as the literal zeros of the causal padding (eab_stream_reset_one); null = all zero
    int out_RT, resid_RT;        // streaming: ring sizes of out / resid
};
int launch_conv(const ConvArgs& a, cudaStream_t st);

// tcgen05 implicit-GEMM variant of ConvArgs (conv_umma.cu).  Weights come as 128B-swizzled K-major images
// [ntaps][nslab][N rows][32 k] already rounded to TF32 (hi) plus the rounding residual (lo) for 3xTF32.
struct UmmaConvArgs {
    ConvSrc src[2];
    int nsrc;
    int B, T, Fin, E;
    int in_stride, out_stride, out_off, Fout;
    int ntaps;
    int dt[kMaxTaps], df[kMaxTaps];
    int wide, kwidth;            // first-layer mode: a tap is a window of kwidth contiguous floats (kf positions x C)
    int nslab;                   // 32-element K slabs per tap
    int ncoef;                   // total input channels (transform coefficients held in shared memory)
    int npass;                   // 1: TF32, 3: 3xTF32 (hi*hi + lo*hi + hi*lo)
    const float* Whi;
    const float* Wlo;
    const float* bias;           // [N] (value | gate) or null
    int Cout, N, gate_off;       // plain: N == Cout; gated: N == 2*Cout, gate_off == Cout
    int relu;
    float algo_frac;
    float* out;
    int out_ld, out_coff;        // output row stride (channels) and channel offset
    const float* resid;          // optional, addressed like out
    double* stats[2];
    const float* stat_alpha[2];
    int nstats;
    int tiles_per_b;             // ceil(T*E / 128)
    unsigned long long* dbg;     // optional [16] cycle counters of CTA 0 (diagnostics), else null
    int stats_ld, stats_coff;    // statistics arrays are [B][stats_ld][2], this launch owns channels stats_coff.. (0 = [B][Cout][2])
    // streaming (one frame per stream and launch): B == 1 and T == number of STREAMS, so a GEMM row is (stream, e) and all
    // streams share one row space; sources / out / resid are rings [stream][RT][F][C], frame n = *step in slot n % RT, a tap dt
    // reads frame n - dt (zeros before the stream's start frame)
    const int* step; const int* start;
    int out_RT, resid_RT;
    // second variant of the same launch (launch_conv_umma_pair: both output parities of a streaming transposed conv): CTAs
    // [0, grid0) run the fields above, the rest these; everything else is shared.  grid0 == 0: one variant
    // post != 0 (streaming, static normalisation): the epilogue applies post_xf to the layer's own result and adds resid_xf(resid)
    // - the residual sum of a U-Net module (x0 + y, EaBNet.py:372-388) without a combine launch; the output then carries no transform
    int post;
    Xform post_xf, resid_xf;
    int grid0;
    struct Var { int E, out_off, ntaps, tiles_per_b; int dt[kMaxTaps], df[kMaxTaps]; const float* Whi; const float* Wlo; } var1;
};
bool umma_conv_supported(const UmmaConvArgs& a);
int launch_conv_umma_pair(const UmmaConvArgs& a0, const UmmaConvArgs& a1, cudaStream_t st);      // streaming only
int launch_conv_umma(const UmmaConvArgs& a, cudaStream_t st);

// "Stage once, shift by descriptor": GEMM rows are (t, e') with a padded pitch P >= E rows per frame; the transformed fp16
// input of a tile (+ halo) exists ONCE per tile as 1-2 column planes with the same pitch, and every tap is the same
// shared-memory plane viewed through a row-shifted UMMA descriptor (conv_staged.cu: planes staged through HBM by
// stage_kernel; conv_raw.cu: planes built in shared memory from raw tiles).
struct PlaneConvArgs {
    ConvSrc src[2];
    int nsrc;
    int B, T, Fin;
    int E, P;                    // valid rows per frame / padded pitch
    int nplanes;                 // 1, or 2 for stride-2 convs (even / odd input columns)
    int plane_cols[2];           // real columns per plane (the rest up to P are zero)
    int col_stride, col_off[2];  // fi = col * col_stride + col_off[plane]
    int ntaps;
    int tap_plane[kMaxTaps], tap_shift[kMaxTaps];   // A rows of a tap = plane rows + tap_shift
    int back, fwd;               // halo rows before / after the 128 tile rows
    int out_stride, out_off, Fout;
    int nslab, ncoef, npass;
    const float* Whi;
    const float* Wlo;
    const float* bias;
    int Cout, N, gate_off, relu;
    float algo_frac;
    float* out;
    int out_ld, out_coff;
    const float* resid;
    double* stats[2];
    const float* stat_alpha[2];
    int nstats;
    int tiles_per_b;             // ceil(T*P / 128)
    unsigned int p_magic;        // floor(2^32 / P) + 1  (division by P in the producers)
    int nbuf;                    // plane double-buffering (1 or 2), chosen by the launcher from the smem budget
    int resident;                // conv_tma: all weight images stay in shared memory (else a ring of nsb stages)
    int nsb;
    unsigned long long* dbg;     // optional [16] cycle counters of CTA 0 (diagnostics)
    // "staged" variant (conv_tma_kernel): the planes were written to HBM by stage_kernel as fp16 images
    // np[(plane * nslab + slab) * npb + hl] : [B][np_rows][64] halves, row = np_front + t*P + col, 128B-swizzled by row & 7
    const void* np[16];
    int np_rows, np_front;
    int stats_ld, stats_coff;    // see UmmaConvArgs
    // conv_tma: per-unit (tap, slab, pass) operand offsets, filled by the launcher: A offset / 16 relative to the tile's
    // plane origin, weight image index.  Kernel parameters live in the constant bank => warp-uniform loads.
    unsigned int unit_a[kMaxConvUnits];
    unsigned short unit_b[kMaxConvUnits];
    // STFT epilogue mode (stft_M > 0, see stft.cu): batch items are (b, mic) sequences, rows are hops, the 128 columns
    // of the launch are interleaved (re, im) pairs of bins out_coff/2 ..; the epilogue applies the square-root
    // compression and writes spec[b][t][f][mic][2] for t < stft_T, f < stft_F.
    int stft_M, stft_T, stft_F;
    // "wide" staging (first layer, 2M input channels): a plane row (t, e) is the whole kf x C tap window, wide_k contiguous
    // floats starting at column e*col_stride of frame t, zero-padded to nslab*64; the time taps are then plain row shifts
    int wide_k;
    // wide_kt > 1: wide_kt frames stacked along K - K block j (wide_k floats) of row (t, e) is the window of frame
    // t - (wide_kt - 1 - j) (zeros before the first frame): the time taps cost no MMAs of their own
    int wide_kt;
    int ksteps;                  // 16-column K steps of a slab that carry data (0 = all 4; 3 when wide_k <= 48: the rest is zero padding)
    float algo_in_share;         // profiling only: share of the layer's input bytes this launch accounts for (0 = all)
};
bool plane_conv_supported(const PlaneConvArgs& a);      // shape rules of the padded-pitch row space (pure check)
// staged variant: geometry helpers + the two launches (stage = normalise + fp16 + layout, conv = TMA-fed GEMM)
int staged_rows(const PlaneConvArgs& a, int* front);    // rows per batch item of a staged plane array (multiple of 8)
bool staged_conv_supported(const PlaneConvArgs& a);
bool staged_conv_fits(const PlaneConvArgs& a);          // shared-memory / unit-table budget only (caller vouches for the geometry)
int launch_stage(const PlaneConvArgs& a, cudaStream_t st);
int launch_conv_staged(PlaneConvArgs a, cudaStream_t st);
extern bool g_stft_tc;           // STFT as a tcgen05 GEMM (default) / fp32 CUDA-core kernel
extern bool g_istft_tc;          // iSTFT as a tcgen05 GEMM (option) / fused fp32 CUDA-core kernel (default)

// ---------------------------------------------------------------------------------------------------
// conv_raw.cu: the conv layer as ONE kernel, no staged planes in HBM.  The loader bulk-copies the RAW fp32 rows a
// tile needs (a contiguous (t, fi) range per source tensor) into a shared-memory ring, transform warps apply the
// producer's norm + PReLU (+ the lazy residual addend), split to fp16 hi (/lo) and write the 128-byte-swizzled operand
// planes in shared memory, the MMA warp runs the row-shifted-descriptor GEMM of conv_tma on them.  The two output
// parities of a transposed conv are two "variants" of one launch: they share the operand planes and differ in taps,
// weights and accumulator columns, so the input is read and normalised once.
// ---------------------------------------------------------------------------------------------------
constexpr int kRawMaxSlabs = 4;          // 64-channel K slabs = source tensors (each [B][T][Fin][64] fp32, one or two addends)
struct RawConvArgs {
    int nslab;
    const float* x0[kRawMaxSlabs];       // first addend of the slab's source
    const float* x1[kRawMaxSlabs];       // second addend (lazy residual sum of a module) or null
    Xform xf0[kRawMaxSlabs], xf1[kRawMaxSlabs];
    int mode0[kRawMaxSlabs], mode1[kRawMaxSlabs];    // 0 none, 1 norm -> PReLU, 2 PReLU -> norm, 3 norm -> PReLU with slopes in [0, 1]
    // ring stage = the raw rows one group of G operand rows needs from ONE slab: addend 0 at byte 0, addend 1 at add1_off
    int G, stage_bytes, add1_off;
    int q32, r32, qG, rG;                // divmod(32, P), divmod(G, P): row / group advance of the transform threads
    unsigned int dual_mask;              // bit s: slab s has a second addend
    int exp_flags;                       // timing experiments (EAB_RAW_EXP): bit 0 no transform, bit 1 no tail rows
    int B, T, Fin, P;
    int nplanes, plane_cols[2], col_stride, col_off[2];
    int back, fwd, tiles_per_b;
    unsigned int p_magic;
    int npass;
    int nvar;                            // 1, or 2 output parities of a transposed conv
    int E[2], out_off[2], out_stride, Fout;
    int ntaps[2];
    const float* Whi[2];
    const float* Wlo[2];
    const float* bias;
    int Cout, N, gate_off;               // Cout == 64 ; N = 64, or 128 when gated (gate_off == 64)
    float* out;                          // [B][T][Fout][Cout]
    double* stats;                       // [B][Cout][2] or null
    int nunits;                          // K units of a tile: (variant, tap, slab, pass)
    unsigned int unit_a[kMaxConvUnits];  // A operand offset / 16 from the operand buffer origin
    unsigned short unit_b[kMaxConvUnits];// resident weight image slot
    unsigned char unit_c[kMaxConvUnits]; // bit 0: variant, bit 7: first unit of the variant, bit 6: last
    int nbuf, nstage, resident, nsb;     // shared-memory plan (chosen by the launcher)
    float algo_frac;
    unsigned long long* dbg;
};
bool raw_conv_supported(const PlaneConvArgs* p, int n);
// max_grid > 0 caps the number of CTAs (tests: every CTA then walks many tiles even on small inputs)
int launch_conv_raw(const PlaneConvArgs* p, int n, cudaStream_t st, unsigned long long* dbg, int max_grid = 0);

struct CombineArgs {
    ConvSrc src[3];
    int nsrc;
    int B, P, C;                 // P = positions per batch item (T*F)
    float* out;
    const int* step;             // streaming (see ConvArgs): P = F, sources / out are rings
    int out_RT;
};
int launch_combine(const CombineArgs& a, cudaStream_t st);

struct LstmArgs {
    ConvSrc src;                 // [B][T][F][E] ; xf applied on load
    int layer_norm;              // LayerNorm(E) after the transform (bf_map.norm, EaBNet.py:598,608)
    const float* ln_g; const float* ln_b;
    const float* Wx;             // packed [E][64][4]
    const float* Wh;             // packed [64][64][4]
    const float* bias;           // packed [64][4]  (b_ih + b_hh)  (CUDA-core kernel) / [256] in image row order (tcgen05)
    const float* Wimg;           // tcgen05 kernel: fp16 images [hi|lo][x|h][256 rows][64 k], 128B-swizzled (or null)
    int B, T, F, E;
    float* out;                  // [B][T][F][64]
    unsigned long long* dbg;     // optional cycle counters (diagnostics)
    // streaming (CUDA-core kernel only): T == 1, carried state [B*F][64] read before / written after the step
    const int* step;
    float* h_state; float* c_state;
    int out_RT;
    int exp_flags;               // diagnostics builds only
    int rows_per_cta;            // tcgen05 kernel, diagnostics: live rows per CTA (0 = 96)
};
int launch_lstm(const LstmArgs& a, cudaStream_t st);

// streaming LSTM step on the tensor cores: the two elementwise kernels around the gate GEMM (lstm.cu)
struct LstmFrameArgs {
    const float* x; int x_RT; Xform xf; int layer_norm; const float* ln_g; const float* ln_b;     // ln: source ring -> out [rows][64]
    const float* gates; int gates_ld;                                                               // cell: [rows][gates_ld]
    float* c_state; float* h_state;                                                                 // [rows][64]
    float* out; int out_RT;                                                                         // ln: [rows][64]; cell: ring [S][out_RT][F][64]
    const int* step;
    int rows, F;                                                                                    // rows = streams x F
};
int launch_lstm_ln_frame(const LstmFrameArgs& a, cudaStream_t st);
int launch_lstm_cell_frame(const LstmFrameArgs& a, cudaStream_t st);
bool lstm_umma_supported(const LstmArgs& a);
int launch_lstm_umma(const LstmArgs& a, cudaStream_t st);
bool lstm_pp_supported(const LstmArgs& a);      // lstm_pp.cu: 128 sequences per CTA as two interleaved sub-batches
int launch_lstm_pp(const LstmArgs& a, cudaStream_t st);

struct BeamArgs {
    const float* w;              // [B][T][F][w_ld]: first 2M (mimo) / 2 (miso) channels used, channel = m*2 + ri
    int w_ld;
    const float* inpt;           // [B][T][F][M][2]
    int B, T, F, M, miso;
    float* out;                  // mimo [B][2][T][F];  miso [B][2][T]
    const int* step;             // streaming: w / inpt are rings of w_RT / inpt_RT frames, T == 1
    int w_RT, inpt_RT;
};
int launch_beam(const BeamArgs& a, cudaStream_t st);

// streaming TCM stack in one launch (tcm_stream.cu); descriptors live in the stream state blob (device memory)
struct TcmStreamDesc {
    // offsets in floats into the model's weight blob (so a re-upload of the weights never invalidates a state)
    long long W_in, W_dil, W_out;                        // fp32 CUDA-core layouts: [256][64], [kd][128][128], [64][256]
    long long sL, hL, aL, sR, hR, aR, sO, hO, aO;        // folded BatchNorm scale / shift and PReLU slope per branch
    long long ring_off;                                  // floats from act_base: history ring [S][RT][64] of the squeezed tensor
    int RT;
    int dt[8];
    int pad_[5];
};
struct TcmStreamArgs {
    const TcmStreamDesc* desc;
    const float* blob;
    int ntcm, p, kd, S;
    const int* step;
    const int* start;                                    // [S] first frame of each stream (see ConvArgs::start), or null
    float* act_base;
    const float* x;  int x_RT;                           // [S][x_RT][256] residual stream entering the stack
    float* out;      int out_RT;                         // [S][out_RT][256] sum of the group outputs
};
bool tcm_stream_supported(int cd1, int d_feat, int kd1);
int launch_tcm_stream(const TcmStreamArgs& a, cudaStream_t st);

// fused w_dnn (Linear-ReLU-Linear) + filter-and-sum (head_fused.cu); weights are the fp16 hi/lo images of the tcgen05 path
struct HeadArgs {
    const float* h;              // [rows][64] second LSTM layer output, rows = B*T*F
    const float* inpt;           // [rows][M][2]
    float* out;                  // [B][2][T][F]
    float* w_out;                // optional [rows][w_ld] copy of the beam weights (debug tap), else null
    int w_ld;
    long long rows;
    int T, F, M;
    const void *W1hi, *W1lo, *W2hi, *W2lo;      // [64][64] and [32][64] K-major 128B-swizzled fp16 images
    const float *b1, *b2;        // [64], [32]
};
bool head_fused_supported(const HeadArgs& a);
int launch_head_fused(const HeadArgs& a, cudaStream_t st);

// A chain of single-branch squeezed TCMs (GaGNet) as one persistent cooperative launch (tcm_chain.cu).  Offsets are in
// floats into the weight blob / in doubles into the statistics arena.
constexpr int kMaxChainLayers = 24;
struct TcmChainLayer {
    unsigned win_hi, win_lo;             // [4 slabs][64 rows][64 k] fp16 images of the 1x1 squeeze (hi / lo)
    unsigned wd_hi, wd_lo;               // [kd taps][64 rows][64 k] dilated conv
    unsigned wo_hi[2], wo_lo[2];         // 2 x [128 rows][64 k] 1x1 expand
    unsigned sc_d, sh_d, al_d;           // norm gamma / beta (or folded BatchNorm scale / shift) and PReLU slope in front of the dilated conv
    unsigned sc_o, sh_o, al_o;           // ... in front of the expand conv
    unsigned st_d, st_o;                 // [B][64][2] statistics accumulators (InstanceNorm)
    // gated TCMs (EaBNet.py:532-578): *_d above is the value (left) branch, these are the gate (right) branch
    unsigned wr_hi, wr_lo, sc_r, sh_r, al_r, st_r;
    short dt[8];                         // tap k reads frame t - dt[k]
};
struct TcmChainArgs {
    const float* blob;
    double* stats;
    unsigned* barrier;                   // zeroed grid-barrier counter
    const float* x_in[3];                // [B][T][256] chain inputs (read by the first layer only)
    float* x_buf[3];                     // [B][T][256] residual streams (chain outputs)
    float* y[3];                         // [B][T][64] scratch
    float* z[3];
    int nchains, nlayers, kd;
    int B, T, tiles_per_b;
    int instance_norm;
    int gated;                           // two dilated branches, value * sigmoid(gate)
    int no_cluster;                      // force the cooperative (grid-barrier) form (option tcm_chain = 2; diagnostics / tests)
    float inv_count;                     // 1 / T
    unsigned long long* dbg;             // optional [16] cycle counters of CTA 0 (diagnostics), else null
    TcmChainLayer L[kMaxChainLayers];    // [chain][layer]
};
bool tcm_chain_supported(const TcmChainArgs& a);
int launch_tcm_chain(const TcmChainArgs& a, cudaStream_t st);

// GaGNet post-filter glue (gag_elementwise.cu)
struct GagPackArgs {
    const float* inpt;           // reference-microphone spectrum, element (b, c, t, f) at b*sb + c*sc + t*st + f*sf (floats)
    long long sb, sc, st, sf;
    const float* pre;            // [B][2][T][F] previous estimate (EaBNet output layout)
    float* x4;                   // [B][T][F][4]  channels (inpt_r, pre_r, inpt_i, pre_i)
    float* pre_row;              // [B][T][KP]    channel ri*F + f, zeros from 2F on
    int B, T, F, KP, KP2;        // KP2 = KP / 2
    // streaming (T == 1): frame *step of every stream; inpt may be a ring of in_RT frames (slot stride in_slot floats: the
    // spectrum ring of an EaBNet stream state), x4 is a ring of x_RT frames [S][x_RT][F][4]
    const int* step;
    int in_RT; long long in_slot;
    int x_RT;
};
int launch_gag_pack(const GagPackArgs& a, cudaStream_t st);

struct GagCrmArgs {
    const float* pre_row;        // [B][T][KP]
    const float* gain;           // [B][T][ld_g] raw linear_g output (activation applied here)
    const float* res_r;          // [B][T][ld_r] linear_r / linear_i outputs
    const float* res_i;
    int ld_g, ld_r;
    int acti;                    // 0 sigmoid, 1 tanh, 2 relu (GaGNet.py:165-172)
    float* next_row;             // [B][T][KP] the estimate as the next module's pre_x row, or null for the last module
    float* out;                  // [B][2][T][F]
    int B, T, F, KP, KP2;
};
int launch_gag_crm(const GagCrmArgs& a, cudaStream_t st);

// scratch: optional device buffer for the tensor-core path's hop planes (else a library-owned grow-only buffer, one per device)
int launch_stft(const float* wave, float* spec, int B, int M, int L, cudaStream_t st, void* scratch = nullptr, size_t scratch_bytes = 0);
// 16-bit PCM source (file channel order, microphone m = file channel order[m], null = identity) for the same spectrum
int launch_stft_pcm16(const short* pcm, const int* order, float* spec, int B, int M, int L, cudaStream_t st, void* scratch = nullptr,
                      size_t scratch_bytes = 0);
// wave16 != null: write int16(clip(y, -1, 1) * 32767) there instead of fp32 samples to `wave`
int launch_istft(const float* spec, float* wave, int B, int T, cudaStream_t st, short* wave16 = nullptr);
// streaming front/back end (stft.cu): one hop in -> spectrum frame *step into a ring; spectrum frame -> one hop out
// (delayed by one hop: overlap-add needs the next frame), carried state in prev_hop [S][M][160] / tail [S][160]
// hop16 / hop_out16 non-null: the hop arrives / leaves as 16-bit PCM (sample / 32768 in, int16(clip(y) * 32767) out);
// start: [S] first frame of each stream (eab_stream_reset_one), or null
int launch_stft_frame(const float* hop, const short* hop16, float* prev_hop, float* spec_ring, int spec_RT, const int* step,
                      const int* start, int S, int M, cudaStream_t st);
int launch_istft_frame(const float* frame, float* tail, float* hop_out, short* hop_out16, const int* step, const int* start, int S,
                       cudaStream_t st);
int launch_stream_restart(int* start, const int* step, int idx, cudaStream_t st);       // start[idx] = *step
int launch_step_advance(int* step, cudaStream_t st);

}  // namespace eab
