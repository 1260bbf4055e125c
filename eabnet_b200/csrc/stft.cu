// 320/160 STFT with square-root compression, and the matching iSTFT, as windowed real-DFT matrix products in
// fp32 (the compression amplifies operand rounding, so no reduced-precision operands here: SURVEY.md 7.2).
//
//   forward (test.py:32-43):  frame g[n] = x_reflect[160 t + n - 160] * hann[n],  X[f] = sum_n g[n] e^{-i th n f}
//       folded with the even/odd symmetry of cos/sin about n = 160 (halves the multiply-adds):
//         Re X[f] =  sum_{k=0..160} Ev[k] cos(th k f),   Ev[k] = g[k] + g[320-k]  (Ev[0]=g[0], Ev[160]=g[160])
//         Im X[f] = -sum_{k=1..159} Od[k] sin(th k f),   Od[k] = g[k] - g[320-k]
//       then z * |z|^{-1/2}   ( == |z|^0.5 * (cos, sin)(atan2(im, re)) ).
//   inverse (enhance.py:59-61): s[n] = (P[n] - Q[n]) / 320, s[320-n] = (P[n] + Q[n]) / 320 for n <= 160 with
//         P[n] = sum_f a_f Yr[f] cos(th n f) (a_0 = a_160 = 1, else 2),  Q[n] = sum_{f=1..159} 2 Yi[f] sin(th n f)
//       windowed overlap-add of the two frames covering each output sample, divided by the squared-window
//       envelope, with the 160-sample centre padding trimmed  -> 160 (T-1) samples.
//
// Both directions use the same "column GEMM": a thread owns one output column (322 of them) and RB = 24 rows
// (frames x mics, or frames) as register accumulators; the A operand sits in shared memory [k][RB], the
// twiddle table (built in double on the host) streams from L2.
#include <math.h>
#include <string.h>
#include <mutex>
#include <vector>

#include <cuda_fp16.h>

#include "common.cuh"

namespace eab {

bool g_stft_tc = true;

namespace {

constexpr int NFFT = 320, HOP = 160, NF = 161;
constexpr int NCOL = 2 * NF;          // 322 output columns
constexpr int LDT = 352;              // padded table row
constexpr int RB = 24;                // accumulator rows per thread
constexpr int THREADS = 352;          // 11 warps, one column per thread

struct Tables {
    float* fwd = nullptr;    // [161][352]: col<161: cos(th k col); col>=161: -sin(th k (col-161))
    float* inv = nullptr;    // [161][352]: col<161: a_f cos(th f n)/320 ; col>=161: 2 sin(th f n)/320 (f in 1..159)
    float* win = nullptr;    // [320] periodic hann
    float* ienv = nullptr;   // [160] 1 / (w[n]^2 + w[n+160]^2)
    // tensor-core forward transform: windowed DFT matrix as fp16 hi / lo weight images of the conv_tma kernel,
    // [3 column groups][2 taps (first / second hop of a frame)][3 K slabs][128 rows][64 k], 128B-swizzled;
    // column 2f = Re X[f], 2f+1 = Im X[f]
    void* dft_hi = nullptr;
    void* dft_lo = nullptr;
    // tensor-core inverse transform (conv_umma weight images, [2 taps][6 K slabs][N rows][64 k], 128B-swizzled, fp16 hi / lo):
    // row n = output sample of a hop, K = (Re X[0..160] | Im X[0..160] | zeros to 384) of frame t (tap 0: its first half lands
    // in hop t - 1) and frame t - 1 (tap 1: its second half); synthesis window, 1 / 320, the half-spectrum weights and the
    // window envelope folded in.  Column groups: samples 0..127 (N = 128) and 128..159 (N = 32).
    void* idft_hi[2] = {nullptr, nullptr};
    void* idft_lo[2] = {nullptr, nullptr};
    void* planes = nullptr;  // scratch: fp16 hi/lo hop planes of the waveforms (grow-only)
    size_t planes_bytes = 0;
};
constexpr int DFT_GROUPS = 3, DFT_SLABS = 3;
constexpr int IDFT_SLABS = 6, IDFT_K = IDFT_SLABS * 64;      // 322 -> 384
constexpr size_t DFT_GROUP_HALVES = (size_t)2 * DFT_SLABS * 128 * 64;
Tables g_tab[64];
std::mutex g_tab_mu;

int get_tables(Tables** out) {
    int dev = 0;
    EAB_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return fail("stft: device index out of range");
    std::lock_guard<std::mutex> lk(g_tab_mu);
    Tables& t = g_tab[dev];
    if (!t.fwd) {
        std::vector<float> fwd((size_t)NF * LDT, 0.f), inv((size_t)NF * LDT, 0.f), win(NFFT), ienv(HOP);
        const double th = 2.0 * M_PI / NFFT;
        for (int k = 0; k < NF; ++k)
            for (int f = 0; f < NF; ++f) {
                const int kf = (k * f) % NFFT;             // exact argument reduction
                const double c = cos(th * kf), s = sin(th * kf);
                fwd[(size_t)k * LDT + f] = (float)c;
                fwd[(size_t)k * LDT + NF + f] = (k == 0 || k == 160) ? 0.f : (float)(-s);
                // inverse: row = frequency k, column = sample f
                const double af = (k == 0 || k == 160) ? 1.0 : 2.0;
                inv[(size_t)k * LDT + f] = (float)(af * c / NFFT);
                inv[(size_t)k * LDT + NF + f] = (k == 0 || k == 160) ? 0.f : (float)(2.0 * s / NFFT);
            }
        std::vector<double> w(NFFT);
        for (int n = 0; n < NFFT; ++n) {
            w[n] = 0.5 - 0.5 * cos(th * n);
            win[n] = (float)w[n];
        }
        for (int n = 0; n < HOP; ++n) {
            const double e = (double)win[n] * win[n] + (double)win[n + HOP] * win[n + HOP];
            ienv[n] = (float)(1.0 / e);
        }
        EAB_CUDA(cudaMalloc(&t.fwd, fwd.size() * 4));
        EAB_CUDA(cudaMalloc(&t.inv, inv.size() * 4));
        EAB_CUDA(cudaMalloc(&t.win, win.size() * 4));
        EAB_CUDA(cudaMalloc(&t.ienv, ienv.size() * 4));
        EAB_CUDA(cudaMemcpy(t.fwd, fwd.data(), fwd.size() * 4, cudaMemcpyHostToDevice));
        EAB_CUDA(cudaMemcpy(t.inv, inv.data(), inv.size() * 4, cudaMemcpyHostToDevice));
        EAB_CUDA(cudaMemcpy(t.win, win.data(), win.size() * 4, cudaMemcpyHostToDevice));
        EAB_CUDA(cudaMemcpy(t.ienv, ienv.data(), ienv.size() * 4, cudaMemcpyHostToDevice));
        // DFT weight images (double precision twiddles, exact argument reduction)
        std::vector<__half> hi(DFT_GROUPS * DFT_GROUP_HALVES), lo(DFT_GROUPS * DFT_GROUP_HALVES);
        for (int g = 0; g < DFT_GROUPS; ++g)
            for (int tap = 0; tap < 2; ++tap)
                for (int sl = 0; sl < DFT_SLABS; ++sl)
                    for (int n = 0; n < 128; ++n)
                        for (int k = 0; k < 64; ++k) {
                            const int col = g * 128 + n, f = col >> 1, im = col & 1;
                            const int hs = sl * 64 + k;                       // sample inside the hop
                            double wv = 0.0;
                            if (f < NF && hs < HOP) {
                                const int ns = tap * HOP + hs;                // sample inside the frame
                                const int kf = (ns * f) % NFFT;
                                wv = w[ns] * (im ? -sin(th * kf) : cos(th * kf));
                                if (im && (f == 0 || f == 160)) wv = 0.0;
                            }
                            const float wf = (float)wv;
                            const __half h = __float2half_rn(wf);
                            const size_t idx = g * DFT_GROUP_HALVES + ((size_t)(tap * DFT_SLABS + sl) * 128 + n) * 64 +
                                               (size_t)((((k >> 3) ^ (n & 7)) << 3) | (k & 7));
                            hi[idx] = h;
                            lo[idx] = __float2half_rn(wf - __half2float(h));
                        }
        EAB_CUDA(cudaMalloc(&t.dft_hi, hi.size() * 2));
        EAB_CUDA(cudaMalloc(&t.dft_lo, lo.size() * 2));
        EAB_CUDA(cudaMemcpy(t.dft_hi, hi.data(), hi.size() * 2, cudaMemcpyHostToDevice));
        EAB_CUDA(cudaMemcpy(t.dft_lo, lo.data(), lo.size() * 2, cudaMemcpyHostToDevice));
        for (int g = 0; g < 2; ++g) {
            const int N = g ? 32 : 128, n0 = g ? 128 : 0;
            std::vector<__half> ih((size_t)2 * IDFT_SLABS * N * 64), il(ih.size());
            for (int tap = 0; tap < 2; ++tap)
                for (int sl = 0; sl < IDFT_SLABS; ++sl)
                    for (int n = 0; n < N; ++n)
                        for (int k = 0; k < 64; ++k) {
                            const int kk = sl * 64 + k, smp = n0 + n;
                            double wv = 0.0;
                            if (kk < 2 * NF) {
                                const int f = kk < NF ? kk : kk - NF;
                                const int ns = smp + (tap ? HOP : 0);        // sample inside the frame
                                const int kf = (ns * f) % NFFT;
                                const double cf = (f == 0 || f == 160) ? 1.0 : 2.0;
                                const double basis = kk < NF ? cos(th * kf) : -sin(th * kf);
                                const double env = (double)win[smp] * win[smp] + (double)win[smp + HOP] * win[smp + HOP];
                                wv = (double)win[ns] * cf * basis / NFFT / env;
                            }
                            const float wf = (float)wv;
                            const __half h = __float2half_rn(wf);
                            const size_t idx = ((size_t)(tap * IDFT_SLABS + sl) * N + n) * 64 + (size_t)((((k >> 3) ^ (n & 7)) << 3) | (k & 7));
                            ih[idx] = h;
                            il[idx] = __float2half_rn(wf - __half2float(h));
                        }
            EAB_CUDA(cudaMalloc(&t.idft_hi[g], ih.size() * 2));
            EAB_CUDA(cudaMalloc(&t.idft_lo[g], il.size() * 2));
            EAB_CUDA(cudaMemcpy(t.idft_hi[g], ih.data(), ih.size() * 2, cudaMemcpyHostToDevice));
            EAB_CUDA(cudaMemcpy(t.idft_lo[g], il.data(), il.size() * 2, cudaMemcpyHostToDevice));
        }
    }
    *out = &t;
    return 0;
}

// acc[r] = sum_k A[k][r] * tab[k][col]   (A in shared memory, 16-byte aligned rows of RB floats)
// The table column is fetched KB values at a time, one block ahead of the arithmetic: with one load per k the loop runs at
// the L2 latency of that load (a streaming frame kernel - few CTAs, the 207 KB table read by each - took 58 us for 161 steps).
// R <= RB live rows (the rows of A keep their pitch of RB floats).
constexpr int KB = 8;
template <int R = RB>
__device__ __forceinline__ void column_gemm(const float* __restrict__ A, const float* __restrict__ tab, int col,
                                            float (&acc)[R]) {
#pragma unroll
    for (int r = 0; r < R; ++r) acc[r] = 0.f;
    const float* tc = tab + col;
    float cur[KB], nxt[KB];
#pragma unroll
    for (int j = 0; j < KB; ++j) cur[j] = __ldg(tc + (size_t)j * LDT);
#pragma unroll 1
    for (int k0 = 0; k0 < NF; k0 += KB) {
#pragma unroll
        for (int j = 0; j < KB; ++j) nxt[j] = k0 + KB + j < NF ? __ldg(tc + (size_t)(k0 + KB + j) * LDT) : 0.f;
#pragma unroll
        for (int j = 0; j < KB; ++j) {
            if (k0 + j < NF) {
                const float tv = cur[j];
                const float4* a4 = reinterpret_cast<const float4*>(A + (k0 + j) * RB);
#pragma unroll
                for (int u = 0; u < R / 4; ++u) {
                    const float4 v = a4[u];
                    acc[4 * u + 0] = fmaf(v.x, tv, acc[4 * u + 0]);
                    acc[4 * u + 1] = fmaf(v.y, tv, acc[4 * u + 1]);
                    acc[4 * u + 2] = fmaf(v.z, tv, acc[4 * u + 2]);
                    acc[4 * u + 3] = fmaf(v.w, tv, acc[4 * u + 3]);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < KB; ++j) cur[j] = nxt[j];
    }
}

constexpr int FR = 8;    // frames per CTA in the forward transform

// grid (ceil(T/FR), B).  Rows = (frame r, mic m) pairs, processed RB at a time.
__global__ void __launch_bounds__(THREADS) stft_kernel(const float* __restrict__ wave, float* __restrict__ spec,
                                                       const float* __restrict__ tab, const float* __restrict__ win,
                                                       int B, int M, int L, int T) {
    extern __shared__ __align__(16) float dsm[];
    float* Aev = dsm;                                   // [NF][RB]
    float* Aod = Aev + NF * RB;                         // [NF][RB]
    float2* Xs = reinterpret_cast<float2*>(Aod + NF * RB);   // [RB][NF]
    const int b = blockIdx.y;
    const int t0 = blockIdx.x * FR;
    const int nrows = FR * M;
    const int col = threadIdx.x;
    for (int rb0 = 0; rb0 < nrows; rb0 += RB) {
        // ---- build the folded, windowed frames
        for (int i = threadIdx.x; i < RB * NF; i += THREADS) {
            const int r = i / NF, k = i - r * NF;
            const int row = rb0 + r;
            float ev = 0.f, od = 0.f;
            const int fr = row / M, m = row - fr * M;
            const int t = t0 + fr;
            if (row < nrows && t < T) {
                const float* x = wave + ((size_t)b * M + m) * L;
                auto sample = [&](int n) -> float {
                    int j = HOP * t + n - HOP;            // centre padding = 160, reflect (test.py:35)
                    if (j < 0) j = -j;
                    if (j >= L) j = 2 * (L - 1) - j;
                    return __ldg(x + j) * __ldg(win + n);
                };
                const float g0 = sample(k);
                if (k == 0 || k == HOP) {
                    ev = g0;
                } else {
                    const float g1 = sample(NFFT - k);
                    ev = g0 + g1;
                    od = g0 - g1;
                }
            }
            Aev[k * RB + r] = ev;
            Aod[k * RB + r] = od;
        }
        __syncthreads();
        // ---- 161 real + 161 imaginary columns
        if (col < NCOL) {
            float acc[RB];
            column_gemm(col < NF ? Aev : Aod, tab, col, acc);
            const int f = col < NF ? col : col - NF;
            float* xs = reinterpret_cast<float*>(Xs);
#pragma unroll
            for (int r = 0; r < RB; ++r) xs[(r * NF + f) * 2 + (col < NF ? 0 : 1)] = acc[r];
        }
        __syncthreads();
        // ---- compression z |z|^{-1/2} and store to [B,T,F,M,2]
        for (int i = threadIdx.x; i < RB * NF; i += THREADS) {
            const int r = i / NF, f = i - r * NF;
            const int row = rb0 + r;
            const int fr = row / M, m = row - fr * M;
            const int t = t0 + fr;
            if (row < nrows && t < T) {
                float2 z = Xs[r * NF + f];
                const float mag = sqrtf(z.x * z.x + z.y * z.y);
                const float sc = mag > 0.f ? rsqrtf(mag) : 0.f;
                z.x *= sc;
                z.y *= sc;
                reinterpret_cast<float2*>(spec)[(((size_t)b * T + t) * NF + f) * M + m] = z;
            }
        }
        __syncthreads();
    }
}

constexpr int IFR = RB;          // frames per CTA in the inverse transform -> IFR-1 output hops

// grid (ceil((T-1)/(IFR-1)), B)
// wave16 != null: the reference's int16 writer instead of fp32 samples, int16(clip(y, -1, 1) * 32767) truncated toward zero
// (dataset/mcse_dataset_offline_gen.py:38-39)
__global__ void __launch_bounds__(THREADS) istft_kernel(const float* __restrict__ spec, float* __restrict__ wave,
                                                        short* __restrict__ wave16, const float* __restrict__ tab,
                                                        const float* __restrict__ win, const float* __restrict__ ienv, int B, int T) {
    extern __shared__ __align__(16) float dsm[];
    float* Are = dsm;                    // [NF][RB]
    float* Aim = Are + NF * RB;          // [NF][RB]
    float* PQ = Aim + NF * RB;           // [frame][col]: P[n] (col<161), Q[n] (col>=161)
    const int b = blockIdx.y;
    const int t0 = blockIdx.x * (IFR - 1);
    const size_t plane = (size_t)T * NF;
    const float* yr = spec + (size_t)b * 2 * plane;
    const float* yi = yr + plane;
    for (int i = threadIdx.x; i < RB * NF; i += THREADS) {
        const int r = i / NF, f = i - r * NF;
        const int t = t0 + r;
        float re = 0.f, im = 0.f;
        if (t < T) {
            re = __ldg(yr + (size_t)t * NF + f);
            im = __ldg(yi + (size_t)t * NF + f);
        }
        Are[f * RB + r] = re;
        Aim[f * RB + r] = im;
    }
    __syncthreads();
    const int col = threadIdx.x;
    if (col < NCOL) {
        float acc[RB];
        column_gemm(col < NF ? Are : Aim, tab, col, acc);
#pragma unroll
        for (int r = 0; r < RB; ++r) PQ[r * NCOL + col] = acc[r];
    }
    __syncthreads();
    const size_t out_len = (size_t)HOP * (T - 1);
    for (int i = threadIdx.x; i < (IFR - 1) * HOP; i += THREADS) {
        const int u = i / HOP, n = i - u * HOP;
        const int ta = t0 + u;              // earlier frame contributes its second half, sample n + 160
        if (ta + 1 > T - 1) continue;
        const float* Pa = PQ + u * NCOL;
        const float* Pb = PQ + (u + 1) * NCOL;
        const float sa = Pa[HOP - n] + Pa[NF + HOP - n];          // s_a[n+160] * 320 / 320 (scale in table)
        const float sb = Pb[n] - Pb[NF + n];                      // s_b[n]
        const float v = (__ldg(win + n + HOP) * sa + __ldg(win + n) * sb) * __ldg(ienv + n);
        const size_t o = (size_t)b * out_len + (size_t)ta * HOP + n;
        if (wave16) wave16[o] = (short)(int)(fminf(fmaxf(v, -1.f), 1.f) * 32767.f);
        else wave[o] = v;
    }
}

constexpr size_t kSmemBytes = (size_t)(2 * NF * RB + RB * NCOL) * sizeof(float);   // all four kernels

// ------------------------------------------------------------------------------------------------ streaming
// Frame n of the centred STFT covers samples [160 n - 160, 160 n + 160): the previous hop (carried) and the new one.
// Frame 0's first half is the reflection of hop 0 (sample 160 - k at position k; position 0 is multiplied by
// hann[0] = 0, so the one sample that would come from hop 1 never matters).  grid (ceil(S / SG)), rows = (stream, mic).
// hop16 != null: the hop is 16-bit PCM (sample / 32768).  start[s] = absolute frame at which stream s (re)started: that
// frame is the stream's frame 0 (reflection instead of the carried hop).
__global__ void __launch_bounds__(THREADS) stft_frame_kernel(const float* __restrict__ hop, const short* __restrict__ hop16,
                                                             float* __restrict__ prev, float* __restrict__ spec, int spec_RT,
                                                             const int* __restrict__ step_p, const int* __restrict__ start,
                                                             const float* __restrict__ tab, const float* __restrict__ win, int S,
                                                             int M, int SG) {
    extern __shared__ __align__(16) float dsm[];
    float* Aev = dsm;
    float* Aod = Aev + NF * RB;
    float2* Xs = reinterpret_cast<float2*>(Aod + NF * RB);
    const int step = *step_p;
    const int s0 = blockIdx.x * SG;
    const int nrows = SG * M;
    const int col = threadIdx.x;
    for (int i = threadIdx.x; i < RB * NF; i += THREADS) {
        const int r = i / NF, k = i - r * NF;
        float ev = 0.f, od = 0.f;
        const int sl = r / M, m = r - sl * M;
        if (r < nrows && s0 + sl < S) {
            const size_t cbase = ((size_t)(s0 + sl) * M + m) * HOP;
            const float* pv = prev + cbase;
            const bool first = step <= (start ? __ldg(start + s0 + sl) : 0);      // the stream's frame 0
            auto cur = [&](int n) -> float { return hop16 ? (float)__ldg(hop16 + cbase + n) * (1.f / 32768.f) : __ldg(hop + cbase + n); };
            auto sample = [&](int n) -> float {
                float x;
                if (n >= HOP) x = cur(n - HOP);
                else if (!first) x = pv[n];
                else x = n == 0 ? 0.f : cur(HOP - n);
                return x * __ldg(win + n);
            };
            const float g0 = sample(k);
            if (k == 0 || k == HOP) {
                ev = g0;
            } else {
                const float g1 = sample(NFFT - k);
                ev = g0 + g1;
                od = g0 - g1;
            }
        }
        Aev[k * RB + r] = ev;
        Aod[k * RB + r] = od;
    }
    __syncthreads();
    // the carried hop: every read of `prev` above is complete
    for (int i = threadIdx.x; i < nrows * HOP; i += THREADS) {
        const int r = i / HOP, n = i - r * HOP;
        const int sl = r / M, m = r - sl * M;
        if (s0 + sl < S) {
            const size_t o = ((size_t)(s0 + sl) * M + m) * HOP + n;
            prev[o] = hop16 ? (float)__ldg(hop16 + o) * (1.f / 32768.f) : __ldg(hop + o);
        }
    }
    if (col < NCOL) {
        float acc[RB];
        column_gemm(col < NF ? Aev : Aod, tab, col, acc);
        const int f = col < NF ? col : col - NF;
        float* xs = reinterpret_cast<float*>(Xs);
#pragma unroll
        for (int r = 0; r < RB; ++r) xs[(r * NF + f) * 2 + (col < NF ? 0 : 1)] = acc[r];
    }
    __syncthreads();
    const int slot = ring_slot(step, spec_RT);
    for (int i = threadIdx.x; i < RB * NF; i += THREADS) {
        const int r = i / NF, f = i - r * NF;
        const int sl = r / M, m = r - sl * M;
        if (r < nrows && s0 + sl < S) {
            float2 z = Xs[r * NF + f];
            const float mag = sqrtf(z.x * z.x + z.y * z.y);
            const float sc = mag > 0.f ? rsqrtf(mag) : 0.f;
            z.x *= sc;
            z.y *= sc;
            reinterpret_cast<float2*>(spec)[(((size_t)(s0 + sl) * spec_RT + slot) * NF + f) * M + m] = z;
        }
    }
}

// Spectrum frame n [S][2][161] -> windowed inverse DFT; output hop n-1 = (second half of frame n-1, carried in `tail`,
// + first half of frame n) / envelope.  The first call (n == 0) emits zeros.  grid (ceil(S / RBF)), rows = streams.
// hop_out16 != null: the hop leaves as int16(clip(y, -1, 1) * 32767); start: see stft_frame_kernel (a stream's first frame emits
// zeros: its overlap-add partner does not exist yet).
constexpr int RBF = 8;           // streams per CTA of the inverse frame kernel (one frame each: 256 streams = 32 CTAs, not 11)
__global__ void __launch_bounds__(THREADS) istft_frame_kernel(const float* __restrict__ frame, float* __restrict__ tail,
                                                              float* __restrict__ hop_out, short* __restrict__ hop_out16,
                                                              const int* __restrict__ step_p, const int* __restrict__ start,
                                                              const float* __restrict__ tab, const float* __restrict__ win,
                                                              const float* __restrict__ ienv, int S) {
    extern __shared__ __align__(16) float dsm[];
    float* Are = dsm;
    float* Aim = Are + NF * RB;
    float* PQ = Aim + NF * RB;
    const int step = *step_p;
    const int s0 = blockIdx.x * RBF;
    for (int i = threadIdx.x; i < RBF * NF; i += THREADS) {
        const int r = i / NF, f = i - r * NF;
        float re = 0.f, im = 0.f;
        if (s0 + r < S) {
            re = __ldg(frame + ((size_t)(s0 + r) * 2 + 0) * NF + f);
            im = __ldg(frame + ((size_t)(s0 + r) * 2 + 1) * NF + f);
        }
        Are[f * RB + r] = re;
        Aim[f * RB + r] = im;
    }
    __syncthreads();
    const int col = threadIdx.x;
    if (col < NCOL) {
        float acc[RBF];
        column_gemm<RBF>(col < NF ? Are : Aim, tab, col, acc);
#pragma unroll
        for (int r = 0; r < RBF; ++r) PQ[r * NCOL + col] = acc[r];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < RBF * HOP; i += THREADS) {
        const int r = i / HOP, n = i - r * HOP;
        if (s0 + r >= S) continue;
        const float* P = PQ + r * NCOL;
        const float first = P[n] - P[NF + n];                       // s[n]
        const float second = P[HOP - n] + P[NF + HOP - n];          // s[n + 160]
        const size_t o = (size_t)(s0 + r) * HOP + n;
        const bool begun = step > (start ? __ldg(start + s0 + r) : 0);
        const float y = begun ? (tail[o] + __ldg(win + n) * first) * __ldg(ienv + n) : 0.f;
        if (hop_out16) hop_out16[o] = (short)(int)(fminf(fmaxf(y, -1.f), 1.f) * 32767.f);
        else hop_out[o] = y;
        tail[o] = __ldg(win + n + HOP) * second;
    }
}

__global__ void step_advance_kernel(int* step) {
    if (threadIdx.x == 0) *step += 1;
}
__global__ void stream_restart_kernel(int* start, const int* step, int idx) {
    if (threadIdx.x == 0) start[idx] = *step;
}

}  // namespace

// ------------------------------------------------------------------------------------------------ tensor-core STFT
// A centred frame t is two consecutive hops of the reflect-padded signal p: X_t = p_hop[t] W_first + p_hop[t+1] W_second,
// i.e. a two-tap "convolution" over hop rows with K = 160 (+32 zero) channels per tap and 322 output columns: exactly
// the GEMM the conv_tma kernel runs.  This kernel writes the hop rows as the kernel's fp16 hi / lo plane images
// (np[slab*2 + hl] : [B*M][np_rows][64] halves, 128B-swizzled by row & 7; row = np_front + hop index, hops 0..T).
// The waveform is fp32 [B][M][L], or (pcm.data != null) the 16-bit PCM wire format [B][M][L] int16 in file channel order:
// sample / 32768 (what torchaudio.load does, enhance.py:35) of file channel pcm.order[mic] (enhance.py:41-42) - the int16 ->
// float conversion and the microphone permutation cost nothing here, the stand-alone pcm16_to_float pass is gone.
struct PcmSrc { const short* data; int order[64]; };
__global__ void __launch_bounds__(256) stft_stage_kernel(const float* __restrict__ wave, const PcmSrc pcm, uint8_t* __restrict__ planes,
                                                         size_t image_bytes, int np_rows, int np_front, int L, int T, int M) {
    const int b = blockIdx.z, slab = blockIdx.y;
    const int c8 = threadIdx.x & 7;
    for (int it = 0; it < 8; ++it) {
        const int rho = (blockIdx.x * 8 + it) * 32 + (threadIdx.x >> 3);
        if (rho >= np_rows) break;
        const int rr = rho - np_front;                      // plane row = hop j * M + mic
        const int j = rr >= 0 ? rr / M : -1;
        const int mic = rr - j * M;
        const int mic_ok = j >= 0 ? mic : 0;
        const float* x = wave + ((size_t)b * M + mic_ok) * L;
        const short* xp = pcm.data ? pcm.data + ((size_t)b * M + pcm.order[mic_ok]) * L : nullptr;
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int hs = slab * 64 + c8 * 8 + i;
            float s = 0.f;
            if (j >= 0 && j <= T && hs < HOP) {
                int q = HOP * j + hs - HOP;                 // centre padding = 160, reflect (test.py:35)
                if (q < 0) q = -q;
                if (q >= L) q = 2 * (L - 1) - q;
                s = xp ? (float)__ldg(xp + q) * (1.f / 32768.f) : __ldg(x + q);
            }
            v[i] = s;
        }
        uint32_t h[4], l[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const __half2 hh = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
            const float2 hf = __half22float2(hh);
            const __half2 ll = __floats2half2_rn(v[2 * i] - hf.x, v[2 * i + 1] - hf.y);
            h[i] = *reinterpret_cast<const uint32_t*>(&hh);
            l[i] = *reinterpret_cast<const uint32_t*>(&ll);
        }
        const size_t off = ((size_t)b * np_rows + rho) * 128 + (size_t)((c8 ^ (rho & 7)) << 4);
        *reinterpret_cast<uint4*>(planes + (size_t)(slab * 2 + 0) * image_bytes + off) = make_uint4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<uint4*>(planes + (size_t)(slab * 2 + 1) * image_bytes + off) = make_uint4(l[0], l[1], l[2], l[3]);
    }
}

int launch_stft_tc(Tables* t, const float* wave, const PcmSrc& pcm, float* spec, int B, int M, int L, int T, cudaStream_t st,
                   void* scratch, size_t scratch_bytes) {
    PlaneConvArgs a;
    memset(&a, 0, sizeof(a));
    // rows of batch item b are (hop j, mic) with pitch M, so that the epilogue's lanes write consecutive mics
    a.B = B; a.T = T + 1; a.Fin = M; a.E = M; a.P = M;
    a.nplanes = 1; a.plane_cols[0] = M; a.col_stride = 1;
    a.ntaps = 2;
    a.tap_plane[0] = 0; a.tap_shift[0] = 0;                 // first half of the frame: hop t
    a.tap_plane[1] = 0; a.tap_shift[1] = M;                 // second half: hop t + 1
    a.back = 0; a.fwd = M;
    a.out_stride = 1; a.Fout = 1;
    a.nslab = DFT_SLABS; a.ncoef = 0; a.npass = 3;
    a.Cout = 128; a.N = 128; a.gate_off = 0; a.algo_frac = 322.f / 384.f;
    a.out = spec; a.out_ld = 128;
    a.tiles_per_b = (int)(((long long)a.T * a.P + 127) / 128);
    a.stft_M = M; a.stft_T = T; a.stft_F = NF;
    int front = 0;
    a.np_rows = staged_rows(a, &front);
    a.np_front = front;
    const size_t image_bytes = (size_t)a.B * a.np_rows * 128;
    const size_t need = image_bytes * DFT_SLABS * 2;
    // hop planes: the caller's scratch when it is large enough (eab_enhance lends the forward workspace, which is idle until
    // the STFT has finished - concurrent calls on different streams then share nothing), else the library's grow-only buffer
    // hop planes: the caller's scratch when it is large enough (eab_enhance / eab_enhance_postnet lend the forward workspace,
    // which is idle until the STFT has finished), else a stream-ordered allocation private to this call (legal during stream
    // capture, nothing shared between streams or models, nothing a captured graph could see freed under it)
    uint8_t* planes = nullptr;
    bool own = false;
    if (scratch && scratch_bytes >= need) planes = static_cast<uint8_t*>(scratch);
    if (!planes) {
        void* q = nullptr;
        EAB_CUDA(cudaMallocAsync(&q, need, st));
        planes = static_cast<uint8_t*>(q);
        own = true;
    }
    for (int i = 0; i < DFT_SLABS * 2; ++i) a.np[i] = planes + (size_t)i * image_bytes;
    {
        ProfScope ps("stft_stage", 0.0, 0.0, st, (pcm.data ? 2.0 : 4.0) * (double)B * M * L + (double)need);
        dim3 grid((a.np_rows + 255) / 256, DFT_SLABS, a.B);
        EAB_CUDA(launch_k(stft_stage_kernel, grid, dim3(256), (size_t)0, st, wave, pcm, planes, image_bytes,
                          a.np_rows, a.np_front, L, T, M));
        EAB_LAUNCH_CHECK("stft_stage_kernel");
    }
    for (int g = 0; g < DFT_GROUPS; ++g) {
        PlaneConvArgs ag = a;
        ag.Whi = reinterpret_cast<const float*>(static_cast<const __half*>(t->dft_hi) + g * DFT_GROUP_HALVES);
        ag.Wlo = reinterpret_cast<const float*>(static_cast<const __half*>(t->dft_lo) + g * DFT_GROUP_HALVES);
        ag.out_coff = g * 128;
        if (launch_conv_staged(ag, st)) { if (own) cudaFreeAsync(planes, st); return 1; }
    }
    if (own) EAB_CUDA(cudaFreeAsync(planes, st));
    return 0;
}

int launch_stft(const float* wave, float* spec, int B, int M, int L, cudaStream_t st, void* scratch, size_t scratch_bytes) {
    if (L < HOP + 1) return fail("stft: need at least 161 samples (reflect padding of 160)");
    if (B <= 0 || M <= 0) return fail("stft: bad shape");
    Tables* t;
    EAB_TRY(get_tables(&t));
    const int T = 1 + L / HOP;
    if (g_stft_tc) { PcmSrc none; memset(&none, 0, sizeof(none)); return launch_stft_tc(t, wave, none, spec, B, M, L, T, st, scratch, scratch_bytes); }
    EAB_TRY(ensure_dynamic_smem(reinterpret_cast<const void*>(stft_kernel), (int)kSmemBytes));
    EAB_TRY(ensure_dynamic_smem(reinterpret_cast<const void*>(istft_kernel), (int)kSmemBytes));
    ProfScope ps("stft", 2.0 * NF * NCOL * (double)B * M * T, 4.0 * ((double)B * M * L + (double)B * T * NF * M * 2), st);
    EAB_CUDA(launch_k(stft_kernel, dim3((T + FR - 1) / FR, B), dim3(THREADS), kSmemBytes, st, wave, spec, (const float*)t->fwd, (const float*)t->win, B, M, L, T));
    EAB_LAUNCH_CHECK("stft_kernel");
    return 0;
}

int launch_stft_pcm16(const short* pcm, const int* order, float* spec, int B, int M, int L, cudaStream_t st, void* scratch,
                      size_t scratch_bytes) {
    if (L < HOP + 1) return fail("stft: need at least 161 samples (reflect padding of 160)");
    if (B <= 0 || M <= 0 || M > 64) return fail("stft (pcm16): bad shape (at most 64 microphones)");
    if (!g_stft_tc) return fail("the 16-bit PCM front door runs on the tensor-core STFT (option stft_tc = 1)");
    Tables* t;
    EAB_TRY(get_tables(&t));
    PcmSrc src;
    memset(&src, 0, sizeof(src));
    src.data = pcm;
    for (int i = 0; i < M; ++i) {
        src.order[i] = order ? order[i] : i;
        if (src.order[i] < 0 || src.order[i] >= M) return fail("pcm16: mic_order entries must be in [0, M)");
    }
    return launch_stft_tc(t, nullptr, src, spec, B, M, L, 1 + L / HOP, st, scratch, scratch_bytes);
}

static int configure_smem() {
    EAB_TRY(ensure_dynamic_smem(reinterpret_cast<const void*>(stft_frame_kernel), (int)kSmemBytes));
    EAB_TRY(ensure_dynamic_smem(reinterpret_cast<const void*>(istft_frame_kernel), (int)kSmemBytes));
    return 0;
}

int launch_stft_frame(const float* hop, const short* hop16, float* prev_hop, float* spec_ring, int spec_RT, const int* step,
                      const int* start, int S, int M, cudaStream_t st) {
    if (S <= 0 || M <= 0 || M > RB) return fail("stft_frame: bad shape (at most 24 microphones)");
    Tables* t;
    EAB_TRY(get_tables(&t));
    EAB_TRY(configure_smem());
    const int SG = RB / M;
    ProfScope ps("stft_frame", 2.0 * NF * NCOL * (double)S * M, 4.0 * ((double)S * M * HOP * 3 + (double)S * NF * M * 2), st);
    EAB_CUDA(launch_k(stft_frame_kernel, dim3((S + SG - 1) / SG), dim3(THREADS), kSmemBytes, st, hop, hop16, prev_hop, spec_ring,
                      spec_RT, step, start, (const float*)t->fwd, (const float*)t->win, S, M, SG));
    EAB_LAUNCH_CHECK("stft_frame_kernel");
    return 0;
}

int launch_istft_frame(const float* frame, float* tail, float* hop_out, short* hop_out16, const int* step, const int* start, int S,
                       cudaStream_t st) {
    if (S <= 0) return fail("istft_frame: bad shape");
    Tables* t;
    EAB_TRY(get_tables(&t));
    EAB_TRY(configure_smem());
    ProfScope ps("istft_frame", 2.0 * NF * NCOL * (double)S, 4.0 * ((double)S * 2 * NF + (double)S * HOP * 3), st);
    EAB_CUDA(launch_k(istft_frame_kernel, dim3((S + RBF - 1) / RBF), dim3(THREADS), kSmemBytes, st, frame, tail, hop_out, hop_out16, step, start,
                      (const float*)t->inv, (const float*)t->win, (const float*)t->ienv, S));
    EAB_LAUNCH_CHECK("istft_frame_kernel");
    return 0;
}

int launch_stream_restart(int* start, const int* step, int idx, cudaStream_t st) {
    EAB_CUDA(launch_k(stream_restart_kernel, dim3(1), dim3(32), (size_t)0, st, start, step, idx));
    EAB_LAUNCH_CHECK("stream_restart_kernel");
    return 0;
}

int launch_step_advance(int* step, cudaStream_t st) {
    EAB_CUDA(launch_k(step_advance_kernel, dim3(1), dim3(32), (size_t)0, st, step));
    EAB_LAUNCH_CHECK("step_advance_kernel");
    return 0;
}

bool g_istft_tc = false;         // iSTFT as a tcgen05 GEMM (option istft_tc; measured slower: 0.245 vs 0.209 ms) / fused fp32 CUDA-core kernel (default)

namespace {

// spec [B][2][T][161] -> rows [B][T][384] = (Re | Im | zeros): the A operand of the inverse-DFT GEMM
__global__ void __launch_bounds__(256) istft_pack_kernel(const float* __restrict__ spec, float* __restrict__ rows, int B, int T) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t n = (size_t)B * T * IDFT_K;
    if (i >= n) return;
    const size_t bt = i / IDFT_K;
    const int kk = (int)(i - bt * IDFT_K);
    const size_t b = bt / T, t = bt - b * T;
    float v = 0.f;
    if (kk < 2 * NF) {
        const int ri = kk >= NF, f = kk - ri * NF;
        v = __ldg(spec + ((b * 2 + ri) * T + t) * NF + f);
    }
    rows[i] = v;
}

// iSTFT as a two-tap tensor-core GEMM (north_star: "windowed DFT-as-GEMM on tensor cores with fused overlap-add"): output row t
// = hop t - 1 = first half of frame t (tap 0) + second half of frame t - 1 (tap 1), K = 2 x 322, N = 160, 3-pass fp16 split;
// window, 1/320, half-spectrum weights and the overlap-add envelope live in the weight images, so the GEMM's rows ARE the
// output samples; a strided device copy drops the unused row 0 of every utterance.
int launch_istft_tc(Tables* t, const float* spec, float* wave, int B, int T, cudaStream_t st) {
    float* rows = nullptr;
    float* tmp = nullptr;
    const size_t nrows = (size_t)B * T;
    EAB_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&rows), nrows * IDFT_K * sizeof(float), st));
    EAB_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&tmp), nrows * HOP * sizeof(float), st));
    auto done = [&](int rc) { cudaFreeAsync(rows, st); cudaFreeAsync(tmp, st); return rc; };
    {
        ProfScope ps("istft", 0.0, 4.0 * (double)B * 2 * T * NF, st, 4.0 * ((double)B * 2 * T * NF + (double)nrows * IDFT_K));
        const size_t n = nrows * IDFT_K;
        if (launch_k(istft_pack_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), (size_t)0, st, spec, rows, B, T) != cudaSuccess)
            return done(fail("istft_pack_kernel launch failed"));
        count_launch();
    }
    for (int g = 0; g < 2; ++g) {
        UmmaConvArgs u;
        memset(&u, 0, sizeof(u));
        u.nsrc = 1;
        u.src[0].x = rows; u.src[0].C = IDFT_K; u.src[0].xf = xform_identity();
        u.B = B; u.T = T; u.Fin = 1; u.E = 1; u.in_stride = 1; u.out_stride = 1; u.out_off = 0; u.Fout = 1;
        u.ntaps = 2; u.dt[0] = 0; u.dt[1] = 1;
        u.nslab = IDFT_SLABS; u.ncoef = IDFT_K; u.npass = 3;
        u.Whi = static_cast<const float*>(t->idft_hi[g]); u.Wlo = static_cast<const float*>(t->idft_lo[g]);
        u.Cout = g ? 32 : 128; u.N = u.Cout;
        u.algo_frac = (float)(2 * NF) / (float)IDFT_K;
        u.out = tmp; u.out_ld = HOP; u.out_coff = g ? 128 : 0;
        u.tiles_per_b = (T + 127) / 128;
        if (launch_conv_umma(u, st)) return done(1);
    }
    // hop h = row h + 1: wave [B][160 (T - 1)] <- tmp [B][T][160] without its first row
    if (cudaMemcpy2DAsync(wave, (size_t)(T - 1) * HOP * sizeof(float), tmp + HOP, (size_t)T * HOP * sizeof(float),
                          (size_t)(T - 1) * HOP * sizeof(float), (size_t)B, cudaMemcpyDeviceToDevice, st) != cudaSuccess)
        return done(fail("istft: strided copy failed"));
    return done(0);
}

}  // namespace

int launch_istft(const float* spec, float* wave, int B, int T, cudaStream_t st, short* wave16) {
    if (T < 2) return fail("istft: need at least 2 frames");
    Tables* t;
    EAB_TRY(get_tables(&t));
    if (g_istft_tc && !wave16) return launch_istft_tc(t, spec, wave, B, T, st);        // (the int16 writer stays with the fused kernel)
    EAB_TRY(ensure_dynamic_smem(reinterpret_cast<const void*>(stft_kernel), (int)kSmemBytes));
    EAB_TRY(ensure_dynamic_smem(reinterpret_cast<const void*>(istft_kernel), (int)kSmemBytes));
    ProfScope ps("istft", 2.0 * NF * NCOL * (double)B * T, 4.0 * ((double)B * 2 * T * NF + (double)B * HOP * (T - 1)), st);
    EAB_CUDA(launch_k(istft_kernel, dim3((T - 1 + IFR - 2) / (IFR - 1), B), dim3(THREADS), kSmemBytes, st, spec, wave, wave16, (const float*)t->inv, (const float*)t->win, (const float*)t->ienv, B, T));
    EAB_LAUNCH_CHECK("istft_kernel");
    return 0;
}

}  // namespace eab
