// Generic fp32 implicit-GEMM convolution on CUDA cores (sm_100a).
//
// One kernel covers every conv-shaped layer of the model for ARBITRARY constructor arguments:
//   Conv2d stride (1,2) plain/gated (EaBNet.py:402,450), ConvTranspose2d stride (1,2) plain/gated with the
//   causal chomp (EaBNet.py:423-425,478-480,624) as two output-parity launches, the TCM 1x1 and dilated
//   Conv1d's (EaBNet.py:549-570) and the w_dnn / 1x1 bf_map linears (EaBNet.py:79-81,593-597).
// The producer's InstanceNorm/BatchNorm + PReLU is applied while the A operand is staged (Xform), the
// channel concat of skip connections is two K-slabs (never materialised), bias / gate / ReLU / residual and
// the per-(b,c) statistics of the *output* are fused into the epilogue.
//
// Tiling: CTA = (16*RPT) output rows x N columns (N = 64*NV, value|gate halves side by side), 256 threads,
// thread tile RPT x 4*NV, K streamed in 32-channel slabs per tap through double-buffered shared memory
// (A: register-staged + transformed, k-major with an XOR swizzle; B: cp.async from the packed weights).
#include "common.cuh"

namespace eab {

namespace {

constexpr int KC = 32;       // channels per K slab
constexpr int NTHREADS = 256;

template <int NV, int RPT, bool GATED>
__global__ void __launch_bounds__(NTHREADS) conv_generic_kernel(const ConvArgs a) {
    constexpr int TM = 16 * RPT;
    constexpr int N = 64 * NV;
    constexpr int NLD = TM / 32;                 // float4 A loads per thread per slab
    extern __shared__ __align__(16) float smem[];
    float* As = smem;                            // [2][KC][TM]
    float* Bs = As + 2 * KC * TM;                // [2][KC][N]
    float* coef = Bs + 2 * KC * N;               // [3][Ctot]

    const int tid = threadIdx.x;
    // streaming: rows of all streams share one row space (r = stream * E + e, one frame each); offline: rows (t, e) of item b
    const bool streaming = a.step != nullptr;
    const int step = streaming ? *a.step : 0;
    const int b = streaming ? 0 : blockIdx.y;
    const int rows_per_b = streaming ? a.B * a.E : a.T * a.E;
    const int row0 = blockIdx.x * TM;
    const int C0 = a.src[0].C;
    const int C1 = a.nsrc > 1 ? a.src[1].C : 0;
    const int Ctot = C0 + C1;

    for (int i = tid; i < Ctot; i += NTHREADS) {
        const int s = i < C0 ? 0 : 1;
        const int c = s ? i - C0 : i;
        float cs, ch, ca;
        xform_coeffs(a.src[s].xf, b, a.src[s].C, c, cs, ch, ca);
        coef[i] = cs;
        coef[Ctot + i] = ch;
        coef[2 * Ctot + i] = ca;
    }

    // loader geometry: thread -> (row = (tid>>3) + 32*i, channel quad c4 = tid&7)
    const int c4 = tid & 7;
    int lt[NLD], le[NLD];
    bool lvalid[NLD];
#pragma unroll
    for (int i = 0; i < NLD; ++i) {
        const int r = row0 + (tid >> 3) + 32 * i;
        lvalid[i] = r < rows_per_b;
        lt[i] = r / a.E;                          // streaming: the stream index
        le[i] = r - lt[i] * a.E;
    }

    // compute geometry
    const int cg = tid & 15;
    const int rg = tid >> 4;

    float acc[RPT][4 * NV];
#pragma unroll
    for (int r = 0; r < RPT; ++r)
#pragma unroll
        for (int j = 0; j < 4 * NV; ++j) acc[r][j] = 0.f;

    const int nslab0 = (C0 + KC - 1) / KC;
    const int nslab1 = (C1 + KC - 1) / KC;
    const int nchunks = a.ntaps * (nslab0 + nslab1);

    __syncthreads();     // coefficients visible

    float4 areg[NLD];

    auto chunk_decode = [&](int ch, int& tap, int& s, int& c0) {
        const int per_tap = nslab0 + nslab1;
        tap = ch / per_tap;
        int r = ch - tap * per_tap;
        if (r < nslab0) { s = 0; c0 = r * KC; } else { s = 1; c0 = (r - nslab0) * KC; }
    };

    auto load_a = [&](int ch) {
        int tap, s, c0;
        chunk_decode(ch, tap, s, c0);
        const ConvSrc& src = a.src[s];
        const int C = src.C;
        const int soff = s ? C0 : 0;
        const int c = c0 + c4 * 4;
        const int dtv = a.dt[tap], dfv = a.df[tap];
        const int prelu = src.xf.prelu;
        const bool vec = (C & 3) == 0;
#pragma unroll
        for (int i = 0; i < NLD; ++i) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            const int fi = le[i] * a.in_stride + dfv;
            size_t frame;                         // index of the source frame in units of [Fin][C]
            bool tok;
            if (streaming) {
                const int n = step - dtv;
                tok = n >= (a.start ? __ldg(a.start + lt[i]) : 0) && n <= step;       // lt = stream index when streaming
                frame = (size_t)lt[i] * src.RT + ring_slot(n, src.RT);
            } else {
                const int tt = lt[i] - dtv;
                tok = tt >= 0 && tt < a.T;
                frame = (size_t)b * a.T + tt;
            }
            if (lvalid[i] && tok && fi >= 0 && fi < a.Fin && c < C) {
                const float* p = src.x + (frame * a.Fin + fi) * C + c;
                float x[4] = {0.f, 0.f, 0.f, 0.f};
                if (vec) {
                    const float4 q = __ldg(reinterpret_cast<const float4*>(p));
                    x[0] = q.x; x[1] = q.y; x[2] = q.z; x[3] = q.w;
                } else {
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        if (c + q < C) x[q] = __ldg(p + q);
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if (c + q < C) {
                        const int ci = soff + c + q;
                        x[q] = xform_apply(x[q], coef[ci], coef[Ctot + ci], coef[2 * Ctot + ci], prelu);
                    }
                }
                v = make_float4(x[0], x[1], x[2], x[3]);
            }
            areg[i] = v;
        }
    };

    auto store_a = [&](int buf) {
        float* A = As + buf * KC * TM;
#pragma unroll
        for (int i = 0; i < NLD; ++i) {
            const int row = ((tid >> 3) + 32 * i) ^ (c4 << 2);
            A[(c4 * 4 + 0) * TM + row] = areg[i].x;
            A[(c4 * 4 + 1) * TM + row] = areg[i].y;
            A[(c4 * 4 + 2) * TM + row] = areg[i].z;
            A[(c4 * 4 + 3) * TM + row] = areg[i].w;
        }
    };

    auto load_b = [&](int ch, int buf) {
        int tap, s, c0;
        chunk_decode(ch, tap, s, c0);
        const int C = a.src[s].C;
        const int soff = s ? C0 : 0;
        float* B = Bs + buf * KC * N;
        const float* wbase = a.W + ((size_t)tap * Ctot + soff + c0) * N;
#pragma unroll
        for (int i = 0; i < 2 * NV; ++i) {
            const int idx = tid + NTHREADS * i;
            const int row = idx / (N / 4);
            const int col4 = idx - row * (N / 4);
            const bool ok = c0 + row < C;
            cp_async16(B + row * N + col4 * 4, ok ? wbase + (size_t)row * N + col4 * 4 : a.W, ok);
        }
        cp_async_commit();
    };

    load_a(0);
    load_b(0, 0);
    store_a(0);
    cp_async_wait_all();
    __syncthreads();

    int buf = 0;
    for (int ch = 0; ch < nchunks; ++ch) {
        const bool more = ch + 1 < nchunks;
        if (more) {
            load_a(ch + 1);
            load_b(ch + 1, buf ^ 1);
        }
        const float* A = As + buf * KC * TM;
        const float* B = Bs + buf * KC * N;
#pragma unroll 4
        for (int kk = 0; kk < KC; ++kk) {
            const int sw = ((kk >> 2) & 7) << 2;
            float av[RPT];
#pragma unroll
            for (int u = 0; u < RPT / 4; ++u) {
                const float4 t = *reinterpret_cast<const float4*>(A + kk * TM + ((rg * RPT + 4 * u) ^ sw));
                av[4 * u + 0] = t.x; av[4 * u + 1] = t.y; av[4 * u + 2] = t.z; av[4 * u + 3] = t.w;
            }
            float bv[4 * NV];
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                const float4 t = *reinterpret_cast<const float4*>(B + kk * N + v * 64 + cg * 4);
                bv[4 * v + 0] = t.x; bv[4 * v + 1] = t.y; bv[4 * v + 2] = t.z; bv[4 * v + 3] = t.w;
            }
#pragma unroll
            for (int r = 0; r < RPT; ++r)
#pragma unroll
                for (int j = 0; j < 4 * NV; ++j) acc[r][j] = fmaf(av[r], bv[j], acc[r][j]);
        }
        if (more) {
            store_a(buf ^ 1);
            cp_async_wait_all();
        }
        __syncthreads();
        buf ^= 1;
    }

    // ------------------------------------------------------------------ epilogue
    constexpr bool gated = GATED;
    constexpr int nvv = GATED ? NV / 2 : NV;      // value slots (gate slot = v + nvv)
    constexpr int NVV_MAX = nvv;
    float psum[2][4 * NVV_MAX], psq[2][4 * NVV_MAX];
#pragma unroll
    for (int s = 0; s < 2; ++s)
#pragma unroll
        for (int j = 0; j < 4 * NVV_MAX; ++j) { psum[s][j] = 0.f; psq[s][j] = 0.f; }

    const bool vec_out = (a.Cout & 3) == 0;
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
        const int row = row0 + rg * RPT + r;
        if (row >= rows_per_b) continue;
        const int t = row / a.E;                  // streaming: the stream index
        const int e = row - t * a.E;
        const int fo = e * a.out_stride + a.out_off;
        const size_t oframe = streaming ? (size_t)t * a.out_RT + ring_slot(step, a.out_RT) : (size_t)b * a.T + t;
        const size_t obase = (oframe * a.Fout + fo) * a.Cout;
        const size_t rbase = streaming ? (((size_t)t * a.resid_RT + ring_slot(step, a.resid_RT)) * a.Fout + fo) * a.Cout : obase;
#pragma unroll
        for (int v = 0; v < NVV_MAX; ++v) {
            const int col = v * 64 + cg * 4;
            if (col >= a.Cout) continue;
            float o[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float val = acc[r][4 * v + q];
                if (a.bias) val += a.bias[col + q];
                if (gated) {
                    const int gslot = GATED ? v + nvv : v;
                    float g = acc[r][4 * gslot + q];
                    if (a.bias) g += a.bias[a.gate_off + col + q];
                    val *= sigmoid_f(g);
                }
                if (a.relu) val = fmaxf(val, 0.f);
                o[q] = val;
            }
            if (a.resid) {
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (col + q < a.Cout) o[q] += __ldg(a.resid + rbase + col + q);
            }
            if (vec_out) {
                *reinterpret_cast<float4*>(a.out + obase + col) = make_float4(o[0], o[1], o[2], o[3]);
            } else {
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (col + q < a.Cout) a.out[obase + col + q] = o[q];
            }
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                if (s >= a.nstats) continue;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    float u = o[q];
                    if (a.stat_alpha[s] && col + q < a.Cout) u = prelu_f(u, a.stat_alpha[s][col + q]);
                    psum[s][4 * v + q] += u;
                    psq[s][4 * v + q] += u * u;
                }
            }
        }
    }

    if (a.nstats > 0) {
        // lanes l and l^16 share cg; fold them, then one partial per warp -> smem -> double atomics
        float* red = smem;                         // [8 warps][2][N][2]   (A/B tiles are dead)
        const int warp = tid >> 5, lane = tid & 31;
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            if (s >= a.nstats) continue;
#pragma unroll
            for (int j = 0; j < 4 * NVV_MAX; ++j) {
                float x = psum[s][j], y = psq[s][j];
                x += __shfl_xor_sync(0xffffffffu, x, 16);
                y += __shfl_xor_sync(0xffffffffu, y, 16);
                if (lane < 16) {
                    const int col = (j >> 2) * 64 + lane * 4 + (j & 3);
                    red[((warp * 2 + s) * N + col) * 2 + 0] = x;
                    red[((warp * 2 + s) * N + col) * 2 + 1] = y;
                }
            }
        }
        __syncthreads();
        for (int i = tid; i < a.nstats * a.Cout; i += NTHREADS) {
            const int s = i / a.Cout;
            const int col = i - s * a.Cout;
            float x = 0.f, y = 0.f;
#pragma unroll
            for (int w = 0; w < 8; ++w) {
                x += red[((w * 2 + s) * N + col) * 2 + 0];
                y += red[((w * 2 + s) * N + col) * 2 + 1];
            }
            double* dst = a.stats[s] + ((size_t)b * a.Cout + col) * 2;
            atomicAdd(dst, (double)x);
            atomicAdd(dst + 1, (double)y);
        }
    }
}

template <int NV, int RPT, bool GATED>
int launch_inst(const ConvArgs& a, cudaStream_t st) {
    constexpr int TM = 16 * RPT;
    constexpr int N = 64 * NV;
    const int Ctot = a.src[0].C + (a.nsrc > 1 ? a.src[1].C : 0);
    size_t smem = (size_t)(2 * KC * TM + 2 * KC * N + 3 * Ctot) * sizeof(float);
    const size_t red = (size_t)8 * 2 * N * 2 * sizeof(float);
    if (smem < red) smem = red;
    EAB_TRY(ensure_dynamic_smem(reinterpret_cast<const void*>(conv_generic_kernel<NV, RPT, GATED>), (int)smem));
    const bool streaming = a.step != nullptr;
    const int rows = streaming ? a.B * a.E : a.T * a.E;
    dim3 grid((rows + TM - 1) / TM, streaming ? 1 : a.B);
    const double pos = streaming ? (double)rows : (double)a.B * rows;
    const int ncol = a.Cout * (GATED ? 2 : 1);
    ProfScope ps("conv_generic", 2.0 * pos * a.ntaps * Ctot * ncol * a.algo_frac,
                 4.0 * (pos * a.in_stride * Ctot / (a.out_stride > 1 ? 2.0 : 1.0) + pos * a.Cout * (a.resid ? 2 : 1) +
                        (double)a.ntaps * Ctot * ncol), st);
    EAB_CUDA(launch_k(conv_generic_kernel<NV, RPT, GATED>, grid, dim3(NTHREADS), smem, st, a));
    EAB_LAUNCH_CHECK("conv_generic_kernel");
    return 0;
}

}  // namespace

int launch_conv(const ConvArgs& a, cudaStream_t st) {
    if (a.ntaps < 1 || a.ntaps > kMaxTaps) return fail("conv: unsupported tap count");
    const int Ctot = a.src[0].C + (a.nsrc > 1 ? a.src[1].C : 0);
    if (Ctot > kMaxCin) return fail("conv: too many input channels");
    if (a.N != 64 && a.N != 128 && a.N != 256) return fail("conv: unsupported output channel count (max 128 gated / 256 plain)");
    if (a.gate_off > 0 && a.gate_off * 2 != a.N) return fail("conv: bad gate offset");
    if (a.B <= 0 || a.T <= 0 || a.E <= 0) return 0;
    if (a.step) {
        if (a.T != 1 || a.nstats != 0) return fail("conv: a streaming launch computes one frame and takes no statistics");
        for (int i = 0; i < a.nsrc; ++i) {
            if (a.src[i].xf.affine == 1) return fail("conv: streaming needs static normalisation (norm_type 'BN')");
            int back = 0;
            for (int k = 0; k < a.ntaps; ++k) { if (a.dt[k] < 0) return fail("conv: streaming needs causal taps"); back = a.dt[k] > back ? a.dt[k] : back; }
            if (a.src[i].RT < back + 1) return fail("conv: source ring shorter than the receptive field");
        }
        if (a.out_RT < 1 || (a.resid && a.resid_RT < 1)) return fail("conv: bad output ring");
    }
    const bool g = a.gate_off > 0;
    // streaming launches have few rows (streams x E): 64-row tiles put twice as many CTAs in flight, which is what a
    // latency-bound step wants; the offline cold path keeps the 128-row tiles
    if (a.step && (long long)a.B * a.E <= 148ll * 128 * 2) {
        switch (a.N / 64) {
            case 1: if (!g) return launch_inst<1, 4, false>(a, st); break;
            case 2: return g ? launch_inst<2, 4, true>(a, st) : launch_inst<2, 4, false>(a, st);
            default: break;
        }
    }
    switch (a.N / 64) {
        case 1: return g ? fail("conv: gated layer needs N >= 128") : launch_inst<1, 8, false>(a, st);
        case 2: return g ? launch_inst<2, 8, true>(a, st) : launch_inst<2, 8, false>(a, st);
        case 4: return g ? launch_inst<4, 4, true>(a, st) : launch_inst<4, 4, false>(a, st);
        default: return fail("conv: unsupported output channel count");
    }
}

}  // namespace eab
