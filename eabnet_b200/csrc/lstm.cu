// One nn.LSTM(E -> 64, batch_first) layer of LSTM_BF (EaBNet.py:591-592, 610-611) over the B*F independent
// (b,f) sequences, zero initial state, gate order i,f,g,o, b_ih + b_hh.
//
// Persistent-in-time kernel: a CTA owns 4*SPT sequences for all T steps.  W_ih and W_hh stay resident in
// shared memory (fp32, [k][unit][gate] so that one 128-bit load feeds the four gates of a unit), the cell
// state lives in registers, h is exchanged through shared memory once per step, and the next frame's input
// is prefetched from HBM while the current step's gates are computed.  For the first layer the producer's
// InstanceNorm/BatchNorm + PReLU and the head's LayerNorm(E) (EaBNet.py:598,608) are applied on load.
// Thread (j = tid & 63, sg = tid >> 6) computes the 4 gates of hidden unit j for SPT sequences.
#include "common.cuh"

namespace eab {

namespace {

constexpr int H = 64;

template <int SPT, int E>
__global__ void __launch_bounds__(256, 1) lstm_kernel(const LstmArgs a) {
    constexpr int S = 4 * SPT;                    // sequences per CTA
    constexpr int SP = ((SPT + 3) / 4) * 4;       // padded per-group stride (128-bit loads)
    constexpr int EPT = E / 4;                    // input channels per loader thread
    static_assert(E % 16 == 0, "E must be a multiple of 16");
    static_assert(S <= 64, "loader mapping covers 64 sequences");

    extern __shared__ __align__(16) float smem[];
    float4* Wx = reinterpret_cast<float4*>(smem);           // [E][64]
    float4* Wh = Wx + E * H;                                // [64][64]
    float* xin = reinterpret_cast<float*>(Wh + H * H);      // [E][4][SP]
    float* hbuf = xin + E * 4 * SP;                         // [64][4][SP]
    float* coef = hbuf + H * 4 * SP;                        // [2 batches][3][E]
    float* lng = coef + 2 * 3 * E;                          // [E]
    float* lnb = lng + E;                                   // [E]

    const int tid = threadIdx.x;
    const int j = tid & 63;
    const int sg = tid >> 6;
    const int NQ = a.B * a.F;
    const int q0 = blockIdx.x * S;
    const int b0 = q0 / a.F;

    {
        const float4* gx = reinterpret_cast<const float4*>(a.Wx);
        const float4* gh = reinterpret_cast<const float4*>(a.Wh);
        for (int i = tid; i < E * H; i += 256) Wx[i] = __ldg(gx + i);
        for (int i = tid; i < H * H; i += 256) Wh[i] = __ldg(gh + i);
        for (int i = tid; i < 2 * E; i += 256) {
            const int bb = i / E, c = i - bb * E;
            float cs = 1.f, ch = 0.f, ca = 1.f;
            if (b0 + bb < a.B) xform_coeffs(a.src.xf, b0 + bb, E, c, cs, ch, ca);
            coef[(bb * 3 + 0) * E + c] = cs;
            coef[(bb * 3 + 1) * E + c] = ch;
            coef[(bb * 3 + 2) * E + c] = ca;
        }
        for (int i = tid; i < E; i += 256) {
            lng[i] = a.layer_norm ? a.ln_g[i] : 1.f;
            lnb[i] = a.layer_norm ? a.ln_b[i] : 0.f;
        }
        for (int i = tid; i < H * 4 * SP; i += 256) hbuf[i] = 0.f;
        for (int i = tid; i < E * 4 * SP; i += 256) xin[i] = 0.f;
    }
    const bool streaming = a.step != nullptr;
    const int step = streaming ? *a.step : 0;
    const int src_slot = streaming ? ring_slot(step, a.src.RT) : 0;
    const int src_T = streaming ? a.src.RT : a.T;        // frames per batch item in the source / output tensors
    const int out_T = streaming ? a.out_RT : a.T;
    const int out_slot = streaming ? ring_slot(step, a.out_RT) : 0;

    // loader role: 4 threads per sequence, EPT consecutive channels each
    const int ls = tid >> 2;
    const int part = tid & 3;
    const int lq = q0 + ls;
    const bool lactive = ls < S && lq < NQ;
    const int lb = lactive ? lq / a.F : 0;
    const int lf = lactive ? lq - lb * a.F : 0;
    const float* lsrc = a.src.x + (((size_t)lb * src_T + src_slot) * a.F + lf) * E + part * EPT;
    const size_t step_stride = (size_t)a.F * E;
    const float* lcoef = coef + (lactive ? (lb - b0) : 0) * 3 * E;
    const int lsg = ls / SPT, lss = ls - lsg * SPT;
    const int prelu = a.src.xf.prelu;

    float xr[EPT];
#pragma unroll
    for (int i = 0; i < EPT; ++i) xr[i] = 0.f;
    auto prefetch = [&](int t) {
        if (lactive) {
            const float4* p = reinterpret_cast<const float4*>(lsrc + (size_t)t * step_stride);
#pragma unroll
            for (int i = 0; i < EPT / 4; ++i) {
                const float4 v = __ldg(p + i);
                xr[4 * i + 0] = v.x; xr[4 * i + 1] = v.y; xr[4 * i + 2] = v.z; xr[4 * i + 3] = v.w;
            }
        }
    };
    auto publish = [&]() {                       // transform (+ LayerNorm) and write to xin
        // every lane executes the shuffles (full-mask); only the final store is predicated
        float v[EPT];
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < EPT; ++i) {
            const int c = part * EPT + i;
            v[i] = lactive ? xform_apply(xr[i], lcoef[c], lcoef[E + c], lcoef[2 * E + c], prelu) : 0.f;
            sum += v[i];
        }
        if (a.layer_norm) {                      // uniform branch
            sum += __shfl_xor_sync(0xffffffffu, sum, 1);
            sum += __shfl_xor_sync(0xffffffffu, sum, 2);
            const float mean = sum * (1.f / E);
            float sq = 0.f;
#pragma unroll
            for (int i = 0; i < EPT; ++i) { const float d = v[i] - mean; sq += d * d; }
            sq += __shfl_xor_sync(0xffffffffu, sq, 1);
            sq += __shfl_xor_sync(0xffffffffu, sq, 2);
            const float rstd = rsqrtf(sq * (1.f / E) + 1e-5f);
#pragma unroll
            for (int i = 0; i < EPT; ++i) {
                const int c = part * EPT + i;
                v[i] = (v[i] - mean) * rstd * lng[c] + lnb[c];
            }
        }
        if (ls < S) {
#pragma unroll
            for (int i = 0; i < EPT; ++i) xin[((part * EPT + i) * 4 + lsg) * SP + lss] = lactive ? v[i] : 0.f;
        }
    };

    // compute role
    float cst[SPT];
#pragma unroll
    for (int s = 0; s < SPT; ++s) cst[s] = 0.f;
    const float4 bias = __ldg(reinterpret_cast<const float4*>(a.bias) + j);
    size_t obase[SPT];
    bool oval[SPT];
#pragma unroll
    for (int s = 0; s < SPT; ++s) {
        const int q = q0 + sg * SPT + s;
        oval[s] = q < NQ;
        const int bq = oval[s] ? q / a.F : 0;
        const int fq = oval[s] ? q - bq * a.F : 0;
        obase[s] = (((size_t)bq * out_T + out_slot) * a.F + fq) * H + j;
    }
    const size_t ostep = (size_t)a.F * H;

    __syncthreads();
    if (streaming && step > 0) {                 // carried state (zero at the first frame, like the offline path)
#pragma unroll
        for (int s = 0; s < SPT; ++s) {
            const int q = q0 + sg * SPT + s;
            if (oval[s]) {
                cst[s] = a.c_state[(size_t)q * H + j];
                hbuf[(j * 4 + sg) * SP + s] = a.h_state[(size_t)q * H + j];
            }
        }
    }
    prefetch(0);
    publish();
    __syncthreads();

    for (int t = 0; t < a.T; ++t) {
        if (t + 1 < a.T) prefetch(t + 1);

        float gi[SPT], gf[SPT], gg[SPT], go[SPT];
#pragma unroll
        for (int s = 0; s < SPT; ++s) { gi[s] = bias.x; gf[s] = bias.y; gg[s] = bias.z; go[s] = bias.w; }

#pragma unroll 4
        for (int k = 0; k < E; ++k) {
            const float4 w = Wx[k * H + j];
            float xs[SP];
#pragma unroll
            for (int u = 0; u < SP / 4; ++u) {
                const float4 v = *reinterpret_cast<const float4*>(xin + (k * 4 + sg) * SP + 4 * u);
                xs[4 * u + 0] = v.x; xs[4 * u + 1] = v.y; xs[4 * u + 2] = v.z; xs[4 * u + 3] = v.w;
            }
#pragma unroll
            for (int s = 0; s < SPT; ++s) {
                gi[s] = fmaf(w.x, xs[s], gi[s]);
                gf[s] = fmaf(w.y, xs[s], gf[s]);
                gg[s] = fmaf(w.z, xs[s], gg[s]);
                go[s] = fmaf(w.w, xs[s], go[s]);
            }
        }
#pragma unroll 4
        for (int k = 0; k < H; ++k) {
            const float4 w = Wh[k * H + j];
            float hs[SP];
#pragma unroll
            for (int u = 0; u < SP / 4; ++u) {
                const float4 v = *reinterpret_cast<const float4*>(hbuf + (k * 4 + sg) * SP + 4 * u);
                hs[4 * u + 0] = v.x; hs[4 * u + 1] = v.y; hs[4 * u + 2] = v.z; hs[4 * u + 3] = v.w;
            }
#pragma unroll
            for (int s = 0; s < SPT; ++s) {
                gi[s] = fmaf(w.x, hs[s], gi[s]);
                gf[s] = fmaf(w.y, hs[s], gf[s]);
                gg[s] = fmaf(w.z, hs[s], gg[s]);
                go[s] = fmaf(w.w, hs[s], go[s]);
            }
        }

        float hn[SPT];
#pragma unroll
        for (int s = 0; s < SPT; ++s) {
            cst[s] = sigmoid_f(gf[s]) * cst[s] + sigmoid_f(gi[s]) * tanh_f(gg[s]);
            hn[s] = sigmoid_f(go[s]) * tanh_f(cst[s]);
        }
        __syncthreads();                         // every thread is done reading xin / hbuf of step t
#pragma unroll
        for (int s = 0; s < SPT; ++s) {
            hbuf[(j * 4 + sg) * SP + s] = hn[s];
            if (oval[s]) a.out[obase[s] + (size_t)t * ostep] = hn[s];
        }
        if (t + 1 < a.T) publish();
        __syncthreads();
    }
    if (streaming) {
#pragma unroll
        for (int s = 0; s < SPT; ++s) {
            const int q = q0 + sg * SPT + s;
            if (oval[s]) {
                a.c_state[(size_t)q * H + j] = cst[s];
                a.h_state[(size_t)q * H + j] = hbuf[(j * 4 + sg) * SP + s];
            }
        }
    }
}

template <int SPT, int E>
int launch_inst(const LstmArgs& a, cudaStream_t st) {
    constexpr int S = 4 * SPT;
    constexpr int SP = ((SPT + 3) / 4) * 4;
    const size_t smem = ((size_t)(E * H + H * H) * 4 + (size_t)(E + H) * 4 * SP + 2 * 3 * E + 2 * E) * sizeof(float);
    EAB_TRY(ensure_dynamic_smem(reinterpret_cast<const void*>(lstm_kernel<SPT, E>), (int)smem));
    const int NQ = a.B * a.F;
    ProfScope ps("lstm", 2.0 * NQ * a.T * (E + H) * 4.0 * H, 4.0 * NQ * a.T * (E + H), st);
    EAB_CUDA(launch_k(lstm_kernel<SPT, E>, dim3((NQ + S - 1) / S), dim3(256), smem, st, a));
    EAB_LAUNCH_CHECK("lstm_kernel");
    return 0;
}

template <int E>
int launch_e(const LstmArgs& a, cudaStream_t st) {
    // pick the sequences-per-thread that minimises (waves over the SMs) x (work per step)
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int NQ = a.B * a.F;
    const int cand[3] = {2, 4, 9};
    int best = 2;
    long best_cost = -1;
    for (int i = 0; i < 3; ++i) {
        const int S = 4 * cand[i];
        const long ctas = (NQ + S - 1) / S;
        const long waves = (ctas + sms - 1) / sms;
        const long cost = waves * (cand[i] * 4 + 6);        // + fixed per-step overhead
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = cand[i]; }
    }
    if (a.F < 36 && best == 9) best = 4;                     // a CTA may straddle at most two batch items
    if (a.F < 16 && best == 4) best = 2;
    if (E > 64 && best == 9) best = 4;                       // shared-memory budget with a 128-wide W_ih
    switch (best) {
        case 2: return launch_inst<2, E>(a, st);
        case 4: return launch_inst<4, E>(a, st);
        default: return launch_inst<9, E>(a, st);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Streaming step on the tensor cores (one frame per stream): the gate pre-activations [x_t, h_{t-1}] [W_ih ; W_hh]^T + b are
// one tcgen05 GEMM over all (stream, f) rows (conv_umma, two sources, K = 64 + 64, N = 256: run_pointwise in model_run.cu);
// these two bandwidth-bound kernels do what surrounds it.
//   lstm_ln_frame_kernel    x_t of every stream from its ring slot -> norm + PReLU (Xform) -> LayerNorm(64) -> [rows][64]
//   lstm_cell_frame_kernel  gates [rows][256] (i | f | g | o, 64 each) + c_{t-1} -> c_t, h_t -> state and the output ring slot
__global__ void __launch_bounds__(256) lstm_ln_frame_kernel(const LstmFrameArgs a) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);         // a warp per (stream, f) row, 2 channels per lane
    if (row >= a.rows) return;
    const int lane = threadIdx.x & 31;
    const int s = row / a.F, f = row - s * a.F;
    const int n = __ldg(a.step);
    const float* xp = a.x + (((size_t)s * a.x_RT + ring_slot(n, a.x_RT)) * a.F + f) * 64 + lane * 2;
    float2 v = *reinterpret_cast<const float2*>(xp);
    float vv[2] = {v.x, v.y};
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        float cs, ch, ca;
        xform_coeffs(a.xf, s, 64, lane * 2 + k, cs, ch, ca);
        float z = vv[k];
        if (a.xf.prelu == 1) { z = prelu_f(z, ca); z = fmaf(z, cs, ch); }
        else { z = fmaf(z, cs, ch); if (a.xf.prelu == 2) z = prelu_f(z, ca); }
        vv[k] = z;
    }
    if (a.layer_norm) {
        float sum = vv[0] + vv[1];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        const float mean = sum * (1.f / 64.f);
        const float d0 = vv[0] - mean, d1 = vv[1] - mean;
        float sq = d0 * d0 + d1 * d1;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
        const float rstd = rsqrtf(sq * (1.f / 64.f) + 1e-5f);
        vv[0] = d0 * rstd * __ldg(a.ln_g + lane * 2) + __ldg(a.ln_b + lane * 2);
        vv[1] = d1 * rstd * __ldg(a.ln_g + lane * 2 + 1) + __ldg(a.ln_b + lane * 2 + 1);
    }
    *reinterpret_cast<float2*>(a.out + (size_t)row * 64 + lane * 2) = make_float2(vv[0], vv[1]);
}

__global__ void __launch_bounds__(256) lstm_cell_frame_kernel(const LstmFrameArgs a) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;      // (row, 4 hidden units)
    if (i >= (long long)a.rows * 16) return;
    const int row = (int)(i >> 4), j = (int)(i & 15) * 4;
    const float* g = a.gates + (size_t)row * a.gates_ld;
    const float4 gi = *reinterpret_cast<const float4*>(g + j), gf = *reinterpret_cast<const float4*>(g + 64 + j),
                 gg = *reinterpret_cast<const float4*>(g + 128 + j), go = *reinterpret_cast<const float4*>(g + 192 + j);
    float4 c = *reinterpret_cast<const float4*>(a.c_state + (size_t)row * 64 + j);
    float4 h;
    c.x = sigmoid_f(gf.x) * c.x + sigmoid_f(gi.x) * tanh_f(gg.x); h.x = sigmoid_f(go.x) * tanh_f(c.x);
    c.y = sigmoid_f(gf.y) * c.y + sigmoid_f(gi.y) * tanh_f(gg.y); h.y = sigmoid_f(go.y) * tanh_f(c.y);
    c.z = sigmoid_f(gf.z) * c.z + sigmoid_f(gi.z) * tanh_f(gg.z); h.z = sigmoid_f(go.z) * tanh_f(c.z);
    c.w = sigmoid_f(gf.w) * c.w + sigmoid_f(gi.w) * tanh_f(gg.w); h.w = sigmoid_f(go.w) * tanh_f(c.w);
    *reinterpret_cast<float4*>(a.c_state + (size_t)row * 64 + j) = c;
    *reinterpret_cast<float4*>(a.h_state + (size_t)row * 64 + j) = h;
    const int s = row / a.F, f = row - s * a.F;
    const int n = __ldg(a.step);
    *reinterpret_cast<float4*>(a.out + (((size_t)s * a.out_RT + ring_slot(n, a.out_RT)) * a.F + f) * 64 + j) = h;
}

}  // namespace

int launch_lstm_ln_frame(const LstmFrameArgs& a, cudaStream_t st) {
    if (a.rows <= 0) return 0;
    if (!a.step || !a.x || !a.out || a.x_RT < 1) return fail("lstm_ln_frame: bad arguments");
    ProfScope ps("lstm_frame", 0.0, 8.0 * a.rows * 64, st);
    EAB_CUDA(launch_k(lstm_ln_frame_kernel, dim3((a.rows + 7) / 8), dim3(256), (size_t)0, st, a));
    EAB_LAUNCH_CHECK("lstm_ln_frame_kernel");
    return 0;
}

int launch_lstm_cell_frame(const LstmFrameArgs& a, cudaStream_t st) {
    if (a.rows <= 0) return 0;
    if (!a.step || !a.gates || !a.c_state || !a.h_state || !a.out || a.out_RT < 1 || a.gates_ld < 256) return fail("lstm_cell_frame: bad arguments");
    ProfScope ps("lstm_frame", 0.0, 4.0 * a.rows * (256 + 4 * 64), st);
    EAB_CUDA(launch_k(lstm_cell_frame_kernel, dim3((unsigned)(((long long)a.rows * 16 + 255) / 256)), dim3(256), (size_t)0, st, a));
    EAB_LAUNCH_CHECK("lstm_cell_frame_kernel");
    return 0;
}

int launch_lstm(const LstmArgs& a, cudaStream_t st) {
    if (a.F < 8) return fail("lstm: F < 8 is not supported");
    if (a.step && (a.T != 1 || !a.h_state || !a.c_state || a.src.RT < 1 || a.out_RT < 1 || a.src.xf.affine == 1))
        return fail("lstm: bad streaming launch");
    switch (a.E) {
        case 32: return launch_e<32>(a, st);
        case 64: return launch_e<64>(a, st);
        case 128: return launch_e<128>(a, st);
        default: return fail("lstm head: embed_dim must be 32, 64 or 128");
    }
}

}  // namespace eab
