// Streaming form of the whole squeezed-TCM stack (EaBNet.py:99-106, 506-578): ONE launch per frame step.
//
// Offline, a TCM is three GEMM launches over 38 464 rows.  In a streaming step every stream contributes a single row,
// so the q*p = 18 TCMs are 54 matrix-vector products per stream that depend on each other: launch latency, not
// arithmetic, is the cost (measured: 54 launches x ~0.1 ms of the 7.3 ms step).  Streams are independent, so a CTA
// keeps SPC streams' residual vector x[256] in shared memory and walks all TCMs itself:
//     y = W_in x                       (256 -> 64, no bias)            -> written to this TCM's history ring
//     zL = sum_k W_L,k  nL(y[n - dt_k]),  zR = sum_k W_R,k nR(y[n - dt_k])   (n* = PReLU -> BatchNorm, zeros for n - dt < 0)
//     x += W_out nO(zL * sigmoid(zR))  (64 -> 256)
// and accumulates the group outputs (x after every p-th TCM).  Weights stream from L2 (fp32, the CUDA-core layouts).
#include "common.cuh"

namespace eab {

namespace {

constexpr int SPC = 2;          // streams per CTA
constexpr int CD = 64;          // squeezed channels
constexpr int DF = 256;         // feature channels
constexpr int NT = 1024;        // the step is L2-latency bound: many threads = many weight loads in flight
constexpr int KG1 = NT / CD;            // 16 K groups of the squeeze   (16 k each)
constexpr int KG2 = NT / (2 * CD);      // 8 K groups of the dilated convs
constexpr int KG3 = NT / DF;            // 4 K groups of the expand     (16 k each)

__device__ __forceinline__ float prelu_norm(float v, float a, float s, float h) {
    v = v > 0.f ? v : a * v;
    return fmaf(v, s, h);
}

__global__ void __launch_bounds__(NT) tcm_stream_kernel(const TcmStreamArgs a) {
    __shared__ float xs[SPC][DF];
    __shared__ float accs[SPC][DF];
    __shared__ float ys[SPC][CD];
    __shared__ float red[NT * SPC];            // partial sums of the current phase: [K group][stream][output]
    __shared__ float u[2][8][SPC][CD];         // [branch][tap][stream][channel] normalised dilated-conv inputs
    __shared__ float uo[SPC][CD];
    const int tid = threadIdx.x;
    const int step = *a.step;
    const int s0 = blockIdx.x * SPC;
    const int xslot = ring_slot(step, a.x_RT);
    for (int i = tid; i < SPC * DF; i += NT) {
        const int s = i / DF, j = i - s * DF;
        xs[s][j] = s0 + s < a.S ? a.x[((size_t)(s0 + s) * a.x_RT + xslot) * DF + j] : 0.f;
        accs[s][j] = 0.f;
    }
    __syncthreads();
    const int kper2 = (a.kd * CD + KG2 - 1) / KG2;         // (tap, channel) pairs per K group of the dilated convs
    // The weights do not depend on the activations: every phase fetches the NEXT phase's first PF weights of this thread before
    // its own reduction and block barriers, so that phase starts on registers instead of an L2 round trip (54 dependent phases).
    constexpr int PF = 16;
    float pf[PF];
    {
        const int n = tid & (CD - 1), kg = tid / CD;
        const float* w = a.blob + a.desc[0].W_in + (size_t)(kg * (DF / KG1)) * CD + n;
#pragma unroll
        for (int k = 0; k < PF; ++k) pf[k] = __ldg(w + (size_t)k * CD);
    }
    for (int l = 0; l < a.ntcm; ++l) {
        const TcmStreamDesc& d = a.desc[l];
        // ---------------------------------------------------------------- squeeze 1x1: 64 outputs x 16 K groups
        {
            const int n = tid & (CD - 1), kg = tid / CD;
            constexpr int KPER = DF / KG1;
            float p[SPC];
#pragma unroll
            for (int s = 0; s < SPC; ++s) p[s] = 0.f;
            static_assert(KPER == PF, "the squeeze weights of a thread are exactly the prefetch registers");
#pragma unroll
            for (int k = 0; k < KPER; ++k) {
                const float wv = pf[k];
#pragma unroll
                for (int s = 0; s < SPC; ++s) p[s] = fmaf(xs[s][kg * KPER + k], wv, p[s]);
            }
#pragma unroll
            for (int s = 0; s < SPC; ++s) red[(kg * SPC + s) * CD + n] = p[s];
        }
        {   // prefetch: the first PF (tap, channel) weights of this thread's dilated-conv column
            const int col = tid & (2 * CD - 1), br = col / CD, kg = tid / (2 * CD);
            const int j0 = kg * kper2;
#pragma unroll
            for (int k = 0; k < PF; ++k) {
                const int j = min(j0 + k, a.kd * CD - 1);
                const int tap = j / CD, c = j & (CD - 1);
                pf[k] = __ldg(a.blob + d.W_dil + ((size_t)tap * 2 * CD + br * CD + c) * (2 * CD) + col);
            }
        }
        __syncthreads();
        float* ring = a.act_base + d.ring_off;
        if (tid < SPC * CD) {
            const int s = tid / CD, n = tid & (CD - 1);
            float y = 0.f;
#pragma unroll
            for (int kg = 0; kg < KG1; ++kg) y += red[(kg * SPC + s) * CD + n];
            ys[s][n] = y;
            if (s0 + s < a.S) ring[((size_t)(s0 + s) * d.RT + ring_slot(step, d.RT)) * CD + n] = y;
        }
        __syncthreads();
        // ---------------------------------------------------------------- normalised taps of both branches
        for (int i = tid; i < a.kd * SPC * CD; i += NT) {
            const int tap = i / (SPC * CD);
            const int r = i - tap * (SPC * CD);
            const int s = r / CD, c = r & (CD - 1);
            const int n = step - d.dt[tap];
            float vl = 0.f, vr = 0.f;
            if (s0 + s < a.S && n >= (a.start ? __ldg(a.start + s0 + s) : 0)) {
                const float v = d.dt[tap] == 0 ? ys[s][c] : ring[((size_t)(s0 + s) * d.RT + ring_slot(n, d.RT)) * CD + c];
                vl = prelu_norm(v, __ldg(a.blob + d.aL + c), __ldg(a.blob + d.sL + c), __ldg(a.blob + d.hL + c));
                vr = prelu_norm(v, __ldg(a.blob + d.aR + c), __ldg(a.blob + d.sR + c), __ldg(a.blob + d.hR + c));
            }
            u[0][tap][s][c] = vl;
            u[1][tap][s][c] = vr;
        }
        __syncthreads();
        // ---------------------------------------------------------------- dilated convs: 128 outputs x 8 K groups
        {
            const int col = tid & (2 * CD - 1);        // 0..63 left outputs, 64..127 right outputs (block-diagonal packing)
            const int br = col / CD;
            const int kg = tid / (2 * CD);
            float p[SPC];
#pragma unroll
            for (int s = 0; s < SPC; ++s) p[s] = 0.f;
            const int j0 = kg * kper2, j1 = min(a.kd * CD, j0 + kper2);
#pragma unroll
            for (int k = 0; k < PF; ++k) {
                const int j = j0 + k;
                if (j < j1) {
                    const int tap = j / CD, c = j & (CD - 1);
#pragma unroll
                    for (int s = 0; s < SPC; ++s) p[s] = fmaf(u[br][tap][s][c], pf[k], p[s]);
                }
            }
#pragma unroll 24
            for (int j = j0 + PF; j < j1; ++j) {
                const int tap = j / CD, c = j & (CD - 1);
                const float wv = __ldg(a.blob + d.W_dil + ((size_t)tap * 2 * CD + br * CD + c) * (2 * CD) + col);
#pragma unroll
                for (int s = 0; s < SPC; ++s) p[s] = fmaf(u[br][tap][s][c], wv, p[s]);
            }
#pragma unroll
            for (int s = 0; s < SPC; ++s) red[(kg * SPC + s) * (2 * CD) + col] = p[s];
        }
        {   // prefetch: this thread's expand weights
            const int j = tid & (DF - 1), kg = tid / DF;
            const float* w = a.blob + d.W_out + (size_t)(kg * (CD / KG3)) * DF + j;
#pragma unroll
            for (int c = 0; c < PF; ++c) pf[c] = __ldg(w + (size_t)c * DF);
        }
        __syncthreads();
        if (tid < SPC * CD) {
            const int s = tid / CD, n = tid & (CD - 1);
            float zl = 0.f, zr = 0.f;
#pragma unroll
            for (int kg = 0; kg < KG2; ++kg) {
                zl += red[(kg * SPC + s) * (2 * CD) + n];
                zr += red[(kg * SPC + s) * (2 * CD) + CD + n];
            }
            const float z = zl * sigmoid_f(zr);
            uo[s][n] = prelu_norm(z, __ldg(a.blob + d.aO + n), __ldg(a.blob + d.sO + n), __ldg(a.blob + d.hO + n));
        }
        __syncthreads();
        // ---------------------------------------------------------------- expand 1x1: 256 outputs x 4 K groups
        {
            const int j = tid & (DF - 1), kg = tid / DF;
            constexpr int KPER = CD / KG3;
            float p[SPC];
#pragma unroll
            for (int s = 0; s < SPC; ++s) p[s] = 0.f;
            static_assert(KPER == PF, "the expand weights of a thread are exactly the prefetch registers");
#pragma unroll
            for (int c = 0; c < KPER; ++c) {
                const float wv = pf[c];
#pragma unroll
                for (int s = 0; s < SPC; ++s) p[s] = fmaf(uo[s][kg * KPER + c], wv, p[s]);
            }
#pragma unroll
            for (int s = 0; s < SPC; ++s) red[(kg * SPC + s) * DF + j] = p[s];
        }
        if (l + 1 < a.ntcm) {   // prefetch: the next TCM's squeeze weights
            const int n = tid & (CD - 1), kg = tid / CD;
            const float* w = a.blob + a.desc[l + 1].W_in + (size_t)(kg * (DF / KG1)) * CD + n;
#pragma unroll
            for (int k = 0; k < PF; ++k) pf[k] = __ldg(w + (size_t)k * CD);
        }
        __syncthreads();
        if (tid < SPC * DF) {
            const int s = tid / DF, j = tid & (DF - 1);
            float o = xs[s][j];
#pragma unroll
            for (int kg = 0; kg < KG3; ++kg) o += red[(kg * SPC + s) * DF + j];
            xs[s][j] = o;
            if ((l + 1) % a.p == 0) accs[s][j] += o;
        }
        __syncthreads();
    }
    const int oslot = ring_slot(step, a.out_RT);
    for (int i = tid; i < SPC * DF; i += NT) {
        const int s = i / DF, j = i - s * DF;
        if (s0 + s < a.S) a.out[((size_t)(s0 + s) * a.out_RT + oslot) * DF + j] = accs[s][j];
    }
}

}  // namespace

bool tcm_stream_supported(int cd1, int d_feat, int kd1) { return cd1 == CD && d_feat == DF && kd1 >= 1 && kd1 <= 8; }

int launch_tcm_stream(const TcmStreamArgs& a, cudaStream_t st) {
    if (a.S < 1 || a.ntcm < 1 || a.p < 1) return fail("tcm_stream: bad arguments");
    double macs = (double)a.ntcm * (DF * CD + 2.0 * a.kd * CD * CD + CD * DF);
    ProfScope ps("tcm_stream", 2.0 * macs * a.S, 4.0 * macs * ((a.S + SPC - 1) / SPC), st);
    EAB_CUDA(launch_k(tcm_stream_kernel, dim3((a.S + SPC - 1) / SPC), dim3(NT), (size_t)0, st, a));
    EAB_LAUNCH_CHECK("tcm_stream_kernel");
    return 0;
}

}  // namespace eab
