// First slice of the training path (BASELINE configs[4]; train_distributed.py:214-230): the
// beamforming head's tail and the loss, forward AND backward, hand-written.
//
//   head:  w = Linear(ReLU(Linear(h2)))  (LSTM_BF.w_dnn, EaBNet.py:593-597, 612-613), out = sum_m w_m * x_m  (complex
//          filter-and-sum, EaBNet.py:114-117).  Backward: d h2, d W1, d b1, d W2, d b2 from d out (the spectrum x is data).
//   loss:  com_mag_mse_loss (EaBNet.py:627-640): 0.5 (masked MSE of the magnitudes + masked MSE of the complex parts).
//
// One CTA walks 128-row tiles (row = (b, t, f)); a tile's h2 rows, ReLU activations and gradients live in shared memory, the
// weight gradients accumulate in registers over all the CTA's tiles (4 x 4 register blocks of dW1, 2 x 4 of dW2) and leave
// as per-CTA partial sums that a second kernel adds up in a fixed order: results do not depend on the launch schedule.
// CUDA-core fp32 throughout (FFMA): this is the correctness-first slice of the backward pass - 32 kFLOP per row, 0.2 TFLOP
// per 64 x 6 s - the tensor-core version belongs with the rest of the backward kernels.
#include <algorithm>
#include <cstring>

#include "../../include/eabnet_b200.h"
#include "common.cuh"

namespace eab {

namespace {

constexpr int HR = 128;           // rows per tile
constexpr int HT = 256;           // threads
constexpr int LD = 68;            // row pitch of the [rows][64] tiles (floats): 16-byte aligned, rows 0..7 hit distinct banks
constexpr int NJ_MAX = 32;        // 2 M <= 32
constexpr int LDJ = 33;

struct HeadGradArgs {
    const float* W1; const float* b1; const float* W2; const float* b2;     // torch layouts: W1 [64][64] (out, in), W2 [2M][64]
    const float* h2;              // [rows][64]
    const float* spec;            // [rows][M][2]
    const float* d_out;           // [B][2][T F]   (backward)
    float* out;                   // [B][2][T F]   (forward)
    float* d_h2;                  // [rows][64]
    float* partial;               // [grid][NG]
    long long rows, TF;
    int M, NJ;
};

__host__ __device__ inline int grad_floats(int NJ) { return 64 * 64 + 64 + NJ * 64 + NJ; }

struct HeadSmem {
    float* W1t;   // [64 i][64 k]
    float* W1s;   // [64 k][64 i]
    float* W2s;   // [NJ][64]
    float* b1s; float* b2s;
    float* hs;    // [HR][LD]
    float* a1;    // [HR][LD]
    float* dz;    // [HR][LD]
    float* dw;    // [HR][LDJ]
    float* xs;    // [HR][LDJ]
    float* gs;    // [HR][2]
};
__host__ __device__ inline size_t head_smem_floats() { return 2 * 4096 + NJ_MAX * 64 + 64 + NJ_MAX + 3 * HR * LD + 2 * HR * LDJ + HR * 2; }
__device__ inline HeadSmem carve(float* p) {
    HeadSmem s;
    s.W1t = p; p += 4096; s.W1s = p; p += 4096; s.W2s = p; p += NJ_MAX * 64; s.b1s = p; p += 64; s.b2s = p; p += NJ_MAX;
    s.hs = p; p += HR * LD; s.a1 = p; p += HR * LD; s.dz = p; p += HR * LD; s.dw = p; p += HR * LDJ; s.xs = p; p += HR * LDJ; s.gs = p;
    return s;
}

__device__ inline void load_weights(const HeadGradArgs& a, const HeadSmem& s) {
    for (int i = threadIdx.x; i < 4096; i += HT) {
        const float w = __ldg(a.W1 + i);                     // W1[k][i']: k = i / 64, i' = i % 64
        s.W1s[i] = w;
        s.W1t[(i & 63) * 64 + (i >> 6)] = w;
    }
    for (int i = threadIdx.x; i < a.NJ * 64; i += HT) s.W2s[i] = __ldg(a.W2 + i);
    for (int i = threadIdx.x; i < 64; i += HT) s.b1s[i] = __ldg(a.b1 + i);
    for (int i = threadIdx.x; i < a.NJ; i += HT) s.b2s[i] = __ldg(a.b2 + i);
}

// tile rows -> shared memory (zeros past the end), then a1 = ReLU(W1 h + b1)
__device__ inline void load_tile_and_hidden(const HeadGradArgs& a, const HeadSmem& s, long long row0) {
    for (int i = threadIdx.x; i < HR * 16; i += HT) {        // 16 float4 per row
        const int r = i >> 4, c = i & 15;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row0 + r < a.rows) v = __ldg(reinterpret_cast<const float4*>(a.h2 + (row0 + r) * 64) + c);
        *reinterpret_cast<float4*>(s.hs + r * LD + c * 4) = v;
    }
    for (int i = threadIdx.x; i < HR * a.NJ; i += HT) {
        const int r = i / a.NJ, j = i - r * a.NJ;
        s.xs[r * LDJ + j] = row0 + r < a.rows ? __ldg(a.spec + (row0 + r) * a.NJ + j) : 0.f;
    }
    __syncthreads();
    const int rp = threadIdx.x >> 2, kq = threadIdx.x & 3;   // rows rp, rp + 64 ; hidden units kq + 4 j
    float acc0[16], acc1[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) { acc0[j] = s.b1s[kq + 4 * j]; acc1[j] = acc0[j]; }
    for (int i = 0; i < 64; ++i) {
        const float h0 = s.hs[rp * LD + i], h1 = s.hs[(rp + 64) * LD + i];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const float w = s.W1t[i * 64 + kq + 4 * j];
            acc0[j] = fmaf(h0, w, acc0[j]);
            acc1[j] = fmaf(h1, w, acc1[j]);
        }
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        s.a1[rp * LD + kq + 4 * j] = fmaxf(acc0[j], 0.f);
        s.a1[(rp + 64) * LD + kq + 4 * j] = fmaxf(acc1[j], 0.f);
    }
    __syncthreads();
}

__global__ void __launch_bounds__(HT, 1) head_fwd_kernel(const HeadGradArgs a) {
    extern __shared__ float smem_f[];
    const HeadSmem s = carve(smem_f);
    load_weights(a, s);
    __syncthreads();
    const long long ntiles = (a.rows + HR - 1) / HR;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long row0 = tile * HR;
        load_tile_and_hidden(a, s, row0);
        for (int i = threadIdx.x; i < HR * a.NJ; i += HT) {  // beam weights w[r][j] (into the dw buffer)
            const int r = i / a.NJ, j = i - r * a.NJ;
            float acc = s.b2s[j];
            for (int k = 0; k < 64; ++k) acc = fmaf(s.a1[r * LD + k], s.W2s[j * 64 + k], acc);
            s.dw[r * LDJ + j] = acc;
        }
        __syncthreads();
        if (threadIdx.x < HR && row0 + threadIdx.x < a.rows) {
            const int r = threadIdx.x;
            float yr = 0.f, yi = 0.f;
            for (int m = 0; m < a.M; ++m) {
                const float wr = s.dw[r * LDJ + 2 * m], wi = s.dw[r * LDJ + 2 * m + 1];
                const float xr = s.xs[r * LDJ + 2 * m], xi = s.xs[r * LDJ + 2 * m + 1];
                yr += wr * xr - wi * xi;                     // (bf_w_r*inpt_r - bf_w_i*inpt_i).sum(-1), EaBNet.py:116
                yi += wr * xi + wi * xr;
            }
            const long long row = row0 + r, b = row / a.TF, p = row - b * a.TF;
            a.out[(b * 2 + 0) * a.TF + p] = yr;
            a.out[(b * 2 + 1) * a.TF + p] = yi;
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(HT, 1) head_bwd_kernel(const HeadGradArgs a) {
    extern __shared__ float smem_f[];
    const HeadSmem s = carve(smem_f);
    load_weights(a, s);
    __syncthreads();
    const int tid = threadIdx.x;
    const int rp = tid >> 2, kq = tid & 3;
    const int kb = tid >> 4, ib = tid & 15;                  // dW1 block rows 4 kb.., columns 4 ib.. ; dW2 rows kb (, kb + 16), columns 4 ib..
    float gW1[4][4], gW2[2][4], gb = 0.f;
#pragma unroll
    for (int x = 0; x < 4; ++x)
#pragma unroll
        for (int y = 0; y < 4; ++y) gW1[x][y] = 0.f;
#pragma unroll
    for (int x = 0; x < 2; ++x)
#pragma unroll
        for (int y = 0; y < 4; ++y) gW2[x][y] = 0.f;
    const long long ntiles = (a.rows + HR - 1) / HR;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long row0 = tile * HR;
        if (tid < HR) {
            const long long row = row0 + tid;
            float gr = 0.f, gi = 0.f;
            if (row < a.rows) {
                const long long b = row / a.TF, p = row - b * a.TF;
                gr = __ldg(a.d_out + (b * 2 + 0) * a.TF + p);
                gi = __ldg(a.d_out + (b * 2 + 1) * a.TF + p);
            }
            s.gs[tid * 2] = gr; s.gs[tid * 2 + 1] = gi;
        }
        load_tile_and_hidden(a, s, row0);                    // (its first barrier also publishes gs)
        // d w[r][m] = conj-product of the output gradient with the spectrum
        for (int i = tid; i < HR * a.M; i += HT) {
            const int r = i / a.M, m = i - r * a.M;
            const float gr = s.gs[r * 2], gi = s.gs[r * 2 + 1];
            const float xr = s.xs[r * LDJ + 2 * m], xi = s.xs[r * LDJ + 2 * m + 1];
            s.dw[r * LDJ + 2 * m] = gr * xr + gi * xi;
            s.dw[r * LDJ + 2 * m + 1] = gi * xr - gr * xi;
        }
        __syncthreads();
        {   // d z1 = (W2^T d w) * [z1 > 0]
            float acc0[16], acc1[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) { acc0[j] = 0.f; acc1[j] = 0.f; }
            for (int jj = 0; jj < a.NJ; ++jj) {
                const float d0 = s.dw[rp * LDJ + jj], d1 = s.dw[(rp + 64) * LDJ + jj];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float w = s.W2s[jj * 64 + kq + 4 * j];
                    acc0[j] = fmaf(d0, w, acc0[j]);
                    acc1[j] = fmaf(d1, w, acc1[j]);
                }
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int k = kq + 4 * j;
                s.dz[rp * LD + k] = s.a1[rp * LD + k] > 0.f ? acc0[j] : 0.f;
                s.dz[(rp + 64) * LD + k] = s.a1[(rp + 64) * LD + k] > 0.f ? acc1[j] : 0.f;
            }
        }
        __syncthreads();
        // weight gradients over the tile's rows (rows past the end hold zeros)
        for (int r = 0; r < HR; ++r) {
            const float4 dzv = *reinterpret_cast<const float4*>(s.dz + r * LD + 4 * kb);
            const float4 hv = *reinterpret_cast<const float4*>(s.hs + r * LD + 4 * ib);
            const float dzr[4] = {dzv.x, dzv.y, dzv.z, dzv.w}, hr[4] = {hv.x, hv.y, hv.z, hv.w};
#pragma unroll
            for (int x = 0; x < 4; ++x)
#pragma unroll
                for (int y = 0; y < 4; ++y) gW1[x][y] = fmaf(dzr[x], hr[y], gW1[x][y]);
            const float4 av = *reinterpret_cast<const float4*>(s.a1 + r * LD + 4 * ib);
            const float ar[4] = {av.x, av.y, av.z, av.w};
            if (kb < a.NJ) {
                const float d = s.dw[r * LDJ + kb];
#pragma unroll
                for (int y = 0; y < 4; ++y) gW2[0][y] = fmaf(d, ar[y], gW2[0][y]);
            }
            if (kb + 16 < a.NJ) {
                const float d = s.dw[r * LDJ + kb + 16];
#pragma unroll
                for (int y = 0; y < 4; ++y) gW2[1][y] = fmaf(d, ar[y], gW2[1][y]);
            }
        }
        if (tid < 64) { for (int r = 0; r < HR; ++r) gb += s.dz[r * LD + tid]; }
        else if (tid - 64 < a.NJ) { for (int r = 0; r < HR; ++r) gb += s.dw[r * LDJ + tid - 64]; }
        __syncthreads();
        {   // d h2 = W1^T d z1, staged through the h tile for coalesced stores
            float acc0[16], acc1[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) { acc0[j] = 0.f; acc1[j] = 0.f; }
            for (int k = 0; k < 64; ++k) {
                const float d0 = s.dz[rp * LD + k], d1 = s.dz[(rp + 64) * LD + k];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float w = s.W1s[k * 64 + kq + 4 * j];
                    acc0[j] = fmaf(d0, w, acc0[j]);
                    acc1[j] = fmaf(d1, w, acc1[j]);
                }
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                s.hs[rp * LD + kq + 4 * j] = acc0[j];
                s.hs[(rp + 64) * LD + kq + 4 * j] = acc1[j];
            }
        }
        __syncthreads();
        for (int i = tid; i < HR * 16; i += HT) {
            const int r = i >> 4, c = i & 15;
            if (row0 + r < a.rows)
                *(reinterpret_cast<float4*>(a.d_h2 + (row0 + r) * 64) + c) = *reinterpret_cast<const float4*>(s.hs + r * LD + c * 4);
        }
        __syncthreads();
    }
    // per-CTA partial sums: [dW1 64x64 | db1 64 | dW2 NJ x 64 | db2 NJ]
    float* P = a.partial + (size_t)blockIdx.x * grad_floats(a.NJ);
#pragma unroll
    for (int x = 0; x < 4; ++x)
#pragma unroll
        for (int y = 0; y < 4; ++y) P[(4 * kb + x) * 64 + 4 * ib + y] = gW1[x][y];
    if (kb < a.NJ)
#pragma unroll
        for (int y = 0; y < 4; ++y) P[4096 + 64 + kb * 64 + 4 * ib + y] = gW2[0][y];
    if (kb + 16 < a.NJ)
#pragma unroll
        for (int y = 0; y < 4; ++y) P[4096 + 64 + (kb + 16) * 64 + 4 * ib + y] = gW2[1][y];
    if (tid < 64) P[4096 + tid] = gb;
    else if (tid - 64 < a.NJ) P[4096 + 64 + a.NJ * 64 + tid - 64] = gb;
}

__global__ void head_bwd_reduce_kernel(const float* __restrict__ partial, float* __restrict__ grads, int nblocks, int ng) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ng) return;
    float acc = 0.f;
    for (int b = 0; b < nblocks; ++b) acc += partial[(size_t)b * ng + i];        // fixed order
    grads[i] = acc;
}

// ------------------------------------------------------------------------------------------------------------ loss
struct LossArgs {
    const float* esti; const float* label;        // [B][2][T][F]
    const int* frames;                            // [B] valid frames per utterance (device) or null = T
    int B, T, F;
    int freq_major;                               // 0: [B][2][T][F] (EaBNet's estimate) ; 1: [B][2][F][T] (GaGNet's stage estimates)
    double* sums;                                 // [2]: sum of masked (|e|-|l|)^2, sum of masked |e-l|^2
    const float* gscale;                          // backward: upstream gradient of the scalar loss (device) or null = 1
    float* d_esti;
    float* loss;
    double n_mask;                                // sum_b frames[b] * F
};

__global__ void __launch_bounds__(256) loss_fwd_kernel(const LossArgs a) {
    const long long TF = (long long)a.T * a.F, n = (long long)a.B * TF;
    double s1 = 0.0, s2 = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const long long b = i / TF, p = i - b * TF;
        const int t = a.freq_major ? (int)(p % a.T) : (int)(p / a.F);
        if (a.frames && t >= __ldg(a.frames + b)) continue;
        const float er = __ldg(a.esti + (b * 2) * TF + p), ei = __ldg(a.esti + (b * 2 + 1) * TF + p);
        const float lr = __ldg(a.label + (b * 2) * TF + p), li = __ldg(a.label + (b * 2 + 1) * TF + p);
        const float dm = sqrtf(er * er + ei * ei) - sqrtf(lr * lr + li * li);
        s1 += (double)(dm * dm);
        s2 += (double)((er - lr) * (er - lr)) + (double)((ei - li) * (ei - li));
    }
    __shared__ double sh[2][8];
    for (int o = 16; o > 0; o >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
    if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = s1; sh[1][threadIdx.x >> 5] = s2; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double t1 = 0.0, t2 = 0.0;
        for (int w = 0; w < 8; ++w) { t1 += sh[0][w]; t2 += sh[1][w]; }
        atomicAdd(a.sums, t1);
        atomicAdd(a.sums + 1, t2);
    }
}
__global__ void loss_finalize_kernel(const LossArgs a) {
    *a.loss = (float)(0.5 * (a.sums[0] / a.n_mask + a.sums[1] / (2.0 * a.n_mask)));
}
// d loss / d esti = [ (|e| - |l|) e / |e| + 0.5 (e - l) ] mask / n_mask   (torch.norm's gradient at 0 is 0)
__global__ void __launch_bounds__(256) loss_bwd_kernel(const LossArgs a) {
    const long long TF = (long long)a.T * a.F, n = (long long)a.B * TF;
    const float g = (a.gscale ? __ldg(a.gscale) : 1.f) / (float)a.n_mask;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const long long b = i / TF, p = i - b * TF;
        const int t = a.freq_major ? (int)(p % a.T) : (int)(p / a.F);
        float dr = 0.f, di = 0.f;
        if (!(a.frames && t >= __ldg(a.frames + b))) {
            const float er = __ldg(a.esti + (b * 2) * TF + p), ei = __ldg(a.esti + (b * 2 + 1) * TF + p);
            const float lr = __ldg(a.label + (b * 2) * TF + p), li = __ldg(a.label + (b * 2 + 1) * TF + p);
            const float me = sqrtf(er * er + ei * ei), ml = sqrtf(lr * lr + li * li);
            const float q = me > 0.f ? (me - ml) / me : 0.f;
            dr = g * (q * er + 0.5f * (er - lr));
            di = g * (q * ei + 0.5f * (ei - li));
        }
        a.d_esti[(b * 2) * TF + p] = dr;
        a.d_esti[(b * 2 + 1) * TF + p] = di;
    }
}

int head_args(HeadGradArgs* a, const float* W1, const float* b1, const float* W2, const float* b2, const float* h2, const float* spec, int B,
              int T, int F, int M) {
    if (!W1 || !b1 || !W2 || !b2 || !h2 || !spec) return fail("eab_head: null argument");
    if (B < 1 || T < 1 || F < 1 || M < 1 || 2 * M > NJ_MAX) return fail("eab_head: bad shape (2 M <= 32)");
    memset(a, 0, sizeof(*a));
    a->W1 = W1; a->b1 = b1; a->W2 = W2; a->b2 = b2; a->h2 = h2; a->spec = spec;
    a->TF = (long long)T * F; a->rows = (long long)B * a->TF; a->M = M; a->NJ = 2 * M;
    if ((reinterpret_cast<uintptr_t>(h2) & 15) != 0) return fail("eab_head: h2 must be 16-byte aligned");
    return 0;
}

}  // namespace

}  // namespace eab

using namespace eab;

extern "C" {

int eab_head_forward(const float* W1, const float* b1, const float* W2, const float* b2, const float* h2, const float* spec, float* out,
                     int B, int T, int F, int M, void* stream) {
    HeadGradArgs a;
    EAB_TRY(head_args(&a, W1, b1, W2, b2, h2, spec, B, T, F, M));
    if (!out) return fail("eab_head_forward: null argument");
    a.out = out;
    const size_t smem = head_smem_floats() * sizeof(float);
    EAB_TRY(ensure_dynamic_smem(reinterpret_cast<const void*>(head_fwd_kernel), (int)smem));
    int sms = 0;
    EAB_TRY(device_sm_count(&sms));
    const long long ntiles = (a.rows + HR - 1) / HR;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    ProfScope ps("head_fwd", 2.0 * a.rows * (64.0 * 64 + a.NJ * 64 + 4.0 * M), 4.0 * a.rows * (64 + a.NJ + 2), st);
    EAB_CUDA(launch_k(head_fwd_kernel, dim3((unsigned)std::min<long long>(ntiles, sms)), dim3(HT), smem, st, a));
    EAB_LAUNCH_CHECK("head_fwd_kernel");
    return 0;
}

size_t eab_head_backward_workspace_bytes(int M) {
    int sms = 0;
    if (M < 1 || 2 * M > NJ_MAX || device_sm_count(&sms)) return 0;
    return (size_t)sms * grad_floats(2 * M) * sizeof(float);
}

int eab_head_backward(const float* W1, const float* b1, const float* W2, const float* b2, const float* h2, const float* spec,
                      const float* d_out, float* d_h2, float* grads, int B, int T, int F, int M, void* workspace, size_t workspace_bytes,
                      void* stream) {
    HeadGradArgs a;
    EAB_TRY(head_args(&a, W1, b1, W2, b2, h2, spec, B, T, F, M));
    if (!d_out || !d_h2 || !grads || !workspace) return fail("eab_head_backward: null argument");
    if ((reinterpret_cast<uintptr_t>(d_h2) & 15) != 0) return fail("eab_head_backward: d_h2 must be 16-byte aligned");
    const size_t need = eab_head_backward_workspace_bytes(M);
    if (!need || workspace_bytes < need) return fail("eab_head_backward: workspace too small");
    a.d_out = d_out; a.d_h2 = d_h2; a.partial = static_cast<float*>(workspace);
    const size_t smem = head_smem_floats() * sizeof(float);
    EAB_TRY(ensure_dynamic_smem(reinterpret_cast<const void*>(head_bwd_kernel), (int)smem));
    int sms = 0;
    EAB_TRY(device_sm_count(&sms));
    const long long ntiles = (a.rows + HR - 1) / HR;
    const int grid = (int)std::min<long long>(ntiles, sms);
    const int ng = grad_floats(a.NJ);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    {
        ProfScope ps("head_bwd", 2.0 * a.rows * (3.0 * 64 * 64 + 3.0 * a.NJ * 64 + 4.0 * M), 4.0 * a.rows * (64 + a.NJ + 2 + 64), st);
        EAB_CUDA(launch_k(head_bwd_kernel, dim3(grid), dim3(HT), smem, st, a));
        EAB_LAUNCH_CHECK("head_bwd_kernel");
    }
    head_bwd_reduce_kernel<<<(ng + 255) / 256, 256, 0, st>>>(a.partial, grads, grid, ng);
    EAB_LAUNCH_CHECK("head_bwd_reduce_kernel");
    return 0;
}

static int loss_fwd(const float* esti, const float* label, const int* frames_dev, int64_t frames_total, int B, int T, int F, int freq_major,
                    float* loss, void* scratch16, void* stream);
static int loss_bwd(const float* esti, const float* label, const int* frames_dev, int64_t frames_total, int B, int T, int F, int freq_major,
                    const float* grad_loss_dev, float* d_esti, void* stream);
int eab_loss_com_mag_mse(const float* esti, const float* label, const int* frames_dev, int64_t frames_total, int B, int T, int F,
                         float* loss, void* scratch16, void* stream) {
    return loss_fwd(esti, label, frames_dev, frames_total, B, T, F, 0, loss, scratch16, stream);
}
int eab_loss_com_mag_mse_backward(const float* esti, const float* label, const int* frames_dev, int64_t frames_total, int B, int T, int F,
                                  const float* grad_loss_dev, float* d_esti, void* stream) {
    return loss_bwd(esti, label, frames_dev, frames_total, B, T, F, 0, grad_loss_dev, d_esti, stream);
}
// the same loss on frequency-major tensors [B][2][F][T]: one stage of stagewise_com_mag_mse_loss (GaGNet.py:601-619)
int eab_loss_com_mag_mse_fm(const float* esti, const float* label, const int* frames_dev, int64_t frames_total, int B, int T, int F,
                            float* loss, void* scratch16, void* stream) {
    return loss_fwd(esti, label, frames_dev, frames_total, B, T, F, 1, loss, scratch16, stream);
}
int eab_loss_com_mag_mse_fm_backward(const float* esti, const float* label, const int* frames_dev, int64_t frames_total, int B, int T, int F,
                                     const float* grad_loss_dev, float* d_esti, void* stream) {
    return loss_bwd(esti, label, frames_dev, frames_total, B, T, F, 1, grad_loss_dev, d_esti, stream);
}

static int loss_fwd(const float* esti, const float* label, const int* frames_dev, int64_t frames_total, int B, int T, int F, int freq_major,
                    float* loss, void* scratch16, void* stream) {
    if (!esti || !label || !loss || !scratch16) return fail("eab_loss_com_mag_mse: null argument");
    if (B < 1 || T < 1 || F < 1) return fail("eab_loss_com_mag_mse: bad shape");
    LossArgs a;
    memset(&a, 0, sizeof(a));
    a.esti = esti; a.label = label; a.frames = frames_dev; a.B = B; a.T = T; a.F = F; a.loss = loss; a.freq_major = freq_major;
    a.sums = static_cast<double*>(scratch16);
    a.n_mask = (double)(frames_dev ? frames_total : (int64_t)B * T) * F;
    if (a.n_mask <= 0) return fail("eab_loss_com_mag_mse: empty mask");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    EAB_CUDA(cudaMemsetAsync(scratch16, 0, 16, st));
    int sms = 0;
    EAB_TRY(device_sm_count(&sms));
    loss_fwd_kernel<<<sms * 4, 256, 0, st>>>(a);
    EAB_LAUNCH_CHECK("loss_fwd_kernel");
    loss_finalize_kernel<<<1, 1, 0, st>>>(a);
    EAB_LAUNCH_CHECK("loss_finalize_kernel");
    return 0;
}

static int loss_bwd(const float* esti, const float* label, const int* frames_dev, int64_t frames_total, int B, int T, int F, int freq_major,
                    const float* grad_loss_dev, float* d_esti, void* stream) {
    if (!esti || !label || !d_esti) return fail("eab_loss_com_mag_mse_backward: null argument");
    if (B < 1 || T < 1 || F < 1) return fail("eab_loss_com_mag_mse_backward: bad shape");
    LossArgs a;
    memset(&a, 0, sizeof(a));
    a.esti = esti; a.label = label; a.frames = frames_dev; a.B = B; a.T = T; a.F = F; a.d_esti = d_esti; a.gscale = grad_loss_dev;
    a.freq_major = freq_major;
    a.n_mask = (double)(frames_dev ? frames_total : (int64_t)B * T) * F;
    if (a.n_mask <= 0) return fail("eab_loss_com_mag_mse_backward: empty mask");
    int sms = 0;
    EAB_TRY(device_sm_count(&sms));
    loss_bwd_kernel<<<sms * 4, 256, 0, static_cast<cudaStream_t>(stream)>>>(a);
    EAB_LAUNCH_CHECK("loss_bwd_kernel");
    return 0;
}

}  // extern "C"
