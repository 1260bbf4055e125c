// Beam-weight MLP + filter-and-sum in ONE kernel (sm_100a):
//     w = Linear(64 -> 2M)(ReLU(Linear(64 -> 64)(h2)))            (LSTM_BF.w_dnn, EaBNet.py:593-597, 613)
//     y[b,:,t,f] = sum_m w[b,t,f,m] * x[b,t,f,m]   (complex)       (EaBNet.py:114-117)
// As separate layers this was two stage + two GEMM launches + the beam kernel: 11.5 GB of HBM traffic for tensors that
// are produced and consumed row by row (2.4 ms per 64 x 6 s step).  Fused, a row of h2 (256 B) is read once, both
// GEMMs run on the tensor cores through shared memory / TMEM, and only the 8-byte result leaves: 2.1 GB, HBM-bound.
//
// Persistent CTA, tiles of 128 rows (row = flattened (b,t,f)), software-pipelined over two buffers of everything:
//   warps 0-7   loaders   : 2 threads per row: h2 row -> fp16 hi/lo -> swizzled A1 operand
//   warps 8-11  epilogue 1: thread = row (TMEM lane): D1 (64 cols) + b1 -> ReLU -> fp16 hi/lo -> swizzled A2 operand
//   warps 12-15 epilogue 2: thread = row: D2 (2M cols) + b2 = w; complex MAC against the row's 2M input values; store
//   warp 16     MMA issuer: G1(i+1) = A1 x W1 (3 fp16 passes, N = 64), then G2(i) = A2 x W2 (3 passes, N = 32)
#include "common.cuh"
#include "umma.cuh"

namespace eab {

namespace {

using namespace umma;

constexpr int TM = 128;
constexpr int NLOAD = 256, NEPI = 128;
constexpr int NTHREADS = NLOAD + 2 * NEPI + 32;
constexpr int MMA_WARP = (NLOAD + 2 * NEPI) / 32;
constexpr int SLAB = TM * 128;                       // one 64-wide fp16 K slab of an A operand (16 KB)
constexpr int A_BYTES = 2 * SLAB;                    // [hi|lo]
constexpr int W1_BYTES = 2 * 64 * 128;               // [hi|lo][64 rows][64 k]
constexpr int W2_BYTES = 2 * 32 * 128;               // [hi|lo][32 rows][64 k]
constexpr int SMEM_BYTES = 4 * A_BYTES + W1_BYTES + W2_BYTES + (64 + 32) * 4 + 256 + 1024;

__global__ void __launch_bounds__(NTHREADS, 1) head_fused_kernel(const HeadArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* A1 = smem;                               // [2 buffers][hi|lo]
    uint8_t* A2 = smem + 2 * A_BYTES;
    uint8_t* W1 = smem + 4 * A_BYTES;
    uint8_t* W2 = W1 + W1_BYTES;
    float* sb1 = reinterpret_cast<float*>(W2 + W2_BYTES);
    float* sb2 = sb1 + 64;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sb2 + 32);
    uint64_t* a1_full = bars;         // [2] loaders -> MMA
    uint64_t* a1_empty = bars + 2;    // [2] MMA (commit) -> loaders
    uint64_t* d1_full = bars + 4;     // [2] MMA (commit) -> epilogue 1
    uint64_t* d1_empty = bars + 6;    // [2] epilogue 1 -> MMA
    uint64_t* a2_full = bars + 8;     // [2] epilogue 1 -> MMA
    uint64_t* a2_empty = bars + 10;   // [2] MMA (commit) -> epilogue 1
    uint64_t* d2_full = bars + 12;    // [2] MMA (commit) -> epilogue 2
    uint64_t* d2_empty = bars + 14;   // [2] epilogue 2 -> MMA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;
    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&a1_full[i], NLOAD); mbar_init(&a1_empty[i], 1);
            mbar_init(&d1_full[i], 1);     mbar_init(&d1_empty[i], NEPI);
            mbar_init(&a2_full[i], NEPI);  mbar_init(&a2_empty[i], 1);
            mbar_init(&d2_full[i], 1);     mbar_init(&d2_empty[i], NEPI);
        }
        fence_barrier_init();
    }
    if (warp == MMA_WARP) tmem_alloc(tmem_slot, 256);      // D1: cols [0,64) [64,128);  D2: cols [128,160) [160,192)
    {
        const uint4* s1 = reinterpret_cast<const uint4*>(a.W1hi);
        const uint4* s1l = reinterpret_cast<const uint4*>(a.W1lo);
        const uint4* s2 = reinterpret_cast<const uint4*>(a.W2hi);
        const uint4* s2l = reinterpret_cast<const uint4*>(a.W2lo);
        uint4* d1 = reinterpret_cast<uint4*>(W1);
        uint4* d2 = reinterpret_cast<uint4*>(W2);
        for (int i = tid; i < 64 * 128 / 16; i += NTHREADS) { d1[i] = __ldg(s1 + i); d1[64 * 128 / 16 + i] = __ldg(s1l + i); }
        for (int i = tid; i < 32 * 128 / 16; i += NTHREADS) { d2[i] = __ldg(s2 + i); d2[32 * 128 / 16 + i] = __ldg(s2l + i); }
        for (int i = tid; i < 64; i += NTHREADS) sb1[i] = __ldg(a.b1 + i);
        for (int i = tid; i < 32; i += NTHREADS) sb2[i] = __ldg(a.b2 + i);
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const long long ntiles = (a.rows + TM - 1) / TM;
    const long long tile_begin = ntiles * blockIdx.x / gridDim.x;
    const long long tile_end = ntiles * (blockIdx.x + 1) / gridDim.x;
    const int nt = (int)(tile_end - tile_begin);

    if (warp < NLOAD / 32) {
        // =========================================================================== loaders (2 threads per row)
        const int row = tid >> 1, half = tid & 1;
        float4 xr[8];
        auto load = [&](int i) {
            const long long r = (tile_begin + i) * TM + row;
            if (r < a.rows) {
                const float4* p = reinterpret_cast<const float4*>(a.h + r * 64 + half * 32);
#pragma unroll
                for (int k = 0; k < 8; ++k) xr[k] = __ldg(p + k);
            } else {
#pragma unroll
                for (int k = 0; k < 8; ++k) xr[k] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        };
        if (nt > 0) load(0);
        for (int i = 0; i < nt; ++i) {
            const int buf = i & 1;
            const uint32_t ph = (uint32_t)((i >> 1) & 1);
            mbar_wait(&a1_empty[buf], ph ^ 1);
            uint8_t* hi_row = A1 + buf * A_BYTES + row * 128;
            uint8_t* lo_row = hi_row + SLAB;
            const float* x = reinterpret_cast<const float*>(xr);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float* v = x + k * 8;
                uint4 hi, lo;
                hi.x = pack_h2(v[0], v[1]); hi.y = pack_h2(v[2], v[3]); hi.z = pack_h2(v[4], v[5]); hi.w = pack_h2(v[6], v[7]);
                lo.x = pack_lo_h2(v[0], v[1], hi.x); lo.y = pack_lo_h2(v[2], v[3], hi.y);
                lo.z = pack_lo_h2(v[4], v[5], hi.z); lo.w = pack_lo_h2(v[6], v[7], hi.w);
                const int off = ((half * 4 + k) ^ (row & 7)) << 4;
                *reinterpret_cast<uint4*>(hi_row + off) = hi;
                *reinterpret_cast<uint4*>(lo_row + off) = lo;
            }
            if (i + 1 < nt) load(i + 1);             // in flight while the rest of the pipeline works on tile i
            fence_proxy_async();
            mbar_arrive(&a1_full[buf]);
        }
    } else if (warp < (NLOAD + NEPI) / 32) {
        // =========================================================================== epilogue 1: bias + ReLU -> A2
        const int quad = warp & 3;
        const int row = quad * 32 + lane;
        for (int i = 0; i < nt; ++i) {
            const int buf = i & 1;
            const uint32_t ph = (uint32_t)((i >> 1) & 1);
            mbar_wait(&d1_full[buf], ph);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * 64);
            uint4 hi[8], lo[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                uint32_t rv[8];
                tmem_ld8_nowait(taddr + k * 8, rv);
                tmem_wait_ld();
                float v[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) v[e] = fmaxf(__uint_as_float(rv[e]) + sb1[k * 8 + e], 0.f);
                hi[k].x = pack_h2(v[0], v[1]); hi[k].y = pack_h2(v[2], v[3]); hi[k].z = pack_h2(v[4], v[5]); hi[k].w = pack_h2(v[6], v[7]);
                lo[k].x = pack_lo_h2(v[0], v[1], hi[k].x); lo[k].y = pack_lo_h2(v[2], v[3], hi[k].y);
                lo[k].z = pack_lo_h2(v[4], v[5], hi[k].z); lo[k].w = pack_lo_h2(v[6], v[7], hi[k].w);
            }
            tc_fence_before();
            mbar_arrive(&d1_empty[buf]);
            mbar_wait(&a2_empty[buf], ph ^ 1);
            uint8_t* hi_row = A2 + buf * A_BYTES + row * 128;
            uint8_t* lo_row = hi_row + SLAB;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int off = (k ^ (row & 7)) << 4;
                *reinterpret_cast<uint4*>(hi_row + off) = hi[k];
                *reinterpret_cast<uint4*>(lo_row + off) = lo[k];
            }
            fence_proxy_async();
            mbar_arrive(&a2_full[buf]);
        }
    } else if (warp < MMA_WARP) {
        // =========================================================================== epilogue 2: + b2, filter-and-sum
        const int quad = warp & 3;
        const int row = quad * 32 + lane;
        const long long TF = (long long)a.T * a.F;
        for (int i = 0; i < nt; ++i) {
            const int buf = i & 1;
            const uint32_t ph = (uint32_t)((i >> 1) & 1);
            const long long r = (tile_begin + i) * TM + row;
            const bool valid = r < a.rows;
            // this row's input spectrum values: issued before the accumulator wait
            float2 xv[16];
            if (valid) {
                const float2* xp = reinterpret_cast<const float2*>(a.inpt) + r * a.M;
#pragma unroll
                for (int m = 0; m < 16; ++m)
                    if (m < a.M) xv[m] = __ldg(xp + m);
            }
            mbar_wait(&d2_full[buf], ph);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(128 + buf * 32);
            uint32_t rw[32];
#pragma unroll
            for (int k = 0; k < 4; ++k) tmem_ld8_nowait(taddr + k * 8, *reinterpret_cast<uint32_t(*)[8]>(&rw[k * 8]));
            tmem_wait_ld();
            tc_fence_before();
            mbar_arrive(&d2_empty[buf]);
            if (valid) {
                if (a.w_out) {                       // optional copy of the beam weights (debug tap), channels m*2+ri
                    float* wp = a.w_out + r * a.w_ld;
#pragma unroll
                    for (int c = 0; c < 32; ++c)
                        if (c < 2 * a.M) wp[c] = __uint_as_float(rw[c]) + sb2[c];
                }
                float yr = 0.f, yi = 0.f;
#pragma unroll
                for (int m = 0; m < 16; ++m) {
                    if (m < a.M) {
                        const float wr = __uint_as_float(rw[2 * m]) + sb2[2 * m];
                        const float wi = __uint_as_float(rw[2 * m + 1]) + sb2[2 * m + 1];
                        yr += wr * xv[m].x - wi * xv[m].y;
                        yi += wr * xv[m].y + wi * xv[m].x;
                    }
                }
                const long long b = r / TF, p = r - b * TF;
                a.out[(b * 2 + 0) * TF + p] = yr;
                a.out[(b * 2 + 1) * TF + p] = yi;
            }
        }
    } else {
        // =========================================================================== MMA issuer
        const uint32_t idesc1 = make_idesc(64), idesc2 = make_idesc(32);
        auto g1 = [&](int i) {
            const int buf = i & 1;
            const uint32_t ph = (uint32_t)((i >> 1) & 1);
            mbar_wait(&a1_full[buf], ph);
            mbar_wait(&d1_empty[buf], ph ^ 1);
            tc_fence_after();
            if (lane == 0) {
                const uint32_t ah = smem_u32(A1 + buf * A_BYTES), al = ah + SLAB;
                const uint32_t wh = smem_u32(W1), wl = wh + 64 * 128;
#pragma unroll
                for (int pass = 0; pass < 3; ++pass)
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_f16(tmem_base + buf * 64, make_desc((pass == 1 ? al : ah) + k * 32),
                                 make_desc((pass == 2 ? wl : wh) + k * 32), idesc1, (pass | k) ? 1u : 0u);
                umma_commit(&d1_full[buf]);
                umma_commit(&a1_empty[buf]);
            }
            __syncwarp();
        };
        auto g2 = [&](int i) {
            const int buf = i & 1;
            const uint32_t ph = (uint32_t)((i >> 1) & 1);
            mbar_wait(&a2_full[buf], ph);
            mbar_wait(&d2_empty[buf], ph ^ 1);
            tc_fence_after();
            if (lane == 0) {
                const uint32_t ah = smem_u32(A2 + buf * A_BYTES), al = ah + SLAB;
                const uint32_t wh = smem_u32(W2), wl = wh + 32 * 128;
#pragma unroll
                for (int pass = 0; pass < 3; ++pass)
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_f16(tmem_base + 128 + buf * 32, make_desc((pass == 1 ? al : ah) + k * 32),
                                 make_desc((pass == 2 ? wl : wh) + k * 32), idesc2, (pass | k) ? 1u : 0u);
                umma_commit(&d2_full[buf]);
                umma_commit(&a2_empty[buf]);
            }
            __syncwarp();
        };
        if (nt > 0) g1(0);
        for (int i = 0; i < nt; ++i) {
            if (i + 1 < nt) g1(i + 1);
            g2(i);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) tmem_dealloc(tmem_base, 256);
}

}  // namespace

bool head_fused_supported(const HeadArgs& a) {
    return a.M >= 1 && a.M <= 16 && a.W1hi && a.W1lo && a.W2hi && a.W2lo && a.b1 && a.b2 && a.rows > 0;
}

int launch_head_fused(const HeadArgs& a, cudaStream_t st) {
    if (!head_fused_supported(a)) return fail("head_fused: unsupported shape");
    EAB_TRY(ensure_dynamic_smem(reinterpret_cast<const void*>(head_fused_kernel), SMEM_BYTES));
    int sms = 0;
    EAB_TRY(device_sm_count(&sms));
    const long long ntiles = (a.rows + TM - 1) / TM;
    const int grid = (int)(ntiles < sms ? ntiles : sms);
    const double rows = (double)a.rows;
    ProfScope ps("head_fused", 2.0 * rows * (64.0 * 64 + 64.0 * 2 * a.M) + 8.0 * rows * a.M,
                 rows * (256.0 + 8.0 * a.M + 8.0), st);
    EAB_CUDA(launch_k(head_fused_kernel, dim3(grid), dim3(NTHREADS), (size_t)SMEM_BYTES, st, a));
    EAB_LAUNCH_CHECK("head_fused_kernel");
    return 0;
}

}  // namespace eab
