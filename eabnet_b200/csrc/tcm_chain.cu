// A whole chain of squeezed TCMs (GaGNet.py:285-326: 1x1 squeeze 256->64, PReLU -> norm -> dilated conv 64->64,
// PReLU -> norm -> 1x1 expand 64->256, + residual) as ONE persistent cooperative launch.
//
// Per layer the layer-by-layer path costs three stage launches and three GEMM launches of 10-50 us each on tensors
// (39 MB residual stream, 10 MB squeezed) that fit the 126 MB L2 several times over: the chain is launch- and
// round-trip-bound, not bandwidth- or FLOP-bound.  Here a CTA owns a fixed set of 128-row tiles (rows = frames of one
// batch item) for the whole chain and walks the layers in three phases:
//   A  x tile -> fp16 hi/lo operand -> tcgen05 GEMM with W_in (K 256, N 64) -> y rows + statistics of PReLU(y)
//   B  y rows (+ dilated halo rows of neighbouring tiles) -> PReLU -> norm -> operand, one K slab per tap -> GEMM with
//      W_dil (N 64) -> z rows + statistics of PReLU(z)
//   C  z rows -> PReLU -> norm -> operand -> GEMM with W_out (K 64, N 256) -> x += result (the CTA's own rows)
// InstanceNorm needs the statistics of ALL frames of a batch item (and phase B reads rows written by other CTAs), so
// A->B and B->C are grid-wide barriers (a counter in the workspace; the launch is cooperative, so every CTA is
// resident); C -> next layer's A needs none.  Up to three independent chains (glance, gaze-real, gaze-imaginary) share
// a launch so that 3 x B x ceil(T/128) tiles fill the 148 persistent CTAs evenly.  All GEMMs are 3-pass fp16 splits
// (hi*hi + lo*hi + hi*lo, fp32 accumulation in TMEM): fp32-grade, like the layer-by-layer path.
// One 512-thread CTA per SM.  A tile's whole operand (4 K slabs in phase A, one per tap in phase B) is staged at once,
// so a tile-phase is one memory round trip, one block barrier, one convergent burst of MMAs and one accumulator wait
// (the first version staged slab by slab with 8 warps per SM: every slab exposed a full L2 / HBM round trip, 19 us per
// tile-phase; a version that kept the NEXT tile's rows in flight in registers across the epilogue spilled 1.3 KB per
// thread at the 128-register cap of a 512-thread CTA and was slower still - profiles/r01_v6_chain_*).
// Everything that crosses CTAs inside the launch (x, y, z, statistics) is read with ld.global.cg (L2): L1 is not
// coherent across SMs.
#include "common.cuh"
#include "umma.cuh"

namespace eab {

namespace {

using namespace umma;

constexpr int TM = 128;
constexpr int NT = 512;
constexpr int TPR = NT / TM;                   // loader threads per operand row = epilogue column groups
constexpr int CPT = 64 / TPR;                  // channels of a 64-channel slab per loader thread / accumulator columns of N = 64 per epilogue thread
constexpr int CF4 = CPT / 4;                   // ... in float4
constexpr int SLAB = TM * 128;                 // one 64-channel fp16 K slab of an A operand (16 KB)
constexpr int A_UNIT = 2 * SLAB;               // hi | lo of one slab
constexpr int A_UNITS = 4;                     // phase A: 4 channel slabs; phase B: one unit per tap (kd <= 4); phase C: 1
constexpr int A_BYTES = A_UNITS * A_UNIT;      // 128 KB
constexpr int W_BYTES = 64 * 1024;             // phase A: 4 slabs x (hi 8 KB | lo 8 KB); B: taps; C: 256 rows x (hi | lo)
constexpr int W_HALF = W_BYTES / 2;
// 194.3 KB: stays under the 196 KB shared-memory carve-out step, which leaves the L1 (register spills, weights) its ~60 KB;
// a version with 5 KB more shared memory dropped to the next step and ran 25 % slower
constexpr int SMEM_BYTES = A_BYTES + W_BYTES + (3 * 64 + 2 * 64) * 4 + 128 + 1024;     // (10 mbarriers and the TMEM slot inside the 128)

__device__ __forceinline__ void grid_barrier(unsigned* ctr, unsigned& target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        target += gridDim.x;
        __threadfence();
        atomicAdd(ctr, 1u);
        unsigned spins = 0;
        while (true) {
            unsigned v;
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
            if (v >= target) break;
            __nanosleep(32);
            if (++spins > (1u << 26)) __trap();          // a protocol bug must fault, never hang the GPU
        }
        __threadfence();
    }
    __syncthreads();
}

// column sums over the 32 lanes of a warp for 8 columns per lane: every lane ends with the total of column lane & 7
// (halving butterfly: 7 + 2 shuffles instead of 5 per column)
__device__ __forceinline__ float warp_column_sums8(float (&v)[8], int lane) {
#pragma unroll
    for (int s = 4; s >= 1; s >>= 1) {
        const bool up = (lane & s) != 0;
#pragma unroll
        for (int i = 0; i < s; ++i) {
            const float mine = up ? v[i + s] : v[i];
            const float other = up ? v[i] : v[i + s];
            v[i] = mine + __shfl_xor_sync(0xffffffffu, other, s);
        }
    }
    float r = v[0];
    r += __shfl_xor_sync(0xffffffffu, r, 8);
    r += __shfl_xor_sync(0xffffffffu, r, 16);
    return r;
}

__device__ __forceinline__ float4 ldcg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }

// CPT fp32 values (channels q*CPT .. of a 64-channel slab row) -> fp16 hi / lo, 16-byte chunks of the swizzled row
__device__ __forceinline__ void store_part(uint8_t* unit, int row, int q, const float (&v)[CPT]) {
    uint8_t* hi_row = unit + row * 128;
    uint8_t* lo_row = hi_row + SLAB;
#pragma unroll
    for (int k = 0; k < CPT / 8; ++k) {
        const float* p = v + k * 8;
        uint4 hi, lo;
        hi.x = pack_h2(p[0], p[1]); hi.y = pack_h2(p[2], p[3]); hi.z = pack_h2(p[4], p[5]); hi.w = pack_h2(p[6], p[7]);
        lo.x = pack_lo_h2(p[0], p[1], hi.x); lo.y = pack_lo_h2(p[2], p[3], hi.y);
        lo.z = pack_lo_h2(p[4], p[5], hi.z); lo.w = pack_lo_h2(p[6], p[7], hi.w);
        const int off = ((q * (CPT / 8) + k) ^ (row & 7)) << 4;
        *reinterpret_cast<uint4*>(hi_row + off) = hi;
        *reinterpret_cast<uint4*>(lo_row + off) = lo;
    }
}

// InstanceNorm's statistics and the dilated halo couple the frames of ONE utterance only (EaBNet.py:554-571), so with
// CLUSTER the launch is not cooperative: a thread-block cluster of ceil(T / 128) CTAs owns one (chain, utterance) unit, a
// CTA one 128-frame tile of it for the whole chain, and the two synchronisations per layer are cluster barriers among those
// few CTAs instead of barriers across the whole grid (measured on the grid-barrier form: 0.81 ms per 6-layer launch at
// 6 % tensor-pipe and 3 % DRAM utilisation - barrier skew of 148 CTAs holding 2 or 3 tiles each).  The hardware schedules
// the clusters as SMs free up.  T > 1024 (more than 8 tiles per utterance) keeps the cooperative form.
__device__ __forceinline__ void cluster_barrier() {
    __threadfence();
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <int KD, bool GATED, bool CLUSTER>
__global__ void __launch_bounds__(NT, 1) tcm_chain_kernel(const TcmChainArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* As = smem;
    uint8_t* Ws = smem + A_BYTES;
    float* coef = reinterpret_cast<float*>(Ws + W_BYTES);        // [3][64] scale, shift, PReLU slope of the consumer's transform
    float* coefR = coef + 3 * 64;                                // [2][64] scale, shift of the gate (right) branch (gated TCMs)
    uint64_t* bar = reinterpret_cast<uint64_t*>(coefR + 2 * 64);
    uint64_t* wbar = bar + 1;                                    // weight images have landed in Ws (bulk copies)
    uint64_t* wfull = bar + 2;                                   // [4] gated phase B: weight ring slot filled / consumed
    uint64_t* wempty = bar + 6;                                  // [4]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 10);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        mbar_init(bar, 1); mbar_init(wbar, 1);
        for (int i = 0; i < 4; ++i) { mbar_init(&wfull[i], 1); mbar_init(&wempty[i], 1); }
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(tmem_slot, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int tiles_per_chain = a.B * a.tiles_per_b;
    const int total = a.nchains * tiles_per_chain;
    // CLUSTER: one tile per CTA, the tiles_per_b CTAs of a cluster are the consecutive tiles of one (chain, utterance)
    const int tile_begin = CLUSTER ? (int)blockIdx.x : (int)((long long)total * blockIdx.x / gridDim.x);
    const int tile_end = CLUSTER ? (int)blockIdx.x + 1 : (int)((long long)total * (blockIdx.x + 1) / gridDim.x);
    const bool in_stats = a.instance_norm != 0;

    uint32_t par = 0;                  // parity of the next completion of `bar`
    uint32_t wpar = 0;                 // ... of `wbar`
    uint32_t wf_par = 0, we_par = 0;   // ... of wfull[i] / wempty[i] (bit i; tracked by warp 0 only)
    unsigned bar_target = 0;
    // loader role: TPR threads per operand row (CPT channels of every slab each); epilogue role: thread = (TMEM lane = row,
    // column group cg of TPR)
    const int lrow = tid / TPR, lq = tid % TPR;
    const int quad = warp & 3, cg = warp >> 2;
    const int erow = quad * 32 + lane;
    const uint32_t idesc64 = make_idesc(64), idesc256 = make_idesc(256);

    struct TileId { int chain, b, t0; };
    auto tile_of = [&](int ti) {
        TileId t;
        t.chain = ti / tiles_per_chain;
        const int rem = ti - t.chain * tiles_per_chain;
        t.b = rem / a.tiles_per_b;
        t.t0 = (rem - t.b * a.tiles_per_b) * TM;
        return t;
    };

    // `units` K slabs in As against `units` weight slabs of `w_stride` bytes (hi at Ws, lo at Ws + W_HALF): per slab three
    // passes (hi*hi, lo*hi, hi*lo) of four K = 16 steps, issued by warp 0 as one convergent loop
    auto issue = [&](uint32_t idesc, int units, int w_stride, unsigned dmask = 0u, unsigned fmask = 1u) {
        fence_proxy_async();
        __syncthreads();
        if (warp == 0) {
            tc_fence_after();
            const uint32_t a0 = desc_lo(smem_u32(As)), w0 = desc_lo(smem_u32(Ws));
            for (int u = 0; u < units; ++u) {
                const uint32_t ah = a0 + (uint32_t)((u * A_UNIT) >> 4), al = ah + (SLAB >> 4);
                const uint32_t wh = w0 + (uint32_t)((u * w_stride) >> 4), wl = wh + (W_HALF >> 4);
                const uint32_t d = tmem_base + (((dmask >> u) & 1u) ? 64u : 0u);       // gated: value cols 0-63, gate cols 64-127
                umma_f16_lo_elect_x4(d, ah, wh, idesc, ((fmask >> u) & 1u) ? 0u : 1u);
                umma_f16_lo_elect_x4(d, al, wh, idesc, 1u);
                umma_f16_lo_elect_x4(d, ah, wl, idesc, 1u);
            }
            umma_commit_elect(bar);
        }
    };
    auto wait_mma = [&]() { mbar_wait(bar, par); par ^= 1u; };
    // Weight images -> Ws as bulk copies issued by ONE thread (they complete on `wbar` while the other threads load and
    // transform the operand rows): w_begin(total bytes), any number of w_copy, then w_wait() by everybody before the MMAs.
    // (The first version copied them with per-thread loads: one exposed L2 round trip per image, 8 per round of phase B -
    // 38 % of a layer's time in "B load + store".)  The caller guarantees that the MMAs that last read Ws have completed.
    auto w_begin = [&](int bytes) { if (tid == 0) mbar_arrive_expect_tx(wbar, (uint32_t)bytes); };
    auto w_copy = [&](const float* src, int dst_off, int bytes) { if (tid == 0) bulk_copy_g2s(Ws + dst_off, src, (uint32_t)bytes, wbar); };
    auto w_wait = [&]() { mbar_wait_spin(wbar, wpar); wpar ^= 1u; };
    // per-channel transform of the consumer: PReLU(alpha) then x*s + h, (s, h) from instance statistics or precomputed
    auto coef_of = [&](const double* stats, int b, int c, unsigned off_sc, unsigned off_sh, float& s, float& h) {
        if (in_stats) {
            const double* st = stats + ((size_t)b * 64 + c) * 2;
            const double mean = __ldcg(st) * (double)a.inv_count;
            double var = __ldcg(st + 1) * (double)a.inv_count - mean * mean;
            if (var < 0.0) var = 0.0;
            const double rstd = rsqrt(var + 1e-5);
            const double g = (double)__ldg(a.blob + off_sc + c);
            s = (float)(g * rstd);
            h = (float)((double)__ldg(a.blob + off_sh + c) - mean * g * rstd);
        } else {
            s = __ldg(a.blob + off_sc + c);
            h = __ldg(a.blob + off_sh + c);
        }
    };
    auto load_coef = [&](const double* stats, int b, unsigned off_sc, unsigned off_sh, unsigned off_al) {
        if (tid < 64) {
            float s, h;
            coef_of(stats, b, tid, off_sc, off_sh, s, h);
            coef[tid] = s; coef[64 + tid] = h; coef[128 + tid] = __ldg(a.blob + off_al + tid);
        }
    };
    auto load_coef_right = [&](const double* stats, int b, unsigned off_sc, unsigned off_sh) {      // threads 64..127
        if (tid >= 64 && tid < 128) {
            float s, h;
            coef_of(stats, b, tid - 64, off_sc, off_sh, s, h);
            coefR[tid - 64] = s; coefR[tid] = h;
        }
    };
    auto transform_right = [&](float (&v)[CPT], const float* alpha) {
#pragma unroll
        for (int i = 0; i < CPT; ++i) {
            const int c = lq * CPT + i;
            v[i] = fmaf(prelu_f(v[i], __ldg(alpha + c)), coefR[c], coefR[64 + c]);
        }
    };
    auto transform = [&](float (&v)[CPT]) {
#pragma unroll
        for (int i = 0; i < CPT; ++i) {
            const int c = lq * CPT + i;
            v[i] = fmaf(prelu_f(v[i], coef[128 + c]), coef[c], coef[64 + c]);
        }
    };
    // epilogue of phases A / B: CPT accumulator columns of this thread's row -> global rows + statistics of PReLU(value)
    // stats2 / off_alpha2: a second statistics set of the same values under another PReLU (the gate branch of a gated TCM);
    // gate: the stored value is acc[col] * sigmoid(acc[64 + col]) (phase B of a gated TCM, EaBNet.py:575)
    auto epilogue64 = [&](float* dst, int b, int t0, double* stats, unsigned off_alpha, double* stats2, unsigned off_alpha2, bool gate) {
        const int t = t0 + erow;
        const bool valid = t < a.T;
        // 8 columns at a time: the next tile's prefetched rows stay in registers across this epilogue
#pragma unroll
        for (int k = 0; k < CPT / 8; ++k) {
            uint32_t rv[8];
            const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(cg * CPT + k * 8);
            tmem_ld8_nowait(taddr, rv);
            tmem_wait_ld();
            float v[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = __uint_as_float(rv[e]);
            if (GATED && gate) {
                uint32_t rg[8];
                tmem_ld8_nowait(taddr + 64, rg);
                tmem_wait_ld();
#pragma unroll
                for (int e = 0; e < 8; ++e) v[e] *= sigmoid_f(__uint_as_float(rg[e]));
            }
            if (valid) st_global_256(dst + ((size_t)b * a.T + t) * 64 + cg * CPT + k * 8, v);
            if (in_stats) {
                const float* al = a.blob + off_alpha + cg * CPT + k * 8;
                float p[8], w[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) { p[i] = valid ? prelu_f(v[i], __ldg(al + i)) : 0.f; w[i] = p[i]; }
                const float s1 = warp_column_sums8(w, lane);
#pragma unroll
                for (int i = 0; i < 8; ++i) w[i] = p[i] * p[i];
                const float s2 = warp_column_sums8(w, lane);
                if (lane < 8) {
                    // per-warp partial sums (fixed shuffle order) straight to the fp64 accumulators: an fp32 shared-memory
                    // accumulation across warps would make the statistics depend on the warps' arrival order
                    double* d1 = stats + ((size_t)b * 64 + cg * CPT + k * 8 + lane) * 2;
                    atomicAdd(d1, (double)s1);
                    atomicAdd(d1 + 1, (double)s2);
                }
                if (GATED && stats2) {
                    const float* al2 = a.blob + off_alpha2 + cg * CPT + k * 8;
#pragma unroll
                    for (int i = 0; i < 8; ++i) { p[i] = valid ? prelu_f(v[i], __ldg(al2 + i)) : 0.f; w[i] = p[i]; }
                    const float r1 = warp_column_sums8(w, lane);
#pragma unroll
                    for (int i = 0; i < 8; ++i) w[i] = p[i] * p[i];
                    const float r2 = warp_column_sums8(w, lane);
                    if (lane < 8) {                        // second set: straight to the global accumulators (one per warp and column)
                        double* d2 = stats2 + ((size_t)b * 64 + cg * CPT + k * 8 + lane) * 2;
                        atomicAdd(d2, (double)r1);
                        atomicAdd(d2 + 1, (double)r2);
                    }
                }
            }
        }
        tc_fence_before();
        __syncthreads();
    };

    // diagnostics (builds with -DEAB_CHAIN_DEBUG only: the counters cost 30 registers, and at the 128-register cap of a
    // 512-thread CTA every spilled byte goes through an L1 that is almost entirely carved out as shared memory):
    // cycles of CTA 0 per section [0] A load+store [1] A mma [2] A epilogue [3] barrier 1 [4] B load+store [5] B mma
    // [6] B epilogue [7] barrier 2 [8] C load+store [9] C mma [10] C epilogue [11] total [12] tiles
#ifdef EAB_CHAIN_DEBUG
    const bool dbg_on = a.dbg != nullptr && blockIdx.x == 0 && tid == 0;
    long long dc[13];
#pragma unroll
    for (int i = 0; i < 13; ++i) dc[i] = 0;
    long long tk = clock64();
    const long long t_start = tk;
    auto tick = [&](int slot) { if (dbg_on) { const long long n = clock64(); dc[slot] += n - tk; tk = n; } };
#else
    auto tick = [&](int) {};
#endif

    for (int l = 0; l < a.nlayers; ++l) {
        // ======================================================================= phase A: y = W_in x
        {
            int loaded = -1;
            float4 xv[4][CF4];                             // [slab][CPT channels] of this thread's row
            auto fetch = [&](int ti) {
                const TileId tl = tile_of(ti);
                const int t = tl.t0 + lrow;
                const float* xsrc = (l == 0 ? a.x_in[tl.chain] : a.x_buf[tl.chain]);
                // rows past T: a clamped row is loaded and multiplied into nothing (their outputs are never stored and
                // are excluded from the statistics), so the loads stay unconditional and in registers
                const int tc = t < a.T ? t : a.T - 1;
                const float* xrow = xsrc + ((size_t)tl.b * a.T + tc) * 256 + lq * CPT;
#pragma unroll
                for (int s = 0; s < 4; ++s)
#pragma unroll
                    for (int k = 0; k < CF4; ++k) xv[s][k] = ldcg4(xrow + s * 64 + k * 4);
            };
            // (Tried in the cluster form: phase C's epilogue leaving the new x rows in As as the next layer's operand and
            // starting its weight copies - bit-exact, "A load + store" 6.1k -> 2.5k cycles per layer, but the time reappeared in
            // the waits of phases B and C: 1.90 ms either way.  What bounds a tile-phase is the L2 -> shared-memory latency of
            // the 288 KB of weight images every CTA streams per layer.)
            constexpr bool fused_in = false;
            if (!fused_in && tile_begin < tile_end) fetch(tile_begin);
            for (int ti = tile_begin; ti < tile_end; ++ti) {
                const TileId tl = tile_of(ti);
                const TcmChainLayer& L = a.L[tl.chain * a.nlayers + l];
                const bool neww = !fused_in && loaded != tl.chain;
                if (neww) {
                    w_begin(2 * 4 * 64 * 128);
                    w_copy(a.blob + L.win_hi, 0, 4 * 64 * 128);
                    w_copy(a.blob + L.win_lo, W_HALF, 4 * 64 * 128);
                    loaded = tl.chain;
                }
                if (!fused_in) {
#pragma unroll
                    for (int s = 0; s < 4; ++s) store_part(As + s * A_UNIT, lrow, lq, *reinterpret_cast<const float(*)[CPT]>(&xv[s][0]));
                }
                if (neww || fused_in) w_wait();
                tick(0);
                issue(idesc64, 4, 64 * 128);
                if (!fused_in && ti + 1 < tile_end) fetch(ti + 1);      // the next tile's rows: in flight under the MMAs and the epilogue
                wait_mma();
                tc_fence_after();
                tick(1);
                epilogue64(a.y[tl.chain], tl.b, tl.t0, a.stats + L.st_d, L.al_d, GATED ? a.stats + L.st_r : nullptr, L.al_r, false);
                tick(2);
            }
        }
        if (CLUSTER) cluster_barrier(); else grid_barrier(a.barrier, bar_target);
        tick(3);
        // ======================================================================= phase B: z = W_dil * norm(PReLU(y)) (dilated)
        if constexpr (GATED) {
            // two branches (EaBNet.py:554-566, 575): value = W_left * norm_l(PReLU_l(y)), gate = W_right * norm_r(PReLU_r(y)), both
            // dilated.  "Stage once, shift by descriptor": the tile's y rows plus the dilated halo (frames t0 - back ..
            // t0 + 127 + fwd, at most 256 rows) are normalised ONCE per branch into a plane [hi | lo] in shared memory, and every
            // one of the 2 KD (branch, tap) units is that plane viewed through a row-shifted UMMA descriptor.  (The first
            // version re-staged the 128 operand rows of every unit: 10 transforms of the tile per layer instead of 2-4,
            // 38 % of a layer's time.)  The units' weights (16 KB each) stream through a 4-slot ring of bulk copies that warp 0
            // keeps three units ahead of its own MMAs.
            constexpr int NU = 2 * KD;
            for (int ti = tile_begin; ti < tile_end; ++ti) {
                const TileId tl = tile_of(ti);
                const TcmChainLayer& L = a.L[tl.chain * a.nlayers + l];
                int back = 0, fwd = 0;
#pragma unroll
                for (int k = 0; k < KD; ++k) { back = max(back, (int)L.dt[k]); fwd = max(fwd, -(int)L.dt[k]); }
                const int R = TM + back + fwd;                 // <= 256 (checked by the launcher)
                const int PB = ((R + 7) & ~7) * 128;           // bytes of one plane image
                // weights of the first four units: in flight under the transform
                if (warp == 0) {
#pragma unroll
                    for (int g = 0; g < 4 && g < NU; ++g) {
                        if (lane == 0) {
                            const bool right = g >= KD;
                            const int tap = g % KD;
                            mbar_arrive_expect_tx(&wfull[g], 2 * 64 * 128);
                            bulk_copy_g2s(Ws + g * 64 * 128, a.blob + (right ? L.wr_hi : L.wd_hi) + tap * (64 * 128 / 4), 64 * 128, &wfull[g]);
                            bulk_copy_g2s(Ws + W_HALF + g * 64 * 128, a.blob + (right ? L.wr_lo : L.wd_lo) + tap * (64 * 128 / 4), 64 * 128, &wfull[g]);
                        }
                    }
                }
                load_coef(a.stats + L.st_d, tl.b, L.sc_d, L.sh_d, L.al_d);
                load_coef_right(a.stats + L.st_r, tl.b, L.sc_r, L.sh_r);
                __syncthreads();
                {   // thread = (plane row, half of the 64 channels)
                    const int rho = tid >> 1, half = tid & 1;
                    if (rho < R) {
                        const int t = tl.t0 - back + rho;
                        const bool ok = t >= 0 && t < a.T;
                        const int tc = t < 0 ? 0 : (t >= a.T ? a.T - 1 : t);
                        const float* yrow = a.y[tl.chain] + ((size_t)tl.b * a.T + tc) * 64 + half * 32;
                        float4 yv[8];
#pragma unroll
                        for (int q = 0; q < 8; ++q) yv[q] = ldcg4(yrow + q * 4);
                        const float* alr = a.blob + L.al_r + half * 32;
                        uint8_t* rowL = As + rho * 128;                       // plane (left, hi); lo at + PB
                        uint8_t* rowR = As + 2 * PB + rho * 128;              // plane (right, hi)
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const float y8[8] = {yv[2 * k].x, yv[2 * k].y, yv[2 * k].z, yv[2 * k].w, yv[2 * k + 1].x, yv[2 * k + 1].y, yv[2 * k + 1].z, yv[2 * k + 1].w};
                            float vl[8], vr[8];
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const int c = half * 32 + k * 8 + i;
                                // literal zeros AFTER the norm for frames outside the utterance (EaBNet.py:557)
                                vl[i] = ok ? fmaf(prelu_f(y8[i], coef[128 + c]), coef[c], coef[64 + c]) : 0.f;
                                vr[i] = ok ? fmaf(prelu_f(y8[i], __ldg(alr + k * 8 + i)), coefR[c], coefR[64 + c]) : 0.f;
                            }
                            const int off = ((half * 4 + k) ^ (rho & 7)) << 4;
                            uint4 hi, lo;
                            hi.x = pack_h2(vl[0], vl[1]); hi.y = pack_h2(vl[2], vl[3]); hi.z = pack_h2(vl[4], vl[5]); hi.w = pack_h2(vl[6], vl[7]);
                            lo.x = pack_lo_h2(vl[0], vl[1], hi.x); lo.y = pack_lo_h2(vl[2], vl[3], hi.y);
                            lo.z = pack_lo_h2(vl[4], vl[5], hi.z); lo.w = pack_lo_h2(vl[6], vl[7], hi.w);
                            *reinterpret_cast<uint4*>(rowL + off) = hi;
                            *reinterpret_cast<uint4*>(rowL + PB + off) = lo;
                            hi.x = pack_h2(vr[0], vr[1]); hi.y = pack_h2(vr[2], vr[3]); hi.z = pack_h2(vr[4], vr[5]); hi.w = pack_h2(vr[6], vr[7]);
                            lo.x = pack_lo_h2(vr[0], vr[1], hi.x); lo.y = pack_lo_h2(vr[2], vr[3], hi.y);
                            lo.z = pack_lo_h2(vr[4], vr[5], hi.z); lo.w = pack_lo_h2(vr[6], vr[7], hi.w);
                            *reinterpret_cast<uint4*>(rowR + off) = hi;
                            *reinterpret_cast<uint4*>(rowR + PB + off) = lo;
                        }
                    }
                }
                tick(4);
                fence_proxy_async();
                __syncthreads();
                if (warp == 0) {
                    tc_fence_after();
                    const uint32_t a0 = desc_lo(smem_u32(As)), w0 = desc_lo(smem_u32(Ws));
#pragma unroll 1
                    for (int g = 0; g < NU; ++g) {
                        const int slot = g & 3;
                        const bool right = g >= KD;
                        const int tap = g % KD;
                        mbar_wait_spin(&wfull[slot], (wf_par >> slot) & 1u);
                        wf_par ^= 1u << slot;
                        tc_fence_after();
                        const uint32_t ah = a0 + (uint32_t)(((right ? 2 : 0) * PB + (back - (int)L.dt[tap]) * 128) >> 4), al = ah + (uint32_t)(PB >> 4);
                        const uint32_t wh = w0 + (uint32_t)((slot * 64 * 128) >> 4), wl = wh + (W_HALF >> 4);
                        const uint32_t d = tmem_base + (right ? 64u : 0u);             // value cols 0-63, gate cols 64-127
                        umma_f16_lo_elect_x4(d, ah, wh, idesc64, tap == 0 ? 0u : 1u);
                        umma_f16_lo_elect_x4(d, al, wh, idesc64, 1u);
                        umma_f16_lo_elect_x4(d, ah, wl, idesc64, 1u);
                        umma_commit_elect(&wempty[slot]);
                        if (g >= 1 && g + 3 < NU) {            // unit g + 3 -> the slot unit g - 1 has just finished with
                            const int ps = (g - 1) & 3, gn = g + 3;
                            mbar_wait_spin(&wempty[ps], (we_par >> ps) & 1u);
                            we_par ^= 1u << ps;
                            if (lane == 0) {
                                const bool rn = gn >= KD;
                                const int tn = gn % KD;
                                mbar_arrive_expect_tx(&wfull[ps], 2 * 64 * 128);
                                bulk_copy_g2s(Ws + ps * 64 * 128, a.blob + (rn ? L.wr_hi : L.wd_hi) + tn * (64 * 128 / 4), 64 * 128, &wfull[ps]);
                                bulk_copy_g2s(Ws + W_HALF + ps * 64 * 128, a.blob + (rn ? L.wr_lo : L.wd_lo) + tn * (64 * 128 / 4), 64 * 128, &wfull[ps]);
                            }
                            __syncwarp();
                        }
                    }
                    umma_commit_elect(bar);
                    // drain the wempty completions nobody waited for, so that every slot's parity is known at the next use
#pragma unroll 1
                    for (int g = (NU > 4 ? NU - 4 : 0); g < NU; ++g) {
                        const int ps = g & 3;
                        mbar_wait_spin(&wempty[ps], (we_par >> ps) & 1u);
                        we_par ^= 1u << ps;
                    }
                }
                wait_mma();
                tc_fence_after();
                tick(5);
                epilogue64(a.z[tl.chain], tl.b, tl.t0, a.stats + L.st_o, L.al_o, nullptr, 0u, true);
                tick(6);
            }
        } else
        {
            int loaded = -1;
            float4 yv[KD][CF4];                            // [tap][CPT channels]
            unsigned okmask = 0;
            auto fetch = [&](int ti) {
                const TileId tl = tile_of(ti);
                const TcmChainLayer& L = a.L[tl.chain * a.nlayers + l];
                const int t = tl.t0 + lrow;
                okmask = 0;
#pragma unroll
                for (int k = 0; k < KD; ++k) {
                    {
                        const int ts = t - L.dt[k];
                        if (t < a.T && ts >= 0 && ts < a.T) okmask |= 1u << k;
                        // unconditional loads from a clamped row (zeroed below): keeps the rows in registers
                        const int tc = ts < 0 ? 0 : (ts >= a.T ? a.T - 1 : ts);
                        const float* yrow = a.y[tl.chain] + ((size_t)tl.b * a.T + tc) * 64 + lq * CPT;
#pragma unroll
                        for (int q = 0; q < CF4; ++q) yv[k][q] = ldcg4(yrow + q * 4);
                    }
                }
            };
            unsigned okcur = 0;
            if (tile_begin < tile_end) fetch(tile_begin);
            for (int ti = tile_begin; ti < tile_end; ++ti) {
                okcur = okmask;
                const TileId tl = tile_of(ti);
                const TcmChainLayer& L = a.L[tl.chain * a.nlayers + l];
                const bool neww = loaded != tl.chain;
                if (neww) {
                    w_begin(2 * KD * 64 * 128);
                    w_copy(a.blob + L.wd_hi, 0, KD * 64 * 128);
                    w_copy(a.blob + L.wd_lo, W_HALF, KD * 64 * 128);
                    loaded = tl.chain;
                }
                load_coef(a.stats + L.st_d, tl.b, L.sc_d, L.sh_d, L.al_d);
                __syncthreads();
#pragma unroll
                for (int k = 0; k < KD; ++k) {
                    {
                        float v[CPT];
#pragma unroll
                        for (int q = 0; q < CF4; ++q) *reinterpret_cast<float4*>(&v[q * 4]) = yv[k][q];
                        transform(v);
                        const bool ok = (okcur >> k) & 1u;
#pragma unroll
                        for (int i = 0; i < CPT; ++i) v[i] = ok ? v[i] : 0.f;       // literal zeros AFTER the norm (GaGNet.py:313)
                        store_part(As + k * A_UNIT, lrow, lq, v);
                    }
                }
                if (neww) w_wait();
                tick(4);
                issue(idesc64, KD, 64 * 128);
                if (ti + 1 < tile_end) fetch(ti + 1);      // the next tile's rows: in flight under the MMAs and the epilogue
                wait_mma();
                tc_fence_after();
                tick(5);
                epilogue64(a.z[tl.chain], tl.b, tl.t0, a.stats + L.st_o, L.al_o, nullptr, 0u, false);
                tick(6);
            }
        }
        if (CLUSTER) cluster_barrier(); else grid_barrier(a.barrier, bar_target);
        tick(7);
        // ======================================================================= phase C: x += W_out * norm(PReLU(z))
        {
            int loaded = -1;
            float4 zv[CF4];
            bool zok = false;
            auto fetch = [&](int ti) {
                const TileId tl = tile_of(ti);
                const int t = tl.t0 + lrow;
                zok = t < a.T;
                const int tc = zok ? t : a.T - 1;
                const float* zrow = a.z[tl.chain] + ((size_t)tl.b * a.T + tc) * 64 + lq * CPT;
#pragma unroll
                for (int q = 0; q < CF4; ++q) zv[q] = ldcg4(zrow + q * 4);
            };
            if (tile_begin < tile_end) fetch(tile_begin);
            for (int ti = tile_begin; ti < tile_end; ++ti) {
                const bool zcur = zok;
                const TileId tl = tile_of(ti);
                const TcmChainLayer& L = a.L[tl.chain * a.nlayers + l];
                const bool neww = loaded != tl.chain;
                if (neww) {
                    w_begin(4 * 128 * 128);
                    w_copy(a.blob + L.wo_hi[0], 0, 128 * 128);
                    w_copy(a.blob + L.wo_hi[1], 128 * 128, 128 * 128);
                    w_copy(a.blob + L.wo_lo[0], W_HALF, 128 * 128);
                    w_copy(a.blob + L.wo_lo[1], W_HALF + 128 * 128, 128 * 128);
                    loaded = tl.chain;
                }
                load_coef(a.stats + L.st_o, tl.b, L.sc_o, L.sh_o, L.al_o);
                __syncthreads();
                {
                    float v[CPT];
#pragma unroll
                    for (int q = 0; q < CF4; ++q) *reinterpret_cast<float4*>(&v[q * 4]) = zv[q];
                    transform(v);
#pragma unroll
                    for (int i = 0; i < CPT; ++i) v[i] = zcur ? v[i] : 0.f;
                    store_part(As, lrow, lq, v);
                }
                if (neww) w_wait();
                tick(8);
                issue(idesc256, 1, 0);
                // this row's residual (256 / TPR columns of x) and the next tile's z rows: in flight under the MMAs
                const int t = tl.t0 + erow;
                const bool valid = t < a.T;
                float* xdst = a.x_buf[tl.chain] + ((size_t)tl.b * a.T + t) * 256 + cg * (4 * CPT);
                // the whole residual of this thread (4 * CPT columns) in flight under the MMAs; unconditional loads from a
                // clamped row keep it in registers
                const int tclamp = t < a.T ? t : a.T - 1;
                const float* xres = (l == 0 ? a.x_in[tl.chain] : a.x_buf[tl.chain]) + ((size_t)tl.b * a.T + tclamp) * 256 + cg * (4 * CPT);
                float4 xr[CPT];
#pragma unroll
                for (int k = 0; k < CPT; ++k) xr[k] = ldcg4(xres + k * 4);
                if (ti + 1 < tile_end) fetch(ti + 1);      // the next tile's z rows
                wait_mma();
                tc_fence_after();
                tick(9);
                const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(cg * 4 * CPT);
#pragma unroll
                for (int g = 0; g < CPT / 2; ++g) {
                    uint32_t rv[8];
                    tmem_ld8_nowait(taddr + g * 8, rv);
                    tmem_wait_ld();
                    const float4 x0 = xr[2 * g], x1 = xr[2 * g + 1];
                    float o[8];
                    o[0] = __uint_as_float(rv[0]) + x0.x; o[1] = __uint_as_float(rv[1]) + x0.y;
                    o[2] = __uint_as_float(rv[2]) + x0.z; o[3] = __uint_as_float(rv[3]) + x0.w;
                    o[4] = __uint_as_float(rv[4]) + x1.x; o[5] = __uint_as_float(rv[5]) + x1.y;
                    o[6] = __uint_as_float(rv[6]) + x1.z; o[7] = __uint_as_float(rv[7]) + x1.w;
                    if (valid) st_global_256(xdst + g * 8, o);
                }
                tc_fence_before();
                __syncthreads();
                tick(10);
            }
        }
        // no grid barrier: the next layer's phase A reads only this CTA's own rows of x_buf (ld.cg, after the block barrier)
    }
#ifdef EAB_CHAIN_DEBUG
    if (dbg_on) {
        dc[11] = clock64() - t_start;
        dc[12] = tile_end - tile_begin;
        for (int i = 0; i < 13; ++i) a.dbg[i] = (unsigned long long)dc[i];
    }
#endif
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 256);
}

}  // namespace

bool tcm_chain_supported(const TcmChainArgs& a) {
    if (a.nchains < 1 || a.nchains > 3 || a.nlayers < 1 || a.nchains * a.nlayers > kMaxChainLayers || a.B < 1 || a.T < 1) return false;
    return a.gated ? (a.kd == 3 || a.kd == 5) : (a.kd >= 1 && a.kd <= A_UNITS);
}

int launch_tcm_chain(const TcmChainArgs& a_in, cudaStream_t st) {
    TcmChainArgs a = a_in;
    if (!tcm_chain_supported(a)) return fail("tcm_chain: unsupported shape");
    a.tiles_per_b = (a.T + TM - 1) / TM;
    const bool cluster = a.tiles_per_b <= 8 && !a.no_cluster;
    const void* kernel = nullptr;
#define EAB_CHAIN_PICK(CL)                                                                               \
    (a.gated ? (a.kd == 3 ? reinterpret_cast<const void*>(tcm_chain_kernel<3, true, CL>)                  \
                : a.kd == 5 ? reinterpret_cast<const void*>(tcm_chain_kernel<5, true, CL>) : nullptr)     \
             : (a.kd == 1 ? reinterpret_cast<const void*>(tcm_chain_kernel<1, false, CL>)                 \
                : a.kd == 2 ? reinterpret_cast<const void*>(tcm_chain_kernel<2, false, CL>)               \
                : a.kd == 3 ? reinterpret_cast<const void*>(tcm_chain_kernel<3, false, CL>)               \
                : a.kd == 4 ? reinterpret_cast<const void*>(tcm_chain_kernel<4, false, CL>) : nullptr))
    kernel = cluster ? EAB_CHAIN_PICK(true) : EAB_CHAIN_PICK(false);
#undef EAB_CHAIN_PICK
    if (!kernel) return fail("tcm_chain: unsupported kernel size");
    EAB_TRY(ensure_dynamic_smem(kernel, SMEM_BYTES));
    const long long total = (long long)a.nchains * a.B * a.tiles_per_b;
    if (total >= (1ll << 30)) return fail("tcm_chain: too many tiles");
    const double rows = (double)a.nchains * a.B * a.T * a.nlayers;
    const double nb = a.gated ? 2.0 : 1.0;
    // algorithmic bytes: per TCM the residual stream in and out (256 channels) + the squeezed tensor once
    ProfScope ps("tcm_chain", 2.0 * rows * (256.0 * 64 + nb * a.kd * 64.0 * 64 + 64.0 * 256), 4.0 * rows * (256 + 64 + 256), st,
                 4.0 * rows * (256 + 64 + 64 * a.kd * nb + 64 + 64 + 256 + 256));
    if (cluster) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)total); cfg.blockDim = dim3(NT); cfg.dynamicSmemBytes = SMEM_BYTES; cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = (unsigned)a.tiles_per_b; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        void* params[1] = {&a};
        EAB_CUDA(cudaLaunchKernelExC(&cfg, kernel, params));
    } else {
        int sms = 0, dev = 0, coop = 0;
        EAB_TRY(device_sm_count(&sms));
        EAB_CUDA(cudaGetDevice(&dev));
        EAB_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
        if (!coop) return fail("tcm_chain: the device does not support cooperative launches");
        const int grid = (int)(total < sms ? total : sms);      // one persistent CTA per SM (194 KB of shared memory each)
        void* params[1] = {&a};
        EAB_CUDA(cudaLaunchCooperativeKernel(kernel, dim3(grid), dim3(NT), params, (size_t)SMEM_BYTES, st));
    }
    EAB_LAUNCH_CHECK("tcm_chain_kernel");
    return 0;
}

}  // namespace eab
