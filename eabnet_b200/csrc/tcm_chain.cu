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
// a launch so that 3 x B x ceil(T/128) tiles fill the 2 x 148 resident CTAs.  All GEMMs are 3-pass fp16 splits
// (hi*hi + lo*hi + hi*lo, fp32 accumulation in TMEM): fp32-grade, like the layer-by-layer path.
// Everything that crosses CTAs inside the launch (x, y, z, statistics) is read with ld.global.cg (L2): L1 is not
// coherent across SMs.
#include "common.cuh"
#include "umma.cuh"

namespace eab {

namespace {

using namespace umma;

constexpr int TM = 128;
constexpr int NT = 256;
constexpr int SLAB = TM * 128;                 // one 64-channel fp16 K slab of an A operand (16 KB)
constexpr int A_BYTES = 2 * SLAB;              // hi | lo
constexpr int W_BYTES = 64 * 1024;             // phase A: 4 slabs x (hi 8 KB | lo 8 KB); B: taps; C: 256 rows x (hi | lo)
constexpr int W_HALF = W_BYTES / 2;
constexpr int SMEM_BYTES = A_BYTES + W_BYTES + (3 * 64 + 2 * 64) * 4 + 64 + 1024;

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
            __nanosleep(64);
            if (++spins > (1u << 26)) __trap();          // a protocol bug must fault, never hang the GPU
        }
        __threadfence();
    }
    __syncthreads();
}

// column sums over the 32 lanes of a warp: lane l ends with the total of v[l] (31 shuffles instead of 160)
__device__ __forceinline__ float warp_column_sums(float (&v)[32], int lane) {
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
        const bool up = (lane & s) != 0;
#pragma unroll
        for (int i = 0; i < s; ++i) {
            const float mine = up ? v[i + s] : v[i];
            const float other = up ? v[i] : v[i + s];
            v[i] = mine + __shfl_xor_sync(0xffffffffu, other, s);
        }
    }
    return v[0];
}

__device__ __forceinline__ float4 ldcg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }

// 32 fp32 values of one operand row -> fp16 hi / lo, written as four 16-byte chunks of the 128-byte swizzled row
__device__ __forceinline__ void store_operand(uint8_t* A, int row, int half, const float (&v)[32]) {
    uint8_t* hi_row = A + row * 128;
    uint8_t* lo_row = hi_row + SLAB;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float* p = v + k * 8;
        uint4 hi, lo;
        hi.x = pack_h2(p[0], p[1]); hi.y = pack_h2(p[2], p[3]); hi.z = pack_h2(p[4], p[5]); hi.w = pack_h2(p[6], p[7]);
        lo.x = pack_lo_h2(p[0], p[1], hi.x); lo.y = pack_lo_h2(p[2], p[3], hi.y);
        lo.z = pack_lo_h2(p[4], p[5], hi.z); lo.w = pack_lo_h2(p[6], p[7], hi.w);
        const int off = ((half * 4 + k) ^ (row & 7)) << 4;
        *reinterpret_cast<uint4*>(hi_row + off) = hi;
        *reinterpret_cast<uint4*>(lo_row + off) = lo;
    }
}

__global__ void __launch_bounds__(NT, 2) tcm_chain_kernel(const TcmChainArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* As = smem;
    uint8_t* Ws = smem + A_BYTES;
    float* coef = reinterpret_cast<float*>(Ws + W_BYTES);        // [3][64] scale, shift, PReLU slope of the consumer's transform
    float* sstat = coef + 3 * 64;                                // [2][64] per-tile column sums
    uint64_t* bar = reinterpret_cast<uint64_t*>(sstat + 2 * 64);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) { mbar_init(bar, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(tmem_slot, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int tiles_per_chain = a.B * a.tiles_per_b;
    const int total = a.nchains * tiles_per_chain;
    const int tile_begin = (int)((long long)total * blockIdx.x / gridDim.x);
    const int tile_end = (int)((long long)total * (blockIdx.x + 1) / gridDim.x);
    const bool in_stats = a.instance_norm != 0;

    uint32_t par = 0;                  // parity of the next completion of `bar`
    unsigned bar_target = 0;
    // loader role: 2 threads per operand row; epilogue role: thread = (TMEM lane = row, column half)
    const int lrow = tid >> 1, lhalf = tid & 1;
    const int quad = warp & 3, chalf = warp >> 2;
    const int erow = quad * 32 + lane;
    const uint32_t idesc64 = make_idesc(64), idesc256 = make_idesc(256);

    // one K slab (already in As) against weight slab `wslab` (hi at Ws + wslab*stride, lo at + W_HALF): 3 passes x 4 K steps
    auto issue = [&](uint32_t idesc, int w_off, bool first) {
        fence_proxy_async();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            const uint32_t ah = smem_u32(As), al = ah + SLAB;
            const uint32_t wh = smem_u32(Ws) + (uint32_t)w_off, wl = wh + W_HALF;
#pragma unroll
            for (int pass = 0; pass < 3; ++pass)
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_f16(tmem_base, make_desc((pass == 1 ? al : ah) + k * 32), make_desc((pass == 2 ? wl : wh) + k * 32), idesc,
                             (!first || pass || k) ? 1u : 0u);
            umma_commit(bar);
        }
    };
    auto wait_mma = [&]() { mbar_wait(bar, par); par ^= 1u; };
    // copy `bytes` of a weight image into the hi half (lo = false) or lo half of Ws at byte offset `dst_off`
    auto load_w = [&](const float* src, int dst_off, int bytes) {
        const uint4* s = reinterpret_cast<const uint4*>(src);
        uint4* d = reinterpret_cast<uint4*>(Ws + dst_off);
        for (int i = tid; i < bytes / 16; i += NT) d[i] = __ldg(s + i);
    };
    // per-channel transform of the consumer: PReLU(alpha) then x*s + h, (s, h) from instance statistics or precomputed
    auto load_coef = [&](const double* stats, int b, unsigned off_sc, unsigned off_sh, unsigned off_al) {
        if (tid < 64) {
            float s, h;
            if (in_stats) {
                const double* st = stats + ((size_t)b * 64 + tid) * 2;
                const double mean = __ldcg(st) * (double)a.inv_count;
                double var = __ldcg(st + 1) * (double)a.inv_count - mean * mean;
                if (var < 0.0) var = 0.0;
                const double rstd = rsqrt(var + 1e-5);
                const double g = (double)__ldg(a.blob + off_sc + tid);
                s = (float)(g * rstd);
                h = (float)((double)__ldg(a.blob + off_sh + tid) - mean * g * rstd);
            } else {
                s = __ldg(a.blob + off_sc + tid);
                h = __ldg(a.blob + off_sh + tid);
            }
            coef[tid] = s; coef[64 + tid] = h; coef[128 + tid] = __ldg(a.blob + off_al + tid);
        }
    };
    // epilogue of phases A / B: 32 accumulator columns of this thread's row -> global rows + statistics of PReLU(value)
    auto epilogue64 = [&](float* dst, int b, int t0, double* stats, unsigned off_alpha) {
        const int t = t0 + erow;
        const bool valid = t < a.T;
        float v[32];
        {
            uint32_t rv[4][8];
            const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(chalf * 32);
#pragma unroll
            for (int k = 0; k < 4; ++k) tmem_ld8_nowait(taddr + k * 8, rv[k]);
            tmem_wait_ld();
#pragma unroll
            for (int k = 0; k < 4; ++k)
#pragma unroll
                for (int e = 0; e < 8; ++e) v[k * 8 + e] = __uint_as_float(rv[k][e]);
        }
        if (valid) {
            float* p = dst + ((size_t)b * a.T + t) * 64 + chalf * 32;
#pragma unroll
            for (int k = 0; k < 4; ++k) st_global_256(p + k * 8, *reinterpret_cast<const float(*)[8]>(&v[k * 8]));
        }
        if (in_stats) {
            float w[32];
            const float* al = a.blob + off_alpha + chalf * 32;
#pragma unroll
            for (int i = 0; i < 32; ++i) { v[i] = valid ? prelu_f(v[i], __ldg(al + i)) : 0.f; w[i] = v[i]; }
            const float s1 = warp_column_sums(w, lane);
#pragma unroll
            for (int i = 0; i < 32; ++i) w[i] = v[i] * v[i];
            const float s2 = warp_column_sums(w, lane);
            atomicAdd(&sstat[chalf * 32 + lane], s1);
            atomicAdd(&sstat[64 + chalf * 32 + lane], s2);
        }
        tc_fence_before();
        __syncthreads();
        if (in_stats && tid < 128) {
            const int c = tid & 63, which = tid >> 6;
            atomicAdd(stats + ((size_t)b * 64 + c) * 2 + which, (double)sstat[which * 64 + c]);
        }
    };

    for (int l = 0; l < a.nlayers; ++l) {
        // ======================================================================= phase A: y = W_in x
        int loaded = -1;
        for (int ti = tile_begin; ti < tile_end; ++ti) {
            const int chain = ti / tiles_per_chain, rem = ti - chain * tiles_per_chain;
            const int b = rem / a.tiles_per_b, t0 = (rem - b * a.tiles_per_b) * TM;
            const TcmChainLayer& L = a.L[chain * a.nlayers + l];
            if (loaded != chain) {
                load_w(a.blob + L.win_hi, 0, 4 * 64 * 128);
                load_w(a.blob + L.win_lo, W_HALF, 4 * 64 * 128);
                loaded = chain;
            }
            if (tid < 128) sstat[tid] = 0.f;
            const float* xsrc = (l == 0 ? a.x_in[chain] : a.x_buf[chain]);
            const int t = t0 + lrow;
            const float* xrow = xsrc + ((size_t)b * a.T + t) * 256 + lhalf * 32;
            float v[32];
            auto load_x = [&](int s) {
                if (t < a.T) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) *reinterpret_cast<float4*>(&v[k * 4]) = ldcg4(xrow + s * 64 + k * 4);
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = 0.f;
                }
            };
            load_x(0);
            for (int s = 0; s < 4; ++s) {
                if (s > 0) wait_mma();                     // the previous slab's MMAs have read As
                store_operand(As, lrow, lhalf, v);
                if (s < 3) load_x(s + 1);                  // in flight while this slab's MMAs run
                issue(idesc64, s * 64 * 128, s == 0);
            }
            wait_mma();
            tc_fence_after();
            epilogue64(a.y[chain], b, t0, a.stats + L.st_d, L.al_d);
        }
        grid_barrier(a.barrier, bar_target);
        // ======================================================================= phase B: z = W_dil * norm(PReLU(y)) (dilated)
        loaded = -1;
        for (int ti = tile_begin; ti < tile_end; ++ti) {
            const int chain = ti / tiles_per_chain, rem = ti - chain * tiles_per_chain;
            const int b = rem / a.tiles_per_b, t0 = (rem - b * a.tiles_per_b) * TM;
            const TcmChainLayer& L = a.L[chain * a.nlayers + l];
            if (loaded != chain) {
                load_w(a.blob + L.wd_hi, 0, a.kd * 64 * 128);
                load_w(a.blob + L.wd_lo, W_HALF, a.kd * 64 * 128);
                loaded = chain;
            }
            if (tid < 128) sstat[tid] = 0.f;
            load_coef(a.stats + L.st_d, b, L.sc_d, L.sh_d, L.al_d);
            __syncthreads();
            const int t = t0 + lrow;
            float v[32];
            bool ok = false;
            auto load_y = [&](int k) {
                const int ts = t - L.dt[k];
                ok = t < a.T && ts >= 0 && ts < a.T;
                if (ok) {
                    const float* yrow = a.y[chain] + ((size_t)b * a.T + ts) * 64 + lhalf * 32;
#pragma unroll
                    for (int q = 0; q < 8; ++q) *reinterpret_cast<float4*>(&v[q * 4]) = ldcg4(yrow + q * 4);
                }
            };
            load_y(0);
            for (int k = 0; k < a.kd; ++k) {
                if (ok) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const int c = lhalf * 32 + i;
                        v[i] = fmaf(prelu_f(v[i], coef[128 + c]), coef[c], coef[64 + c]);
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = 0.f;        // literal zeros AFTER the norm (GaGNet.py:313)
                }
                if (k > 0) wait_mma();
                store_operand(As, lrow, lhalf, v);
                if (k + 1 < a.kd) load_y(k + 1);           // in flight while this tap's MMAs run
                issue(idesc64, k * 64 * 128, k == 0);
            }
            wait_mma();
            tc_fence_after();
            epilogue64(a.z[chain], b, t0, a.stats + L.st_o, L.al_o);
        }
        grid_barrier(a.barrier, bar_target);
        // ======================================================================= phase C: x += W_out * norm(PReLU(z))
        loaded = -1;
        for (int ti = tile_begin; ti < tile_end; ++ti) {
            const int chain = ti / tiles_per_chain, rem = ti - chain * tiles_per_chain;
            const int b = rem / a.tiles_per_b, t0 = (rem - b * a.tiles_per_b) * TM;
            const TcmChainLayer& L = a.L[chain * a.nlayers + l];
            if (loaded != chain) {
                load_w(a.blob + L.wo_hi[0], 0, 128 * 128);
                load_w(a.blob + L.wo_hi[1], 128 * 128, 128 * 128);
                load_w(a.blob + L.wo_lo[0], W_HALF, 128 * 128);
                load_w(a.blob + L.wo_lo[1], W_HALF + 128 * 128, 128 * 128);
                loaded = chain;
            }
            load_coef(a.stats + L.st_o, b, L.sc_o, L.sh_o, L.al_o);
            __syncthreads();
            {
                const int t = t0 + lrow;
                float v[32];
                if (t < a.T) {
                    const float* zrow = a.z[chain] + ((size_t)b * a.T + t) * 64 + lhalf * 32;
#pragma unroll
                    for (int q = 0; q < 8; ++q) *reinterpret_cast<float4*>(&v[q * 4]) = ldcg4(zrow + q * 4);
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const int c = lhalf * 32 + i;
                        v[i] = fmaf(prelu_f(v[i], coef[128 + c]), coef[c], coef[64 + c]);
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = 0.f;
                }
                store_operand(As, lrow, lhalf, v);
                issue(idesc256, 0, true);
            }
            wait_mma();
            tc_fence_after();
            {
                const int t = t0 + erow;
                const bool valid = t < a.T;
                const float* xsrc = (l == 0 ? a.x_in[chain] : a.x_buf[chain]) + ((size_t)b * a.T + t) * 256 + chalf * 128;
                float* xdst = a.x_buf[chain] + ((size_t)b * a.T + t) * 256 + chalf * 128;
                const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(chalf * 128);
                float4 xr[2][4];                   // residual of the next 16 columns in flight while this group is added / stored
                auto load_res = [&](int g, float4 (&dst)[4]) {
                    if (valid) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) dst[k] = ldcg4(xsrc + g * 16 + k * 4);
                    }
                };
                load_res(0, xr[0]);
#pragma unroll
                for (int g = 0; g < 8; ++g) {
                    uint32_t rv[2][8];
#pragma unroll
                    for (int k = 0; k < 2; ++k) tmem_ld8_nowait(taddr + g * 16 + k * 8, rv[k]);
                    if (g < 7) load_res(g + 1, xr[(g + 1) & 1]);
                    tmem_wait_ld();
                    if (valid) {
#pragma unroll
                        for (int k = 0; k < 2; ++k) {
                            const float4 x0 = xr[g & 1][2 * k], x1 = xr[g & 1][2 * k + 1];
                            float o[8];
                            o[0] = __uint_as_float(rv[k][0]) + x0.x; o[1] = __uint_as_float(rv[k][1]) + x0.y;
                            o[2] = __uint_as_float(rv[k][2]) + x0.z; o[3] = __uint_as_float(rv[k][3]) + x0.w;
                            o[4] = __uint_as_float(rv[k][4]) + x1.x; o[5] = __uint_as_float(rv[k][5]) + x1.y;
                            o[6] = __uint_as_float(rv[k][6]) + x1.z; o[7] = __uint_as_float(rv[k][7]) + x1.w;
                            st_global_256(xdst + g * 16 + k * 8, o);
                        }
                    }
                }
            }
            tc_fence_before();
            __syncthreads();
        }
        // no grid barrier: the next layer's phase A reads only this CTA's own rows of x_buf (ld.cg, after the block barrier)
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 256);
}

}  // namespace

bool tcm_chain_supported(const TcmChainArgs& a) {
    return a.nchains >= 1 && a.nchains <= 3 && a.nlayers >= 1 && a.nchains * a.nlayers <= kMaxChainLayers && a.kd >= 1 &&
           a.kd * 64 * 128 <= W_HALF && a.B >= 1 && a.T >= 1;
}

int launch_tcm_chain(const TcmChainArgs& a_in, cudaStream_t st) {
    TcmChainArgs a = a_in;
    if (!tcm_chain_supported(a)) return fail("tcm_chain: unsupported shape");
    static int max_ctas = 0;
    if (!max_ctas) {
        EAB_CUDA(cudaFuncSetAttribute(tcm_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        EAB_CUDA(cudaFuncSetAttribute(tcm_chain_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        int dev = 0, sms = 0, per_sm = 0, coop = 0;
        EAB_CUDA(cudaGetDevice(&dev));
        EAB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        EAB_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
        if (!coop) return fail("tcm_chain: the device does not support cooperative launches");
        EAB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tcm_chain_kernel, NT, SMEM_BYTES));
        if (per_sm < 1) return fail("tcm_chain: kernel does not fit an SM");
        if (per_sm > 2) per_sm = 2;                      // 256 TMEM columns per CTA
        max_ctas = per_sm * sms;
    }
    a.tiles_per_b = (a.T + TM - 1) / TM;
    const long long total = (long long)a.nchains * a.B * a.tiles_per_b;
    if (total >= (1ll << 30)) return fail("tcm_chain: too many tiles");
    const int grid = (int)(total < max_ctas ? total : max_ctas);
    const double rows = (double)a.nchains * a.B * a.T * a.nlayers;
    ProfScope ps("tcm_chain", 2.0 * rows * (256.0 * 64 + a.kd * 64.0 * 64 + 64.0 * 256),
                 4.0 * rows * (256 + 64 + 64 * a.kd + 64 + 64 + 256 + 256), st);
    void* params[1] = {&a};
    EAB_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(tcm_chain_kernel), dim3(grid), dim3(NT), params,
                                         (size_t)SMEM_BYTES, st));
    EAB_LAUNCH_CHECK("tcm_chain_kernel");
    return 0;
}

}  // namespace eab
