// Two-kernel form of the tensor-core convolution (sm_100a):
//
//   stage_kernel     : bandwidth-bound.  Reads the raw fp32 activation(s) once, applies the producer layer's norm + PReLU
//                      (Xform), converts to fp16 (hi, and lo = fp16(x - hi) for the 3-pass layers) and writes "plane
//                      images" to HBM: rows of 64 channels = 128 bytes in the padded-pitch row space of the consumer
//                      GEMM (row = front + t*P + col; zero pad rows/columns written explicitly), already in the
//                      128-byte-swizzled order the tensor core reads.
//   conv_tma_kernel  : the GEMM.  Because a tile's plane rows are now one contiguous, pre-swizzled byte range, the
//                      whole A operand of a tile is fetched by ONE elected thread with cp.async.bulk (a few copies of
//                      ~30 KB) straight into shared memory - no register staging, no conversion, no proxy fence, the
//                      copy engine hides the latency.  Taps are row-shifted UMMA descriptors into those planes (see
//                      conv_plane.cu), weights stream through a small ring, accumulators are double-buffered in TMEM, and
//                      8 epilogue warps do bias / gate / ReLU / residual / statistics / bulk row stores.
//
// conv_plane.cu (fused gather + transform producers) measured ~11k cycles per 128-row tile in the producer warps against
// ~1k cycles of MMA; splitting the elementwise work out trades ~40 % more activation bytes for a conv kernel whose
// critical path is the epilogue / MMA only.
#include <cuda_fp16.h>

#include "common.cuh"
#include "umma.cuh"

namespace eab {

namespace {

using namespace umma;

constexpr int TM = 128;
constexpr int KC = 64;
constexpr int NSB = 3;                          // weight ring stages
constexpr int NEPI = 256;                       // 8 epilogue warps per group
constexpr int NGRP = 1;                         // epilogue groups (2 = alternate tiles between two groups: measured no gain, costs a plane buffer)
constexpr int NTHREADS = 128 + NGRP * NEPI;     // warp 0 plane copies, 1 MMA, 2 weight loader, 3 idle, 4-11 / 12-19 epilogue

inline int ceil8(int x) { return (x + 7) & ~7; }
constexpr int STAGE_RB = 8;                     // 32-row blocks per stage CTA

// ------------------------------------------------------------------------------------------------ stage kernel
// grid (row blocks, nplanes*nslab, B), 256 threads: 8 threads per plane row (8 channels = one 16-byte fp16 chunk each)
__global__ void __launch_bounds__(256) stage_kernel(const PlaneConvArgs a) {
    __shared__ float coef[6 * 64];
    pdl_trigger();
    pdl_wait();
    const int b = blockIdx.z;
    const int ps = blockIdx.y;
    const int plane = ps / a.nslab;
    const int slab = ps - plane * a.nslab;
    const int C0 = a.src[0].C;
    const bool second = slab * KC >= C0;
    const ConvSrc& src = second ? a.src[1] : a.src[0];
    const int cbase = second ? slab * KC - C0 : slab * KC;
    if (threadIdx.x < 64) {
        float cs, ch, ca;
        xform_coeffs(src.xf, b, src.C, cbase + threadIdx.x, cs, ch, ca);
        coef[threadIdx.x] = cs;
        coef[64 + threadIdx.x] = ch;
        coef[128 + threadIdx.x] = src.xf.prelu ? ca : 1.f;
        if (src.x2) {
            xform_coeffs(src.xf2, b, src.C, cbase + threadIdx.x, cs, ch, ca);
            coef[192 + threadIdx.x] = cs;
            coef[256 + threadIdx.x] = ch;
            coef[320 + threadIdx.x] = src.xf2.prelu ? ca : 1.f;
        }
    }
    __syncthreads();
    const int mode = (src.xf.affine == 0 && src.xf.prelu == 0) ? 0 : (src.xf.prelu == 1 ? 2 : 1);
    const int mode2 = (src.xf2.affine == 0 && src.xf2.prelu == 0) ? 0 : (src.xf2.prelu == 1 ? 2 : 1);
    const int npb = a.npass == 3 ? 2 : 1;
    const int c8 = threadIdx.x & 7;
    uint8_t* dst_hi = static_cast<uint8_t*>(const_cast<void*>(a.np[ps * npb]));
    uint8_t* dst_lo = npb == 2 ? static_cast<uint8_t*>(const_cast<void*>(a.np[ps * npb + 1])) : nullptr;
    const float* cs1 = coef + c8 * 8;
    const float* cs2 = coef + 192 + c8 * 8;
    // STAGE_RB row blocks of 32 rows per CTA: the (double precision) coefficient set-up above is paid once per 256 rows
#pragma unroll 2
    for (int it = 0; it < STAGE_RB; ++it) {
        const int rho = (blockIdx.x * STAGE_RB + it) * 32 + (threadIdx.x >> 3);
        if (rho >= a.np_rows) break;
        const int r = rho - a.np_front;
        float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
        if (r >= 0 && r < a.T * a.P) {
            const int t = a.P == 1 ? r : (int)__umulhi((unsigned)r, a.p_magic);
            const int col = r - t * a.P;
            if (col < a.plane_cols[plane]) {
                const int fi = col * a.col_stride + a.col_off[plane];
                const size_t eoff = (((size_t)b * a.T + t) * a.Fin + fi) * src.C + cbase + c8 * 8;
                const float4* p = reinterpret_cast<const float4*>(src.x + eoff);
                v0 = __ldg(p);
                v1 = __ldg(p + 1);
                float4 w0, w1;
                if (src.x2) {
                    const float4* p2 = reinterpret_cast<const float4*>(src.x2 + eoff);
                    w0 = __ldg(p2); w1 = __ldg(p2 + 1);
                }
                if (mode != 0) {
                    float x[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        if (mode == 1) { const float z = fmaf(x[i], cs1[i], cs1[64 + i]); x[i] = fmaxf(z, 0.f) + cs1[128 + i] * fminf(z, 0.f); }
                        else x[i] = fmaf(fmaxf(x[i], 0.f) + cs1[128 + i] * fminf(x[i], 0.f), cs1[i], cs1[64 + i]);
                    }
                    v0 = make_float4(x[0], x[1], x[2], x[3]);
                    v1 = make_float4(x[4], x[5], x[6], x[7]);
                }
                if (src.x2) {                        // + the second addend of a module's residual sum
                    float x[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
                    if (mode2 != 0) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            if (mode2 == 1) { const float z = fmaf(x[i], cs2[i], cs2[64 + i]); x[i] = fmaxf(z, 0.f) + cs2[128 + i] * fminf(z, 0.f); }
                            else x[i] = fmaf(fmaxf(x[i], 0.f) + cs2[128 + i] * fminf(x[i], 0.f), cs2[i], cs2[64 + i]);
                        }
                    }
                    v0.x += x[0]; v0.y += x[1]; v0.z += x[2]; v0.w += x[3];
                    v1.x += x[4]; v1.y += x[5]; v1.z += x[6]; v1.w += x[7];
                }
            }
        }
        uint4 hi;
        hi.x = pack_h2(v0.x, v0.y); hi.y = pack_h2(v0.z, v0.w); hi.z = pack_h2(v1.x, v1.y); hi.w = pack_h2(v1.z, v1.w);
        const size_t row_g = (size_t)b * a.np_rows + rho;
        const size_t off = row_g * 128 + (size_t)((c8 ^ (rho & 7)) << 4);
        *reinterpret_cast<uint4*>(dst_hi + off) = hi;
        if (npb == 2) {
            uint4 lo;
            lo.x = pack_lo_h2(v0.x, v0.y, hi.x); lo.y = pack_lo_h2(v0.z, v0.w, hi.y);
            lo.z = pack_lo_h2(v1.x, v1.y, hi.z); lo.w = pack_lo_h2(v1.z, v1.w, hi.w);
            *reinterpret_cast<uint4*>(dst_lo + off) = lo;
        }
    }
}

// ------------------------------------------------------------------------------------------------ conv kernel
struct Plan {
    int R;                           // plane rows a tile needs = 128 + back + fwd
    int plane_bytes;                 // (R + 8 rounded up to 8) * 128 : the copy starts on an 8-row (1024 B) boundary
    int buf_bytes;                   // nplanes * nslab * npb * plane_bytes
    int b_stage_bytes, stg_ld;
    int b_off, stg_off, rowoff_off, bias_off, bar_off, total;
};

__host__ __device__ inline Plan make_plan(const PlaneConvArgs& a) {
    Plan p;
    p.R = TM + a.back + a.fwd;
    p.plane_bytes = ((p.R + 7 + 7) & ~7) * 128;
    const int npb = a.npass == 3 ? 2 : 1;
    p.buf_bytes = a.nplanes * a.nslab * npb * p.plane_bytes;
    p.b_stage_bytes = a.N * 128;
    p.stg_ld = a.Cout + 4;
    p.b_off = a.nbuf * p.buf_bytes;
    p.stg_off = p.b_off + NSB * p.b_stage_bytes;
    p.rowoff_off = p.stg_off + NGRP * TM * p.stg_ld * 4;
    p.bias_off = p.rowoff_off + NGRP * TM * 8;
    p.bar_off = p.bias_off + a.N * 4;
    p.total = p.bar_off + 256 + 1024;
    return p;
}

__global__ void __launch_bounds__(NTHREADS, 1) conv_tma_kernel(const PlaneConvArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const Plan pl = make_plan(a);
    uint8_t* planes = smem;
    uint8_t* Bs = smem + pl.b_off;
    float* stg = reinterpret_cast<float*>(smem + pl.stg_off);
    long long* rowoff = reinterpret_cast<long long*>(smem + pl.rowoff_off);
    float* sbias = reinterpret_cast<float*>(smem + pl.bias_off);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + pl.bar_off);
    uint64_t* plane_full = bars;            // [3]
    uint64_t* plane_empty = bars + 3;       // [3]
    uint64_t* b_full = bars + 6;            // [NSB]
    uint64_t* b_empty = bars + 9;           // [NSB]
    uint64_t* acc_full = bars + 12;         // [2]
    uint64_t* acc_empty = bars + 14;        // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;
    const uint32_t tmem_cols = a.N <= 64 ? 128u : (a.N <= 128 ? 256u : 512u);
    const int npb = a.npass == 3 ? 2 : 1;
    const int nimg = a.nplanes * a.nslab * npb;

    if (tid == 0) {
        for (int i = 0; i < 3; ++i) { mbar_init(&plane_full[i], 1); mbar_init(&plane_empty[i], 1); }
        for (int i = 0; i < NSB; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], NEPI); }
        fence_barrier_init();
    }
    pdl_trigger();
    if (warp == 1) tmem_alloc(tmem_slot, tmem_cols);
    for (int i = tid; i < a.N; i += NTHREADS) sbias[i] = a.bias ? __ldg(a.bias + i) : 0.f;
    pdl_wait();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const long long ntiles = (long long)a.B * a.tiles_per_b;
    const int tile_begin = (int)(ntiles * blockIdx.x / gridDim.x);
    const int tile_end = (int)(ntiles * (blockIdx.x + 1) / gridDim.x);
    const int rows_per_b = a.T * a.P;
    const int units_per_tile = a.ntaps * a.nslab * a.npass;

    if (warp == 0) {
        // =========================================================================== plane copies (one lane)
        int buf = 0;
        uint32_t bphase = 0;
        for (int tile = tile_begin; tile < tile_end; ++tile) {
            const int b = tile / a.tiles_per_b;
            const int row0 = (tile - b * a.tiles_per_b) * TM;
            const int rs = a.np_front + row0 - a.back;          // first plane row the tile needs (>= 0)
            const int ra = rs & ~7;                              // copy from the enclosing 8-row boundary: swizzle phases match
            const uint32_t bytes = (uint32_t)(((rs - ra) + pl.R + 7) & ~7) * 128u;
            mbar_wait(&plane_empty[buf], bphase ^ 1);
            if (lane == 0) {
                mbar_arrive_expect_tx(&plane_full[buf], bytes * (uint32_t)nimg);
                for (int i = 0; i < nimg; ++i) {
                    const uint8_t* src = static_cast<const uint8_t*>(a.np[i]) + ((size_t)b * a.np_rows + ra) * 128;
                    bulk_copy_g2s(planes + buf * pl.buf_bytes + i * pl.plane_bytes, src, bytes, &plane_full[buf]);
                }
            }
            __syncwarp();
            if (++buf == a.nbuf) { buf = 0; bphase ^= 1; }
        }
    } else if (warp == 1) {
        // =========================================================================== MMA issuer
        const uint32_t idesc = make_idesc(a.N);
        int buf = 0, stage = 0, acc = 0;
        uint32_t bphase = 0, sphase = 0, aphase = 0;
        const bool dbg_on = a.dbg != nullptr && blockIdx.x == 0 && lane == 0;
        long long t_wacc = 0, t_wplane = 0, t_wb = 0;
        const long long t_start = dbg_on ? clock64() : 0;
        for (int tile = tile_begin; tile < tile_end; ++tile) {
            const int b = tile / a.tiles_per_b;
            const int row0 = (tile - b * a.tiles_per_b) * TM;
            const int rs = a.np_front + row0 - a.back;
            const int lead = rs & 7;                             // rows between the copy start and the tile's first plane row
            const long long w0 = dbg_on ? clock64() : 0;
            mbar_wait(&acc_empty[acc], aphase ^ 1);
            const long long w1 = dbg_on ? clock64() : 0;
            mbar_wait(&plane_full[buf], bphase);
            if (dbg_on) { t_wacc += w1 - w0; t_wplane += clock64() - w1; }
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)(acc * a.N);
            const uint32_t pbase = smem_u32(planes + buf * pl.buf_bytes);
            int unit = 0;
            for (int tap = 0; tap < a.ntaps; ++tap) {
                const int arow = lead + a.back + a.tap_shift[tap];
                for (int slab = 0; slab < a.nslab; ++slab) {
                    const uint32_t a_hi = pbase + (uint32_t)(((a.tap_plane[tap] * a.nslab + slab) * npb) * pl.plane_bytes + arow * 128);
                    for (int pass = 0; pass < a.npass; ++pass, ++unit) {
                        const long long w2 = dbg_on ? clock64() : 0;
                        mbar_wait(&b_full[stage], sphase);
                        if (dbg_on) t_wb += clock64() - w2;
                        tc_fence_after();
                        if (lane == 0) {
                            const uint32_t a_addr = a_hi + (pass == 1 ? (uint32_t)pl.plane_bytes : 0u);
                            const uint32_t b_addr = smem_u32(Bs + stage * pl.b_stage_bytes);
#pragma unroll
                            for (int k = 0; k < KC / 16; ++k)
                                umma_f16(d_tmem, make_desc(a_addr + k * 32), make_desc(b_addr + k * 32), idesc, (unit | k) ? 1u : 0u);
                            umma_commit(&b_empty[stage]);
                            if (unit == units_per_tile - 1) {
                                umma_commit(&acc_full[acc]);
                                umma_commit(&plane_empty[buf]);
                            }
                        }
                        __syncwarp();
                        if (++stage == NSB) { stage = 0; sphase ^= 1; }
                    }
                }
            }
            if (++acc == 2) { acc = 0; aphase ^= 1; }
            if (++buf == a.nbuf) { buf = 0; bphase ^= 1; }
        }
        if (dbg_on) { a.dbg[4] = clock64() - t_start; a.dbg[5] = t_wacc; a.dbg[6] = t_wplane; a.dbg[7] = t_wb; a.dbg[2] = tile_end - tile_begin; a.dbg[0] = clock64() - t_start; }
    } else if (warp == 2) {
        // =========================================================================== B (weight) loader
        int stage = 0;
        uint32_t sphase = 0;
        const uint32_t bytes = (uint32_t)pl.b_stage_bytes;
        for (int tile = tile_begin; tile < tile_end; ++tile) {
            for (int ts = 0; ts < a.ntaps * a.nslab; ++ts) {
                for (int pass = 0; pass < a.npass; ++pass) {
                    mbar_wait(&b_empty[stage], sphase ^ 1);
                    if (lane == 0) {
                        const float* img = (pass == 2 ? a.Wlo : a.Whi) + (size_t)ts * a.N * 32;
                        mbar_arrive_expect_tx(&b_full[stage], bytes);
                        bulk_copy_g2s(Bs + stage * pl.b_stage_bytes, img, bytes, &b_full[stage]);
                    }
                    __syncwarp();
                    if (++stage == NSB) { stage = 0; sphase ^= 1; }
                }
            }
        }
    } else if (warp >= 4) {
        // =========================================================================== epilogue (8 warps)
        const int grp = (warp - 4) >> 3;        // epilogue group = TMEM accumulator it drains
        const int et = (tid - 128) & (NEPI - 1);        // 0..255 within the group
        stg += grp * TM * pl.stg_ld;
        rowoff += grp * TM;
        const int quad = warp & 3;              // TMEM lane quadrant (warp % 4)
        const int chalf = ((warp - 4) >> 2) & 1;        // which half of the output channels this warp converts
        const int row = quad * 32 + lane;
        const bool gated = a.gate_off > 0;
        const int ld = pl.stg_ld;
        const int cper = a.Cout >> 1;           // channels per half (multiple of 8)
        int acc = grp;                          // NGRP == 2: fixed per group; NGRP == 1: alternates
        uint32_t aphase = 0;
        const bool dbg_on = a.dbg != nullptr && blockIdx.x == 0 && et == 0 && grp == 0;
        long long t_wfull = 0, t_tmem = 0, t_store = 0, t_stats = 0;
        const long long t_start = dbg_on ? clock64() : 0;
        for (int tile = tile_begin + grp; tile < tile_end; tile += NGRP) {
            const int b = tile / a.tiles_per_b;
            const int row0 = (tile - b * a.tiles_per_b) * TM;
            {
                const int r = row0 + row;
                long long off = -1;
                if (r < rows_per_b) {
                    const int t = a.P == 1 ? r : (int)__umulhi((unsigned)r, a.p_magic);
                    const int e = r - t * a.P;
                    if (e < a.E) off = ((((long long)b * a.T + t) * a.Fout) + (e * a.out_stride + a.out_off)) * a.out_ld + a.out_coff;
                }
                if (chalf == 0) rowoff[row] = off;
            }
            const long long e0 = dbg_on ? clock64() : 0;
            mbar_wait(&acc_full[acc], aphase);
            const long long e1 = dbg_on ? clock64() : 0;
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * a.N);
            // ---- phase 1: TMEM -> bias / gate / ReLU -> fp32 staging tile (thread = row; 16-byte stores, conflict-free)
            for (int c0 = chalf * cper; c0 < (chalf + 1) * cper; c0 += 8) {
                uint32_t rv[8], rg[8];
                tmem_ld8_nowait(taddr + c0, rv);
                if (gated) tmem_ld8_nowait(taddr + a.gate_off + c0, rg);
                tmem_wait_ld();
                float v[8];
                const float4 b0 = *reinterpret_cast<const float4*>(sbias + c0);
                const float4 b1 = *reinterpret_cast<const float4*>(sbias + c0 + 4);
                v[0] = __uint_as_float(rv[0]) + b0.x; v[1] = __uint_as_float(rv[1]) + b0.y; v[2] = __uint_as_float(rv[2]) + b0.z; v[3] = __uint_as_float(rv[3]) + b0.w;
                v[4] = __uint_as_float(rv[4]) + b1.x; v[5] = __uint_as_float(rv[5]) + b1.y; v[6] = __uint_as_float(rv[6]) + b1.z; v[7] = __uint_as_float(rv[7]) + b1.w;
                if (gated) {
                    const float4 g0 = *reinterpret_cast<const float4*>(sbias + a.gate_off + c0);
                    const float4 g1 = *reinterpret_cast<const float4*>(sbias + a.gate_off + c0 + 4);
                    v[0] *= sigmoid_f(__uint_as_float(rg[0]) + g0.x); v[1] *= sigmoid_f(__uint_as_float(rg[1]) + g0.y);
                    v[2] *= sigmoid_f(__uint_as_float(rg[2]) + g0.z); v[3] *= sigmoid_f(__uint_as_float(rg[3]) + g0.w);
                    v[4] *= sigmoid_f(__uint_as_float(rg[4]) + g1.x); v[5] *= sigmoid_f(__uint_as_float(rg[5]) + g1.y);
                    v[6] *= sigmoid_f(__uint_as_float(rg[6]) + g1.z); v[7] *= sigmoid_f(__uint_as_float(rg[7]) + g1.w);
                }
                if (a.relu) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) v[i] = fmaxf(v[i], 0.f);
                }
                *reinterpret_cast<float4*>(stg + row * ld + c0) = make_float4(v[0], v[1], v[2], v[3]);
                *reinterpret_cast<float4*>(stg + row * ld + c0 + 4) = make_float4(v[4], v[5], v[6], v[7]);
            }
            tc_fence_before();
            mbar_arrive(&acc_empty[acc]);
            named_bar_sync(2 + grp, NEPI);            // staging tile + row offsets complete
            const long long e2 = dbg_on ? clock64() : 0;
            // ---- phase 2: coalesced copy-out (a warp writes 512 contiguous bytes per instruction), residual add, and the
            //      per-channel statistics of exactly what was stored.  16-byte chunk g = i*NEPI + et: row g / tpr, and the
            //      channel quad cq = 4 (g % tpr) is the SAME for every i when tpr divides NEPI (Cout 16/32/64/128).
            const int tpr = a.Cout >> 2;        // 16-byte chunks per row
            const int cq = (et % tpr) * 4;
            const int rows_per_it = NEPI / tpr;
            float ssum[2][4], ssq[2][4];
#pragma unroll
            for (int s = 0; s < 2; ++s)
#pragma unroll
                for (int j = 0; j < 4; ++j) { ssum[s][j] = 0.f; ssq[s][j] = 0.f; }
            float al[2][4];
#pragma unroll
            for (int s = 0; s < 2; ++s)
#pragma unroll
                for (int j = 0; j < 4; ++j) al[s][j] = (s < a.nstats && a.stat_alpha[s]) ? __ldg(a.stat_alpha[s] + cq + j) : 1.f;
            for (int r = et / tpr; r < TM; r += rows_per_it) {
                const long long off = rowoff[r];
                if (off < 0) continue;          // dummy / ragged rows: not stored, not counted
                float4 o = *reinterpret_cast<const float4*>(stg + r * ld + cq);
                if (a.resid) {
                    const float4 q = __ldg(reinterpret_cast<const float4*>(a.resid + off + cq));
                    o.x += q.x; o.y += q.y; o.z += q.z; o.w += q.w;
                }
                *reinterpret_cast<float4*>(a.out + off + cq) = o;
                const float ov[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
                for (int s = 0; s < 2; ++s) {
                    if (s >= a.nstats) continue;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float u = fmaxf(ov[j], 0.f) + al[s][j] * fminf(ov[j], 0.f);
                        ssum[s][j] += u;
                        ssq[s][j] = fmaf(u, u, ssq[s][j]);
                    }
                }
            }
            const long long e3 = dbg_on ? clock64() : 0;
            if (a.nstats) {
                // threads with equal et % tpr hold partial sums of the same channels: fold across the warp (lane strides
                // of tpr), then across the 8 warps through the (now dead) staging tile, then one fp64 atomic per value
                named_bar_sync(2 + grp, NEPI);        // every thread has finished reading the staging tile
                for (int s = 0; s < a.nstats; ++s)
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        float x = ssum[s][j], y = ssq[s][j];
                        for (int o = tpr; o < 32; o <<= 1) {
                            x += __shfl_xor_sync(0xffffffffu, x, o);
                            y += __shfl_xor_sync(0xffffffffu, y, o);
                        }
                        ssum[s][j] = x; ssq[s][j] = y;
                    }
                const int wv = et >> 5;
                const int nlead = tpr < 32 ? tpr : 32;             // lanes holding distinct channel quads
                if (lane < nlead) {
                    for (int s = 0; s < a.nstats; ++s)
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            // red[warp][s][channel][2]
                            float* rp = stg + (((wv * 2 + s) * a.Cout) + cq + j) * 2;
                            rp[0] = ssum[s][j]; rp[1] = ssq[s][j];
                        }
                }
                named_bar_sync(2 + grp, NEPI);
                const int nsc = a.nstats * a.Cout;
                for (int i = et; i < nsc; i += NEPI) {
                    const int s = i / a.Cout, c = i - s * a.Cout;
                    float x = 0.f, y = 0.f;
#pragma unroll
                    for (int w = 0; w < NEPI / 32; ++w) { x += stg[(((w * 2 + s) * a.Cout) + c) * 2]; y += stg[(((w * 2 + s) * a.Cout) + c) * 2 + 1]; }
                    double* dstp = a.stats[s] + ((size_t)b * a.Cout + c) * 2;
                    atomicAdd(dstp, (double)x);
                    atomicAdd(dstp + 1, (double)y);
                }
            }
            named_bar_sync(2 + grp, NEPI);            // staging free for the next tile
            if (dbg_on) { const long long e4 = clock64(); t_wfull += e1 - e0; t_tmem += e2 - e1; t_store += e3 - e2; t_stats += e4 - e3; }
            if (NGRP == 2) aphase ^= 1;
            else if (++acc == 2) { acc = 0; aphase ^= 1; }
        }
        if (dbg_on) { a.dbg[8] = clock64() - t_start; a.dbg[9] = t_wfull; a.dbg[10] = t_tmem; a.dbg[11] = t_store; a.dbg[12] = t_stats; }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, tmem_cols);
}

int choose_nbuf(PlaneConvArgs& a) {
    for (int nb = 3; nb >= 1; --nb) {
        a.nbuf = nb;
        if (make_plan(a).total <= 227 * 1024) return nb;
    }
    return 0;
}

}  // namespace

int staged_rows(const PlaneConvArgs& a, int* front) {
    const int f = ceil8(a.back);
    if (front) *front = f;
    return ceil8(f + a.tiles_per_b * TM + a.fwd + 8);
}

bool staged_conv_supported(const PlaneConvArgs& a_in) {
    if (!plane_conv_supported(a_in)) return false;
    PlaneConvArgs a = a_in;
    if ((a.Cout != 16 && a.Cout != 32 && a.Cout != 64 && a.Cout != 128) || a.nplanes * a.nslab * (a.npass == 3 ? 2 : 1) > 16) return false;
    return choose_nbuf(a) > 0;
}

int launch_stage(const PlaneConvArgs& a_in, cudaStream_t st) {
    PlaneConvArgs a = a_in;
    a.p_magic = a.P == 1 ? 0u : (unsigned)((1ull << 32) / (unsigned)a.P) + 1u;
    dim3 grid((a.np_rows + 32 * STAGE_RB - 1) / (32 * STAGE_RB), a.nplanes * a.nslab, a.B);
    double cin = 0;
    for (int i = 0; i < a.nsrc; ++i) cin += a.src[i].C;
    const double elems = (double)a.B * a.T * a.Fin * cin;
    ProfScope ps("stage", 4.0 * elems, elems * 4.0 + (double)a.B * a.np_rows * 128.0 * a.nplanes * a.nslab * (a.npass == 3 ? 2 : 1), st);
    EAB_CUDA(launch_k(stage_kernel, grid, dim3(256), (size_t)0, st, a));
    EAB_LAUNCH_CHECK("stage_kernel");
    return 0;
}

int launch_conv_staged(PlaneConvArgs a, cudaStream_t st) {
    if (a.B <= 0 || a.T <= 0 || a.E <= 0) return 0;
    a.p_magic = a.P == 1 ? 0u : (unsigned)((1ull << 32) / (unsigned)a.P) + 1u;
    if (!choose_nbuf(a)) return fail("conv_staged: shared-memory budget");
    const Plan pl = make_plan(a);
    static int configured = 0;
    if (pl.total > configured) {
        EAB_CUDA(cudaFuncSetAttribute(conv_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, pl.total));
        configured = pl.total;
    }
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        EAB_CUDA(cudaGetDevice(&dev));
        EAB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    }
    const long long ntiles = (long long)a.B * a.tiles_per_b;
    if (ntiles >= (1ll << 30)) return fail("conv_staged: too many tiles");
    const int grid = (int)(ntiles < sms ? ntiles : sms);
    const double pos = (double)a.B * a.T * a.E;
    double kreal = 0;
    for (int i = 0; i < a.nsrc; ++i) kreal += a.src[i].C;
    ProfScope ps("conv_tma", 2.0 * pos * a.ntaps * kreal * a.N * a.algo_frac,
                 (double)a.B * a.np_rows * 128.0 * a.nplanes * a.nslab * (a.npass == 3 ? 2 : 1) + 4.0 * pos * a.Cout * (a.resid ? 2 : 1) +
                     4.0 * a.ntaps * kreal * a.N,
                 st);
    EAB_CUDA(launch_k(conv_tma_kernel, dim3(grid), dim3(NTHREADS), (size_t)pl.total, st, a));
    EAB_LAUNCH_CHECK("conv_tma_kernel");
    return 0;
}

}  // namespace eab
