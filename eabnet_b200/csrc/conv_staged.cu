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
//                      copy engine hides the latency.  Taps are row-shifted UMMA descriptors into those planes, weights are
//                      resident or stream through a small ring, accumulators are double-buffered in TMEM, and
//                      8 epilogue warps do bias / gate / ReLU / residual / statistics / 32-byte row stores.
//
// Since round 2 the 2-D conv layers with 64-channel sources run on conv_raw.cu (no stage pass, no planes in HBM); this pair
// remains for what that kernel does not take: the first layer (2M input channels, frequency-pair rows), the STFT-as-GEMM,
// pointwise / dilated GEMMs with a residual, ReLU, two statistics or wide sources (TCM layer path, w_dnn, GaGNet heads).
#include <cstdlib>
#include <cuda_fp16.h>

#include "common.cuh"
#include "umma.cuh"

namespace eab {

namespace {

using namespace umma;

constexpr int TM = 128;
constexpr int KC = 64;
constexpr int NSB_MAX = 8;                      // weight ring stages: a.nsb in [2, 8], as many as fit next to two plane buffers
constexpr int NEPI = 256;                       // 8 epilogue warps per group
// warp 0 plane copies, 1 MMA issuer, 2 weight loader, 3 idle, 4-11 epilogue.  One epilogue group: its statistics cost
// almost nothing (per-thread running sums), and at 384 threads it can afford the 64 accumulator registers.
constexpr int NTHREADS_STAGED = 128 + NEPI;

inline int ceil8(int x) { return (x + 7) & ~7; }
constexpr int STAGE_RB = 8;                     // 32-row blocks per stage CTA

// 8 consecutive channels (one 32-byte sector, one instruction) starting at element offset `eoff`
__device__ __forceinline__ void load8(const float* base, size_t eoff, float4& v0, float4& v1) { ld_global_nc_256(base + eoff, v0, v1); }

// ------------------------------------------------------------------------------------------------ stage kernel
// grid (row blocks, nplanes*nslab, B), 256 threads: 8 threads per plane row (8 channels = one 16-byte fp16 chunk each)
__global__ void __launch_bounds__(256) stage_kernel(const PlaneConvArgs a) {
    __shared__ float coef[6 * 64];
    const int b = blockIdx.z;
    const int ps = blockIdx.y;
    const int plane = ps / a.nslab;
    const int slab = ps - plane * a.nslab;
    const int C0 = a.src[0].C;
    const bool second = a.wide_k == 0 && slab * KC >= C0;
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
        if (a.wide_k > 0) {
            // first layer: the row is the raw kf x C window itself (no transform: the network input has none)
            if (r >= 0 && r < a.T * a.P) {
                const int t = (int)__umulhi((unsigned)r, a.p_magic);
                const int col = r - t * a.P;
                int k0 = slab * KC + c8 * 8;
                int tt = t;
                if (a.wide_kt > 1) {               // frames stacked along K (wide_k % 8 == 0: a chunk stays inside its frame)
                    const int j = k0 / a.wide_k;
                    k0 -= j * a.wide_k;
                    tt = j < a.wide_kt ? t - (a.wide_kt - 1 - j) : -1;
                }
                // the window never leaves its frame: in the pair layout the last row's second position is past the last bin
                const int wk = min(a.wide_k, (a.Fin - col * a.col_stride) * src.C);
                if (col < a.plane_cols[0] && k0 < wk && tt >= 0) {
                    const float* p = src.x + (((size_t)b * a.T + tt) * a.Fin + (size_t)col * a.col_stride) * src.C + k0;     // 8-byte aligned
                    float x[8];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        float2 q = make_float2(0.f, 0.f);
                        if (k0 + 2 * i < wk) q = __ldg(reinterpret_cast<const float2*>(p) + i);
                        x[2 * i] = q.x; x[2 * i + 1] = k0 + 2 * i + 1 < wk ? q.y : 0.f;
                    }
                    v0 = make_float4(x[0], x[1], x[2], x[3]);
                    v1 = make_float4(x[4], x[5], x[6], x[7]);
                }
            }
        } else if (r >= 0 && r < a.T * a.P) {
            const int t = a.P == 1 ? r : (int)__umulhi((unsigned)r, a.p_magic);
            const int col = r - t * a.P;
            if (col < a.plane_cols[plane]) {
                const int fi = col * a.col_stride + a.col_off[plane];
                const size_t eoff = (((size_t)b * a.T + t) * a.Fin + fi) * src.C + cbase + c8 * 8;
                load8(src.x, eoff, v0, v1);
                float4 w0, w1;
                if (src.x2) load8(src.x2, eoff, w0, w1);
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
    int b_stage_bytes, stg_ld, w_images;
    int b_off, stg_off, rowoff_off, bias_off, utab_off, bar_off, total;
};

__host__ __device__ inline Plan make_plan(const PlaneConvArgs& a) {
    Plan p;
    p.R = TM + a.back + a.fwd;
    p.plane_bytes = ((p.R + 7 + 7) & ~7) * 128;
    const int npb = a.npass == 3 ? 2 : 1;
    p.buf_bytes = a.nplanes * a.nslab * npb * p.plane_bytes;
    p.b_stage_bytes = a.N * 128;
    // resident weights: every (tap, slab) image, hi (+ lo for the 3-pass layers), stays in shared memory for the whole
    // launch.  Streaming them per tile through a 3-deep ring is bound by L2 latency (3 x 8..16 KB in flight ~ 16 B/clk),
    // which is what limited this kernel (measured: 8.8 of 14.4 ms with MMA and epilogue both disabled).
    p.w_images = a.ntaps * a.nslab * npb;
    p.stg_ld = a.Cout + 4;
    p.b_off = a.nbuf * p.buf_bytes;
    // ring mode, 3 passes: a stage holds the W-hi AND the W-lo image of one (tap, slab) - one wait and one commit per three units
    p.stg_off = p.b_off + (a.resident ? p.w_images : a.nsb * npb) * p.b_stage_bytes;
    p.rowoff_off = p.stg_off;
    p.bias_off = p.rowoff_off;
    p.utab_off = p.bias_off + a.N * 4;
    p.bar_off = (p.utab_off + 15) & ~15;
    p.total = p.bar_off + 256 + 1024;
    return p;
}

// fold 8 per-row values across the 32 rows (lanes) of a warp with a halving butterfly (9 shuffles): every lane gets the
// total of column ((lane>>4)&1)*4 + ((lane>>3)&1)*2 + ((lane>>2)&1)
__device__ __forceinline__ float fold8(const float (&u)[8], int lane) {
    const bool h16 = (lane & 16) != 0, h8 = (lane & 8) != 0, h4 = (lane & 4) != 0;
    float u4[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) u4[j] = (h16 ? u[j + 4] : u[j]) + __shfl_xor_sync(0xffffffffu, h16 ? u[j] : u[j + 4], 16);
    float u2[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) u2[j] = (h8 ? u4[j + 2] : u4[j]) + __shfl_xor_sync(0xffffffffu, h8 ? u4[j] : u4[j + 2], 8);
    float u1 = (h4 ? u2[1] : u2[0]) + __shfl_xor_sync(0xffffffffu, h4 ? u2[0] : u2[1], 4);
    u1 += __shfl_xor_sync(0xffffffffu, u1, 2);
    u1 += __shfl_xor_sync(0xffffffffu, u1, 1);
    return u1;
}

__global__ void __launch_bounds__(NTHREADS_STAGED, 1) conv_tma_kernel(const PlaneConvArgs a) {
    constexpr int NTHREADS = NTHREADS_STAGED;
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
    uint64_t* b_full = bars + 6;            // [NSB_MAX]
    uint64_t* b_empty = bars + 14;          // [NSB_MAX]
    uint64_t* acc_full = bars + 22;         // [2]
    uint64_t* acc_empty = bars + 24;        // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 26);

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;
    const uint32_t tmem_cols = a.N <= 64 ? 128u : (a.N <= 128 ? 256u : 512u);
    const int npb = a.npass == 3 ? 2 : 1;
    const int nimg = a.nplanes * a.nslab * npb;

    if (tid == 0) {
        for (int i = 0; i < 3; ++i) { mbar_init(&plane_full[i], 1); mbar_init(&plane_empty[i], 1); }
        for (int i = 0; i < NSB_MAX; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], NEPI); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, tmem_cols);
    for (int i = tid; i < a.N; i += NTHREADS) sbias[i] = a.bias ? __ldg(a.bias + i) : 0.f;
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
        // =========================================================================== MMA issuers
        // One elected thread issues every MMA and the warp is a serial instruction chain: measured ~200 cycles per MMA
        // when the tap / slab / pass geometry was re-derived inside a lane-0-only loop (4x the tensor pipe's 48-64).
        // The loop below is convergent (all lanes, warp-uniform operands from the constant bank), the per-unit operand
        // offsets come tabulated from the launcher, and a K step only adds 2 to the two descriptor low words.
        // (Two issuing warps alternating tiles were measured in round 1: no gain over one issuer + 3 plane buffers.)

        const uint32_t idesc = make_idesc(a.N);
        const bool dbg_on = a.dbg != nullptr && blockIdx.x == 0;
        long long t_wacc = 0, t_wplane = 0, t_wb = 0, t_issue = 0, t_commit = 0;
        const long long t_start = dbg_on ? clock64() : 0;
        const uint32_t bs_lo = desc_lo(smem_u32(Bs));
        const uint32_t bstep = (uint32_t)pl.b_stage_bytes >> 4;
        int stage = 0;
        uint32_t sphase = 0;
        if (a.resident && tile_begin < tile_end) { mbar_wait(&b_full[0], 0u); tc_fence_after(); }
        for (int tile = tile_begin; tile < tile_end; ++tile) {
            const int ord = tile - tile_begin;                   // ordinal of the tile in this CTA: every phase follows from it
            const int acc = ord & 1;
            const int buf = ord % a.nbuf;
            const uint32_t bphase = (uint32_t)((ord / a.nbuf) & 1);
            const uint32_t aphase = (uint32_t)((ord >> 1) & 1);
            const int b = tile / a.tiles_per_b;
            const int row0 = (tile - b * a.tiles_per_b) * TM;
            const int rs = a.np_front + row0 - a.back;
            const int lead = rs & 7;               // rows between the copy start and the tile's first plane row
            const long long w0 = dbg_on ? clock64() : 0;
            mbar_wait(&acc_empty[acc], aphase ^ 1);
            const long long w1 = dbg_on ? clock64() : 0;
            mbar_wait(&plane_full[buf], bphase);
            if (dbg_on) { t_wacc += w1 - w0; t_wplane += clock64() - w1; }
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)(acc * a.N);
            const uint32_t origin = (smem_u32(planes + buf * pl.buf_bytes) + (uint32_t)lead * 128u) >> 4;
            int pass3 = 0;                                       // units run (tap, slab) major, pass minor
            for (int unit = 0; unit < units_per_tile; ++unit) {
                uint32_t blo;
                if (!a.resident) {                               // single issuer in ring mode: the ring position runs on
                    if (pass3 == 0) {
                        const long long w2 = dbg_on ? clock64() : 0;
                        mbar_wait(&b_full[stage], sphase);
                        if (dbg_on) t_wb += clock64() - w2;
                        tc_fence_after();
                    }
                    blo = bs_lo + (uint32_t)(stage * npb + (pass3 == 2 ? 1 : 0)) * bstep;
                } else {
                    blo = bs_lo + (uint32_t)a.unit_b[unit] * bstep;
                }
                const uint32_t alo = ((origin + a.unit_a[unit]) & 0x3FFFu) | (1u << 16);
                const long long w3 = dbg_on ? clock64() : 0;
                if (a.ksteps == 3) umma_f16_lo_elect_x3(d_tmem, alo, blo, idesc, unit ? 1u : 0u);
                else umma_f16_lo_elect_x4(d_tmem, alo, blo, idesc, unit ? 1u : 0u);
                const long long w4 = dbg_on ? clock64() : 0;
                if (++pass3 == a.npass) {
                    pass3 = 0;
                    if (!a.resident) {
                        umma_commit_elect(&b_empty[stage]);
                        if (++stage == a.nsb) { stage = 0; sphase ^= 1; }
                    }
                }
                if (dbg_on) { t_issue += w4 - w3; t_commit += clock64() - w4; }
            }
            umma_commit_elect(&acc_full[acc]);
            umma_commit_elect(&plane_empty[buf]);
        }
        if (dbg_on && lane == 0) { a.dbg[4] = clock64() - t_start; a.dbg[5] = t_wacc; a.dbg[6] = t_wplane; a.dbg[7] = t_wb; a.dbg[2] = tile_end - tile_begin; a.dbg[0] = clock64() - t_start; a.dbg[1] = t_issue; a.dbg[3] = t_commit; a.dbg[13] = a.nsb; a.dbg[14] = a.nbuf; a.dbg[15] = a.resident; }
    } else if (warp == 2) {
        // =========================================================================== B (weight) loader
        int stage = 0;
        uint32_t sphase = 0;
        const uint32_t bytes = (uint32_t)pl.b_stage_bytes;
        if (a.resident) {
            if (lane == 0 && tile_begin < tile_end) {
                mbar_arrive_expect_tx(&b_full[0], bytes * (uint32_t)pl.w_images);
                for (int ts = 0; ts < a.ntaps * a.nslab; ++ts)
                    for (int hl = 0; hl < npb; ++hl)
                        bulk_copy_g2s(Bs + (ts * npb + hl) * pl.b_stage_bytes, (hl ? a.Wlo : a.Whi) + (size_t)ts * a.N * 32, bytes, &b_full[0]);
            }
            __syncwarp();
        } else
        for (int tile = tile_begin; tile < tile_end; ++tile) {
            for (int ts = 0; ts < a.ntaps * a.nslab; ++ts) {
                mbar_wait(&b_empty[stage], sphase ^ 1);
                if (lane == 0) {
                    mbar_arrive_expect_tx(&b_full[stage], bytes * (uint32_t)npb);
                    uint8_t* dst = Bs + (size_t)stage * npb * pl.b_stage_bytes;
                    bulk_copy_g2s(dst, a.Whi + (size_t)ts * a.N * 32, bytes, &b_full[stage]);
                    if (npb == 2) bulk_copy_g2s(dst + pl.b_stage_bytes, a.Wlo + (size_t)ts * a.N * 32, bytes, &b_full[stage]);
                }
                __syncwarp();
                if (++stage == a.nsb) { stage = 0; sphase ^= 1; }
            }
        }
    } else if (warp >= 4) {
        // =========================================================================== epilogue (8 warps per group)
        // Register-only: thread = output row (TMEM lane), warp = (lane quadrant, half of the channels).  Values go
        // TMEM -> registers -> HBM as full 32-byte sectors; the per-channel statistics are folded across the 32 rows of
        // the warp with a halving butterfly (9 shuffles per 8 columns per statistic) into one owner lane per column,
        // accumulated in registers over the CTA's tiles and flushed with fp64 atomics only when the batch item changes.
        // No staging tile, no block barrier: shared memory is left to the operand traffic of the tensor core.
        constexpr int grp = 0;
        const int quad = warp & 3;              // TMEM lane quadrant (warp % 4)
        const int chalf = ((warp - 4) >> 2) & 1;        // which half of the output channels this warp converts
        const int row = quad * 32 + lane;
        const bool gated = a.gate_off > 0;
        const int cper = a.Cout >> 1;           // channels per half (8 .. 64, multiple of 8)
        const int niter = cper >> 3;
        const int own = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);     // column (of 8) this lane ends up owning
        constexpr int ngrp = 1;
        int acc = grp;
        uint32_t aphase = 0;
        const bool dbg_on = a.dbg != nullptr && blockIdx.x == 0 && warp == 4 && lane == 0;
        long long t_wfull = 0, t_tmem = 0;
        const long long t_start = dbg_on ? clock64() : 0;
        // Statistics, fast path (one statistic, <= 32 channels per thread): every thread keeps running sums of ITS row's
        // values per channel over all the tiles of a batch item; the cross-row butterfly runs once per batch item.
        const bool fast_stats = a.nstats == 1 && niter <= 4;
        float rs_sum[32], rs_sq[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) { rs_sum[i] = 0.f; rs_sq[i] = 0.f; }
        float ssum[2][8], ssq[2][8];            // [statistic][iteration]: totals of column chalf*cper + it*8 + own
#pragma unroll
        for (int s2 = 0; s2 < 2; ++s2)
#pragma unroll
            for (int it = 0; it < 8; ++it) { ssum[s2][it] = 0.f; ssq[s2][it] = 0.f; }
        int cur_b = -1;
        auto flush = [&](int b) {
            if (b < 0 || a.nstats == 0) return;
            if (fast_stats) {
#pragma unroll
                for (int it = 0; it < 4; ++it) {
                    if (it >= niter) continue;
                    float u8[8], w8[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) { u8[i] = rs_sum[it * 8 + i]; w8[i] = rs_sq[it * 8 + i]; rs_sum[it * 8 + i] = 0.f; rs_sq[it * 8 + i] = 0.f; }
                    ssum[0][it] = fold8(u8, lane);
                    ssq[0][it] = fold8(w8, lane);
                }
            }
            if ((lane & 3) == 0) {
#pragma unroll
                for (int s2 = 0; s2 < 2; ++s2) {
                    if (s2 >= a.nstats) continue;
#pragma unroll
                    for (int it = 0; it < 8; ++it) {
                        if (it >= niter) continue;
                        double* dstp = a.stats[s2] + ((size_t)b * (a.stats_ld ? a.stats_ld : a.Cout) + a.stats_coff + chalf * cper + it * 8 + own) * 2;
                        atomicAdd(dstp, (double)ssum[s2][it]);
                        atomicAdd(dstp + 1, (double)ssq[s2][it]);
                    }
                }
            }
#pragma unroll
            for (int s2 = 0; s2 < 2; ++s2)
#pragma unroll
                for (int it = 0; it < 8; ++it) { ssum[s2][it] = 0.f; ssq[s2][it] = 0.f; }
        };
        for (int tile = tile_begin + grp; tile < tile_end; tile += ngrp) {
            const int b = tile / a.tiles_per_b;
            if (b != cur_b) { flush(cur_b); cur_b = b; }
            const int row0 = (tile - b * a.tiles_per_b) * TM;
            long long off = -1;
            {
                const int r = row0 + row;
                if (r < rows_per_b) {
                    const int t = a.P == 1 ? r : (int)__umulhi((unsigned)r, a.p_magic);
                    const int e = r - t * a.P;
                    if (e < a.E) off = ((((long long)b * a.T + t) * a.Fout) + (e * a.out_stride + a.out_off)) * a.out_ld + a.out_coff;
                }
            }
            const bool row_valid = off >= 0;
            const long long e0 = dbg_on ? clock64() : 0;
            mbar_wait(&acc_full[acc], aphase);
            const long long e1 = dbg_on ? clock64() : 0;
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * a.N);
#pragma unroll
            for (int it = 0; it < 8; ++it) {
                if (it >= niter) continue;
                const int c0 = chalf * cper + it * 8;
                uint32_t rv[8], rg[8];
                tmem_ld8_nowait(taddr + c0, rv);
                if (gated) tmem_ld8_nowait(taddr + a.gate_off + c0, rg);
                float4 q0 = make_float4(0.f, 0.f, 0.f, 0.f), q1 = q0;
                if (a.resid && row_valid) {          // residual rows: issued under the TMEM load
                    q0 = __ldg(reinterpret_cast<const float4*>(a.resid + off + c0));
                    q1 = __ldg(reinterpret_cast<const float4*>(a.resid + off + c0 + 4));
                }
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
                v[0] += q0.x; v[1] += q0.y; v[2] += q0.z; v[3] += q0.w;
                v[4] += q1.x; v[5] += q1.y; v[6] += q1.z; v[7] += q1.w;
                if (a.stft_M > 0) {
                    // DFT-as-GEMM epilogue: columns are (re, im) pairs; z |z|^-1/2 (test.py:41-43) -> spec[b][t][f][mic][2]
                    // rows are (hop t, mic) with pitch stft_M: consecutive lanes write consecutive mics of one (t, f)
                    const int r = row0 + row;
                    const int t = a.P == 1 ? r : (int)__umulhi((unsigned)r, a.p_magic);
                    const int mic = r - t * a.P;
                    if (t < a.stft_T && mic < a.stft_M) {
                        const int f0 = (a.out_coff + c0) >> 1;
                        float2* sp = reinterpret_cast<float2*>(a.out) + (((size_t)b * a.stft_T + t) * a.stft_F + f0) * a.stft_M + mic;
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            if (f0 + i < a.stft_F) {
                                const float x = v[2 * i], y = v[2 * i + 1];
                                const float mag = sqrtf(x * x + y * y);
                                const float sc = mag > 0.f ? rsqrtf(mag) : 0.f;
                                sp[(size_t)i * a.stft_M] = make_float2(x * sc, y * sc);
                            }
                        }
                    }
                } else if (row_valid) {
                    st_global_256(a.out + off + c0, v);
                }
                if (a.nstats) {
                    if (!row_valid) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) v[i] = 0.f;
                    }
                    if (fast_stats) {
                        if (it < 4) {
                            float u[8];
                            if (a.stat_alpha[0]) {
                                const float4 a0 = __ldg(reinterpret_cast<const float4*>(a.stat_alpha[0] + c0));
                                const float4 a1 = __ldg(reinterpret_cast<const float4*>(a.stat_alpha[0] + c0 + 4));
                                const float al[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
                                for (int i = 0; i < 8; ++i) u[i] = fmaxf(v[i], 0.f) + al[i] * fminf(v[i], 0.f);
                            } else {
#pragma unroll
                                for (int i = 0; i < 8; ++i) u[i] = v[i];
                            }
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                rs_sum[(it & 3) * 8 + i] += u[i];
                                rs_sq[(it & 3) * 8 + i] = fmaf(u[i], u[i], rs_sq[(it & 3) * 8 + i]);
                            }
                        }
                    } else {
#pragma unroll
                        for (int s2 = 0; s2 < 2; ++s2) {
                            if (s2 >= a.nstats) continue;
                            float u[8], w[8];
                            if (a.stat_alpha[s2]) {
                                const float4 a0 = __ldg(reinterpret_cast<const float4*>(a.stat_alpha[s2] + c0));
                                const float4 a1 = __ldg(reinterpret_cast<const float4*>(a.stat_alpha[s2] + c0 + 4));
                                const float al[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
                                for (int i = 0; i < 8; ++i) u[i] = fmaxf(v[i], 0.f) + al[i] * fminf(v[i], 0.f);
                            } else {
#pragma unroll
                                for (int i = 0; i < 8; ++i) u[i] = v[i];
                            }
#pragma unroll
                            for (int i = 0; i < 8; ++i) w[i] = u[i] * u[i];
                            ssum[s2][it] += fold8(u, lane);
                            ssq[s2][it] += fold8(w, lane);
                        }
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(&acc_empty[acc]);
            if (dbg_on) { t_wfull += e1 - e0; t_tmem += clock64() - e1; }
            if (ngrp == 2) aphase ^= 1;
            else if (++acc == 2) { acc = 0; aphase ^= 1; }
        }
        flush(cur_b);
        if (dbg_on) { a.dbg[8] = clock64() - t_start; a.dbg[9] = t_wfull; a.dbg[10] = t_tmem; a.dbg[11] = 0; a.dbg[12] = 0; }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, tmem_cols);
}

int choose_nbuf(PlaneConvArgs& a) {
    a.nsb = 3;
    a.resident = 1;                                      // 1. resident weights next to 3 or 2 plane buffers
    for (int nb = 3; nb >= 2; --nb) {
        a.nbuf = nb;
        if (make_plan(a).total <= 227 * 1024) return nb;
    }
    a.resident = 0;                                      // 2. weight ring: as many plane buffers as fit with a 3-stage ring
    static const int nbuf_env = getenv("EAB_TMA_NBUF") ? atoi(getenv("EAB_TMA_NBUF")) : 3;     // diagnostics
    for (int nb = nbuf_env >= 1 && nbuf_env <= 3 ? nbuf_env : 3; nb >= 1; --nb) {                    //    (measured: 3 buffers + 3 stages beats 2 buffers + 6 stages), then
        a.nbuf = nb;                                     //    deepen the ring into whatever shared memory is left
        a.nsb = 3;
        if (make_plan(a).total > 227 * 1024) continue;
        for (int ns = NSB_MAX; ns > 3; --ns) {
            a.nsb = ns;
            if (make_plan(a).total <= 227 * 1024) return nb;
        }
        a.nsb = 3;
        return nb;
    }
    return 0;
}

}  // namespace

// shape rules of the padded-pitch row space (shared by stage_kernel / conv_tma_kernel / conv_raw_kernel)
bool plane_conv_supported(const PlaneConvArgs& a) {
    if (a.N % 16 != 0 || a.N < 16 || a.N > 256) return false;
    if (a.Cout != 16 && a.Cout != 32 && a.Cout != 64 && a.Cout != 128) return false;
    if (a.gate_off > 0 && (a.gate_off != a.Cout || a.N != 2 * a.Cout)) return false;
    if (a.gate_off == 0 && a.N != a.Cout) return false;
    if (a.ntaps < 1 || a.ntaps > kMaxTaps || a.nplanes < 1 || a.nplanes > 2) return false;
    for (int i = 0; i < a.nsrc; ++i)
        if (a.src[i].C % KC != 0) return false;
    for (int i = 0; i < a.nplanes; ++i)
        if (a.plane_cols[i] > a.P) return false;
    for (int i = 0; i < a.ntaps; ++i)
        if (a.back + a.tap_shift[i] < 0 || a.tap_shift[i] > a.fwd) return false;
    if (a.out_ld % 4 != 0 || a.out_coff % 4 != 0 || a.P < a.E || a.P < 1) return false;
    if (((long long)a.T * a.P + a.back + 4 * TM + 2ll * a.P) * a.P >= (1ll << 31)) return false;     // magic-division range
    return true;
}

int staged_rows(const PlaneConvArgs& a, int* front) {
    const int f = ceil8(a.back);
    if (front) *front = f;
    return ceil8(f + a.tiles_per_b * TM + a.fwd + 8);
}

bool staged_conv_fits(const PlaneConvArgs& a_in) {
    PlaneConvArgs a = a_in;
    if (a.ntaps * a.nslab * a.npass > kMaxConvUnits) return false;
    if (a.nplanes * a.nslab * (a.npass == 3 ? 2 : 1) > 16) return false;
    return choose_nbuf(a) > 0;
}

bool staged_conv_supported(const PlaneConvArgs& a_in) {
    if (!plane_conv_supported(a_in)) return false;
    PlaneConvArgs a = a_in;
    if (a.ntaps * a.nslab * a.npass > kMaxConvUnits) return false;
    if ((a.Cout != 16 && a.Cout != 32 && a.Cout != 64 && a.Cout != 128) || a.nplanes * a.nslab * (a.npass == 3 ? 2 : 1) > 16) return false;
    return choose_nbuf(a) > 0;
}

int launch_stage(const PlaneConvArgs& a_in, cudaStream_t st) {
    PlaneConvArgs a = a_in;
    a.p_magic = a.P == 1 ? 0u : (unsigned)((1ull << 32) / (unsigned)a.P) + 1u;
    dim3 grid((a.np_rows + 32 * STAGE_RB - 1) / (32 * STAGE_RB), a.nplanes * a.nslab, a.B);
    double cin = 0, inb = 0;
    for (int i = 0; i < a.nsrc; ++i) {
        cin += a.src[i].C;
        inb += (double)a.B * a.T * a.Fin * a.src[i].C * 4.0 * (a.src[i].x2 ? 2 : 1);
    }
    const double elems = (double)a.B * a.T * a.Fin * cin;
    // a pure design cost: no algorithmic bytes (SURVEY.md 8d counts the layer's fp32 input once, at the conv launch)
    ProfScope ps("stage", 4.0 * elems, 0.0, st, inb + (double)a.B * a.np_rows * 128.0 * a.nplanes * a.nslab * (a.npass == 3 ? 2 : 1));
    EAB_CUDA(launch_k(stage_kernel, grid, dim3(256), (size_t)0, st, a));
    EAB_LAUNCH_CHECK("stage_kernel");
    return 0;
}

int launch_conv_staged(PlaneConvArgs a, cudaStream_t st) {
    if (a.B <= 0 || a.T <= 0 || a.E <= 0) return 0;
    a.p_magic = a.P == 1 ? 0u : (unsigned)((1ull << 32) / (unsigned)a.P) + 1u;
    static const int k3_env = getenv("EAB_TMA_K3") ? atoi(getenv("EAB_TMA_K3")) : 1;
    if (!k3_env) a.ksteps = 0;
    if (!choose_nbuf(a)) return fail("conv_staged: shared-memory budget");
    const Plan pl = make_plan(a);
    {
        const int npb = a.npass == 3 ? 2 : 1;
        const int units = a.ntaps * a.nslab * a.npass;
        if (units > kMaxConvUnits) return fail("conv_staged: too many K units");
        for (int u = 0; u < units; ++u) {
            const int pass = u % a.npass, ts = u / a.npass;
            const int slab = ts % a.nslab, tap = ts / a.nslab;
            const unsigned a_rel = (unsigned)((((a.tap_plane[tap] * a.nslab + slab) * npb) + (pass == 1 ? 1 : 0)) * pl.plane_bytes +
                                              (a.back + a.tap_shift[tap]) * 128);
            a.unit_a[u] = a_rel >> 4;
            a.unit_b[u] = (unsigned short)((tap * a.nslab + slab) * npb + (pass == 2 ? 1 : 0));
        }
    }
    EAB_TRY(ensure_dynamic_smem(reinterpret_cast<const void*>(conv_tma_kernel), pl.total));
    int sms = 0;
    EAB_TRY(device_sm_count(&sms));
    const long long ntiles = (long long)a.B * a.tiles_per_b;
    if (ntiles >= (1ll << 30)) return fail("conv_staged: too many tiles");
    const int grid = (int)(ntiles < sms ? ntiles : sms);
    const double pos = (double)a.B * a.T * a.E;
    double kreal = 0;
    for (int i = 0; i < a.nsrc; ++i) kreal += a.src[i].C;
    if (a.stft_M > 0) kreal = 160;                           // a hop per tap
    const double in_bytes = (double)a.B * a.np_rows * 128.0 * a.nplanes * a.nslab * (a.npass == 3 ? 2 : 1);
    if (a.stft_M == 0 && ((a.out_ld & 7) || (a.out_coff & 7) || (reinterpret_cast<uintptr_t>(a.out) & 31)))
        return fail("conv_staged: output rows must be 32-byte aligned");
    // algorithmic: the layer's fp32 input(s) once + its output once; the parity launches of a transposed conv each count a
    // share (algo_in_share) of the input they both read.  moved: the staged fp16 planes this launch reads + output + weights
    double in_algo = 0;
    for (int i = 0; i < a.nsrc; ++i) in_algo += 4.0 * a.B * a.T * a.Fin * a.src[i].C;
    if (a.stft_M > 0) in_algo = 4.0 * a.B * a.stft_M * 160.0 * a.T;
    else if (a.wide_k > 0) in_algo = 4.0 * a.B * a.T * a.Fin * a.src[0].C;
    in_algo *= a.algo_in_share > 0.f ? a.algo_in_share : 1.f;
    const double out_b = a.stft_M > 0 ? 8.0 * a.B * a.stft_T * a.stft_M * 64.0 * (a.out_coff / 128 == 2 ? 33.0 / 64.0 : 1.0) : 4.0 * pos * a.Cout;
    ProfScope ps(a.stft_M > 0 ? "stft" : "conv_tma", 2.0 * pos * a.ntaps * kreal * a.N * a.algo_frac,
                 in_algo + out_b + (a.resid ? 4.0 * pos * a.Cout : 0.0), st,
                 in_bytes + 4.0 * pos * a.Cout * (a.resid ? 2 : 1) + 4.0 * a.ntaps * kreal * a.N);
    EAB_CUDA(launch_k(conv_tma_kernel, dim3(grid), dim3(NTHREADS_STAGED), (size_t)pl.total, st, a));
    EAB_LAUNCH_CHECK("conv_tma_kernel");
    return 0;
}

}  // namespace eab
