// tcgen05 / TMEM implicit-GEMM convolution for sm_100a -- the hot-path kernel of the 2-D conv family
// (gated / plain Conv2d and ConvTranspose2d with stride (1,2), EaBNet.py:391-490) and of every other
// conv-shaped layer whose channel counts fit (TCM 1x1 / dilated convs, w_dnn).
//
// GEMM view: rows = output positions (t,e) of one batch item, K = taps x input channels, N = output channels
// (value | gate halves side by side).  One CTA (320 threads) walks a contiguous range of 128-row tiles:
//
//   warps 0-3  A producers : gather the tap's input rows from HBM/L2 (channels-last, 128 B per row per K slab),
//                            apply the producer layer's norm + PReLU (Xform), round to TF32 (or split hi/lo for
//                            3xTF32), store into the 128B-swizzled K-major stage, fence.proxy.async, arrive.
//   warp 4     MMA issuer  : one elected lane issues 4 x tcgen05.mma.kind::tf32 (M128 x N x K8) per stage,
//                            tcgen05.commit frees the stage; a second commit per tile publishes the accumulator.
//   warp 5     B loader    : cp.async.bulk of the pre-swizzled, pre-rounded weight image (N x 128 B) per stage.
//   warps 6-9  epilogue    : tcgen05.ld accumulator rows from TMEM (double-buffered, so the next tile's MMAs
//                            overlap), bias + gate + ReLU, stage in smem, coalesced 128-bit stores (+ residual),
//                            per-(b,c) sum / sum-of-squares of the tile -> fp64 atomics (InstanceNorm statistics).
//
// The channel concat of a skip connection is two K slabs; a transposed conv is two launches (output parity);
// the causal halo and the ragged last tile are literal zero rows.  npass = 3 runs every K slab three times
// (A_hi B_hi + A_lo B_hi + A_hi B_lo) for fp32-grade accuracy where single-pass TF32 is not enough.
#include <cuda_fp16.h>

#include <algorithm>

#include "common.cuh"
#include "umma.cuh"

namespace eab {

namespace {

using namespace umma;

#ifdef EAB_UMMA_DEBUG
#define UDBG(i) do { if (a.dbg && blockIdx.x == 0) a.dbg[i] = clock64(); } while (0)
#else
#define UDBG(i) do { } while (0)
#endif

constexpr int TM = 128;             // rows per tile (UMMA M)
constexpr int KC = 64;              // fp16 elements per K slab = one 128-byte swizzle row
constexpr int NPROD = 256;          // producer threads (8 warps)
constexpr int RPP = TM * 8 / NPROD;  // rows per producer thread per K slab
constexpr int NEPI = 128;           // epilogue threads
constexpr int NTHREADS = NPROD + 64 + NEPI;
constexpr int A_STAGE_BYTES = TM * 128;

struct Smem {
    // offsets (bytes) into the 1024-aligned dynamic shared memory
    int a_off, b_off, stg_off, rowoff_off, coef_off, bias_off, post_off, bar_off, total;
    int nstages, a_stage_bytes, b_stage_bytes, stg_ld;
    int merged;                 // 3-pass layers: ONE ring stage per (tap, slab) holds A hi | A lo and W hi | W lo
};

// A 3-pass layer used to push three stages per (tap, slab) through the ring - A hi twice - each with its own wait, proxy fence
// and arrive; a single-tile CTA (every launch of a streaming step) spent most of its ~14 us on those 27-54 handshakes.  Merged
// stages hold both halves of the split once: one handshake and twelve MMAs per (tap, slab).  Falls back to per-pass stages when
// two merged stages do not fit (N = 256).
__host__ __device__ inline Smem smem_plan(int N, int Cout, int ncoef, int npass, int post = 0) {
    Smem s;
    s.stg_ld = Cout + 4;
    const int post_bytes = post ? 6 * Cout * 4 : 0;          // [2][3][Cout] coefficients of the fused residual sum
    const int fixed = TM * s.stg_ld * 4 + 2 * TM * 8 + 3 * ncoef * 4 + N * 4 + post_bytes + 256 + 64;
    const int avail = 227 * 1024 - 1024 - fixed;
    s.merged = 0;
    s.a_stage_bytes = A_STAGE_BYTES;
    s.b_stage_bytes = N * 128;
    if (npass == 3 && avail / (2 * (A_STAGE_BYTES + N * 128)) >= 2) {
        s.merged = 1;
        s.a_stage_bytes = 2 * A_STAGE_BYTES;
        s.b_stage_bytes = 2 * N * 128;
    }
    const int stage = s.a_stage_bytes + s.b_stage_bytes;
    int ns = avail / stage;
    if (ns > 6) ns = 6;
    s.nstages = ns;
    s.a_off = 0;
    s.b_off = ns * s.a_stage_bytes;
    s.stg_off = s.b_off + ns * s.b_stage_bytes;
    s.rowoff_off = s.stg_off + TM * s.stg_ld * 4;
    s.coef_off = s.rowoff_off + 2 * TM * 8;         // row offsets into out, and into resid (a ring of its own when streaming)
    s.bias_off = (s.coef_off + 3 * ncoef * 4 + 15) / 16 * 16;
    s.post_off = s.bias_off + N * 4;
    s.bar_off = (s.post_off + post_bytes + 15) / 16 * 16;
    s.total = s.bar_off + 256 + 1024;          // + slack for the 1024-byte alignment of the base
    return s;
}

template <bool WIDE>
__global__ void __launch_bounds__(NTHREADS, 1) conv_umma_kernel(const UmmaConvArgs a) {
    extern __shared__ uint8_t smem_raw[];
    // align to 1024 B (128B-swizzle atom) by OFFSETTING the __shared__ array: a round trip through uintptr_t would
    // demote every later access to generic ST.E / LD.E
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const Smem sp = smem_plan(a.N, a.Cout, a.ncoef, a.npass, a.post);
    uint8_t* As = smem + sp.a_off;
    uint8_t* Bs = smem + sp.b_off;
    float* stg = reinterpret_cast<float*>(smem + sp.stg_off);
    long long* rowoff = reinterpret_cast<long long*>(smem + sp.rowoff_off);
    long long* rowoff_r = rowoff + TM;
    float* coef = reinterpret_cast<float*>(smem + sp.coef_off);
    float* sbias = reinterpret_cast<float*>(smem + sp.bias_off);
    float* pcoef = reinterpret_cast<float*>(smem + sp.post_off);       // [post | resid][scale, shift, slope][Cout] (a.post)
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + sp.bar_off);
    uint64_t* full = bars;                      // [nstages]  A stored (128 arrivals) + B bytes landed
    uint64_t* empty = bars + 8;                 // [nstages]  MMAs that read the stage have completed
    uint64_t* acc_full = bars + 16;             // [2]        accumulator complete
    uint64_t* acc_empty = bars + 18;            // [2]        accumulator drained by the epilogue
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;
    if (tid == 0) UDBG(0);
    const int NS = sp.nstages;
    const uint32_t tmem_cols = a.N <= 64 ? 128u : (a.N <= 128 ? 256u : 512u);      // two accumulators

    if (tid == 0) {
        for (int i = 0; i < NS; ++i) { mbar_init(&full[i], NPROD + 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], NEPI); }
        fence_barrier_init();
    }
    if (warp == NPROD / 32) tmem_alloc(tmem_slot, tmem_cols);
    for (int i = tid; i < a.N; i += NTHREADS) sbias[i] = a.bias ? __ldg(a.bias + i) : 0.f;
    if (a.step) {
        // streaming: one coefficient set for the whole launch (B == 1) and static normalisation (BatchNorm), i.e. weights only:
        // computed here, under the predecessor's tail, instead of after the dependency wait (4.8 of a one-tile launch's 12 us)
        const int C0s = a.src[0].C;
        for (int i = tid; i < a.ncoef; i += NTHREADS) {
            const int sidx = i < C0s ? 0 : 1;
            const int c = sidx ? i - C0s : i;
            float cs, ch, ca;
            xform_coeffs(a.src[sidx].xf, 0, a.src[sidx].C, c, cs, ch, ca);
            coef[i] = cs;
            coef[a.ncoef + i] = ch;
            coef[2 * a.ncoef + i] = a.src[sidx].xf.prelu ? ca : 1.f;
        }
        if (a.post)
            for (int i = tid; i < 2 * a.Cout; i += NTHREADS) {
                const int sidx = i / a.Cout, c = i - sidx * a.Cout;
                float cs, ch, ca;
                xform_coeffs(sidx ? a.resid_xf : a.post_xf, 0, a.Cout, c, cs, ch, ca);
                pcoef[(sidx * 3 + 0) * a.Cout + c] = cs;
                pcoef[(sidx * 3 + 1) * a.Cout + c] = ch;
                pcoef[(sidx * 3 + 2) * a.Cout + c] = ca;
            }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (tid == 0) UDBG(1);
    if (a.step) {                 // streaming launches are programmatic dependents (launch_k_pdl): everything above ran under the
        pdl_trigger();            // predecessor's tail; nothing below may start before its results are visible
        pdl_wait();
    }
    if (tid == 0) UDBG(2);

    // Two variants in one grid (streaming: both output parities of a transposed conv - they read the same frames and differ in
    // taps, weights and output columns): CTAs [0, grid0) run variant 0, the rest variant 1.  grid0 == 0: one variant.
    const bool v1 = a.grid0 > 0 && (int)blockIdx.x >= a.grid0;
    const int vE = v1 ? a.var1.E : a.E;
    const int v_ntaps = v1 ? a.var1.ntaps : a.ntaps;
    const int v_out_off = v1 ? a.var1.out_off : a.out_off;
    const int v_tpb = v1 ? a.var1.tiles_per_b : a.tiles_per_b;
    const float* vWhi = v1 ? a.var1.Whi : a.Whi;
    const float* vWlo = v1 ? a.var1.Wlo : a.Wlo;
    const unsigned bx = v1 ? blockIdx.x - (unsigned)a.grid0 : blockIdx.x;
    const unsigned gx = a.grid0 > 0 ? (v1 ? gridDim.x - (unsigned)a.grid0 : (unsigned)a.grid0) : gridDim.x;
    // contiguous tile range of this CTA
    const long long ntiles = (long long)a.B * v_tpb;
    const long long tile_begin = ntiles * bx / gx;
    const long long tile_end = ntiles * (bx + 1) / gx;
    const int rows_per_b = a.T * vE;
    const bool streaming = a.step != nullptr;
    const int nfr = streaming ? __ldg(a.step) : 0;             // absolute frame index of this step
    const int chunks_per_pass_unit = v_ntaps * a.nslab;       // (tap, slab) pairs
    const int nchunks = chunks_per_pass_unit * (sp.merged ? 1 : a.npass);      // ring stages per tile

    if (warp < NPROD / 32) {
        // =========================================================================== A producers
        // Software-pipelined over the flattened (tile, unit) stream: the global loads of units g+1, g+2 are in
        // flight while unit g is transformed and stored, across tile boundaries.  The inner loop is kept free of
        // divisions and 64-bit index arithmetic (it was issue-bound on exactly that): per-tile row coordinates,
        // incremental (tap, slab) counters, 32-bit element offsets from a per-batch base pointer.
        const int c4 = tid & 7;                 // 16-byte chunk inside the 128-byte slab row
        const int rbase = tid >> 3;             // rows rbase + 32 i, i < RPP
        const int upt = chunks_per_pass_unit;   // units per tile
        const long long total_units = (tile_end - tile_begin) * upt;
        const int nslab0 = WIDE ? a.nslab : a.src[0].C / KC;
        const int C0 = a.src[0].C, C1 = a.nsrc > 1 ? a.src[1].C : 0;
        const uint32_t st_off = (uint32_t)(rbase * 128 + ((c4 ^ (rbase & 7)) << 4));   // same swizzle for rows +32 i

        // ---- load-side cursor
        long long l_tile = tile_begin;
        int l_tap = 0, l_slab = 0;
        const float* xb0 = nullptr;
        const float* xb1 = nullptr;
        int rt[RPP], rf[RPP];                   // frame index and e * in_stride of this thread's rows (big negative if ragged)
        int rs[RPP];                            // streaming: rt = the row's STREAM, rs = that stream's start frame
        auto decode_tile = [&](long long tile) {
            const int b = (int)(tile / v_tpb);
            const int row0 = (int)(tile - (long long)b * v_tpb) * TM;
            xb0 = a.src[0].x + (size_t)b * a.T * a.Fin * C0;
            xb1 = a.nsrc > 1 ? a.src[1].x + (size_t)b * a.T * a.Fin * C1 : nullptr;
#pragma unroll
            for (int i = 0; i < RPP; ++i) {
                const int r = row0 + rbase + 32 * i;
                const int t = r / vE;
                rt[i] = r < rows_per_b ? t : -(1 << 28);
                rf[i] = (r - t * vE) * a.in_stride;
                rs[i] = (streaming && a.start && r < rows_per_b) ? __ldg(a.start + t) : 0;
            }
        };
        auto issue_loads = [&](float4 (&v)[2 * RPP], uint32_t& mask) {
            const int dtv = v1 ? a.var1.dt[l_tap] : a.dt[l_tap], dfv = v1 ? a.var1.df[l_tap] : a.df[l_tap];
            const bool second = l_slab >= nslab0;
            const float* __restrict__ xb = second ? xb1 : xb0;
            const int C = second ? C1 : C0;
            const int cc = (second ? l_slab - nslab0 : l_slab) * KC + c4 * 8;      // 8 channels = one 16-byte fp16 chunk
            mask = 0;
#pragma unroll
            for (int i = 0; i < RPP; ++i) {
                v[2 * i] = make_float4(0.f, 0.f, 0.f, 0.f);
                v[2 * i + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
                int tt = rt[i] - dtv;
                const int fi = rf[i] + dfv;
                if (streaming) {                 // frame n - dt of stream rt[i]: its ring slot, or the zeros of the causal padding
                    const int RT = second ? a.src[1].RT : a.src[0].RT;
                    const int fr = nfr - dtv;
                    tt = (rt[i] >= 0 && fr >= rs[i]) ? rt[i] * RT + ring_slot(fr, RT) : -1;
                }
                if ((streaming ? tt >= 0 : (unsigned)tt < (unsigned)a.T) && (unsigned)fi < (unsigned)a.Fin) {
                    const float* p = xb + (uint32_t)((tt * a.Fin + fi) * C + cc);
                    if (!WIDE) {
                        v[2 * i] = __ldg(reinterpret_cast<const float4*>(p));
                        v[2 * i + 1] = __ldg(reinterpret_cast<const float4*>(p) + 1);
                    } else {                     // first layer: window of kwidth floats, 8-byte aligned only
                        const float2* p2 = reinterpret_cast<const float2*>(p);
                        if (cc + 1 < a.kwidth) { const float2 q = __ldg(p2); v[2 * i].x = q.x; v[2 * i].y = q.y; }
                        if (cc + 3 < a.kwidth) { const float2 q = __ldg(p2 + 1); v[2 * i].z = q.x; v[2 * i].w = q.y; }
                        if (cc + 5 < a.kwidth) { const float2 q = __ldg(p2 + 2); v[2 * i + 1].x = q.x; v[2 * i + 1].y = q.y; }
                        if (cc + 7 < a.kwidth) { const float2 q = __ldg(p2 + 3); v[2 * i + 1].z = q.x; v[2 * i + 1].w = q.y; }
                    }
                    mask |= 1u << i;
                }
            }
            if (++l_slab == a.nslab) {
                l_slab = 0;
                if (++l_tap == v_ntaps) {
                    l_tap = 0;
                    if (++l_tile < tile_end) decode_tile(l_tile);
                }
            }
        };

        // ---- store-side state
        int cur_b = streaming ? 0 : -1;          // streaming: the coefficients were set in the prologue
        int stage = 0;
        uint32_t phase = 0;
        int s_tap = 0, s_slab = 0;
        long long s_tile = tile_begin;
        const int mode0 = WIDE ? 0 : (a.src[0].xf.affine == 0 && a.src[0].xf.prelu == 0 ? 0 : (a.src[0].xf.prelu == 1 ? 2 : 1));
        const int mode1 = a.nsrc > 1 ? (a.src[1].xf.affine == 0 && a.src[1].xf.prelu == 0 ? 0 : (a.src[1].xf.prelu == 1 ? 2 : 1)) : 0;
        auto consume = [&](float4 (&v)[2 * RPP], uint32_t mask) {
            if ((s_tap | s_slab) == 0) {
                const int b = (int)(s_tile / v_tpb);
                if (b != cur_b) {
                    named_bar_sync(1, NPROD);   // nobody still reads the previous coefficients
                    for (int i = tid; i < a.ncoef; i += NPROD) {
                        const int s = i < C0 ? 0 : 1;
                        const int c = s ? i - C0 : i;
                        float cs, ch, ca;
                        xform_coeffs(a.src[s].xf, b, a.src[s].C, c, cs, ch, ca);
                        coef[i] = cs;
                        coef[a.ncoef + i] = ch;
                        coef[2 * a.ncoef + i] = a.src[s].xf.prelu ? ca : 1.f;   // slope 1 == no PReLU
                    }
                    named_bar_sync(1, NPROD);
                    cur_b = b;
                    if (tid == 0) UDBG(3);
                }
            }
            const bool second = s_slab >= nslab0;
            const int mode = second ? mode1 : mode0;
            if (mode != 0) {
                const int ci = s_slab * KC + c4 * 8;            // slabs of source 1 follow those of source 0
                const uint32_t tmask = mask == (1u << RPP) - 1 ? 0xFFFFFFFFu : mask;   // interior tiles: no per-row test
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const float4 cs = *reinterpret_cast<const float4*>(coef + ci + 4 * h);
                    const float4 ch = *reinterpret_cast<const float4*>(coef + a.ncoef + ci + 4 * h);
                    const float4 ca = *reinterpret_cast<const float4*>(coef + 2 * a.ncoef + ci + 4 * h);
                    if (mode == 1) {             // norm -> PReLU (2-D blocks)
#pragma unroll
                        for (int i = 0; i < RPP; ++i) {
                            if (tmask & (1u << i)) {
                                float4& q = v[2 * i + h];
                                float x;
                                x = fmaf(q.x, cs.x, ch.x); q.x = fmaxf(x, 0.f) + ca.x * fminf(x, 0.f);
                                x = fmaf(q.y, cs.y, ch.y); q.y = fmaxf(x, 0.f) + ca.y * fminf(x, 0.f);
                                x = fmaf(q.z, cs.z, ch.z); q.z = fmaxf(x, 0.f) + ca.z * fminf(x, 0.f);
                                x = fmaf(q.w, cs.w, ch.w); q.w = fmaxf(x, 0.f) + ca.w * fminf(x, 0.f);
                            }
                        }
                    } else {                     // PReLU -> norm (TCM branches)
#pragma unroll
                        for (int i = 0; i < RPP; ++i) {
                            if (tmask & (1u << i)) {
                                float4& q = v[2 * i + h];
                                q.x = fmaf(fmaxf(q.x, 0.f) + ca.x * fminf(q.x, 0.f), cs.x, ch.x);
                                q.y = fmaf(fmaxf(q.y, 0.f) + ca.y * fminf(q.y, 0.f), cs.y, ch.y);
                                q.z = fmaf(fmaxf(q.z, 0.f) + ca.z * fminf(q.z, 0.f), cs.z, ch.z);
                                q.w = fmaf(fmaxf(q.w, 0.f) + ca.w * fminf(q.w, 0.f), cs.w, ch.w);
                            }
                        }
                    }
                }
            }
            uint4 hi[RPP];
#pragma unroll
            for (int i = 0; i < RPP; ++i)
                hi[i] = make_uint4(pack_h2(v[2 * i].x, v[2 * i].y), pack_h2(v[2 * i].z, v[2 * i].w),
                                   pack_h2(v[2 * i + 1].x, v[2 * i + 1].y), pack_h2(v[2 * i + 1].z, v[2 * i + 1].w));
            if (sp.merged) {
                mbar_wait_backoff(&empty[stage], phase ^ 1, 20u);
                uint8_t* A = As + stage * sp.a_stage_bytes + st_off;
#pragma unroll
                for (int i = 0; i < RPP; ++i) {
                    *reinterpret_cast<uint4*>(A + i * 4096) = hi[i];
                    *reinterpret_cast<uint4*>(A + A_STAGE_BYTES + i * 4096) =
                        make_uint4(pack_lo_h2(v[2 * i].x, v[2 * i].y, hi[i].x), pack_lo_h2(v[2 * i].z, v[2 * i].w, hi[i].y),
                                   pack_lo_h2(v[2 * i + 1].x, v[2 * i + 1].y, hi[i].z), pack_lo_h2(v[2 * i + 1].z, v[2 * i + 1].w, hi[i].w));
                }
                fence_proxy_async();
                mbar_arrive(&full[stage]);
                if (++stage == NS) { stage = 0; phase ^= 1; }
            } else
            for (int pass = 0; pass < a.npass; ++pass) {
                mbar_wait_backoff(&empty[stage], phase ^ 1, 20u);      // (polling: the suspend-hinted wait pays a wake-up latency per handshake,
                                                                       //  15-20 us of a one-tile streaming launch)
                uint8_t* A = As + stage * sp.a_stage_bytes + st_off;
                if (pass == 1) {                 // residual of the fp16 rounding, itself rounded to fp16
#pragma unroll
                    for (int i = 0; i < RPP; ++i)
                        *reinterpret_cast<uint4*>(A + i * 4096) =
                            make_uint4(pack_lo_h2(v[2 * i].x, v[2 * i].y, hi[i].x), pack_lo_h2(v[2 * i].z, v[2 * i].w, hi[i].y),
                                       pack_lo_h2(v[2 * i + 1].x, v[2 * i + 1].y, hi[i].z), pack_lo_h2(v[2 * i + 1].z, v[2 * i + 1].w, hi[i].w));
                } else {
#pragma unroll
                    for (int i = 0; i < RPP; ++i) *reinterpret_cast<uint4*>(A + i * 4096) = hi[i];
                }
                fence_proxy_async();
                mbar_arrive(&full[stage]);
                if (++stage == NS) { stage = 0; phase ^= 1; }
            }
            if (++s_slab == a.nslab) {
                s_slab = 0;
                if (++s_tap == v_ntaps) { s_tap = 0; ++s_tile; }
            }
        };

        float4 v0[2 * RPP], v1[2 * RPP];
        uint32_t m0 = 0, m1 = 0;
        decode_tile(l_tile);
        long long issued = 0;
        if (issued < total_units) { issue_loads(v0, m0); ++issued; }
        for (long long g = 0; g < total_units; g += 2) {
            if (issued < total_units) { issue_loads(v1, m1); ++issued; }
            consume(v0, m0);
            if (g + 1 >= total_units) break;
            if (issued < total_units) { issue_loads(v0, m0); ++issued; }
            consume(v1, m1);
        }
        if (tid == 0) UDBG(4);
    } else if (warp == NPROD / 32) {
        // =========================================================================== MMA issuer
        const uint32_t idesc = make_idesc(a.N);
        int stage = 0;
        uint32_t phase = 0;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (long long tile = tile_begin; tile < tile_end; ++tile) {
            mbar_wait_spin(&acc_empty[acc], acc_phase ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)(acc * a.N);
            for (int ch = 0; ch < nchunks; ++ch) {
                mbar_wait_spin(&full[stage], phase);
                tc_fence_after();
                if (lane == 0) {
                    const uint32_t a_addr = smem_u32(As + stage * sp.a_stage_bytes);
                    const uint32_t b_addr = smem_u32(Bs + stage * sp.b_stage_bytes);
                    if (sp.merged) {            // A hi W hi + A lo W hi + A hi W lo from one stage
#pragma unroll
                        for (int g = 0; g < 3; ++g) {
                            const uint32_t aa = a_addr + (g == 1 ? A_STAGE_BYTES : 0);
                            const uint32_t bb = b_addr + (g == 2 ? a.N * 128 : 0);
#pragma unroll
                            for (int k = 0; k < KC / 16; ++k)
                                umma_f16(d_tmem, make_desc(aa + k * 32), make_desc(bb + k * 32), idesc, (ch | g | k) ? 1u : 0u);
                        }
                    } else
#pragma unroll
                    for (int k = 0; k < KC / 16; ++k) {
                        const uint64_t ad = make_desc(a_addr + k * 32);
                        const uint64_t bd = make_desc(b_addr + k * 32);
                        umma_f16(d_tmem, ad, bd, idesc, (ch | k) ? 1u : 0u);
                    }
                    umma_commit(&empty[stage]);                         // stage reusable once these MMAs retire
                    if (ch == nchunks - 1) umma_commit(&acc_full[acc]);   // accumulator complete
                }
                __syncwarp();
                if (++stage == NS) { stage = 0; phase ^= 1; }
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        if (lane == 0) UDBG(5);
    } else if (warp == NPROD / 32 + 1) {
        // =========================================================================== B (weight) loader
        int stage = 0;
        uint32_t phase = 0;
        const uint32_t bytes = (uint32_t)(a.N * 128);                // one weight image: N rows x 128 B
        for (long long tile = tile_begin; tile < tile_end; ++tile) {
            for (int unit = 0; unit < chunks_per_pass_unit; ++unit) {
                if (sp.merged) {
                    mbar_wait_backoff(&empty[stage], phase ^ 1, 32u);
                    if (lane == 0) {
                        mbar_arrive_expect_tx(&full[stage], 2 * bytes);
                        bulk_copy_g2s(Bs + stage * sp.b_stage_bytes, vWhi + (size_t)unit * a.N * 32, bytes, &full[stage]);
                        bulk_copy_g2s(Bs + stage * sp.b_stage_bytes + bytes, vWlo + (size_t)unit * a.N * 32, bytes, &full[stage]);
                    }
                    __syncwarp();
                    if (++stage == NS) { stage = 0; phase ^= 1; }
                    continue;
                }
                for (int pass = 0; pass < a.npass; ++pass) {
                    mbar_wait_backoff(&empty[stage], phase ^ 1, 32u);
                    if (lane == 0) {
                        const float* img = (pass == 2 ? vWlo : vWhi) + (size_t)unit * a.N * 32;     // N rows x 128 B
                        mbar_arrive_expect_tx(&full[stage], bytes);
                        bulk_copy_g2s(Bs + stage * sp.b_stage_bytes, img, bytes, &full[stage]);
                    }
                    __syncwarp();
                    if (++stage == NS) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else {
        // =========================================================================== epilogue (last 4 warps)
        const int et = tid - (NPROD + 64);      // 0..127
        const int quad = warp & 3;              // TMEM lane quadrant this warp may read
        const int row = quad * 32 + lane;       // accumulator row == tile row
        const bool gated = a.gate_off > 0;
        const int ld = sp.stg_ld;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (long long tile = tile_begin; tile < tile_end; ++tile) {
            const int b = (int)(tile / v_tpb);
            const int row0 = (int)(tile - (long long)b * v_tpb) * TM;
            const int nvalid = min(TM, rows_per_b - row0);
            {
                const int r = row0 + row;
                long long off = -1;
                long long off_r = -1;
                if (r < rows_per_b) {
                    const int t = r / vE, e = r - t * vE;
                    const long long fo = e * a.out_stride + v_out_off;
                    if (streaming) {
                        off = (((long long)t * a.out_RT + ring_slot(nfr, a.out_RT)) * a.Fout + fo) * a.out_ld + a.out_coff;
                        off_r = (((long long)t * a.resid_RT + ring_slot(nfr, a.resid_RT)) * a.Fout + fo) * a.out_ld + a.out_coff;
                    } else {
                        off = ((((long long)b * a.T + t) * a.Fout) + fo) * a.out_ld + a.out_coff;
                        off_r = off;
                    }
                }
                rowoff[row] = off;
                rowoff_r[row] = off_r;
            }
            mbar_wait_backoff(&acc_full[acc], acc_phase, 64u);
            if (et == 0) UDBG(6);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * a.N);
            for (int c0 = 0; c0 < a.Cout; c0 += 16) {
                float v[16];
                tmem_ld16(taddr + c0, v);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float4 bq = *reinterpret_cast<const float4*>(sbias + c0 + 4 * i);
                    v[4 * i] += bq.x; v[4 * i + 1] += bq.y; v[4 * i + 2] += bq.z; v[4 * i + 3] += bq.w;
                }
                if (gated) {
                    float g[16];
                    tmem_ld16(taddr + a.gate_off + c0, g);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float4 bq = *reinterpret_cast<const float4*>(sbias + a.gate_off + c0 + 4 * i);
                        v[4 * i] *= sigmoid_f(g[4 * i] + bq.x);
                        v[4 * i + 1] *= sigmoid_f(g[4 * i + 1] + bq.y);
                        v[4 * i + 2] *= sigmoid_f(g[4 * i + 2] + bq.z);
                        v[4 * i + 3] *= sigmoid_f(g[4 * i + 3] + bq.w);
                    }
                }
                if (a.relu) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
                }
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    *reinterpret_cast<float4*>(stg + row * ld + c0 + 4 * i) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
            }
            tc_fence_before();
            mbar_arrive(&acc_empty[acc]);       // the MMA warp may start overwriting this accumulator
            named_bar_sync(2, NEPI);            // staging tile + row offsets complete
            // ---- coalesced stores (+ residual): Cout/4 threads per row
            const int tpr = a.Cout >> 2;
            const int rows_per_it = NEPI / tpr;
            const int cq = (et % tpr) * 4;
            for (int r = et / tpr; r < nvalid; r += rows_per_it) {
                float4 o = *reinterpret_cast<const float4*>(stg + r * ld + cq);
                const long long off = rowoff[r];
                if (a.post) {                   // module output = xform(this layer) + xform(module input conv): as combine_kernel computes it
                    float4 q = __ldg(reinterpret_cast<const float4*>(a.resid + rowoff_r[r] + cq));
                    const float* p0 = pcoef + cq;
                    const float* p1 = pcoef + 3 * a.Cout + cq;
                    const int m0 = a.post_xf.prelu, m1 = a.resid_xf.prelu;
                    q.x = xform_apply(q.x, p1[0], p1[a.Cout], p1[2 * a.Cout], m1) + xform_apply(o.x, p0[0], p0[a.Cout], p0[2 * a.Cout], m0);
                    q.y = xform_apply(q.y, p1[1], p1[a.Cout + 1], p1[2 * a.Cout + 1], m1) + xform_apply(o.y, p0[1], p0[a.Cout + 1], p0[2 * a.Cout + 1], m0);
                    q.z = xform_apply(q.z, p1[2], p1[a.Cout + 2], p1[2 * a.Cout + 2], m1) + xform_apply(o.z, p0[2], p0[a.Cout + 2], p0[2 * a.Cout + 2], m0);
                    q.w = xform_apply(q.w, p1[3], p1[a.Cout + 3], p1[2 * a.Cout + 3], m1) + xform_apply(o.w, p0[3], p0[a.Cout + 3], p0[2 * a.Cout + 3], m0);
                    o = q;
                } else if (a.resid) {
                    const float4 q = __ldg(reinterpret_cast<const float4*>(a.resid + rowoff_r[r] + cq));
                    o.x += q.x; o.y += q.y; o.z += q.z; o.w += q.w;
                    if (a.nstats) *reinterpret_cast<float4*>(stg + r * ld + cq) = o;
                }
                *reinterpret_cast<float4*>(a.out + off + cq) = o;
            }
            // ---- per-channel statistics of the tile (valid rows only)
            if (a.nstats) {
                if (a.resid) named_bar_sync(2, NEPI);
                const int nsc = a.nstats * a.Cout;
                for (int i = et; i < nsc * 2; i += NEPI) {
                    const int half = i / nsc;
                    const int sc = i - half * nsc;
                    const int s = sc / a.Cout, c = sc - s * a.Cout;
                    const float al = a.stat_alpha[s] ? __ldg(a.stat_alpha[s] + c) : 1.f;
                    const bool pre = a.stat_alpha[s] != nullptr;
                    float sum = 0.f, sq = 0.f;
                    const int r_lo = half * (TM / 2), r_hi = min(nvalid, r_lo + TM / 2);
                    for (int r = r_lo; r < r_hi; ++r) {
                        float u = stg[r * ld + c];
                        if (pre) u = prelu_f(u, al);
                        sum += u;
                        sq += u * u;
                    }
                    if (r_hi > r_lo) {
                        double* dst = a.stats[s] + ((size_t)b * (a.stats_ld ? a.stats_ld : a.Cout) + a.stats_coff + c) * 2;
                        atomicAdd(dst, (double)sum);
                        atomicAdd(dst + 1, (double)sq);
                    }
                }
            }
            if (et == 0) UDBG(7);
            named_bar_sync(2, NEPI);            // staging free for the next tile
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (tid == 0) UDBG(8);
    if (warp == NPROD / 32) tmem_dealloc(tmem_base, tmem_cols);
    if (warp == NPROD / 32 && lane == 0) UDBG(9);
}

}  // namespace

bool umma_conv_supported(const UmmaConvArgs& a) {
    if (a.N % 16 != 0 || a.N < 16 || a.N > 256) return false;
    if (a.Cout != 16 && a.Cout != 32 && a.Cout != 64 && a.Cout != 128) return false;
    if (a.gate_off > 0 && (a.gate_off != a.Cout || a.N != 2 * a.Cout)) return false;
    if (a.gate_off == 0 && a.N != a.Cout) return false;
    if (a.ntaps < 1 || a.ntaps > kMaxTaps) return false;
    if (a.wide) {
        if (a.nsrc != 1 || (a.src[0].C & 1) || (a.kwidth & 1)) return false;
        if (a.src[0].xf.affine != 0 || a.src[0].xf.prelu != 0) return false;     // wide mode reads raw input only
    } else {
        for (int i = 0; i < a.nsrc; ++i)
            if (a.src[i].C % KC != 0) return false;      // whole 64-channel fp16 slabs
    }
    if (a.out_ld % 4 != 0 || a.out_coff % 4 != 0) return false;
    return true;
}

// Both output parities of a streaming transposed conv as ONE launch (a streaming step is a chain of ~15 us launches whatever
// their size: 25 of the 81 conv launches of a step disappear).  The variants must agree in everything but taps, weights, output
// columns and row count.
int launch_conv_umma_pair(const UmmaConvArgs& a0, const UmmaConvArgs& a1, cudaStream_t st) {
    if (!a0.step || a0.wide || a1.wide || a0.grid0 || a1.grid0) return fail("conv_umma_pair: streaming, non-wide launches only");
    if (!umma_conv_supported(a0) || !umma_conv_supported(a1)) return fail("conv_umma_pair: unsupported shape");
#define EAB_PAIR_SAME(f) if (a0.f != a1.f) return fail("conv_umma_pair: the variants differ in " #f " (only taps, weights and output columns may differ)")
    EAB_PAIR_SAME(nsrc); EAB_PAIR_SAME(B); EAB_PAIR_SAME(T); EAB_PAIR_SAME(Fin); EAB_PAIR_SAME(Fout); EAB_PAIR_SAME(in_stride);
    EAB_PAIR_SAME(out_stride); EAB_PAIR_SAME(nslab); EAB_PAIR_SAME(npass); EAB_PAIR_SAME(N); EAB_PAIR_SAME(Cout); EAB_PAIR_SAME(gate_off);
    EAB_PAIR_SAME(out); EAB_PAIR_SAME(bias); EAB_PAIR_SAME(resid); EAB_PAIR_SAME(out_RT); EAB_PAIR_SAME(resid_RT); EAB_PAIR_SAME(ncoef);
    EAB_PAIR_SAME(relu); EAB_PAIR_SAME(out_ld); EAB_PAIR_SAME(out_coff); EAB_PAIR_SAME(post);
#undef EAB_PAIR_SAME
    for (int i = 0; i < a0.nsrc; ++i)
        if (a0.src[i].x != a1.src[i].x || a0.src[i].RT != a1.src[i].RT) return fail("conv_umma_pair: different sources");
    if (a1.E <= 0) return launch_conv_umma(a0, st);
    if (a0.E <= 0) return launch_conv_umma(a1, st);
    UmmaConvArgs a = a0;
    a.var1.E = a1.E; a.var1.out_off = a1.out_off; a.var1.ntaps = a1.ntaps; a.var1.tiles_per_b = a1.tiles_per_b;
    for (int k = 0; k < kMaxTaps; ++k) { a.var1.dt[k] = a1.dt[k]; a.var1.df[k] = a1.df[k]; }
    a.var1.Whi = a1.Whi; a.var1.Wlo = a1.Wlo;
    if (a.B != 1 || a.nstats != 0 || a.out_RT < 1 || (a.resid && a.resid_RT < 1)) return fail("conv_umma: bad streaming launch");
    if (a.post && (!a.resid || a.post_xf.affine == 1 || a.resid_xf.affine == 1)) return fail("conv_umma: a fused residual sum needs a residual and static normalisation");
    for (int i = 0; i < a.nsrc; ++i) {
        if (a.src[i].xf.affine == 1) return fail("conv_umma: a streaming launch needs static normalisation (its coefficients are set before the dependency wait)");
        int back = 0;
        for (int k = 0; k < a0.ntaps; ++k) back = std::max(back, a0.dt[k]);
        for (int k = 0; k < a1.ntaps; ++k) back = std::max(back, a1.dt[k]);
        if (a.src[i].RT < back + 1) return fail("conv_umma: source ring shorter than the receptive field");
    }
    const Smem sp = smem_plan(a.N, a.Cout, a.ncoef, a.npass, a.post);
    if (sp.nstages < 2) return fail("conv_umma: not enough shared memory for two stages");
    EAB_TRY(ensure_dynamic_smem(reinterpret_cast<const void*>(conv_umma_kernel<false>), sp.total));
    int sms = 0;
    EAB_TRY(device_sm_count(&sms));
    const long long nt0 = (long long)a.B * a0.tiles_per_b, nt1 = (long long)a.B * a1.tiles_per_b;
    int g0 = (int)nt0, g1 = (int)nt1;
    if (nt0 + nt1 > sms) {               // one CTA per SM: split the SMs by tile count
        g0 = (int)std::max(1ll, std::min((long long)sms - 1, (sms * nt0 + (nt0 + nt1) / 2) / (nt0 + nt1)));
        g1 = sms - g0;
    }
    a.grid0 = g0;
    double kreal = 0;
    for (int i = 0; i < a.nsrc; ++i) kreal += a.src[i].C;
    const double pos0 = (double)a.B * a.T * a0.E, pos1 = (double)a.B * a.T * a1.E;
    ProfScope ps("conv_umma", 2.0 * (pos0 * a0.ntaps + pos1 * a1.ntaps) * kreal * a.N * a.algo_frac,
                 4.0 * ((pos0 + pos1) * a.in_stride * kreal / 2.0 + (pos0 + pos1) * a.Cout * (a.resid ? 2 : 1) +
                        (double)(a0.ntaps + a1.ntaps) * kreal * a.N),
                 st);
    EAB_CUDA(launch_k_pdl(conv_umma_kernel<false>, dim3(g0 + g1), dim3(NTHREADS), (size_t)sp.total, st, a));
    EAB_LAUNCH_CHECK("conv_umma_kernel");
    return 0;
}

int launch_conv_umma(const UmmaConvArgs& a, cudaStream_t st) {
    if (!umma_conv_supported(a)) return fail("conv_umma: unsupported shape");
    if (a.B <= 0 || a.T <= 0 || a.E <= 0) return 0;
    if (a.post && !a.step) return fail("conv_umma: the fused residual sum is a streaming feature");
    if (a.step) {
        if (a.B != 1 || a.nstats != 0 || a.out_RT < 1 || (a.resid && a.resid_RT < 1)) return fail("conv_umma: bad streaming launch");
        if (a.post && (!a.resid || a.post_xf.affine == 1 || a.resid_xf.affine == 1)) return fail("conv_umma: a fused residual sum needs a residual and static normalisation");
        for (int i = 0; i < a.nsrc; ++i) {
            if (a.src[i].xf.affine == 1) return fail("conv_umma: a streaming launch needs static normalisation (its coefficients are set before the dependency wait)");
            int back = 0;
            for (int k = 0; k < a.ntaps; ++k) back = std::max(back, a.dt[k]);
            if (a.src[i].RT < back + 1) return fail("conv_umma: source ring shorter than the receptive field");
        }
    }
    const Smem sp = smem_plan(a.N, a.Cout, a.ncoef, a.npass, a.post);
    if (sp.nstages < 2) return fail("conv_umma: not enough shared memory for two stages");
    EAB_TRY(ensure_dynamic_smem(reinterpret_cast<const void*>(conv_umma_kernel<false>), sp.total));
    EAB_TRY(ensure_dynamic_smem(reinterpret_cast<const void*>(conv_umma_kernel<true>), sp.total));
    int sms = 0;
    EAB_TRY(device_sm_count(&sms));
    const long long ntiles = (long long)a.B * a.tiles_per_b;
    const int grid = (int)(ntiles < sms ? ntiles : sms);
    const double pos = (double)a.B * a.T * a.E;
    double kreal = 0;
    for (int i = 0; i < a.nsrc; ++i) kreal += a.src[i].C;
    if (a.wide) kreal = a.kwidth;
    ProfScope ps("conv_umma", 2.0 * pos * a.ntaps * kreal * a.N * a.algo_frac,
                 4.0 * (pos * a.in_stride * kreal / (a.out_stride > 1 ? 2.0 : 1.0) / (a.wide ? (double)a.kwidth / a.src[0].C / 2.0 : 1.0) +
                        pos * a.Cout * (a.resid ? 2 : 1) + (double)a.ntaps * kreal * a.N),
                 st);
    if (a.step) {
        if (a.wide) EAB_CUDA(launch_k_pdl(conv_umma_kernel<true>, dim3(grid), dim3(NTHREADS), (size_t)sp.total, st, a));
        else EAB_CUDA(launch_k_pdl(conv_umma_kernel<false>, dim3(grid), dim3(NTHREADS), (size_t)sp.total, st, a));
    } else if (a.wide) EAB_CUDA(launch_k(conv_umma_kernel<true>, dim3(grid), dim3(NTHREADS), (size_t)sp.total, st, a));
    else EAB_CUDA(launch_k(conv_umma_kernel<false>, dim3(grid), dim3(NTHREADS), (size_t)sp.total, st, a));
    EAB_LAUNCH_CHECK("conv_umma_kernel");
    return 0;
}

}  // namespace eab
