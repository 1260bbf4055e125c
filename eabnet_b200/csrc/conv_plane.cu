// tcgen05 implicit-GEMM convolution, "stage once, shift by descriptor" (sm_100a).
//
// conv_umma.cu gathers the A operand once PER TAP; with a 2x3 / 2x5 kernel every input row then crosses the
// L2 -> SM fabric (and the producers' convert path) 6 times, which is what bounded that kernel.  Here the GEMM rows
// are output positions (t, e') laid out with a padded pitch P >= E per frame, and the tile's transformed fp16 input
// (+ halo rows) is staged ONCE as one or two "planes" that use the same pitch:
//     transposed conv (one output parity) / 1-D dilated conv : plane col = fi
//     stride-2 conv                                          : even plane col c = fi/2, odd plane col c = (fi-1)/2
// Then the A operand of tap (dt, df) is the SAME shared-memory plane seen through a UMMA descriptor whose start address
// is shifted by  -dt*P + df'  rows of 128 bytes.  (128B-swizzled K-major operands may start at any 128-byte row: the
// tensor core swizzles on absolute shared-memory address bits - verified by tools/exp/shift_desc.cu on B200, with the
// descriptor's base_offset field left at 0.)  Out-of-range taps land in zero pad columns / zero rows of the plane.
//
// Roles (one CTA per SM, persistent over a contiguous tile range):
//   warps 0-15  plane producers : LDG (each input row once per tile), norm + PReLU, fp16 (hi / lo), swizzled STS
//   warp 16     MMA issuer      : per tile, per (tap, slab, pass): 4 x tcgen05.mma.kind::f16, double-buffered TMEM
//   warp 17     B loader        : cp.async.bulk of the weight image of each (tap, slab, pass) through a small ring
//   warps 18-21 epilogue        : TMEM -> bias/gate/ReLU -> smem staging -> coalesced stores, residual, statistics
#include <cuda_fp16.h>

#include "common.cuh"
#include "umma.cuh"

namespace eab {

namespace {

using namespace umma;

constexpr int TM = 128;
constexpr int KC = 64;
constexpr int NPROD = 512;
constexpr int NEPI = 128;
constexpr int NTHREADS = NPROD + 64 + NEPI;
constexpr int NSB = 3;               // weight ring stages

struct Plan {
    int R;                           // staged rows per plane = 128 + back + fwd
    int plane_bytes;                 // R * 128
    int buf_bytes;                   // nplanes * nslab * npassbuf * plane_bytes
    int b_stage_bytes, stg_ld;
    int planes_off, b_off, stg_off, rowoff_off, coef_off, bias_off, bar_off, total;
};

__host__ __device__ inline Plan make_plan(const PlaneConvArgs& a) {
    Plan p;
    p.R = TM + a.back + a.fwd;
    p.plane_bytes = ((p.R + 7) & ~7) * 128;      // whole 1024-byte swizzle atoms: the swizzle phase is an absolute-address property
    const int npb = a.npass == 3 ? 2 : 1;
    p.buf_bytes = (a.nplanes * a.nslab * npb * p.plane_bytes + 1023) / 1024 * 1024;
    p.b_stage_bytes = a.N * 128;
    p.stg_ld = a.Cout + 4;
    p.planes_off = 0;
    p.b_off = a.nbuf * p.buf_bytes;
    p.stg_off = p.b_off + NSB * p.b_stage_bytes;
    p.rowoff_off = p.stg_off + TM * p.stg_ld * 4;
    p.coef_off = p.rowoff_off + TM * 8;
    p.bias_off = (p.coef_off + 3 * a.ncoef * 4 + 15) / 16 * 16;
    p.bar_off = p.bias_off + a.N * 4;
    p.total = p.bar_off + 256 + 1024;
    return p;
}

__global__ void __launch_bounds__(NTHREADS, 1) conv_plane_kernel(const PlaneConvArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const Plan pl = make_plan(a);
    uint8_t* planes = smem + pl.planes_off;
    uint8_t* Bs = smem + pl.b_off;
    float* stg = reinterpret_cast<float*>(smem + pl.stg_off);
    long long* rowoff = reinterpret_cast<long long*>(smem + pl.rowoff_off);
    float* coef = reinterpret_cast<float*>(smem + pl.coef_off);
    float* sbias = reinterpret_cast<float*>(smem + pl.bias_off);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + pl.bar_off);
    uint64_t* plane_full = bars;            // [2]   producers done (one arrival per producer warp)
    uint64_t* plane_empty = bars + 2;       // [2]   all MMAs of the tile have read the planes
    uint64_t* b_full = bars + 4;            // [NSB]
    uint64_t* b_empty = bars + 8;           // [NSB]
    uint64_t* acc_full = bars + 12;         // [2]
    uint64_t* acc_empty = bars + 14;        // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;
    const uint32_t tmem_cols = a.N <= 64 ? 128u : (a.N <= 128 ? 256u : 512u);
    const int npb = a.npass == 3 ? 2 : 1;

    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&plane_full[i], NPROD / 32);
            mbar_init(&plane_empty[i], 1);
            mbar_init(&acc_full[i], 1);
            mbar_init(&acc_empty[i], NEPI);
        }
        for (int i = 0; i < NSB; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
        fence_barrier_init();
    }
    pdl_trigger();
    if (warp == NPROD / 32) tmem_alloc(tmem_slot, tmem_cols);
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

    if (warp < NPROD / 32) {
        // =========================================================================== plane producers
        const int C0 = a.src[0].C, C1 = a.nsrc > 1 ? a.src[1].C : 0;
        const int nslab0 = C0 / KC;
        const int R = pl.R;
        const int nps = a.nplanes * a.nslab;
        const int items_total = nps * R * 2;                       // work item = half a plane row: 32 channels
        const int mode0 = (a.src[0].xf.affine == 0 && a.src[0].xf.prelu == 0) ? 0 : (a.src[0].xf.prelu == 1 ? 2 : 1);
        const int mode1 = a.nsrc > 1 ? ((a.src[1].xf.affine == 0 && a.src[1].xf.prelu == 0) ? 0 : (a.src[1].xf.prelu == 1 ? 2 : 1)) : 0;
        // frames are shifted by `kf` so that the row coordinate fed to the magic division is never negative
        const int kf = (a.back + a.P - 1) / a.P + 1;
        int cur_b = -1;
        int buf = 0;
        uint32_t bphase = 0;
        const bool dbg_on = a.dbg != nullptr && blockIdx.x == 0 && tid == 0;
        long long t_wait = 0, t_items = 0, t_fence = 0, t_pref = 0;
        const long long t_start = dbg_on ? clock64() : 0;
        // item -> (plane/slab, row, half) decode + global loads; returns false for an out-of-range item
        struct Item { int ps, lam, hf, slab; bool ok, second; };
        auto decode_and_load = [&](int item, int tile, float4 (&v)[8], Item& it) {
            const int b = tile / a.tiles_per_b;
            const int row0 = (tile - b * a.tiles_per_b) * TM;
            it.hf = item & 1;
            const int rid = item >> 1;
            it.ps = rid / R;
            it.lam = rid - it.ps * R;
            const int plane = it.ps / a.nslab;
            it.slab = it.ps - plane * a.nslab;
            const int rho = row0 - a.back + kf * a.P + it.lam;     // >= 0
            const int tq = a.P == 1 ? rho : (int)__umulhi((unsigned)rho, a.p_magic);
            const int col = rho - tq * a.P;
            const int t = tq - kf;
            it.ok = t >= 0 && t < a.T && col < a.plane_cols[plane];
            it.second = it.slab >= nslab0;
            if (it.ok) {
                const int fi = col * a.col_stride + a.col_off[plane];
                const int C = it.second ? C1 : C0;
                const float* xb = (it.second ? a.src[1].x : a.src[0].x) + (size_t)b * a.T * a.Fin * C;
                const int cc = (it.second ? it.slab - nslab0 : it.slab) * KC + it.hf * 32;
                const float4* p = reinterpret_cast<const float4*>(xb + (uint32_t)((t * a.Fin + fi) * C + cc));
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = __ldg(p + j);
            }
        };
        auto rot_of = [&](int tile) { return (int)(((unsigned)(tile - tile_begin) * 97u) % (unsigned)NPROD); };
        auto first_item = [&](int tile) { int vt = tid + rot_of(tile); return vt >= NPROD ? vt - NPROD : vt; };

        float4 v[8];
        Item cur;
        cur.ok = false; cur.ps = 0; cur.lam = 0; cur.hf = 0; cur.slab = 0; cur.second = false;
        bool have = false;                                         // v / cur hold the prefetched first item of `tile`
        if (tile_begin < tile_end && first_item(tile_begin) < items_total) { decode_and_load(first_item(tile_begin), tile_begin, v, cur); have = true; }
        for (int tile = tile_begin; tile < tile_end; ++tile) {
            const int b = tile / a.tiles_per_b;
            if (b != cur_b) {
                named_bar_sync(1, NPROD);
                for (int i = tid; i < a.ncoef; i += NPROD) {
                    const int s = i < C0 ? 0 : 1;
                    const int c = s ? i - C0 : i;
                    float cs, ch, ca;
                    xform_coeffs(a.src[s].xf, b, a.src[s].C, c, cs, ch, ca);
                    coef[i] = cs;
                    coef[a.ncoef + i] = ch;
                    coef[2 * a.ncoef + i] = a.src[s].xf.prelu ? ca : 1.f;
                }
                named_bar_sync(1, NPROD);
                cur_b = b;
            }
            const long long w0 = dbg_on ? clock64() : 0;
            mbar_wait(&plane_empty[buf], bphase ^ 1);
            if (dbg_on) t_wait += clock64() - w0;
            uint8_t* pbuf = planes + buf * pl.buf_bytes;
            const long long i0 = dbg_on ? clock64() : 0;
            for (int item = first_item(tile); item < items_total; item += NPROD) {
                if (!have) decode_and_load(item, tile, v, cur);
                have = false;
                uint8_t* drow = pbuf + (cur.ps * npb) * pl.plane_bytes + cur.lam * 128;
                const int sw = cur.lam & 7;
                if (cur.ok) {
                    const int md = cur.second ? mode1 : mode0;
                    if (md != 0) {
                        const float* cb = coef + cur.slab * KC + cur.hf * 32;
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float4 cs = *reinterpret_cast<const float4*>(cb + 4 * j);
                            const float4 ch = *reinterpret_cast<const float4*>(cb + a.ncoef + 4 * j);
                            const float4 ca = *reinterpret_cast<const float4*>(cb + 2 * a.ncoef + 4 * j);
                            float4& q = v[j];
                            if (md == 1) {
                                float x;
                                x = fmaf(q.x, cs.x, ch.x); q.x = fmaxf(x, 0.f) + ca.x * fminf(x, 0.f);
                                x = fmaf(q.y, cs.y, ch.y); q.y = fmaxf(x, 0.f) + ca.y * fminf(x, 0.f);
                                x = fmaf(q.z, cs.z, ch.z); q.z = fmaxf(x, 0.f) + ca.z * fminf(x, 0.f);
                                x = fmaf(q.w, cs.w, ch.w); q.w = fmaxf(x, 0.f) + ca.w * fminf(x, 0.f);
                            } else {
                                q.x = fmaf(fmaxf(q.x, 0.f) + ca.x * fminf(q.x, 0.f), cs.x, ch.x);
                                q.y = fmaf(fmaxf(q.y, 0.f) + ca.y * fminf(q.y, 0.f), cs.y, ch.y);
                                q.z = fmaf(fmaxf(q.z, 0.f) + ca.z * fminf(q.z, 0.f), cs.z, ch.z);
                                q.w = fmaf(fmaxf(q.w, 0.f) + ca.w * fminf(q.w, 0.f), cs.w, ch.w);
                            }
                        }
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int c = 0; c < 4; ++c) {                      // four 16-byte fp16 chunks of this half row
                    const float4 lo4 = v[2 * c], hi4 = v[2 * c + 1];
                    uint4 hi;
                    hi.x = pack_h2(lo4.x, lo4.y); hi.y = pack_h2(lo4.z, lo4.w);
                    hi.z = pack_h2(hi4.x, hi4.y); hi.w = pack_h2(hi4.z, hi4.w);
                    const int off = ((cur.hf * 4 + c) ^ sw) << 4;
                    *reinterpret_cast<uint4*>(drow + off) = hi;
                    if (npb == 2) {
                        uint4 lo;
                        lo.x = pack_lo_h2(lo4.x, lo4.y, hi.x); lo.y = pack_lo_h2(lo4.z, lo4.w, hi.y);
                        lo.z = pack_lo_h2(hi4.x, hi4.y, hi.z); lo.w = pack_lo_h2(hi4.z, hi4.w, hi.w);
                        *reinterpret_cast<uint4*>(drow + pl.plane_bytes + off) = lo;
                    }
                }
            }
            // the next tile's first item: loads go out BEFORE the proxy fence, so their latency and the fence overlap
            const long long i1 = dbg_on ? clock64() : 0;
            if (tile + 1 < tile_end && first_item(tile + 1) < items_total) { decode_and_load(first_item(tile + 1), tile + 1, v, cur); have = true; }
            const long long i2 = dbg_on ? clock64() : 0;
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&plane_full[buf]);
            if (dbg_on) { t_items += i1 - i0; t_pref += i2 - i1; t_fence += clock64() - i2; }
            if (++buf == a.nbuf) { buf = 0; bphase ^= 1; }
        }
        if (dbg_on) { a.dbg[0] = clock64() - t_start; a.dbg[1] = t_wait; a.dbg[2] = tile_end - tile_begin; a.dbg[3] = items_total; a.dbg[13] = t_items; a.dbg[14] = t_fence; a.dbg[15] = t_pref; }
    } else if (warp == NPROD / 32) {
        // =========================================================================== MMA issuer
        const uint32_t idesc = make_idesc(a.N);
        int buf = 0, stage = 0, acc = 0;
        uint32_t bphase = 0, sphase = 0, aphase = 0;
        const bool dbg_on = a.dbg != nullptr && blockIdx.x == 0 && lane == 0;
        long long t_wacc = 0, t_wplane = 0, t_wb = 0;
        const long long t_start = dbg_on ? clock64() : 0;
        for (int tile = tile_begin; tile < tile_end; ++tile) {
            long long w0 = dbg_on ? clock64() : 0;
            mbar_wait(&acc_empty[acc], aphase ^ 1);
            long long w1 = dbg_on ? clock64() : 0;
            mbar_wait(&plane_full[buf], bphase);
            if (dbg_on) { t_wacc += w1 - w0; t_wplane += clock64() - w1; }
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)(acc * a.N);
            const uint32_t pbase = smem_u32(planes + buf * pl.buf_bytes);
            int unit = 0;
            for (int tap = 0; tap < a.ntaps; ++tap) {
                const int arow = a.back + a.tap_shift[tap];
                for (int slab = 0; slab < a.nslab; ++slab) {
                    const uint32_t a_hi = pbase + (uint32_t)(((a.tap_plane[tap] * a.nslab + slab) * npb) * pl.plane_bytes + arow * 128);
                    for (int pass = 0; pass < a.npass; ++pass, ++unit) {
                        w0 = dbg_on ? clock64() : 0;
                        mbar_wait(&b_full[stage], sphase);
                        if (dbg_on) t_wb += clock64() - w0;
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
        if (dbg_on) { a.dbg[4] = clock64() - t_start; a.dbg[5] = t_wacc; a.dbg[6] = t_wplane; a.dbg[7] = t_wb; }
    } else if (warp == NPROD / 32 + 1) {
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
    } else {
        // =========================================================================== epilogue
        const int et = tid - (NPROD + 64);
        const int quad = warp & 3;
        const int row = quad * 32 + lane;
        const bool gated = a.gate_off > 0;
        const int ld = pl.stg_ld;
        int acc = 0;
        uint32_t aphase = 0;
        const bool dbg_on = a.dbg != nullptr && blockIdx.x == 0 && et == 0;
        long long t_wfull = 0, t_tmem = 0, t_store = 0, t_stats = 0;
        const long long t_start = dbg_on ? clock64() : 0;
        for (int tile = tile_begin; tile < tile_end; ++tile) {
            const int b = tile / a.tiles_per_b;
            const int row0 = (tile - b * a.tiles_per_b) * TM;
            bool row_valid;
            {
                const int r = row0 + row;
                long long off = -1;
                if (r < rows_per_b) {
                    const int t = a.P == 1 ? r : (int)__umulhi((unsigned)r, a.p_magic);
                    const int e = r - t * a.P;
                    if (e < a.E) off = ((((long long)b * a.T + t) * a.Fout) + (e * a.out_stride + a.out_off)) * a.out_ld + a.out_coff;
                }
                rowoff[row] = off;
                row_valid = off >= 0;
            }
            const long long e0 = dbg_on ? clock64() : 0;
            mbar_wait(&acc_full[acc], aphase);
            const long long e1 = dbg_on ? clock64() : 0;
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
                if (!row_valid) {                // dummy / ragged rows contribute exact zeros to the statistics
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = 0.f;
                }
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    *reinterpret_cast<float4*>(stg + row * ld + c0 + 4 * i) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
            }
            tc_fence_before();
            mbar_arrive(&acc_empty[acc]);
            named_bar_sync(2, NEPI);
            const long long e2 = dbg_on ? clock64() : 0;
            if (!a.resid) {
                // one bulk shared -> global copy per valid row (Cout * 4 bytes), issued by the row's own thread
                fence_proxy_async();
                if (row_valid)
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(a.out + rowoff[row]),
                                 "r"(smem_u32(stg + row * ld)), "r"((uint32_t)(a.Cout * 4))
                                 : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            } else {
                const int tpr = a.Cout >> 2;
                const int rows_per_it = NEPI / tpr;
                const int cq = (et % tpr) * 4;
                for (int r0 = et / tpr; r0 < TM; r0 += 4 * rows_per_it) {
                    long long off[4];
                    float4 o[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int r = r0 + j * rows_per_it;
                        off[j] = r < TM ? rowoff[r] : -1;
                        o[j] = r < TM ? *reinterpret_cast<const float4*>(stg + r * ld + cq) : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        if (off[j] < 0) continue;
                        const float4 q = __ldg(reinterpret_cast<const float4*>(a.resid + off[j] + cq));
                        o[j].x += q.x; o[j].y += q.y; o[j].z += q.z; o[j].w += q.w;
                        if (a.nstats) *reinterpret_cast<float4*>(stg + (r0 + j * rows_per_it) * ld + cq) = o[j];
                        *reinterpret_cast<float4*>(a.out + off[j] + cq) = o[j];
                    }
                }
            }
            const long long e3 = dbg_on ? clock64() : 0;
            if (a.nstats) {
                if (a.resid) named_bar_sync(2, NEPI);
                const int nsc = a.nstats * a.Cout;
                for (int i = et; i < nsc * 2; i += NEPI) {
                    const int half = i / nsc;
                    const int sc = i - half * nsc;
                    const int s = sc / a.Cout, c = sc - s * a.Cout;
                    const bool pre = a.stat_alpha[s] != nullptr;
                    const float al = pre ? __ldg(a.stat_alpha[s] + c) : 1.f;
                    const float* col = stg + (half * (TM / 2)) * ld + c;
                    float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
#pragma unroll 8
                    for (int r = 0; r < TM / 2; r += 2) {      // invalid rows were staged as zeros
                        float u0 = col[r * ld], u1 = col[(r + 1) * ld];
                        u0 = fmaxf(u0, 0.f) + al * fminf(u0, 0.f);      // al == 1 when the statistics are of the raw value
                        u1 = fmaxf(u1, 0.f) + al * fminf(u1, 0.f);
                        s0 += u0; q0 = fmaf(u0, u0, q0);
                        s1 += u1; q1 = fmaf(u1, u1, q1);
                    }
                    double* dstp = a.stats[s] + ((size_t)b * (a.stats_ld ? a.stats_ld : a.Cout) + a.stats_coff + c) * 2;
                    atomicAdd(dstp, (double)(s0 + s1));
                    atomicAdd(dstp + 1, (double)(q0 + q1));
                }
            }
            if (!a.resid) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            named_bar_sync(2, NEPI);
            if (dbg_on) { const long long e4 = clock64(); t_wfull += e1 - e0; t_tmem += e2 - e1; t_store += e3 - e2; t_stats += e4 - e3; }
            if (++acc == 2) { acc = 0; aphase ^= 1; }
        }
        if (dbg_on) { a.dbg[8] = clock64() - t_start; a.dbg[9] = t_wfull; a.dbg[10] = t_tmem; a.dbg[11] = t_store; a.dbg[12] = t_stats; }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == NPROD / 32) tmem_dealloc(tmem_base, tmem_cols);
}

bool basic_ok(const PlaneConvArgs& a) {
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

}  // namespace

bool plane_conv_supported(const PlaneConvArgs& a_in) {
    if (!basic_ok(a_in)) return false;
    PlaneConvArgs a = a_in;
    a.nbuf = 1;
    return make_plan(a).total <= 227 * 1024;
}

int launch_conv_plane(PlaneConvArgs a, cudaStream_t st) {
    if (!plane_conv_supported(a)) return fail("conv_plane: unsupported shape");
    if (a.B <= 0 || a.T <= 0 || a.E <= 0) return 0;
    a.p_magic = a.P == 1 ? 0u : (unsigned)((1ull << 32) / (unsigned)a.P) + 1u;     // P == 1 is special-cased in the kernel
    a.nbuf = 2;
    if (make_plan(a).total > 227 * 1024) a.nbuf = 1;
    const Plan pl = make_plan(a);
    static int configured = 0;
    if (pl.total > configured) {
        EAB_CUDA(cudaFuncSetAttribute(conv_plane_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, pl.total));
        configured = pl.total;
    }
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        EAB_CUDA(cudaGetDevice(&dev));
        EAB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    }
    const long long ntiles = (long long)a.B * a.tiles_per_b;
    if (ntiles >= (1ll << 30)) return fail("conv_plane: too many tiles");
    const int grid = (int)(ntiles < sms ? ntiles : sms);
    const double pos = (double)a.B * a.T * a.E;
    double kreal = 0;
    for (int i = 0; i < a.nsrc; ++i) kreal += a.src[i].C;
    ProfScope ps("conv_plane", 2.0 * pos * a.ntaps * kreal * a.N * a.algo_frac,
                 4.0 * ((double)a.B * a.T * a.Fin * kreal + pos * a.Cout * (a.resid ? 2 : 1) + (double)a.ntaps * kreal * a.N), st);
    EAB_CUDA(launch_k(conv_plane_kernel, dim3(grid), dim3(NTHREADS), (size_t)pl.total, st, a));
    EAB_LAUNCH_CHECK("conv_plane_kernel");
    return 0;
}

}  // namespace eab
