// nn.LSTM(64 -> 64) layer of LSTM_BF (EaBNet.py:591-592, 610-611) on the tensor cores (sm_100a), fp32-grade:
// the second generation of lstm_umma.cu.  Same arithmetic (3-pass fp16 split GEMMs into TMEM, 7 SFU operations per hidden
// unit, packed f32x2 cell arithmetic), different schedule: a CTA owns 128 sequences as TWO INTERLEAVED SUB-BATCHES of 64
// that take turns - while the cell warps work on sub-batch A, the tensor core runs the whole gate GEMM of sub-batch B's next
// step (input projection and recurrent half, 24 MMAs), so the ~1 300 cycles the cell warps of lstm_umma.cu wait for the
// recurrent MMAs of every step (tools/lstm_timing.py) are covered by the other sub-batch's cell phase.
//
// What makes the interleave pay is the TMEM access shape.  A warp may only touch the 32 TMEM lanes of its quadrant
// (warp % 4), and the SFU cost of a warp instruction does not depend on how many lanes are live, so both sub-batches
// must live in EVERY quadrant: sub-batch s owns MMA rows 32 q + 16 s + [0, 16) of each quadrant q, and the cell warps read
// their gates with tcgen05.ld.16x256b (16 lanes x 8 columns per repeat; thread i gets rows i/4 and i/4 + 8, columns
// 2 (i%4), 2 (i%4) + 1 of every 8-column group - probed with tools/exp/tmem_layout.cu): all 32 threads of a warp are busy
// on 16 rows, and all four schedulers' SFUs work (lstm_umma.cu: 96 rows on three quadrants, the fourth scheduler idle).
// Every MMA still spans M = 128: the rows of the other sub-batch produce accumulator lanes nobody reads.
//
// MEASURED (B = 64, T = 601, F = 161; profiles/r02_v6_lstm_pp_full.txt) and NOT the default (option lstm_pp): 7 030 cycles per
// step for 128 sequences against 6 500 for 96 in lstm_umma.cu - 19 % fewer cycles per sequence, but 81 CTAs of 128 take 5.3 ms
// for both layers where 108 CTAs of 96 take 4.1.  The schedule is bound by the TENSOR pipe (88 % active): an M = 128 MMA costs
// the same 128 cycles (N = 256, K = 16) whether 64 or 128 of its rows are live, so two sub-batches double the tensor work
// (48 MMAs = 6 100 cycles per step) and that, not the SFU (51 %), sets the step.  It would pay with 256 sequences per CTA
// (two full M = 128 sub-batches), i.e. from ~38 000 sequences (B >= 236 at F = 161) per GPU upwards.
//
//   warps 0-15  cell warps  : (quadrant, quarter of the hidden units); per step and sub-batch 2 rows x 4 units per thread:
//                             tcgen05.ld, bias, sigmoid / tanh, c / h update, h_t -> the A operand (fp16 hi | lo) and -> HBM
//                             (8-byte stores: the four threads of a row fill one 32-byte sector), arrive
//   warps 16-23 x producers : sub-batch (warp & 1), two threads per sequence: prefetch x one step ahead, decoder norm +
//                             PReLU + LayerNorm, fp16 hi | lo -> the A operand
//   warp 24     MMA issuer  : per (step, sub-batch): x_t W_ih^T (fresh) + h_{t-1} W_hh^T, 24 MMAs M128 x N256 x K16, one commit
#include <cstdlib>

#include "common.cuh"
#include "umma.cuh"

namespace eab {

namespace {

using namespace umma;

constexpr int H = 64;
constexpr int ROWS = 128;                               // sequences per CTA = MMA rows
constexpr int NCELLW = 16, NPRODW = 8;
constexpr int MMA_WARP = NCELLW + NPRODW;
constexpr int NTHREADS = (MMA_WARP + 1) * 32;           // 800
constexpr int NCELL = NCELLW * 32;                      // every cell thread holds cells of both sub-batches
constexpr int NXP = NPRODW * 32 / 2;                    // producer threads per sub-batch (two per sequence)
constexpr int SLAB_BYTES = ROWS * 128;                  // one 64-wide fp16 K slab of the A operand
constexpr int A_BYTES = 4 * SLAB_BYTES;                 // [hi|lo][x|h]
constexpr int B_SLAB_BYTES = 256 * 128;                 // one K slab of the weight image (256 gate rows)
constexpr int B_BYTES = 4 * B_SLAB_BYTES;               // [hi|lo][x|h]
constexpr int MISC_FLOATS = 256 + 2 * 3 * 64 + 2 * 64;  // bias, transform coefficients (2 batch items), LayerNorm
constexpr int SMEM_BYTES = A_BYTES + B_BYTES + MISC_FLOATS * 4 + 128 + 1024;

#define MBW mbar_wait_spin

__device__ __forceinline__ float ex2(float x) {          // one SFU op, ~2 ulp
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// two fp32 operations per instruction (sm_100: FFMA2 / FMUL2 / FADD2); each half rounds exactly like the scalar form
struct F2 { float x, y; };
__device__ __forceinline__ uint64_t pk(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ F2 upk(uint64_t r) { F2 o; asm("mov.b64 {%0, %1}, %2;" : "=f"(o.x), "=f"(o.y) : "l"(r)); return o; }
__device__ __forceinline__ F2 fma2(F2 a, F2 b, F2 c) {
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(pk(a.x, a.y)), "l"(pk(b.x, b.y)), "l"(pk(c.x, c.y)));
    return upk(d);
}
__device__ __forceinline__ F2 mul2(F2 a, F2 b) {
    uint64_t d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(pk(a.x, a.y)), "l"(pk(b.x, b.y)));
    return upk(d);
}
__device__ __forceinline__ F2 add2(F2 a, F2 b) {
    uint64_t d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(pk(a.x, a.y)), "l"(pk(b.x, b.y)));
    return upk(d);
}
__device__ __forceinline__ F2 sub2(F2 a, F2 b) {
    uint64_t d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(pk(a.x, a.y)), "l"(pk(b.x, b.y)));
    return upk(d);
}
// 16 TMEM lanes x 32 columns: r[4 g + 2 rs + e] = (lane base + i/4 + 8 rs, column 8 g + 2 (i%4) + e) for thread i of the warp
__device__ __forceinline__ void tmem_ld_16x256b_x4(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
}

__device__ __forceinline__ uint8_t* a_slab(uint8_t* A, int hl, int slab) { return A + (hl * 2 + slab) * SLAB_BYTES; }

__global__ void __launch_bounds__(NTHREADS, 1) lstm_pp_kernel(const LstmArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* As = smem;
    uint8_t* Bs = smem + A_BYTES;
    float* sbias = reinterpret_cast<float*>(smem + A_BYTES + B_BYTES);
    float* coef = sbias + 256;                  // [2][3][64]
    float* lng = coef + 2 * 3 * 64;
    float* lnb = lng + 64;
    uint64_t* bars = reinterpret_cast<uint64_t*>(lnb + 64);
    uint64_t* h_ready = bars;                   // [2]    h_{t-1} of the sub-batch is in the A operand (all cell threads)
    uint64_t* acc_full = bars + 2;              // [2]    the 256 gate columns of the sub-batch's step are complete
    uint64_t* x_ready = bars + 6;               // [2]    x_t of the sub-batch is in the A operand (its producer threads)
    uint64_t* x_free = bars + 8;                // [2]    the sub-batch's MMAs of the step have consumed the x operand
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;
    const int NQ = a.B * a.F;
    const int q0 = blockIdx.x * ROWS;
    const int b0 = q0 / a.F;

    if (tid == 0) {
        for (int s = 0; s < 2; ++s) {
            mbar_init(&h_ready[s], NCELL);
            mbar_init(&acc_full[s], 1);
            mbar_init(&x_ready[s], NXP);
            mbar_init(&x_free[s], 1);
        }
        fence_barrier_init();
    }
    if (warp == MMA_WARP) tmem_alloc(tmem_slot, 512);
    {
        const uint4* src = reinterpret_cast<const uint4*>(a.Wimg);
        uint4* dst = reinterpret_cast<uint4*>(Bs);
        for (int i = tid; i < B_BYTES / 16; i += NTHREADS) dst[i] = __ldg(src + i);
        uint4* az = reinterpret_cast<uint4*>(As);
        for (int i = tid; i < A_BYTES / 16; i += NTHREADS) az[i] = make_uint4(0, 0, 0, 0);      // h_{-1} = 0
        for (int i = tid; i < 256; i += NTHREADS)       // image row order: half*128 + quarter*32 + gate*8 + j; gate 2 (g) feeds tanh
            sbias[i] = __ldg(a.bias + i) * (((i >> 3) & 3) == 2 ? -2.f : -1.f) * 1.4426950408889634f;
        for (int i = tid; i < 2 * 64; i += NTHREADS) {
            const int bb = i >> 6, c = i & 63;
            float cs = 1.f, ch = 0.f, ca = 1.f;
            if (b0 + bb < a.B) xform_coeffs(a.src.xf, b0 + bb, 64, c, cs, ch, ca);
            coef[(bb * 3 + 0) * 64 + c] = cs;
            coef[(bb * 3 + 1) * 64 + c] = ch;
            coef[(bb * 3 + 2) * 64 + c] = a.src.xf.prelu ? ca : 1.f;
        }
        for (int i = tid; i < 64; i += NTHREADS) {
            lng[i] = a.layer_norm ? __ldg(a.ln_g + i) : 1.f;
            lnb[i] = a.layer_norm ? __ldg(a.ln_b + i) : 0.f;
        }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < NCELLW) {
        // ======================================================================= cell warps
        // thread = (quadrant, quarter of the hidden units, i): per sub-batch s the rows 32 quad + 16 s + i/4 (+ 8) and per
        // column half u the units quarter*16 + u*8 + 2 (i%4) + {0, 1}; gate columns u*128 + quarter*32 + gate*8 + ...
        const int quad = warp & 3;
        const int qtr = warp >> 2;
        const int j2 = (lane & 3) * 2;
        const int rr = lane >> 2;
        // output rows of this thread: float index of (row, unit quarter*16 + 2 (i%4)) at t = 0, or -1 past the last sequence
        long long orow[2][2];
#pragma unroll
        for (int s = 0; s < 2; ++s)
#pragma unroll
            for (int rs = 0; rs < 2; ++rs) {
                const int q = q0 + quad * 32 + 16 * s + rr + 8 * rs;
                const int bq = q / a.F;
                orow[s][rs] = q < NQ ? ((long long)bq * a.T * a.F + (q - bq * a.F)) * H + qtr * 16 + j2 : -1;
            }
        const size_t ostep = (size_t)a.F * H;
        float c[2][2][2][2];                    // [sub-batch][half][row][unit of the pair]
#pragma unroll
        for (int i = 0; i < 16; ++i) (&c[0][0][0][0])[i] = 0.f;
        mbar_arrive(&h_ready[0]);               // h_{-1} = 0 is in place
        mbar_arrive(&h_ready[1]);
        const bool dbg_on = a.dbg != nullptr && blockIdx.x == 0 && tid == 0;
        long long t_wait = 0, t_cell = 0, t_w1 = 0;
        const long long t_start = dbg_on ? clock64() : 0;
        constexpr float L2E = 1.4426950408889634f;
        const F2 one = {1.f, 1.f}, nl = {-L2E, -L2E}, nl2 = {-2.f * L2E, -2.f * L2E};
        for (int t = 0; t < a.T; ++t) {
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                const long long c0 = dbg_on ? clock64() : 0;
                long long c1 = 0;
                const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32 + 16 * s) << 16) + (uint32_t)(s * 256 + qtr * 32);
                uint32_t h_hi[2][2], h_lo[2][2];
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    if (u == 0) {
                        MBW(&acc_full[s], (uint32_t)(t & 1));
                        if (dbg_on) c1 = clock64();
                        tc_fence_after();
                    }
                    uint32_t r[16];
                    tmem_ld_16x256b_x4(taddr + u * 128, r);
                    const float* bi = sbias + u * 128 + qtr * 32 + j2;     // pre-scaled by -log2(e) (i, f, o) / -2 log2(e) (g)
                    const float2 b_i = *reinterpret_cast<const float2*>(bi);
                    const float2 b_f = *reinterpret_cast<const float2*>(bi + 8);
                    const float2 b_g = *reinterpret_cast<const float2*>(bi + 16);
                    const float2 b_o = *reinterpret_cast<const float2*>(bi + 24);
                    tmem_wait_ld();
                    // 7 SFU ops per unit instead of 10: the sigmoid / tanh quotients share their reciprocals
                    //   c' = sig(f) c + sig(i) tanh(g) = [c (1+Ei)(1+Eg) + (1-Eg)(1+Ef)] / [(1+Ei)(1+Ef)(1+Eg)]
                    //   h  = sig(o) tanh(c')          = (1-Ec) / [(1+Eo)(1+Ec)]
                    // with Ei = e^-i, Ef = e^-f, Eg = e^-2g, Ec = e^-2c', Eo = e^-o; exponents clamped from above only (2^40)
#pragma unroll
                    for (int rs = 0; rs < 2; ++rs) {
                        const F2 gi = {__uint_as_float(r[0 + 2 * rs]), __uint_as_float(r[1 + 2 * rs])};
                        const F2 gf = {__uint_as_float(r[4 + 2 * rs]), __uint_as_float(r[5 + 2 * rs])};
                        const F2 gg = {__uint_as_float(r[8 + 2 * rs]), __uint_as_float(r[9 + 2 * rs])};
                        const F2 go = {__uint_as_float(r[12 + 2 * rs]), __uint_as_float(r[13 + 2 * rs])};
                        const F2 ti = fma2(gi, nl, {b_i.x, b_i.y});
                        const F2 tf = fma2(gf, nl, {b_f.x, b_f.y});
                        const F2 tg = fma2(gg, nl2, {b_g.x, b_g.y});
                        const F2 to = fma2(go, nl, {b_o.x, b_o.y});
                        const F2 Ei = {ex2(fminf(ti.x, 40.f)), ex2(fminf(ti.y, 40.f))};
                        const F2 Ef = {ex2(fminf(tf.x, 40.f)), ex2(fminf(tf.y, 40.f))};
                        const F2 Eg = {ex2(fminf(tg.x, 40.f)), ex2(fminf(tg.y, 40.f))};
                        const F2 Eo = {ex2(fminf(to.x, 40.f)), ex2(fminf(to.y, 40.f))};
                        const F2 A = add2(one, Ei), Bf = add2(one, Ef), G = add2(one, Eg);
                        const F2 AG = mul2(A, G);
                        const F2 num = fma2({c[s][u][rs][0], c[s][u][rs][1]}, AG, mul2(sub2(one, Eg), Bf));
                        const F2 den = mul2(AG, Bf);
                        const F2 cn = mul2(num, {rcp_approx(den.x), rcp_approx(den.y)});
                        c[s][u][rs][0] = cn.x; c[s][u][rs][1] = cn.y;
                        const F2 tc = mul2(cn, nl2);
                        const F2 Ec = {ex2(fminf(tc.x, 40.f)), ex2(fminf(tc.y, 40.f))};
                        const F2 hden = mul2(add2(one, Eo), add2(one, Ec));
                        const F2 h = mul2(sub2(one, Ec), {rcp_approx(hden.x), rcp_approx(hden.y)});
                        h_hi[u][rs] = pack_h2(h.x, h.y);
                        h_lo[u][rs] = pack_lo_h2(h.x, h.y, h_hi[u][rs]);
                        if (orow[s][rs] >= 0) *reinterpret_cast<float2*>(a.out + orow[s][rs] + (size_t)t * ostep + u * 8) = make_float2(h.x, h.y);
                    }
                }
                // h_t overwrites the operand rows the sub-batch's MMAs of this step read (both halves complete: waited above)
#pragma unroll
                for (int rs = 0; rs < 2; ++rs) {
                    const int row = quad * 32 + 16 * s + rr + 8 * rs;
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        const uint32_t off = (uint32_t)(row * 128 + (((qtr * 2 + u) ^ (row & 7)) << 4) + j2 * 2);
                        *reinterpret_cast<uint32_t*>(a_slab(As, 0, 1) + off) = h_hi[u][rs];
                        *reinterpret_cast<uint32_t*>(a_slab(As, 1, 1) + off) = h_lo[u][rs];
                    }
                }
                tc_fence_before();
                fence_proxy_async();
                mbar_arrive(&h_ready[s]);       // h_t: operand of step t+1
                if (dbg_on) { t_wait += c1 - c0; t_cell += clock64() - c1; }
            }
        }
        if (dbg_on) { a.dbg[0] = clock64() - t_start; a.dbg[1] = t_wait; a.dbg[2] = t_cell; a.dbg[3] = 0; a.dbg[4] = a.T; a.dbg[6] = t_w1; }
    } else if (warp < MMA_WARP) {
        // ======================================================================= x producers (two threads per sequence)
        const int pw = warp - NCELLW;
        const int s = pw & 1;                                        // the sub-batch this warp feeds
        const int row = 32 * (pw >> 1) + 16 * s + (lane >> 1);
        const int half = lane & 1;                                   // channels half*32 .. +32
        const int q = q0 + row;
        const bool valid = q < NQ;
        const int bq = valid ? q / a.F : 0;
        const int fq = valid ? q - bq * a.F : 0;
        const float* xp = a.src.x + (((size_t)bq * a.T) * a.F + fq) * H + half * 32;
        const size_t xstep = (size_t)a.F * H;
        const float* cf = coef + (valid ? bq - b0 : 0) * 3 * 64 + half * 32;
        const float* lg = lng + half * 32;
        const float* lb = lnb + half * 32;
        const int mode = (a.src.xf.affine == 0 && a.src.xf.prelu == 0) ? 0 : (a.src.xf.prelu == 1 ? 2 : 1);
        uint8_t* xrow_hi = a_slab(As, 0, 0) + row * 128;
        uint8_t* xrow_lo = a_slab(As, 1, 0) + row * 128;
        float4 xr[8];
        auto load = [&](int t) {
            if (valid) {
                const float4* p = reinterpret_cast<const float4*>(xp + (size_t)t * xstep);
#pragma unroll
                for (int i = 0; i < 8; ++i) xr[i] = __ldg(p + i);
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) xr[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        };
        auto publish = [&]() {
            float* x = reinterpret_cast<float*>(xr);
            if (valid && mode != 0) {
#pragma unroll
                for (int k = 0; k < 32; ++k) {
                    const float sc = cf[k], sh = cf[64 + k], al = cf[128 + k];
                    float v = x[k];
                    if (mode == 1) { v = fmaf(v, sc, sh); v = fmaxf(v, 0.f) + al * fminf(v, 0.f); }
                    else { v = fmaxf(v, 0.f) + al * fminf(v, 0.f); v = fmaf(v, sc, sh); }
                    x[k] = v;
                }
            }
            if (a.layer_norm) {                  // uniform branch: every lane takes part in the pair shuffles
                float sum = 0.f;
#pragma unroll
                for (int k = 0; k < 32; ++k) sum += x[k];
                sum += __shfl_xor_sync(0xffffffffu, sum, 1);
                const float mean = sum * (1.f / 64.f);
                float sq = 0.f;
#pragma unroll
                for (int k = 0; k < 32; ++k) { const float d = x[k] - mean; sq += d * d; }
                sq += __shfl_xor_sync(0xffffffffu, sq, 1);
                const float rstd = rsqrtf(sq * (1.f / 64.f) + 1e-5f);
                if (valid) {
#pragma unroll
                    for (int k = 0; k < 32; ++k) x[k] = (x[k] - mean) * rstd * lg[k] + lb[k];
                }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float* v = x + i * 8;
                uint4 hi, lo;
                hi.x = pack_h2(v[0], v[1]); hi.y = pack_h2(v[2], v[3]); hi.z = pack_h2(v[4], v[5]); hi.w = pack_h2(v[6], v[7]);
                lo.x = pack_lo_h2(v[0], v[1], hi.x); lo.y = pack_lo_h2(v[2], v[3], hi.y);
                lo.z = pack_lo_h2(v[4], v[5], hi.z); lo.w = pack_lo_h2(v[6], v[7], hi.w);
                const int off = ((half * 4 + i) ^ (row & 7)) << 4;
                *reinterpret_cast<uint4*>(xrow_hi + off) = hi;
                *reinterpret_cast<uint4*>(xrow_lo + off) = lo;
            }
            fence_proxy_async();
        };
        load(0);
        publish();
        mbar_arrive(&x_ready[s]);
        if (a.T > 1) load(1);
        for (int t = 1; t < a.T; ++t) {
            MBW(&x_free[s], (uint32_t)((t - 1) & 1));    // the sub-batch's MMAs of step t-1 have consumed x_{t-1}
            publish();                                   // x_t
            mbar_arrive(&x_ready[s]);
            if (t + 1 < a.T) load(t + 1);                // in flight for a whole step
        }
    } else {
        // ======================================================================= MMA issuer
        // Convergent issue (every lane runs the code, one elected lane issues).  Per (step, sub-batch) and column half:
        // x slab (fresh) then h slab, three passes each (A_hi B_hi, A_lo B_hi, A_hi B_lo), four K steps per pass; all of it
        // executes under the OTHER sub-batch's cell phase.
        const uint32_t idesc = make_idesc(256);
        const uint32_t a_lo0 = desc_lo(smem_u32(As)), b_lo0 = desc_lo(smem_u32(Bs));
        const bool mdbg = a.dbg != nullptr && blockIdx.x == 0 && lane == 0;
        long long m_wh = 0, m_wx = 0;
        for (int t = 0; t < a.T; ++t) {
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                const long long m0 = mdbg ? clock64() : 0;
                MBW(&h_ready[s], (uint32_t)(t & 1));     // h_{t-1} stored, the accumulator of step t-1 read
                const long long m1 = mdbg ? clock64() : 0;
                MBW(&x_ready[s], (uint32_t)(t & 1));
                if (mdbg) { m_wh += m1 - m0; m_wx += clock64() - m1; }
                tc_fence_after();
                {
                    // N = 256 per MMA: 12 KB of operand reads per 128 tensor-pipe cycles (two N = 128 MMAs read 16 KB, which is the
                    // whole shared-memory bandwidth: measured ~135 cycles each next to the cell warps' own traffic)
                    const uint32_t d = tmem_base + (uint32_t)(s * 256);
#pragma unroll
                    for (int slab = 0; slab < 2; ++slab) {
#pragma unroll
                        for (int pass = 0; pass < 3; ++pass) {
                            const int ahl = pass == 1 ? 1 : 0;
                            const int bhl = pass == 2 ? 1 : 0;
                            const uint32_t aa = a_lo0 + (uint32_t)(((ahl * 2 + slab) * SLAB_BYTES) >> 4);
                            const uint32_t bb = b_lo0 + (uint32_t)(((bhl * 2 + slab) * B_SLAB_BYTES) >> 4);
                            umma_f16_lo_elect_x4(d, aa, bb, idesc, (slab | pass) ? 1u : 0u);
                        }
                    }
                    umma_commit_elect(&acc_full[s]);
                }
                umma_commit_elect(&x_free[s]);
            }
        }
        if (mdbg) { a.dbg[5] = m_wh; a.dbg[7] = m_wx; }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) tmem_dealloc(tmem_base, 512);
}

}  // namespace

bool lstm_pp_supported(const LstmArgs& a) { return a.E == 64 && a.F >= ROWS && a.Wimg != nullptr; }

int launch_lstm_pp(const LstmArgs& a, cudaStream_t st) {
    if (!lstm_pp_supported(a)) return fail("lstm_pp: unsupported shape");
    EAB_TRY(ensure_dynamic_smem(reinterpret_cast<const void*>(lstm_pp_kernel), SMEM_BYTES));
    const int NQ = a.B * a.F;
    ProfScope ps("lstm_umma", 2.0 * NQ * a.T * (64 + H) * 4.0 * H, 4.0 * NQ * a.T * (64 + H), st);
    EAB_CUDA(launch_k(lstm_pp_kernel, dim3((NQ + ROWS - 1) / ROWS), dim3(NTHREADS), (size_t)SMEM_BYTES, st, a));
    EAB_LAUNCH_CHECK("lstm_pp_kernel");
    return 0;
}

}  // namespace eab
