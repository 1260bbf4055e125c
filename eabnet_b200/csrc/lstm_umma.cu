// nn.LSTM(64 -> 64) layer of LSTM_BF (EaBNet.py:591-592, 610-611) on the tensor cores (sm_100a), fp32-grade.
//
// One CTA owns 128 (b,f) sequences for all T steps.  Per step the gate pre-activations
//     G[128 x 256] = x_t W_ih^T + h_{t-1} W_hh^T      (K = 64 + 64)
// are three tcgen05.mma.kind::f16 passes each (hi*hi + lo*hi + hi*lo of an fp16 split, i.e. ~22 mantissa bits) into a
// 256-column TMEM accumulator.  The weight images (128 KB) stay resident in shared memory for the whole sequence,
// the cell state lives in registers, and h_t goes straight back into the swizzled A operand of step t+1.
// Only the recurrent half sits on the T-serial chain: the accumulator is double-buffered (2 x 256 TMEM columns) and the
// input projection of step t+1 is issued right behind the recurrent MMAs of step t, so it executes while the cell warps
// work on step t.  Gate columns are ordered so that BOTH column halves hold 8 hidden units of every cell thread: all 16
// cell warps start on the first half while the tensor core finishes the second.
//
//   warps 0-15  cell warps  : four threads per sequence (16 hidden units each): tcgen05.ld gates, bias, sigmoid/tanh,
//                             c/h update, h_t -> HBM (fp32) and -> A operand (fp16 hi/lo), arrive
//   warps 16-19 x producers : one thread per sequence: prefetch x from HBM two steps ahead, apply the decoder's norm +
//                             PReLU and the head's LayerNorm (thread-local over the 64 channels), split to fp16 hi/lo,
//                             store to the A operand once the input-projection MMAs have released it, arrive
//   warp 20     MMA issuer  : h-part of step t (24 MMAs M128 x N128 x K16, two commits), then x-part of step t+1
#include <cstdlib>

#include "common.cuh"
#include "umma.cuh"

namespace eab {

namespace {

using namespace umma;

constexpr int H = 64;
constexpr int ROWS = 128;
#ifdef EAB_LSTM_SUSPEND
#define MBW mbar_wait
#else
#define MBW mbar_wait_spin
#endif

__device__ __forceinline__ float ex2(float x) {          // one SFU op, ~2 ulp
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// two fp32 operations per instruction (sm_100: FFMA2 / FMUL2 / FADD2); each lane rounds exactly like the scalar form
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
__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// timing experiments (diagnostics builds, EAB_NVCC_EXTRA=-DEAB_LSTM_EXPERIMENT; option lstm_exp; wrong results): bit 1 no copy-out
// stores, 2 no SFU work, 4 one bias load per half instead of 32 (measured: no effect), 8 no fp32 staging stores (-3 %)
#ifdef EAB_LSTM_EXPERIMENT
#define EXP_FLAG(bit) ((a.exp_flags & (bit)) != 0)
#else
#define EXP_FLAG(bit) false
#endif
// 96 of the 128 MMA rows carry sequences: 10 304 sequences (64 x 161) then make 108 CTAs - one wave over 148 SMs with
// 25 % less cell work per CTA than 81 CTAs of 128 (the step time is the cell phase, not the tensor pipe).  A warp may
// only touch TMEM lanes 32 (warp % 4) ... + 32, so warps 0-15 with warp % 4 == 3 would own the empty lane quadrant:
// they are x producers instead, next to warps 16-17.
constexpr int RPC = 96;                                 // sequences per CTA
constexpr int NCELL = 4 * RPC, NXP = 2 * RPC;           // 12 cell warps (4 threads per row), 6 producer warps (2 per row)
constexpr int MMA_WARP = 18;
constexpr int NTHREADS = 19 * 32;
constexpr int SLAB_BYTES = ROWS * 128;                 // one 64-wide fp16 K slab of the A operand
constexpr int A_BYTES = 4 * SLAB_BYTES;                // [hi|lo][x|h]
constexpr int B_SLAB_BYTES = 256 * 128;                // one K slab of the weight image (256 gate rows)
constexpr int B_BYTES = 4 * B_SLAB_BYTES;              // [hi|lo][x|h]
constexpr int MISC_FLOATS = 256 + 2 * 3 * 64 + 2 * 64; // bias, transform coefficients (2 batch items), LayerNorm
constexpr int HS_BYTES = RPC * 256;                    // fp32 staging of h_t for the coalesced copy-out (16-byte chunks, XOR-swizzled)
constexpr int SMEM_BYTES = A_BYTES + B_BYTES + HS_BYTES + MISC_FLOATS * 4 + 64 + 1024;

__device__ __forceinline__ uint8_t* a_slab(uint8_t* A, int hl, int slab) { return A + (hl * 2 + slab) * SLAB_BYTES; }

__global__ void __launch_bounds__(NTHREADS, 1) lstm_umma_kernel(const LstmArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* As = smem;
    uint8_t* Bs = smem + A_BYTES;
    uint8_t* Hs = smem + A_BYTES + B_BYTES;
    float* sbias = reinterpret_cast<float*>(smem + A_BYTES + B_BYTES + HS_BYTES);
    float* coef = sbias + 256;                  // [2][3][64]
    float* lng = coef + 2 * 3 * 64;
    float* lnb = lng + 64;
    uint64_t* bars = reinterpret_cast<uint64_t*>(lnb + 64);
    uint64_t* h_ready = bars;                   // h_{t-1} is in the A operand (all cell threads)
    uint64_t* acc_full = bars + 1;              // [2] gate columns [0,128) / [128,256) of the current step complete
    uint64_t* x_ready = bars + 3;               // x_t is in the A operand (all producer threads)
    uint64_t* x_free = bars + 4;                // the input-projection MMAs have consumed the x operand
    uint64_t* hs_free = bars + 5;               // the producers have copied the staged h_t out to HBM
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;
    const int NQ = a.B * a.F;
    const int rows_act = a.rows_per_cta > 0 ? a.rows_per_cta : RPC;      // (experiment: fewer live rows per CTA, more CTAs)
    const int q0 = blockIdx.x * rows_act;
    const int b0 = q0 / a.F;

    if (tid == 0) {
        mbar_init(h_ready, NCELL);
        mbar_init(&acc_full[0], 1);
        mbar_init(&acc_full[1], 1);
        mbar_init(x_ready, NXP);
        mbar_init(x_free, 1);
        mbar_init(hs_free, NXP);
        fence_barrier_init();
    }
    if (warp == MMA_WARP) tmem_alloc(tmem_slot, 512);
    {
        const uint4* src = reinterpret_cast<const uint4*>(a.Wimg);
        uint4* dst = reinterpret_cast<uint4*>(Bs);
        for (int i = tid; i < B_BYTES / 16; i += NTHREADS) dst[i] = __ldg(src + i);     // static weights: before the wait
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

    if (warp < 16 && (warp & 3) != 3) {
        // ======================================================================= cell warps
        // thread = (sequence row, quarter of the hidden units): units quarter*16 + u*8 + e, gate columns
        // u*128 + quarter*32 + gate*8 + e of accumulator buffer t & 1
        const int quad = warp & 3;              // TMEM lane quadrant
        const int qtr = warp >> 2;              // which 16 hidden units
        const int row = quad * 32 + lane;
        const int q = q0 + row;
        const bool valid = q < NQ && row < rows_act;
        const int bq = valid ? q / a.F : 0;
        const int fq = valid ? q - bq * a.F : 0;
        uint8_t* hs_row = Hs + row * 256;
        const uint32_t taddr0 = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(qtr * 32);
        uint8_t* hrow_hi = a_slab(As, 0, 1) + row * 128;
        uint8_t* hrow_lo = a_slab(As, 1, 1) + row * 128;
        float c[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) c[i] = 0.f;
        mbar_arrive(h_ready);                   // h_{-1} = 0 is in place
        const bool dbg_on = a.dbg != nullptr && blockIdx.x == 0 && tid == 0;
        long long t_wait = 0, t_cell = 0, t_fence = 0, t_w1 = 0;
        const long long t_start = dbg_on ? clock64() : 0;
        constexpr float L2E = 1.4426950408889634f;
        for (int t = 0; t < a.T; ++t) {
            const long long c0 = dbg_on ? clock64() : 0;
            long long c1 = 0;
            const uint32_t taddr = taddr0 + (uint32_t)((t & 1) * 256);
            uint4 hst_hi[2], hst_lo[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {       // 8 hidden units at a time: column half u
                const long long w0 = dbg_on ? clock64() : 0;
                MBW(&acc_full[u], (uint32_t)(t & 1));
                if (dbg_on) { if (u == 0) c1 = clock64(); else t_w1 += clock64() - w0; }
                tc_fence_after();
                uint32_t gi[8], gf[8], gg[8], go[8];
                tmem_ld8_nowait(taddr + u * 128 + 0, gi);
                tmem_ld8_nowait(taddr + u * 128 + 8, gf);
                tmem_ld8_nowait(taddr + u * 128 + 16, gg);
                tmem_ld8_nowait(taddr + u * 128 + 24, go);
                tmem_wait_ld();
                float hv[8];
                const float* bi = sbias + (EXP_FLAG(4) ? 0 : u * 128 + qtr * 32);   // biases pre-scaled by -log2(e) (i,f,o) / -2 log2(e) (g)
                const float bconst = EXP_FLAG(4) ? bi[0] : 0.f;          // experiment: one bias load per half instead of 32
                // 7 SFU ops per unit instead of 10: the sigmoid / tanh quotients share their reciprocals
                //   c' = sig(f) c + sig(i) tanh(g) = [c (1+Ei)(1+Eg) + (1-Eg)(1+Ef)] / [(1+Ei)(1+Ef)(1+Eg)]
                //   h  = sig(o) tanh(c')          = (1-Ec) / [(1+Eo)(1+Ec)]
                // with Ei = e^-i, Ef = e^-f, Eg = e^-2g, Ec = e^-2c', Eo = e^-o.  The exponents are clamped from above
                // only (2^40: the functions are saturated to < 1e-12 there and the triple product stays < 2^127).
                // Two hidden units per instruction wherever the ISA has a packed fp32 form (the busiest scheduler of this
                // kernel issues on 80 % of its cycles - ncu, profiles/r02_v3_lstm_full.txt - so the cell phase is bound by
                // instruction issue as much as by the SFU): 43 instructions per unit PAIR instead of 38 per unit.
                const F2 one = {1.f, 1.f}, nl = {-L2E, -L2E}, nl2 = {-2.f * L2E, -2.f * L2E};
#pragma unroll
                for (int e = 0; e < 8; e += 2) {
                    if (EXP_FLAG(2)) {          // experiment: no SFU work at all
#pragma unroll
                        for (int q = 0; q < 2; ++q) {
                            const float cn = fmaf(c[u * 8 + e + q], __uint_as_float(gf[e + q]), __uint_as_float(gi[e + q]) * __uint_as_float(gg[e + q])) * 0.25f;
                            c[u * 8 + e + q] = cn;
                            hv[e + q] = cn * __uint_as_float(go[e + q]) * 0.01f;
                        }
                        continue;
                    }
                    const F2 ti = fma2({__uint_as_float(gi[e]), __uint_as_float(gi[e + 1])}, nl, EXP_FLAG(4) ? F2{bconst, bconst} : F2{bi[e], bi[e + 1]});
                    const F2 tf = fma2({__uint_as_float(gf[e]), __uint_as_float(gf[e + 1])}, nl, EXP_FLAG(4) ? F2{bconst, bconst} : F2{bi[8 + e], bi[8 + e + 1]});
                    const F2 tg = fma2({__uint_as_float(gg[e]), __uint_as_float(gg[e + 1])}, nl2, EXP_FLAG(4) ? F2{bconst, bconst} : F2{bi[16 + e], bi[16 + e + 1]});
                    const F2 to = fma2({__uint_as_float(go[e]), __uint_as_float(go[e + 1])}, nl, EXP_FLAG(4) ? F2{bconst, bconst} : F2{bi[24 + e], bi[24 + e + 1]});
                    const F2 Ei = {ex2(fminf(ti.x, 40.f)), ex2(fminf(ti.y, 40.f))};
                    const F2 Ef = {ex2(fminf(tf.x, 40.f)), ex2(fminf(tf.y, 40.f))};
                    const F2 Eg = {ex2(fminf(tg.x, 40.f)), ex2(fminf(tg.y, 40.f))};
                    const F2 Eo = {ex2(fminf(to.x, 40.f)), ex2(fminf(to.y, 40.f))};
                    const F2 A = add2(one, Ei), Bf = add2(one, Ef), G = add2(one, Eg);
                    const F2 AG = mul2(A, G);
                    const F2 num = fma2({c[u * 8 + e], c[u * 8 + e + 1]}, AG, mul2(sub2(one, Eg), Bf));
                    const F2 den = mul2(AG, Bf);
                    const F2 cn = mul2(num, {rcp_approx(den.x), rcp_approx(den.y)});
                    c[u * 8 + e] = cn.x; c[u * 8 + e + 1] = cn.y;
                    const F2 tc = mul2(cn, nl2);
                    const F2 Ec = {ex2(fminf(tc.x, 40.f)), ex2(fminf(tc.y, 40.f))};
                    const F2 hden = mul2(add2(one, Eo), add2(one, Ec));
                    const F2 h = mul2(sub2(one, Ec), {rcp_approx(hden.x), rcp_approx(hden.y)});
                    hv[e] = h.x; hv[e + 1] = h.y;
                }
                if (!EXP_FLAG(8)) {   // fp32 copy for HBM: chunk c of a row lives at (c ^ (row & 7)) * 16 (conflict-free both ways)
                    if (u == 0 && t > 0) MBW(hs_free, (uint32_t)((t - 1) & 1));    // h_{t-1} has left the staging tile
                    const int c0 = qtr * 4 + u * 2;
                    *reinterpret_cast<float4*>(hs_row + (((c0 + 0) ^ (row & 7)) << 4)) = make_float4(hv[0], hv[1], hv[2], hv[3]);
                    *reinterpret_cast<float4*>(hs_row + (((c0 + 1) ^ (row & 7)) << 4)) = make_float4(hv[4], hv[5], hv[6], hv[7]);
                }
                uint4 hi, lo;
                hi.x = pack_h2(hv[0], hv[1]); hi.y = pack_h2(hv[2], hv[3]); hi.z = pack_h2(hv[4], hv[5]); hi.w = pack_h2(hv[6], hv[7]);
                lo.x = pack_lo_h2(hv[0], hv[1], hi.x); lo.y = pack_lo_h2(hv[2], hv[3], hi.y);
                lo.z = pack_lo_h2(hv[4], hv[5], hi.z); lo.w = pack_lo_h2(hv[6], hv[7], hi.w);
                hst_hi[u] = hi;
                hst_lo[u] = lo;
            }
            // h_t overwrites the operand the recurrent MMAs of this step read: both halves have completed (waited above)
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int chunk = ((qtr * 2 + u) ^ (row & 7)) << 4;
                *reinterpret_cast<uint4*>(hrow_hi + chunk) = hst_hi[u];
                *reinterpret_cast<uint4*>(hrow_lo + chunk) = hst_lo[u];
            }
            const long long c2 = dbg_on ? clock64() : 0;
            tc_fence_before();
            fence_proxy_async();
            mbar_arrive(h_ready);               // h_t: operand for step t+1 and staged for the copy-out
            if (dbg_on) { t_wait += c1 - c0; t_cell += c2 - c1; t_fence += clock64() - c2; }
        }
        if (dbg_on) { a.dbg[0] = clock64() - t_start; a.dbg[1] = t_wait; a.dbg[2] = t_cell; a.dbg[3] = t_fence; a.dbg[4] = a.T; a.dbg[6] = t_w1; }
    } else if (warp < MMA_WARP) {
        // ======================================================================= x producers (two threads per row)
        const int pidx = warp < 16 ? (warp >> 2) : warp - 12;       // 0..5
        const int row = pidx * 16 + (lane >> 1);
        const int half = lane & 1;                                   // channels half*32 .. +32
        const int q = q0 + row;
        const bool valid = q < NQ && row < rows_act;
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
        // copy-out role: 16-byte chunk g = i*NXP + ptid of the staged h tile (row = g >> 4), 512 contiguous bytes per warp
        const int ptid = pidx * 32 + lane;
        // a CTA straddles at most two batch items: row q lives at q*H floats, plus (T-1)*F*H per batch item before it
        const long long jump = (long long)(a.T - 1) * a.F * H;
        const int q_next_b = (b0 + 1) * a.F;
        const size_t ostep = (size_t)a.F * H;
        auto copy_out = [&](int t) {            // h_t: staged by the cell warps, phase t+1 of h_ready
            MBW(h_ready, (uint32_t)((t + 1) & 1));
            float* base = a.out + (long long)b0 * jump + (size_t)t * ostep;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int g = i * NXP + ptid;
                const int r = g >> 4, ch = g & 15;
                const int qq = q0 + r;
                const float4 v = *reinterpret_cast<const float4*>(Hs + r * 256 + ((ch ^ (r & 7)) << 4));
                if (qq < NQ && r < rows_act && !EXP_FLAG(1))
                    *reinterpret_cast<float4*>(base + (long long)qq * H + ch * 4 + (qq >= q_next_b ? jump : 0)) = v;
            }
            mbar_arrive(hs_free);
        };
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
                    const float s = cf[k], h = cf[64 + k], al = cf[128 + k];
                    float v = x[k];
                    if (mode == 1) { v = fmaf(v, s, h); v = fmaxf(v, 0.f) + al * fminf(v, 0.f); }
                    else { v = fmaxf(v, 0.f) + al * fminf(v, 0.f); v = fmaf(v, s, h); }
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
        mbar_arrive(x_ready);
        if (a.T > 1) load(1);
        for (int t = 1; t < a.T; ++t) {
            MBW(x_free, (uint32_t)((t - 1) & 1));        // the input-projection MMAs of step t-1 have consumed x_{t-1}
            publish();                                   // x_t
            mbar_arrive(x_ready);
            if (t + 1 < a.T) load(t + 1);                // in flight for a whole step
            copy_out(t - 1);
        }
        copy_out(a.T - 1);
    } else {
        // ======================================================================= MMA issuer
        const uint32_t idesc = make_idesc(128);
        const bool dbg_on = a.dbg != nullptr && blockIdx.x == 0 && lane == 0;
        long long t_wait = 0, t_wx = 0;
        // one K = 64 operand slab (x: slab 0, h: slab 1) against the matching weight slab, both column halves,
        // three passes per half: A_hi B_hi, A_lo B_hi, A_hi B_lo
        // Convergent issue (every lane runs the code, one elected lane issues): descriptor low words are a base plus
        // compile-time offsets, a K step adds 2 (see umma.cuh).  The recurrent MMAs are ON the serial chain of the step.
        const uint32_t a_lo0 = desc_lo(smem_u32(As)), b_lo0 = desc_lo(smem_u32(Bs));
        auto issue = [&](int slab, int buf, bool fresh, bool commit_halves) {
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
#pragma unroll
                for (int pass = 0; pass < 3; ++pass) {
                    const int ahl = pass == 1 ? 1 : 0;
                    const int bhl = pass == 2 ? 1 : 0;
                    const uint32_t aa = a_lo0 + (uint32_t)(((ahl * 2 + slab) * SLAB_BYTES) >> 4);
                    const uint32_t bb = b_lo0 + (uint32_t)(((bhl * 2 + slab) * B_SLAB_BYTES + hf * 128 * 128) >> 4);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_f16_lo_elect(tmem_base + buf * 256 + hf * 128, aa + 2 * k, bb + 2 * k, idesc, (!fresh || (pass | k)) ? 1u : 0u);
                }
                if (commit_halves) umma_commit_elect(&acc_full[hf]);
            }
        };
        MBW(x_ready, 0u);
        tc_fence_after();
        issue(0, 0, true, false);
        umma_commit_elect(x_free);
        for (int t = 0; t < a.T; ++t) {
            const long long c0 = dbg_on ? clock64() : 0;
            MBW(h_ready, (uint32_t)(t & 1));
            if (dbg_on) t_wait += clock64() - c0;
            tc_fence_after();
            issue(1, t & 1, false, true);                            // recurrent half: the only MMAs on the serial chain
            if (t + 1 < a.T) {
                const long long c2 = dbg_on ? clock64() : 0;
                MBW(x_ready, (uint32_t)((t + 1) & 1));
                if (dbg_on) t_wx += clock64() - c2;
                tc_fence_after();
                issue(0, (t + 1) & 1, true, false);                  // runs under the cell phase of step t
                umma_commit_elect(x_free);
            }
        }
        if (dbg_on) { a.dbg[5] = t_wait; a.dbg[7] = t_wx; }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) tmem_dealloc(tmem_base, 512);
}

}  // namespace

bool lstm_umma_supported(const LstmArgs& a) { return a.E == 64 && a.F >= RPC && a.Wimg != nullptr; }

int launch_lstm_umma(const LstmArgs& a, cudaStream_t st) {
    if (!lstm_umma_supported(a)) return fail("lstm_umma: unsupported shape");
    EAB_TRY(ensure_dynamic_smem(reinterpret_cast<const void*>(lstm_umma_kernel), SMEM_BYTES));
    const int NQ = a.B * a.F;
    ProfScope ps("lstm_umma", 2.0 * NQ * a.T * (64 + H) * 4.0 * H, 4.0 * NQ * a.T * (64 + H), st);
    static const int rows_env = getenv("EAB_LSTM_ROWS") ? atoi(getenv("EAB_LSTM_ROWS")) : 0;      // diagnostics: 1..96 live rows per CTA
    LstmArgs b = a;
    b.rows_per_cta = (rows_env >= 1 && rows_env <= RPC) ? rows_env : 0;
    const int rpc = b.rows_per_cta ? b.rows_per_cta : RPC;
    EAB_CUDA(launch_k(lstm_umma_kernel, dim3((NQ + rpc - 1) / rpc), dim3(NTHREADS), (size_t)SMEM_BYTES, st, b));
    EAB_LAUNCH_CHECK("lstm_umma_kernel");
    return 0;
}

}  // namespace eab
