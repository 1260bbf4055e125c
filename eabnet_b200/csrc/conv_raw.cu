// The 2-D conv layer as ONE kernel (sm_100a): no normalise / stage pass, no fp16 plane images in HBM.
//
// A layer's input is stored RAW (what its producer wrote) next to the per-(b,c) InstanceNorm sums; the reference order is
// conv -> norm -> PReLU (EaBNet.py:402-405, 426-429, 684-686), so the consumer has to apply norm + PReLU before its GEMM.
// Round 1 did that in a stand-alone stage_kernel that wrote fp16 operand planes to HBM (26 % of the step, 2.25x the
// algorithmic traffic of the conv stack).  Here the operand is built on chip:
//
//   warp 0        loader: per group of 32 operand rows of a tile, the raw fp32 rows it needs are ONE contiguous (t, fi)
//                 range of each source tensor -> one cp.async.bulk per (source, addend) into a shared-memory ring
//   warps 12..    transform: ring -> registers -> gamma (x - mean) rstd + beta -> PReLU (+ second addend of a lazy residual
//                 sum) -> fp16 hi (/ lo for the 3-pass layers) -> 128-byte-swizzled operand planes in shared memory
//                 (column-parity planes of a stride-2 conv, literal zeros for pad rows / columns), fence.proxy.async,
//                 mbarrier
//   warp 1        MMA issuer: taps are row-shifted UMMA descriptors into the planes (as conv_tma), accumulators in TMEM
//   warp 2        weight loader (resident images, or a ring when they do not fit)
//   warps 4-11    epilogue: tcgen05.ld -> bias / gate -> 32-byte stores of the RAW output + per-thread running sums for
//                 the output's InstanceNorm statistics
//
// The two output parities of a transposed conv (EaBNet.py:410-431, 463-490) are two variants of the same launch: they
// share the operand planes (read + normalised once) and use their own taps, weights and accumulator columns.
// Arithmetic is the stage_kernel + conv_tma_kernel arithmetic, operation for operation: results are bit-identical.
#include <cuda_fp16.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "common.cuh"
#include "umma.cuh"

namespace eab {

namespace {

using namespace umma;

constexpr int TM = 128;
constexpr int NSTAGE_MAX = 8;
constexpr int NSB_MAX = 8;
constexpr int NCTRL = 128;                       // warps 0-3: loader, MMA issuer, weight loader, spare
constexpr int NEPI = 256;                        // warps 4-11
constexpr int SMEM_LIMIT = 227 * 1024;
constexpr int kTail = 8;                         // operand rows the last group of a tile may hold beyond G (R = 129 -> 2 groups, not 3)

struct RawPlan {
    int R, plane_bytes, buf_bytes, b_stage_bytes, w_images, ngroups;
    int b_off, ring_off, bias_off, coef_off, bar_off, total;
};

__host__ __device__ inline RawPlan make_raw_plan(const RawConvArgs& a) {
    RawPlan p;
    p.R = TM + a.back + a.fwd;
    p.plane_bytes = ((p.R + 7) & ~7) * 128;
    const int npb = a.npass == 3 ? 2 : 1;
    p.buf_bytes = a.nplanes * a.nslab * npb * p.plane_bytes;
    p.b_stage_bytes = a.N * 128;
    p.w_images = 0;
    for (int v = 0; v < a.nvar; ++v) p.w_images += a.ntaps[v] * a.nslab * npb;
    p.ngroups = p.R <= a.G + kTail ? 1 : (p.R - kTail + a.G - 1) / a.G;      // the last group also takes a tail of <= kTail rows
    p.b_off = a.nbuf * p.buf_bytes;
    p.ring_off = p.b_off + (a.resident ? p.w_images : a.nsb) * p.b_stage_bytes;
    p.bias_off = p.ring_off + a.nstage * a.stage_bytes;
    p.coef_off = p.bias_off + a.N * 4;
    p.bar_off = (p.coef_off + a.nslab * 2 * 192 * 4 + 15) & ~15;              // coefficients [slab][addend][s | h | alpha][64]
    p.total = p.bar_off + 512 + 1024;                    // barriers + slack for the 1024-byte alignment of the base
    return p;
}

// raw rows (units of one 64-channel row of a source tensor, relative to the batch item) that operand rows [G j, G j + G) of the
// tile at padded row `row0` read: [g_lo, g_lo + n).  Operand row rho <-> padded row r = row0 - back + rho = t P + col;
// plane p holds input column col * col_stride + col_off[p] (zeros where that is >= Fin or r is outside [0, T P)).
__device__ __forceinline__ void group_range(const RawConvArgs& a, int R, int ngroups, int rows_per_b, int row0, int j, int& g_lo, int& n) {
    int r_a = row0 - a.back + j * a.G;
    int r_b = j == ngroups - 1 ? row0 - a.back + R - 1 : r_a + a.G - 1;        // the last group runs to the end of the tile's rows
    r_a = max(r_a, 0);
    r_b = min(r_b, rows_per_b - 1);
    g_lo = 0; n = 0;
    if (r_a > r_b) return;
    const int t_a = a.P == 1 ? r_a : (int)__umulhi((unsigned)r_a, a.p_magic);
    const int c_a = r_a - t_a * a.P;
    const int t_b = a.P == 1 ? r_b : (int)__umulhi((unsigned)r_b, a.p_magic);
    const int c_b = r_b - t_b * a.P;
    g_lo = t_a * a.Fin + min(c_a * a.col_stride, a.Fin);
    n = t_b * a.Fin + min((c_b + 1) * a.col_stride, a.Fin) - g_lo;
}

__device__ __forceinline__ float4 lds128(const uint8_t* p) { return *reinterpret_cast<const float4*>(p); }

// two fp32 operations per instruction (sm_100: FFMA2 / FMUL2 / FADD2); each lane rounds exactly like the scalar form
__device__ __forceinline__ void ffma2(float& dx, float& dy, float ax, float ay, float bx, float by, float cx, float cy) {
    uint64_t ra, rb, rc, rd;
    asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(ax), "f"(ay));
    asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(bx), "f"(by));
    asm("mov.b64 %0, {%1, %2};" : "=l"(rc) : "f"(cx), "f"(cy));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(dx), "=f"(dy) : "l"(rd));
}
__device__ __forceinline__ void fmul2(float& dx, float& dy, float ax, float ay, float bx, float by) {
    uint64_t ra, rb, rd;
    asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(ax), "f"(ay));
    asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(bx), "f"(by));
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(dx), "=f"(dy) : "l"(rd));
}
__device__ __forceinline__ void fsub2(float& dx, float& dy, float ax, float ay, float bx, float by) {
    uint64_t ra, rb, rd;
    asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(ax), "f"(ay));
    asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(bx), "f"(by));
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(dx), "=f"(dy) : "l"(rd));
}

// norm + PReLU of four channels (2-D blocks: conv -> norm -> PReLU, EaBNet.py:402-405).  FAST: every slope of the launch
// lies in [0, 1], where max(z, 0) + a min(z, 0) == max(z, a z) bit for bit (z > 0: a z <= z; z < 0: a z >= z; the sum's other
// term is an exact zero).  Each channel goes through the operations stage_kernel applies, in the same order and rounding
// (fma, then the PReLU), two channels per instruction where the ISA has a packed form.  A source without norm / PReLU runs
// with the identity coefficients (1, 0, 1), which reproduce x exactly.  The variant is a template parameter: as a run-time
// switch the compiler re-dispatched on it around every quad (branches instead of four independent chains).
template <bool FAST>
__device__ __forceinline__ void xf4(float4& v, const float4& s, const float4& h, const float4& al) {
    float z0, z1, z2, z3;
    ffma2(z0, z1, v.x, v.y, s.x, s.y, h.x, h.y);
    ffma2(z2, z3, v.z, v.w, s.z, s.w, h.z, h.w);
    if (FAST) {
        float a0, a1, a2, a3;
        fmul2(a0, a1, z0, z1, al.x, al.y);
        fmul2(a2, a3, z2, z3, al.z, al.w);
        v.x = fmaxf(z0, a0); v.y = fmaxf(z1, a1); v.z = fmaxf(z2, a2); v.w = fmaxf(z3, a3);
    } else {
        ffma2(v.x, v.y, al.x, al.y, fminf(z0, 0.f), fminf(z1, 0.f), fmaxf(z0, 0.f), fmaxf(z1, 0.f));
        ffma2(v.z, v.w, al.z, al.w, fminf(z2, 0.f), fminf(z3, 0.f), fmaxf(z2, 0.f), fmaxf(z3, 0.f));
    }
}

// fp16 hi pair of two floats and the fp16-rounded residual (3-pass split), residual subtraction as one packed instruction
__device__ __forceinline__ uint32_t pack_lo2(float a, float b, uint32_t hi) {
    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&hi));
    float dx, dy;
    fsub2(dx, dy, a, b, f.x, f.y);
    return pack_h2(dx, dy);
}

// One batch of NI operand chunks (8 channels each) of ONE slab for one thread: every raw load is issued before any
// arithmetic and the NI chains are independent.  soff = byte offset of the item's raw row in the stage (a valid address for
// pad items too, whose result is replaced by zeros), dst = its 16-byte slot in the hi image (lo image plane_bytes further).
// The thread's two 4-channel quads are stored as two 8-byte halves (the quad it loaded first is the chunk's half qa).
// cf = [addend][s | h | alpha][64].  `act` = the lanes of the warp that call this together (for the all-valid vote).
template <int NI, bool DUAL, bool LO, bool FAST>
__device__ __forceinline__ void xf_items(const uint8_t* st0, const uint8_t* st1, const int (&soff)[NI], const bool (&ok)[NI],
                                         uint8_t* const (&dst)[NI], const float* cf, int qa, int chA, int chB, int plane_bytes,
                                         unsigned act) {
    float4 vA[NI], vB[NI];
#pragma unroll
    for (int i = 0; i < NI; ++i) {
        vA[i] = lds128(st0 + soff[i] + qa * 16);
        vB[i] = lds128(st0 + soff[i] + (qa ^ 1) * 16);
    }
    {
        const float4 sA = *reinterpret_cast<const float4*>(cf + chA), hA = *reinterpret_cast<const float4*>(cf + 64 + chA),
                     aA = *reinterpret_cast<const float4*>(cf + 128 + chA);
        const float4 sB = *reinterpret_cast<const float4*>(cf + chB), hB = *reinterpret_cast<const float4*>(cf + 64 + chB),
                     aB = *reinterpret_cast<const float4*>(cf + 128 + chB);
#pragma unroll
        for (int i = 0; i < NI; ++i) { xf4<FAST>(vA[i], sA, hA, aA); xf4<FAST>(vB[i], sB, hB, aB); }
    }
    if (DUAL) {                                  // + the second addend of a module's lazy residual sum (EaBNet.py:386)
        float4 wA[NI], wB[NI];
#pragma unroll
        for (int i = 0; i < NI; ++i) {
            wA[i] = lds128(st1 + soff[i] + qa * 16);
            wB[i] = lds128(st1 + soff[i] + (qa ^ 1) * 16);
        }
        {
            const float* c1 = cf + 192;
            const float4 sA = *reinterpret_cast<const float4*>(c1 + chA), hA = *reinterpret_cast<const float4*>(c1 + 64 + chA),
                         aA = *reinterpret_cast<const float4*>(c1 + 128 + chA);
            const float4 sB = *reinterpret_cast<const float4*>(c1 + chB), hB = *reinterpret_cast<const float4*>(c1 + 64 + chB),
                         aB = *reinterpret_cast<const float4*>(c1 + 128 + chB);
#pragma unroll
            for (int i = 0; i < NI; ++i) { xf4<FAST>(wA[i], sA, hA, aA); xf4<FAST>(wB[i], sB, hB, aB); }
        }
#pragma unroll
        for (int i = 0; i < NI; ++i) {
            vA[i].x += wA[i].x; vA[i].y += wA[i].y; vA[i].z += wA[i].z; vA[i].w += wA[i].w;
            vB[i].x += wB[i].x; vB[i].y += wB[i].y; vB[i].z += wB[i].z; vB[i].w += wB[i].w;
        }
    }
    bool all_ok = true;
#pragma unroll
    for (int i = 0; i < NI; ++i) all_ok = all_ok && ok[i];
    if (!__all_sync(act, all_ok)) {              // rare (pad rows / columns): literal zeros
#pragma unroll
        for (int i = 0; i < NI; ++i)
            if (!ok[i]) { vA[i] = make_float4(0.f, 0.f, 0.f, 0.f); vB[i] = vA[i]; }
    }
#pragma unroll
    for (int i = 0; i < NI; ++i) {
        const uint32_t hA0 = pack_h2(vA[i].x, vA[i].y), hA1 = pack_h2(vA[i].z, vA[i].w);
        const uint32_t hB0 = pack_h2(vB[i].x, vB[i].y), hB1 = pack_h2(vB[i].z, vB[i].w);
        *reinterpret_cast<uint2*>(dst[i] + qa * 8) = make_uint2(hA0, hA1);
        *reinterpret_cast<uint2*>(dst[i] + (qa ^ 1) * 8) = make_uint2(hB0, hB1);
        if (LO) {
            const uint32_t lA0 = pack_lo2(vA[i].x, vA[i].y, hA0), lA1 = pack_lo2(vA[i].z, vA[i].w, hA1);
            const uint32_t lB0 = pack_lo2(vB[i].x, vB[i].y, hB0), lB1 = pack_lo2(vB[i].z, vB[i].w, hB1);
            *reinterpret_cast<uint2*>(dst[i] + plane_bytes + qa * 8) = make_uint2(lA0, lA1);
            *reinterpret_cast<uint2*>(dst[i] + plane_bytes + (qa ^ 1) * 8) = make_uint2(lB0, lB1);
        }
    }
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

// setmaxnreg moves registers inside the CTA's OWN allocation (640 threads x 96 at launch = 61 440): the increases must be
// covered by the decreases of the same CTA (128 x 48 + 256 x 120 + 256 x 96 = 61 440), or the increase spins forever.
// mbarrier waits of this kernel: polling (try_wait without a suspend-time hint; the hinted form compiles to TRYWAIT +
// NANOSLEEP.SYNCS with a wake-up latency paid on every handshake), with a 128 ns back-off for the roles that have slack.
#define RWAIT mbar_wait_spin                         // critical path: the MMA issuer, the transform warps' data waits
#define RWAIT_IDLE(bar, par) mbar_wait_backoff(bar, par, 128u)      // roles with slack: epilogue, loaders, operand-buffer reuse
#ifdef EAB_RAW_DEBUG
constexpr bool kDbg = true;
#else
constexpr bool kDbg = false;
#endif

template <int REGS> __device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS)); }
template <int REGS> __device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS)); }

// NTW transform warps (8: 640 threads, the register file is re-partitioned between the roles with setmaxnreg)
// LO: 3-pass layers (fp16 hi + lo operand images); FAST: every PReLU slope of the inputs in [0, 1] (xf4)
template <int NTW, bool LO, bool FAST>
__global__ void __launch_bounds__(NCTRL + NEPI + NTW * 32, 1) conv_raw_kernel(const RawConvArgs a) {
    constexpr int NTHREADS = NCTRL + NEPI + NTW * 32;
    constexpr int TR0 = NCTRL + NEPI;                        // first transform thread
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const RawPlan pl = make_raw_plan(a);
    uint8_t* opnd = smem;
    uint8_t* Bs = smem + pl.b_off;
    uint8_t* ring = smem + pl.ring_off;
    float* sbias = reinterpret_cast<float*>(smem + pl.bias_off);
    float* coef = reinterpret_cast<float*>(smem + pl.coef_off);            // [slab][addend][s | h | alpha][64]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + pl.bar_off);
    uint64_t* raw_full = bars;              // [NSTAGE_MAX]
    uint64_t* raw_empty = bars + 8;         // [NSTAGE_MAX]
    uint64_t* opnd_full = bars + 16;        // [2]
    uint64_t* opnd_empty = bars + 18;       // [2]
    uint64_t* b_full = bars + 20;           // [NSB_MAX]
    uint64_t* b_empty = bars + 28;          // [NSB_MAX]
    uint64_t* acc_full = bars + 36;         // [2 accumulator sets][2 variants]
    uint64_t* acc_empty = bars + 40;        // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 42);

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;
    const int acc_cols = a.nvar * a.N;                       // TMEM columns of one accumulator set
    const uint32_t tmem_cols = 2 * acc_cols <= 128 ? 128u : (2 * acc_cols <= 256 ? 256u : 512u);
    const int npb = a.npass == 3 ? 2 : 1;

    if (tid == 0) {
        for (int i = 0; i < NSTAGE_MAX; ++i) { mbar_init(&raw_full[i], 1); mbar_init(&raw_empty[i], NTW); }
        for (int i = 0; i < 2; ++i) { mbar_init(&opnd_full[i], NTW); mbar_init(&opnd_empty[i], 1); }
        for (int i = 0; i < NSB_MAX; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
        for (int i = 0; i < 4; ++i) mbar_init(&acc_full[i], 1);
        for (int i = 0; i < 2; ++i) mbar_init(&acc_empty[i], NEPI);
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

    // ----------------------------------------------------------------------------------------------- transform of one tile
    // thread = (row rr of each 32-row block of the group, 16-byte fp16 chunk c8 = 8 channels); a ring stage (group, slab) gives
    // the thread NI = (G / 32) x nplanes chunks, transformed as one batch (xf_items).  The two 16-byte halves of a chunk's 32
    // raw bytes are read in an order that depends on c8 (c8 >= 4: upper half first) so that the eight lanes of a quarter-warp
    // touch all 32 banks in both loads.
    // (Tried: the epilogue warps - idle most of the time on the non-gated layers - transforming every second or third stage
    // of a tile.  Bit-exact, but SLOWER, 14.6 ms of conv_raw per step against 9.9: their 64 statistics registers go to local
    // memory around the transform code and their epilogue, which frees the accumulator the next MMA waits for, runs late.)
    struct XfState { int stage; uint32_t sphase; int cur_b; long long t_wo, t_wr, t_cf; };
    const int pcols0 = a.plane_cols[0], pcols1 = a.plane_cols[1], coff0 = a.col_off[0], coff1 = a.col_off[1];
    const bool np2 = a.nplanes == 2;
    auto xf_tile = [&](const int tile, const int ttid, XfState& xs, const bool dbg_on) {
        const int c8 = ttid & 7;
        const int rr = ttid >> 3;                            // 0..31
        const int qa = (c8 >> 2) & 1;                        // which half this thread loads first
        const int chA = c8 * 8 + qa * 4, chB = c8 * 8 + (qa ^ 1) * 4;
        const int rpt = a.G >> 5;                            // rows per thread per group: 1, 2 or 4
        int& stage = xs.stage;
        uint32_t& sphase = xs.sphase;
        int& cur_b = xs.cur_b;
        long long& t_wo = xs.t_wo; long long& t_wr = xs.t_wr; long long& t_cf = xs.t_cf;
        {
            const int ord = tile - tile_begin;
            const int buf = ord % a.nbuf;
            const uint32_t bphase = (uint32_t)((ord / a.nbuf) & 1);
            const int b = tile / a.tiles_per_b;
            const int row0 = (tile - b * a.tiles_per_b) * TM;
            if (b != cur_b) {
                const long long w0 = dbg_on ? clock64() : 0;
                named_bar_sync(1, 256);
                for (int i = ttid; i < a.nslab * 2 * 64; i += 256) {
                    const int s = i >> 7, ad = (i >> 6) & 1, c = i & 63;
                    float cs = 1.f, ch = 0.f, ca = 1.f;
                    if (ad == 0) { xform_coeffs(a.xf0[s], b, 64, c, cs, ch, ca); if (!a.xf0[s].prelu) ca = 1.f; }
                    else if (a.x1[s]) { xform_coeffs(a.xf1[s], b, 64, c, cs, ch, ca); if (!a.xf1[s].prelu) ca = 1.f; }
                    float* cf = coef + (s * 2 + ad) * 192;
                    cf[c] = cs; cf[64 + c] = ch; cf[128 + c] = ca;
                }
                named_bar_sync(1, 256);
                cur_b = b;
                if (dbg_on) t_cf += clock64() - w0;
            }
            const long long w1 = dbg_on ? clock64() : 0;
            RWAIT_IDLE(&opnd_empty[buf], bphase ^ 1);
            if (dbg_on) t_wo += clock64() - w1;
            uint8_t* obuf = opnd + (size_t)buf * pl.buf_bytes;
            // This thread's operand rows of the tile are rho = rr, rr + 32, rr + 64, ... across the groups; their (frame,
            // column) advance by (q32, r32) = divmod(32, P) per row and the groups' first rows by divmod(G, P): two divisions
            // per tile.  (A per-item set-up - division, indexed constant loads, shuffles between the lanes of a row - cost more
            // than the transform arithmetic: ncu, profiles/r02_raw_l48_*.)
            int t_r, c_r, t_a, c_a;                          // padded row r = row0 - back + rho = t P + c ; the group's first row
            {
                const unsigned u_a = (unsigned)(row0 - a.back + 4 * a.P);      // >= 0: back <= 4 P (launcher)
                const unsigned q_a = a.P == 1 ? u_a : __umulhi(u_a, a.p_magic);
                t_a = (int)q_a - 4; c_a = (int)(u_a - q_a * (unsigned)a.P);
                const unsigned u_r = u_a + (unsigned)rr;
                const unsigned q_r = a.P == 1 ? u_r : __umulhi(u_r, a.p_magic);
                t_r = (int)q_r - 4; c_r = (int)(u_r - q_r * (unsigned)a.P);
            }
            int rho = rr;
            const int pstep = a.nslab * npb * pl.plane_bytes;                         // plane 1 images follow plane 0's
            const int ni = rpt * a.nplanes;                                           // 2 or 4 regular items per thread and group
            for (int j = 0; j < pl.ngroups; ++j) {
                const bool last = j == pl.ngroups - 1;
                const int rho_end = last ? pl.R : (j + 1) * a.G;
                // first raw row of the stage = the group's first padded row clamped to 0 (group_range, what the loader copied from)
                const int g_lo = t_a < 0 ? 0 : t_a * a.Fin + min(c_a * a.col_stride, a.Fin);
                c_a += a.rG; t_a += a.qG;
                if (c_a >= a.P) { c_a -= a.P; ++t_a; }
                // row slots h = 0..rpt-1 (row rho = j G + rr + 32 h) and the tail slot 4 (the <= kTail rows past the last group's
                // G).  sc: column, 0x7fffffff for a row outside [0, T P) (literal zeros), -1 for a slot past the tile's rows
                // (it repeats slot 0: no divergence) ; sb: raw row in the stage of column 0 of the row's frame position
                int sb[5], sc[5];
#pragma unroll
                for (int h = 0; h < 5; ++h) {
                    sb[h] = 0; sc[h] = -1;
                    const bool slot = h < 4 ? h < rpt : (last && rr < kTail);
                    if (slot) {
                        if (rho < rho_end || (h == 4 && rho < pl.R)) {
                            const bool in = t_r >= 0 && t_r < a.T;
                            sc[h] = in ? c_r : 0x7fffffff;
                            sb[h] = in ? t_r * a.Fin + c_r * a.col_stride - g_lo : 0;
                        }
                        if (h < 4) {
                            rho += 32; c_r += a.r32; t_r += a.q32;
                            if (c_r >= a.P) { c_r -= a.P; ++t_r; }
                        }
                    }
                }
                const int drow = (j * a.G + rr) * 128 + ((c8 ^ (rr & 7)) << 4);      // 32 h more rows: + 4096 h (same swizzle phase)
                const bool any = sc[0] >= 0;
                const bool any_tail = sc[4] >= 0;
#pragma unroll 1
                for (int s = 0; s < a.nslab; ++s) {
                    const long long w2 = dbg_on ? clock64() : 0;
                    RWAIT(&raw_full[stage], sphase);
                    if (dbg_on) t_wr += clock64() - w2;
                    const uint8_t* st0 = ring + (size_t)stage * a.stage_bytes + c8 * 32;
                    const uint8_t* st1 = st0 + a.add1_off;
                    const float* cf = coef + s * 2 * 192;
                    uint8_t* ob = obuf + s * npb * pl.plane_bytes + drow;
                    const bool dual = (a.dual_mask >> s) & 1u;
                    const unsigned act = __ballot_sync(0xffffffffu, any);
                    if (any && !(a.exp_flags & 1)) {
                        if (ni == 4) {
                            // (four chunks per thread only in launches without lazy pairs: the launcher halves G for those)
                            uint8_t* dst[4];
                            int so[4]; bool k[4];
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                // (every array index a compile-time constant: a run-time one sends the slot arrays to local memory)
                                const int h = np2 ? (i >> 1) : i;
                                const bool p1 = np2 && (i & 1);
                                const int sch = np2 ? sc[i >> 1] : sc[i], sbh = np2 ? sb[i >> 1] : sb[i];
                                const bool use = sch >= 0;                            // a slot past the tile's rows repeats slot 0
                                const int scx = use ? sch : sc[0], sbx = use ? sbh : sb[0];
                                k[i] = (unsigned)scx < (unsigned)(p1 ? pcols1 : pcols0);
                                so[i] = k[i] ? (sbx + (p1 ? coff1 : coff0)) * 256 : 0;
                                dst[i] = ob + (use ? h * 4096 : 0) + (p1 ? pstep : 0);
                            }
                            xf_items<4, false, LO, FAST>(st0, st1, so, k, dst, cf, qa, chA, chB, pl.plane_bytes, act);
                        } else {
                            uint8_t* dst[2];
                            int so[2]; bool k[2];
#pragma unroll
                            for (int i = 0; i < 2; ++i) {
                                const int h = np2 ? 0 : i;
                                const bool p1 = np2 && i == 1;
                                const int sch = np2 ? sc[0] : sc[i], sbh = np2 ? sb[0] : sb[i];
                                const bool use = sch >= 0;
                                const int scx = use ? sch : sc[0], sbx = use ? sbh : sb[0];
                                k[i] = (unsigned)scx < (unsigned)(p1 ? pcols1 : pcols0);
                                so[i] = k[i] ? (sbx + (p1 ? coff1 : coff0)) * 256 : 0;
                                dst[i] = ob + (use ? h * 4096 : 0) + (p1 ? pstep : 0);
                            }
                            if (dual) xf_items<2, true, LO, FAST>(st0, st1, so, k, dst, cf, qa, chA, chB, pl.plane_bytes, act);
                            else xf_items<2, false, LO, FAST>(st0, st1, so, k, dst, cf, qa, chA, chB, pl.plane_bytes, act);
                        }
                    }
                    const unsigned act_t = __ballot_sync(0xffffffffu, any_tail);
                    if (any_tail && !(a.exp_flags & 3)) {                                                   // the <= kTail rows past the last group's G
                        uint8_t* dst[2];
                        int so[2]; bool k[2];
#pragma unroll
                        for (int i = 0; i < 2; ++i) {
                            const bool p1 = np2 && i == 1;                            // one plane: the second item repeats the first
                            k[i] = (unsigned)sc[4] < (unsigned)(p1 ? pcols1 : pcols0);
                            so[i] = k[i] ? (sb[4] + (p1 ? coff1 : coff0)) * 256 : 0;
                            dst[i] = ob + rpt * 4096 + (p1 ? pstep : 0);
                        }
                        if (dual) xf_items<2, true, LO, FAST>(st0, st1, so, k, dst, cf, qa, chA, chB, pl.plane_bytes, act_t);
                        else xf_items<2, false, LO, FAST>(st0, st1, so, k, dst, cf, qa, chA, chB, pl.plane_bytes, act_t);
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&raw_empty[stage]);
                    if (++stage == a.nstage) { stage = 0; sphase ^= 1; }
                }
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&opnd_full[buf]);
        }
    };

    if (warp < 4) {
        if (NTW == 8) reg_dec<48>();
        if (warp == 0) {
            // =========================================================================== raw loader
            int stage = 0;
            uint32_t sphase = 0;
            const bool dbg_on = kDbg && a.dbg != nullptr && blockIdx.x == 0 && lane == 0;
            long long t_w = 0;
            const long long t_start = dbg_on ? clock64() : 0;
            for (int tile = tile_begin; tile < tile_end; ++tile) {
                const int b = tile / a.tiles_per_b;
                const int row0 = (tile - b * a.tiles_per_b) * TM;
                for (int j = 0; j < pl.ngroups; ++j) {
                    int g_lo, n;
                    group_range(a, pl.R, pl.ngroups, rows_per_b, row0, j, g_lo, n);
                    const size_t goff = ((size_t)b * a.T * a.Fin + g_lo) * 64;
                    for (int s = 0; s < a.nslab; ++s) {
                        const long long w0 = dbg_on ? clock64() : 0;
                        RWAIT_IDLE(&raw_empty[stage], sphase ^ 1);
                        if (dbg_on) t_w += clock64() - w0;
                        if (lane == 0) {
                            if (n > 0 && !(a.exp_flags & 16)) {
                                const bool dual = a.x1[s] != nullptr;
                                mbar_arrive_expect_tx(&raw_full[stage], (uint32_t)(n * 256 * (dual ? 2 : 1)));
                                uint8_t* dst = ring + (size_t)stage * a.stage_bytes;
                                bulk_copy_g2s(dst, a.x0[s] + goff, (uint32_t)(n * 256), &raw_full[stage]);
                                if (dual) bulk_copy_g2s(dst + a.add1_off, a.x1[s] + goff, (uint32_t)(n * 256), &raw_full[stage]);
                            } else {
                                mbar_arrive(&raw_full[stage]);
                            }
                        }
                        __syncwarp();
                        if (++stage == a.nstage) { stage = 0; sphase ^= 1; }
                    }
                }
            }
            if (dbg_on) { a.dbg[4] = t_w; a.dbg[5] = clock64() - t_start; }
        } else if (warp == 1) {
            // =========================================================================== MMA issuer (convergent, see conv_tma)
            const uint32_t idesc = make_idesc(a.N);
            const uint32_t bs_lo = desc_lo(smem_u32(Bs));
            const uint32_t bstep = (uint32_t)pl.b_stage_bytes >> 4;
            int stage = 0;
            uint32_t sphase = 0;
            const bool dbg_on = kDbg && a.dbg != nullptr && blockIdx.x == 0 && lane == 0;
            long long t_wa = 0, t_wo = 0, t_wb = 0;
            const long long t_start = dbg_on ? clock64() : 0;
            if (a.resident && tile_begin < tile_end) { RWAIT(&b_full[0], 0u); tc_fence_after(); }
            for (int tile = tile_begin; tile < tile_end; ++tile) {
                const int ord = tile - tile_begin;
                const int acc = ord & 1;
                const int buf = ord % a.nbuf;
                const uint32_t bphase = (uint32_t)((ord / a.nbuf) & 1);
                const uint32_t aphase = (uint32_t)((ord >> 1) & 1);
                const long long w0 = dbg_on ? clock64() : 0;
                RWAIT(&acc_empty[acc], aphase ^ 1);
                const long long w1 = dbg_on ? clock64() : 0;
                mbar_wait_backoff(&opnd_full[buf], bphase, 20u);                     // a lone poller still takes ~1/4 of its scheduler's issue slots
                if (dbg_on) { t_wa += w1 - w0; t_wo += clock64() - w1; }
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * acc_cols);
                const uint32_t origin = smem_u32(opnd + buf * pl.buf_bytes) >> 4;
                for (int unit = 0; unit < a.nunits; ++unit) {
                    const uint32_t fl = a.unit_c[unit];
                    uint32_t blo;
                    if (!a.resident) {
                        // 3-pass layers: passes 0 (A hi) and 1 (A lo) of a (tap, slab) multiply the SAME weight image (W hi):
                        // one ring stage serves both (bit 5 of the unit flags = "weights already in the stage")
                        if (!(fl & 0x20u)) {
                            const long long w2 = dbg_on ? clock64() : 0;
                            RWAIT(&b_full[stage], sphase);
                            if (dbg_on) t_wb += clock64() - w2;
                            tc_fence_after();
                        }
                        blo = bs_lo + (uint32_t)stage * bstep;
                    } else {
                        blo = bs_lo + (uint32_t)a.unit_b[unit] * bstep;
                    }
                    const uint32_t alo = ((origin + a.unit_a[unit]) & 0x3FFFu) | (1u << 16);
                    if (!(a.exp_flags & 4)) umma_f16_lo_elect_x4(d_tmem + (fl & 1u) * (uint32_t)a.N, alo, blo, idesc, (fl & 0x80u) ? 0u : 1u);
                    if (!a.resident && !(fl & 0x10u)) {                 // (bit 4: the next unit reads the same stage)
                        umma_commit_elect(&b_empty[stage]);
                        if (++stage == a.nsb) { stage = 0; sphase ^= 1; }
                    }
                    if (fl & 0x40u) umma_commit_elect(&acc_full[acc * 2 + (int)(fl & 1u)]);
                }
                umma_commit_elect(&opnd_empty[buf]);
            }
            if (dbg_on) { a.dbg[6] = t_wa; a.dbg[7] = t_wo; a.dbg[8] = t_wb; a.dbg[9] = clock64() - t_start; a.dbg[12] = tile_end - tile_begin; a.dbg[13] = pl.ngroups; }
        } else if (warp == 2) {
            // =========================================================================== weight loader
            const uint32_t bytes = (uint32_t)pl.b_stage_bytes;
            if (a.resident) {
                if (lane == 0 && tile_begin < tile_end) {
                    mbar_arrive_expect_tx(&b_full[0], bytes * (uint32_t)pl.w_images);
                    int slot = 0;
                    for (int v = 0; v < a.nvar; ++v)
                        for (int ts = 0; ts < a.ntaps[v] * a.nslab; ++ts)
                            for (int hl = 0; hl < npb; ++hl, ++slot)
                                bulk_copy_g2s(Bs + (size_t)slot * pl.b_stage_bytes, (hl ? a.Wlo[v] : a.Whi[v]) + (size_t)ts * a.N * 32, bytes, &b_full[0]);
                }
                __syncwarp();
            } else {
                int stage = 0;
                uint32_t sphase = 0;
                for (int tile = tile_begin; tile < tile_end; ++tile)
                    for (int v = 0; v < a.nvar; ++v)
                        for (int ts = 0; ts < a.ntaps[v] * a.nslab; ++ts)
                            for (int hl = 0; hl < npb; ++hl) {               // W hi (passes 0 and 1), W lo (pass 2)
                                if (a.exp_flags & 64) RWAIT(&b_empty[stage], sphase ^ 1); else RWAIT_IDLE(&b_empty[stage], sphase ^ 1);
                                if (lane == 0) {
                                    const float* img = (hl ? a.Wlo[v] : a.Whi[v]) + (size_t)ts * a.N * 32;
                                    if (a.exp_flags & 32) { mbar_arrive(&b_full[stage]); }
                                    else {
                                        mbar_arrive_expect_tx(&b_full[stage], bytes);
                                        bulk_copy_g2s(Bs + (size_t)stage * pl.b_stage_bytes, img, bytes, &b_full[stage]);
                                    }
                                }
                                __syncwarp();
                                if (++stage == a.nsb) { stage = 0; sphase ^= 1; }
                            }
            }
        }
    } else if (warp < 12) {
        // =========================================================================== epilogue (8 warps)
        // thread = output row (TMEM lane), warp = (lane quadrant, half of the 64 output channels); see conv_tma_kernel
        if (NTW == 8) reg_inc<120>();
        const int quad = warp & 3;
        const int chalf = ((warp - 4) >> 2) & 1;
        const int row = quad * 32 + lane;
        const bool gated = a.gate_off > 0;
        const int own = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
        float rs_sum[32], rs_sq[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) { rs_sum[i] = 0.f; rs_sq[i] = 0.f; }
        int cur_b = -1;
        auto flush = [&](int b) {
            if (b < 0 || a.stats == nullptr) return;
#pragma unroll
            for (int it = 0; it < 4; ++it) {
                float u8[8], w8[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) { u8[i] = rs_sum[it * 8 + i]; w8[i] = rs_sq[it * 8 + i]; rs_sum[it * 8 + i] = 0.f; rs_sq[it * 8 + i] = 0.f; }
                const float s1 = fold8(u8, lane);
                const float s2 = fold8(w8, lane);
                if ((lane & 3) == 0) {
                    double* dstp = a.stats + ((size_t)b * a.Cout + chalf * 32 + it * 8 + own) * 2;
                    atomicAdd(dstp, (double)s1);
                    atomicAdd(dstp + 1, (double)s2);
                }
            }
        };
        int acc = 0;
        uint32_t aphase = 0;
        const bool dbg_on = kDbg && a.dbg != nullptr && blockIdx.x == 0 && warp == 4 && lane == 0;
        long long t_wf = 0;
        const long long t_start = dbg_on ? clock64() : 0;
        auto epi_tile = [&](const int tile) {
            const int b = tile / a.tiles_per_b;
            if (b != cur_b) { flush(cur_b); cur_b = b; }
            const int row0 = (tile - b * a.tiles_per_b) * TM;
            const int r = row0 + row;
            int t = 0, e = a.P;                              // e = P: invalid for every variant
            if (r < rows_per_b) {
                t = a.P == 1 ? r : (int)__umulhi((unsigned)r, a.p_magic);
                e = r - t * a.P;
            }
            const long long fbase = ((long long)b * a.T + t) * a.Fout;
#pragma unroll 1
            for (int v = 0; v < a.nvar; ++v) {
                const bool row_valid = e < a.E[v];
                const long long off = (fbase + (e * a.out_stride + a.out_off[v])) * 64;
                const long long w0 = dbg_on ? clock64() : 0;
                RWAIT_IDLE(&acc_full[acc * 2 + v], aphase);
                if (dbg_on) t_wf += clock64() - w0;
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * acc_cols + v * a.N);
#pragma unroll
                for (int it = 0; it < 4; ++it) {
                    const int c0 = chalf * 32 + it * 8;
                    uint32_t rv[8], rg[8];
                    tmem_ld8_nowait(taddr + c0, rv);
                    if (gated) tmem_ld8_nowait(taddr + a.gate_off + c0, rg);
                    tmem_wait_ld();
                    float val[8];
                    const float4 b0 = *reinterpret_cast<const float4*>(sbias + c0);
                    const float4 b1 = *reinterpret_cast<const float4*>(sbias + c0 + 4);
                    val[0] = __uint_as_float(rv[0]) + b0.x; val[1] = __uint_as_float(rv[1]) + b0.y; val[2] = __uint_as_float(rv[2]) + b0.z; val[3] = __uint_as_float(rv[3]) + b0.w;
                    val[4] = __uint_as_float(rv[4]) + b1.x; val[5] = __uint_as_float(rv[5]) + b1.y; val[6] = __uint_as_float(rv[6]) + b1.z; val[7] = __uint_as_float(rv[7]) + b1.w;
                    if (gated) {
                        const float4 g0 = *reinterpret_cast<const float4*>(sbias + a.gate_off + c0);
                        const float4 g1 = *reinterpret_cast<const float4*>(sbias + a.gate_off + c0 + 4);
                        val[0] *= sigmoid_f(__uint_as_float(rg[0]) + g0.x); val[1] *= sigmoid_f(__uint_as_float(rg[1]) + g0.y);
                        val[2] *= sigmoid_f(__uint_as_float(rg[2]) + g0.z); val[3] *= sigmoid_f(__uint_as_float(rg[3]) + g0.w);
                        val[4] *= sigmoid_f(__uint_as_float(rg[4]) + g1.x); val[5] *= sigmoid_f(__uint_as_float(rg[5]) + g1.y);
                        val[6] *= sigmoid_f(__uint_as_float(rg[6]) + g1.z); val[7] *= sigmoid_f(__uint_as_float(rg[7]) + g1.w);
                    }
                    // conv_tma adds a (zero) residual here; x + 0.f is the identity except for -0, which no statistic or consumer sees
                    if (row_valid && !(a.exp_flags & 8)) {
                        st_global_256(a.out + off + c0, val);
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            rs_sum[it * 8 + i] += val[i];
                            rs_sq[it * 8 + i] = fmaf(val[i], val[i], rs_sq[it * 8 + i]);
                        }
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(&acc_empty[acc]);
            if (++acc == 2) { acc = 0; aphase ^= 1; }
        };
        for (int tile = tile_begin; tile < tile_end; ++tile) epi_tile(tile);
        flush(cur_b);
        if (dbg_on) { a.dbg[10] = t_wf; a.dbg[11] = clock64() - t_start; }
    } else {
        // =========================================================================== transform warps
        const int ttid = tid - TR0;
        XfState xs = {0, 0u, -1, 0, 0, 0};
        const bool dbg_on = kDbg && a.dbg != nullptr && blockIdx.x == 0 && ttid == 0;
        const long long t_start = dbg_on ? clock64() : 0;
        for (int tile = tile_begin; tile < tile_end; ++tile) xf_tile(tile, ttid, xs, dbg_on);
        const long long t_wo = xs.t_wo, t_wr = xs.t_wr, t_cf = xs.t_cf;
        if (dbg_on) { a.dbg[0] = clock64() - t_start; a.dbg[1] = t_wo; a.dbg[2] = t_wr; a.dbg[3] = t_cf; }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, tmem_cols);
}

// stage geometry for a group size G: the last group of a tile also takes a tail of <= kTail rows (capacity = largest group)
void set_group(RawConvArgs& a, int G, bool any_dual) {
    a.G = G;
    const int R = TM + a.back + a.fwd;
    const int ng = R <= G + kTail ? 1 : (R - kTail + G - 1) / G;
    const int cap = ng == 1 ? R : std::max(G, R - (ng - 1) * G);
    a.add1_off = cap * a.col_stride * 256;
    a.stage_bytes = a.add1_off * (any_dual ? 2 : 1);
}

// shared-memory plan.  Two operand buffers first (the transform of tile i+1 runs under the MMAs of tile i: measured 1.3-2x
// on the 3-pass layers, whose hi + lo images leave room for one buffer only at the full group size), with resident weights
// before a weight ring; the group size (= ring stage size) is halved when that is what makes the second buffer fit.
bool choose_plan(RawConvArgs& a, int G0, bool any_dual) {
    static const int force_nbuf = getenv("EAB_RAW_NBUF") ? atoi(getenv("EAB_RAW_NBUF")) : 0;      // diagnostics
    static const int force_g = getenv("EAB_RAW_G") ? atoi(getenv("EAB_RAW_G")) : 0;               // 1: always the halved group
    static const int force_ns = getenv("EAB_RAW_NS") ? atoi(getenv("EAB_RAW_NS")) : 0;            // raw ring depth of ring-weight layers
    for (int nbuf = 2; nbuf >= 1; --nbuf) {
        if (force_nbuf && nbuf != force_nbuf) continue;
        for (int resident = 1; resident >= 0; --resident) {
            for (int G = G0; G >= 32 && G * a.nplanes >= 64 && G >= G0 / 2; G >>= 1) {      // >= 2 chunks per transform thread
                if (force_g && G == G0 && G0 / 2 >= 32 && (G0 / 2) * a.nplanes >= 64) continue;
                set_group(a, G, any_dual);
                a.resident = resident; a.nbuf = nbuf; a.nsb = 3; a.nstage = 0;
                const int fixed = make_raw_plan(a).total;
                if (fixed >= SMEM_LIMIT) continue;
                int ns = std::min(NSTAGE_MAX, (SMEM_LIMIT - fixed) / a.stage_bytes);
                if (ns < 2) continue;
                if (!resident && force_ns >= 2 && ns > force_ns) ns = force_ns;
                a.nstage = ns;
                if (!resident)                                   // leftover shared memory deepens the weight ring
                    for (int nsb = NSB_MAX; nsb > 3; --nsb) {
                        a.nsb = nsb;
                        if (make_raw_plan(a).total <= SMEM_LIMIT) break;
                        a.nsb = 3;
                    }
                return true;
            }
        }
    }
    return false;
}

// variants of one layer (plan_staged has already made their plane geometry identical) -> kernel arguments
bool build_args(const PlaneConvArgs* p, int n, RawConvArgs* out) {
    if (n < 1 || n > 2) return false;
    const PlaneConvArgs& q = p[0];
    RawConvArgs a;
    memset(&a, 0, sizeof(a));
    if (q.Cout != 64 || (q.N != 64 && q.N != 128) || (q.gate_off != 0 && (q.gate_off != 64 || q.N != 128)) || (q.gate_off == 0 && q.N != 64))
        return false;
    if (q.relu || q.resid || q.nstats > 1 || q.stat_alpha[0] || q.stft_M > 0 || q.wide_k > 0) return false;
    if (q.out_ld != 64 || q.out_coff != 0 || q.stats_ld != 0 || q.stats_coff != 0) return false;
    if ((reinterpret_cast<uintptr_t>(q.out) & 31) != 0) return false;
    if (q.npass != 1 && q.npass != 3) return false;
    if (q.nplanes < 1 || q.nplanes > 2 || q.P < 1) return false;
    // slabs = sources (64 channels each, one or two addends)
    int nslab = 0;
    bool any_dual = false;
    for (int i = 0; i < q.nsrc; ++i) {
        const ConvSrc& s = q.src[i];
        if (s.C != 64 || s.RT != 0) return false;     // (no null check: planning passes run on offsets)
        if ((reinterpret_cast<uintptr_t>(s.x) & 15) || (s.x2 && (reinterpret_cast<uintptr_t>(s.x2) & 15))) return false;
        if (nslab >= kRawMaxSlabs) return false;
        a.x0[nslab] = s.x; a.xf0[nslab] = s.xf;
        if (s.xf.prelu == 1 || (s.x2 && s.xf2.prelu == 1)) return false;      // PReLU -> norm (TCM order): the staged pair
        a.mode0[nslab] = (s.xf.prelu == 0 || s.xf.alpha01) ? 3 : 1;
        a.x1[nslab] = s.x2;
        if (s.x2) {
            a.xf1[nslab] = s.xf2;
            a.mode1[nslab] = (s.xf2.prelu == 0 || s.xf2.alpha01) ? 3 : 1;
            any_dual = true;
        } else {
            a.xf1[nslab] = xform_identity();
        }
        ++nslab;
    }
    if (nslab != q.nslab) return false;
    a.nslab = nslab;
    if (q.col_stride < 1 || q.col_stride > 2 || q.nplanes != q.col_stride) return false;
    a.B = q.B; a.T = q.T; a.Fin = q.Fin; a.P = q.P;
    a.nplanes = q.nplanes; a.col_stride = q.col_stride;
    for (int i = 0; i < 2; ++i) { a.plane_cols[i] = q.plane_cols[i]; a.col_off[i] = q.col_off[i]; }
    if (a.plane_cols[0] > a.P || (a.nplanes == 2 && a.plane_cols[1] > a.P)) return false;
    if ((long long)a.P * a.col_stride < a.Fin) return false;         // group_range: a frame's columns fit its padded pitch
    for (int pl = 0; pl < a.nplanes; ++pl)                          // planes cover input columns < Fin only
        if (a.plane_cols[pl] > 0 && (a.plane_cols[pl] - 1) * a.col_stride + a.col_off[pl] >= a.Fin) return false;
    a.tiles_per_b = q.tiles_per_b;
    a.p_magic = a.P == 1 ? 0u : (unsigned)((1ull << 32) / (unsigned)a.P) + 1u;
    if (((long long)a.T * a.P + 4 * TM + 6ll * a.P) * a.P >= (1ll << 31)) return false;              // magic-division range
    if ((long long)a.T * a.Fin >= (1ll << 30)) return false;
    a.npass = q.npass; a.nvar = n;
    a.out_stride = q.out_stride; a.Fout = q.Fout;
    a.bias = q.bias; a.Cout = q.Cout; a.N = q.N; a.gate_off = q.gate_off;
    a.out = q.out;
    a.stats = q.nstats ? q.stats[0] : nullptr;
    a.algo_frac = q.algo_frac;
    int back = 0, fwd = 0;
    for (int v = 0; v < n; ++v) {
        const PlaneConvArgs& w = p[v];
        if (w.P != q.P || w.nplanes != q.nplanes || w.nslab != q.nslab || w.npass != q.npass || w.tiles_per_b != q.tiles_per_b ||
            w.Cout != q.Cout || w.N != q.N || w.gate_off != q.gate_off || w.bias != q.bias || w.out != q.out || w.out_stride != q.out_stride ||
            w.Fout != q.Fout || w.B != q.B || w.T != q.T || w.Fin != q.Fin || w.nsrc != q.nsrc || w.col_stride != q.col_stride ||
            w.plane_cols[0] != q.plane_cols[0] || w.plane_cols[1] != q.plane_cols[1] || w.nstats != q.nstats || w.stats[0] != q.stats[0] ||
            w.relu || w.resid || w.stft_M > 0 || w.wide_k > 0 || w.out_ld != 64 || w.out_coff != 0)
            return false;
        for (int i = 0; i < q.nsrc; ++i)
            if (w.src[i].x != q.src[i].x || w.src[i].x2 != q.src[i].x2 || w.src[i].C != q.src[i].C) return false;
        if (w.ntaps < 1 || w.ntaps > kMaxTaps || w.E < 1 || w.E > w.P) return false;
        a.E[v] = w.E; a.out_off[v] = w.out_off; a.ntaps[v] = w.ntaps;
        a.Whi[v] = w.Whi; a.Wlo[v] = w.Wlo;
        if ((reinterpret_cast<uintptr_t>(w.Whi) & 15) || (w.npass == 3 && (reinterpret_cast<uintptr_t>(w.Wlo) & 15))) return false;
        for (int i = 0; i < w.ntaps; ++i) {
            if (w.tap_plane[i] < 0 || w.tap_plane[i] >= w.nplanes) return false;
            back = std::max(back, -w.tap_shift[i]);
            fwd = std::max(fwd, w.tap_shift[i]);
        }
    }
    a.back = back; a.fwd = fwd;
    // group size: a stage holds <= 128 raw rows per addend (32 KB), i.e. 2 or 4 operand chunks per transform thread
    const int G0 = 128 / (q.col_stride * (any_dual ? 2 : 1));
    set_group(a, G0, any_dual);
    // K units: variant-major, then (tap, slab), pass fastest (the weight ring streams images in the same order)
    const int npb = a.npass == 3 ? 2 : 1;
    a.nbuf = 1; a.nstage = 2; a.resident = 0; a.nsb = 3;
    const int plane_bytes = make_raw_plan(a).plane_bytes;
    int nunits = 0, slot_base = 0;
    for (int v = 0; v < n; ++v) {
        const PlaneConvArgs& w = p[v];
        const int uv = w.ntaps * a.nslab * a.npass;
        if (nunits + uv > kMaxConvUnits) return false;
        for (int u = 0; u < uv; ++u) {
            const int pass = u % a.npass, ts = u / a.npass;
            const int slab = ts % a.nslab, tap = ts / a.nslab;
            const unsigned a_rel = (unsigned)(((w.tap_plane[tap] * a.nslab + slab) * npb + (pass == 1 ? 1 : 0)) * plane_bytes +
                                              (back + w.tap_shift[tap]) * 128);
            a.unit_a[nunits] = a_rel >> 4;
            a.unit_b[nunits] = (unsigned short)(slot_base + (tap * a.nslab + slab) * npb + (pass == 2 ? 1 : 0));
            a.unit_c[nunits] = (unsigned char)(v | (u == 0 ? 0x80 : 0) | (u == uv - 1 ? 0x40 : 0) | (a.npass == 3 && pass == 1 ? 0x20 : 0) |
                                               (a.npass == 3 && pass == 0 ? 0x10 : 0));
            ++nunits;
        }
        slot_base += w.ntaps * a.nslab * npb;
    }
    a.nunits = nunits;
    if (!choose_plan(a, G0, any_dual)) return false;
    if (a.back > 4 * a.P) return false;                               // xf_tile: floor division of row0 - back through + 4 P
    a.q32 = 32 / a.P; a.r32 = 32 % a.P; a.qG = a.G / a.P; a.rG = a.G % a.P;
    a.dual_mask = 0;
    for (int i = 0; i < a.nslab; ++i) a.dual_mask |= a.x1[i] ? 1u << i : 0u;
    *out = a;
    return true;
}

}  // namespace

bool raw_conv_supported(const PlaneConvArgs* p, int n) {
    RawConvArgs a;
    return build_args(p, n, &a);
}

int launch_conv_raw(const PlaneConvArgs* p, int n, cudaStream_t st, unsigned long long* dbg, int max_grid) {
    RawConvArgs a;
    if (!build_args(p, n, &a)) return fail("conv_raw: unsupported layer");
    if (a.B <= 0 || a.T <= 0) return 0;
    a.dbg = dbg;
    static const int exp_flags = getenv("EAB_RAW_EXP") ? atoi(getenv("EAB_RAW_EXP")) : 0;      // timing experiments (wrong results)
    a.exp_flags = exp_flags;
    const RawPlan pl = make_raw_plan(a);
    constexpr int NTW = 8;
    bool fast = true;
    for (int s = 0; s < a.nslab; ++s) fast = fast && a.mode0[s] == 3 && (!a.x1[s] || a.mode1[s] == 3);
    const bool lo = a.npass == 3;
    void (*kern)(RawConvArgs) = lo ? (fast ? conv_raw_kernel<NTW, true, true> : conv_raw_kernel<NTW, true, false>)
                                   : (fast ? conv_raw_kernel<NTW, false, true> : conv_raw_kernel<NTW, false, false>);
    EAB_TRY(ensure_dynamic_smem(reinterpret_cast<const void*>(kern), pl.total));
    int sms = 0;
    EAB_TRY(device_sm_count(&sms));
    const long long ntiles = (long long)a.B * a.tiles_per_b;
    if (ntiles >= (1ll << 30)) return fail("conv_raw: too many tiles");
    int grid = (int)(ntiles < sms ? ntiles : sms);
    if (max_grid > 0 && grid > max_grid) grid = max_grid;      // diagnostics / tests: many tiles per CTA at small sizes
    // algorithmic bytes: every input tensor once (a lazy residual pair is ONE tensor), the output once, fp32;
    // moved: both addends of a lazy pair, and the weights
    double pos = 0, kn = 0, in_algo = 0, in_moved = 0;
    for (int v = 0; v < a.nvar; ++v) { pos += (double)a.B * a.T * a.E[v]; kn += (double)a.B * a.T * a.E[v] * a.ntaps[v]; }
    for (int s = 0; s < a.nslab; ++s) {
        in_algo += 4.0 * a.B * a.T * a.Fin * 64.0;
        in_moved += 4.0 * a.B * a.T * a.Fin * 64.0 * (a.x1[s] ? 2 : 1);
    }
    const double kreal = 64.0 * a.nslab;
    double wbytes = 0;
    for (int v = 0; v < a.nvar; ++v) wbytes += 2.0 * (a.npass == 3 ? 2 : 1) * a.ntaps[v] * kreal * a.N;
    static const bool verbose = getenv("EAB_RAW_VERBOSE") != nullptr;     // diagnostics: one line per launch
    if (verbose)
        fprintf(stderr, "conv_raw: Fin %d Fout %d P %d nslab %d dual %d planes %d npass %d N %d nvar %d taps %d+%d R %d G %d groups %d nbuf %d nstage %d "
                        "resident %d nsb %d stage_bytes %d smem %d tiles %lld\n", a.Fin, a.Fout, a.P, a.nslab, a.add1_off != a.stage_bytes, a.nplanes,
                a.npass, a.N, a.nvar, a.ntaps[0], a.nvar > 1 ? a.ntaps[1] : 0, pl.R, a.G, pl.ngroups, a.nbuf, a.nstage, a.resident, a.nsb,
                a.stage_bytes, pl.total, ntiles);
    ProfScope ps("conv_raw", 2.0 * kn * kreal * a.N * a.algo_frac, in_algo + 4.0 * pos * a.Cout, st, in_moved + 4.0 * pos * a.Cout + wbytes);
    EAB_CUDA(launch_k(kern, dim3(grid), dim3(NCTRL + NEPI + NTW * 32), (size_t)pl.total, st, a));
    EAB_LAUNCH_CHECK("conv_raw_kernel");
    return 0;
}

}  // namespace eab
