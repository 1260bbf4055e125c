// tcgen05 / TMEM / mbarrier / bulk-copy PTX wrappers shared by the sm_100a tensor-core kernels.
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

namespace eab {
namespace umma {

constexpr uint32_t SPIN_LIMIT = 1u << 22;     // x the ~10 ms suspend hint: far beyond any legitimate wait
constexpr unsigned long long WAIT_LIMIT_NS = 4000000000ull;      // wall-clock bound of one mbarrier wait (4 s)

__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// true once a wait has lasted WAIT_LIMIT_NS (checked every 256 polls from the 64th on)
__device__ __forceinline__ bool wait_expired(uint32_t spins, unsigned long long& t0) {
    if (spins < 64 || (spins & 255u) != 64u) return false;
    const unsigned long long now = global_ns();
    if (t0 == 0) { t0 = now; return false; }
    return now - t0 > WAIT_LIMIT_NS;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// Blocks in hardware: the suspend-time hint lets the warp sleep until the phase completes (or the hint expires)
// instead of spinning - spinning waiters otherwise steal issue slots from the warps that do the work.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done = 0, spins = 0;
    unsigned long long t0 = 0;
    while (true) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity), "r"(0x989680u)
            : "memory");
        if (done) break;
        if (++spins > SPIN_LIMIT || wait_expired(spins, t0)) __trap();        // a protocol bug must fault, never hang the GPU
    }
}
// Latency-critical variant: no suspend hint, the warp polls (a serial chain such as the LSTM step cannot afford the
// wake-up latency of a suspended waiter; only a handful of warps ever poll at the same time there).
// (Two-level loops: the inner one is try_wait + branch + counter only; the wall-clock check that turns a protocol bug into a
// trap instead of a hang runs once per 256 polls.  A flat loop with the check inlined issued 10-14 instructions per poll, and
// 30 % of conv_raw's issued instructions were polls.)
__device__ __forceinline__ bool mbar_try(uint32_t addr, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    return done != 0;
}
__device__ __forceinline__ void mbar_wait_spin(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    unsigned long long t0 = 0;
    for (uint32_t outer = 0;; ++outer) {
#pragma unroll 1
        for (int i = 0; i < 256; ++i)
            if (mbar_try(addr, parity)) return;
        const unsigned long long now = global_ns();
        if (t0 == 0) t0 = now;
        else if (now - t0 > WAIT_LIMIT_NS || outer > (1u << 20)) __trap();
    }
}
// Polling with a fixed back-off: a waiter that is NOT on the critical path (an epilogue warp waiting for the next
// accumulator, a loader waiting for a free ring stage) must not burn issue slots - measured on conv_raw_kernel: 60 % of all
// issued instructions were try_wait spin loops - and must not pay the wake-up latency of the hinted form either.
__device__ __forceinline__ void mbar_wait_backoff(uint64_t* bar, uint32_t parity, uint32_t sleep_ns) {
    const uint32_t addr = smem_u32(bar);
    unsigned long long t0 = 0;
    for (uint32_t outer = 0;; ++outer) {
#pragma unroll 1
        for (int i = 0; i < 256; ++i) {
            if (mbar_try(addr, parity)) return;
            __nanosleep(sleep_ns);
        }
        const unsigned long long now = global_ns();
        if (t0 == 0) t0 = now;
        else if (now - t0 > WAIT_LIMIT_NS || outer > (1u << 18)) __trap();
    }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

__device__ __forceinline__ void bulk_copy_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// K-major, 128-byte-swizzled operand tile: rows of 128 B, 8-row atoms of 1024 B stacked along M/N (SBO = 1024 B).
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);          // start address            bits [0,14)
    d |= (uint64_t)1 << 16;                              // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;                    // stride byte offset       bits [32,46)
    d |= (uint64_t)1 << 46;                              // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                              // SWIZZLE_128B
    return d;
}

// kind::f16 (fp16 operands, same 10-bit mantissa as TF32 at half the bytes and twice the rate), fp32 accumulate,
// both operands K-major, M = 128
__device__ __forceinline__ uint32_t make_idesc(int N, int M = 128) {
    uint32_t d = 0;
    d |= 1u << 4;                    // D format F32
    d |= 0u << 7;                    // A format F16
    d |= 0u << 10;                   // B format F16
    d |= (uint32_t)(N >> 3) << 17;   // N / 8
    d |= (uint32_t)(M >> 4) << 24;   // M / 16
    return d;
}

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Same MMA with the descriptors given as their LOW words (start address >> 4 | LBO): the high word of a K-major
// SWIZZLE_128B descriptor (SBO = 1024 B, version 1, swizzle mode 2) is the constant DESC_HI, so an issue loop only adds
// 2 (= 32 bytes) per K step instead of rebuilding two 64-bit descriptors (the issuing thread is a serial chain).
constexpr uint32_t DESC_HI = (1024u >> 4) | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint32_t desc_lo(uint32_t smem_addr) { return ((smem_addr >> 4) & 0x3FFFu) | (1u << 16); }
__device__ __forceinline__ void umma_f16_lo(uint32_t tmem_d, uint32_t alo, uint32_t blo, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "mov.b64 da, {%1, %5};\n\t"
        "mov.b64 db, {%2, %5};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t}" ::"r"(tmem_d),
        "r"(alo), "r"(blo), "r"(idesc), "r"(accumulate), "r"(DESC_HI)
        : "memory");
}
// Whole-warp (convergent) forms: every lane executes the surrounding loop with warp-uniform operands, one elected lane
// issues.  Keeping the issue loop convergent lets the compiler hold the descriptors in uniform registers; issuing from
// inside an `if (lane == 0)` region costs an ELECT / R2UR.BROADCAST / BRA.U.ANY waterfall (~15 instructions) per MMA.
__device__ __forceinline__ void umma_f16_lo_elect(uint32_t tmem_d, uint32_t alo, uint32_t blo, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p, e;\n\t.reg .b64 da, db;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "mov.b64 da, {%1, %5};\n\t"
        "mov.b64 db, {%2, %5};\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t}" ::"r"(tmem_d),
        "r"(alo), "r"(blo), "r"(idesc), "r"(accumulate), "r"(DESC_HI)
        : "memory");
}
// One 64-channel K slab = four K = 16 steps in ONE asm block: the descriptors are assembled once and stepped with a
// 64-bit add (the +2 stays inside the 14-bit address field), so the operands cross into uniform registers once per
// slab instead of once per MMA.
__device__ __forceinline__ void umma_f16_lo_elect_x4(uint32_t tmem_d, uint32_t alo, uint32_t blo, uint32_t idesc, uint32_t accumulate_first) {
    asm volatile(
        "{\n\t.reg .pred p, e;\n\t.reg .b64 da, db;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "mov.b64 da, {%1, %5};\n\t"
        "mov.b64 db, {%2, %5};\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t"
        "add.s64 da, da, 2;\n\tadd.s64 db, db, 2;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, 1;\n\t"
        "add.s64 da, da, 2;\n\tadd.s64 db, db, 2;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, 1;\n\t"
        "add.s64 da, da, 2;\n\tadd.s64 db, db, 2;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, 1;\n\t}" ::"r"(tmem_d),
        "r"(alo), "r"(blo), "r"(idesc), "r"(accumulate_first), "r"(DESC_HI)
        : "memory");
}
// three K steps: a slab whose last 16 columns are zero padding (first layer: 36 of 64)
__device__ __forceinline__ void umma_f16_lo_elect_x3(uint32_t tmem_d, uint32_t alo, uint32_t blo, uint32_t idesc, uint32_t accumulate_first) {
    asm volatile(
        "{\n\t.reg .pred p, e;\n\t.reg .b64 da, db;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "mov.b64 da, {%1, %5};\n\t"
        "mov.b64 db, {%2, %5};\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t"
        "add.s64 da, da, 2;\n\tadd.s64 db, db, 2;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, 1;\n\t"
        "add.s64 da, da, 2;\n\tadd.s64 db, db, 2;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, 1;\n\t}" ::"r"(tmem_d),
        "r"(alo), "r"(blo), "r"(idesc), "r"(accumulate_first), "r"(DESC_HI)
        : "memory");
}
__device__ __forceinline__ void umma_commit_elect(uint64_t* bar) {
    asm volatile(
        "{\n\t.reg .pred e;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void tmem_ld8_nowait(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// one full 32-byte sector per lane in ONE store instruction (STG.256, sm_100+): the per-row epilogue stores are
// scattered across rows, so what counts is sectors per instruction, not lanes per line
__device__ __forceinline__ void st_global_256(float* p, const float (&v)[8]) {
    asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]),
                 "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
                 : "memory");
}

__device__ __forceinline__ void ld_global_nc_256(const float* p, float4& v0, float4& v1) {
    asm volatile("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=f"(v0.x), "=f"(v0.y), "=f"(v0.z), "=f"(v0.w), "=f"(v1.x), "=f"(v1.y), "=f"(v1.z), "=f"(v1.w)
                 : "l"(p));
}

// 8 floats -> 8 fp16 (round to nearest even), and the fp16-rounded residual for the 3-pass split
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ uint32_t pack_lo_h2(float a, float b, uint32_t hi) {
    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&hi));
    return pack_h2(a - f.x, b - f.y);
}


}  // namespace umma
}  // namespace eab
