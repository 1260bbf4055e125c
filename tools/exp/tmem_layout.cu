// Probe: what does tcgen05.ld.16x256b.x4 hand to each thread?  TMEM is filled with value = lane * 1000 + column through
// 32x32b stores (thread = lane), then read back with the 16-lane shape at lane bases 0 and 16 of each warp's quadrant.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -o tmem_layout tmem_layout.cu ; run: ./tmem_layout
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void probe(uint32_t* out) {
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(64u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = slot;
    const int row = warp * 32 + lane;
    const uint32_t trow = base + ((uint32_t)(warp * 32) << 16);
    for (int c0 = 0; c0 < 64; c0 += 8) {
        uint32_t v[8];
        for (int i = 0; i < 8; ++i) v[i] = row * 1000 + c0 + i;
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(trow + c0), "r"(v[0]), "r"(v[1]),
                     "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int s = 0; s < 2; ++s) {
        uint32_t r[16];
        const uint32_t ta = base + ((uint32_t)(warp * 32 + 16 * s) << 16) + 32u;     // columns 32..63
        asm volatile("tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                       "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                     : "r"(ta));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int i = 0; i < 16; ++i) out[((s * 128) + tid) * 16 + i] = r[i];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(64u) : "memory");
}
int main() {
    uint32_t* d;
    cudaMalloc(&d, 2 * 128 * 16 * 4);
    cudaMemset(d, 0xff, 2 * 128 * 16 * 4);
    probe<<<1, 128>>>(d);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status %s\n", cudaGetErrorString(e));
    static uint32_t h[2 * 128 * 16];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int s = 0; s < 2; ++s)
        for (int t = 0; t < 128; ++t) {
            const int warp = t >> 5, i = t & 31, j = i & 3;
            for (int k = 0; k < 16; ++k) {
                const int g = k >> 2, rs = (k >> 1) & 1, e = k & 1;
                const int row = warp * 32 + 16 * s + (i >> 2) + 8 * rs, col = 32 + 8 * g + 2 * j + e;
                if (h[(s * 128 + t) * 16 + k] != (uint32_t)(row * 1000 + col)) ++bad;
            }
        }
    printf("expected-layout mismatches: %d\n", bad);
    for (int t : {0, 1, 4, 5, 33}) {
        printf("s0 thread %d:", t);
        for (int k = 0; k < 16; ++k) printf(" %u", h[t * 16 + k]);
        printf("\n");
    }
    printf("s1 thread 0:");
    for (int k = 0; k < 16; ++k) printf(" %u", h[(128 + 0) * 16 + k]);
    printf("\n");
    return 0;
}
