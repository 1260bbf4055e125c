// Experiment: tcgen05.mma kind::f16 throughput (cycles per M128 x N x K16 MMA) vs A-descriptor row shift and N.
#include <cstdio>
#include <cuda_fp16.h>
#include "../../eabnet_b200/csrc/umma.cuh"
using namespace eab::umma;

__global__ void k(int shift, int N, int reps, long long* out) {
    extern __shared__ uint8_t raw[];
    uint8_t* sm = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
    uint8_t* As = sm;                    // 512 rows x 128 B
    uint8_t* Bs = sm + 512 * 128;        // 256 rows x 128 B
    uint64_t* bar = reinterpret_cast<uint64_t*>(Bs + 256 * 128);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
    for (int i = threadIdx.x; i < (512 + 256) * 128 / 16; i += blockDim.x) reinterpret_cast<uint4*>(sm)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
    if (threadIdx.x < 32) tmem_alloc(slot, 256);
    fence_proxy_async(); tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tm = *slot;
    if (threadIdx.x == 0) {
        const uint32_t idesc = make_idesc(N);
        const uint32_t a0 = smem_u32(As) + shift * 128, b0 = smem_u32(Bs);
        const long long t0 = clock64();
        for (int r = 0; r < reps; ++r)
            for (int kk = 0; kk < 4; ++kk)
                umma_f16(tm, make_desc(a0 + ((r & 1) * 128) * 128 + kk * 32), make_desc(b0 + kk * 32), idesc, 1u);
        umma_commit(bar);
        mbar_wait(bar, 0);
        out[0] = clock64() - t0;
    }
    tc_fence_before(); __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tm, 256);
}

int main() {
    long long* d; cudaMalloc(&d, 8);
    const int smem = (512 + 256) * 128 + 64 + 1024;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int reps = 2000;
    for (int N : {64, 128, 256})
        for (int s : {0, 8, 1, 3, 5, 41, 83}) {
            k<<<1, 128, smem>>>(s, N, reps, d);
            if (cudaDeviceSynchronize() != cudaSuccess) { printf("error\n"); return 1; }
            long long c; cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
            printf("N %3d shift %3d : %.1f cycles per MMA (M128 x N x K16)\n", N, s, (double)c / (reps * 4));
        }
    return 0;
}
