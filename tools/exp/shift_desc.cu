// Experiment: does tcgen05.mma accept a 128B-swizzled K-major A operand whose start address is shifted by an
// arbitrary number of 128-byte rows (not a multiple of the 1024-byte swizzle atom)?  Variants: base_offset = 0 or
// (addr >> 7) & 7.   nvcc -gencode arch=compute_100a,code=sm_100a -o shift_desc shift_desc.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_fp16.h>
#include "../../eabnet_b200/csrc/umma.cuh"
using namespace eab::umma;

__device__ __forceinline__ uint64_t make_desc_bo(uint32_t addr, int base_offset) {
    uint64_t d = make_desc(addr);
    d |= (uint64_t)(base_offset & 7) << 49;
    return d;
}

__global__ void k(const __half* A_all, const __half* B, float* D, int shift, int mode) {
    extern __shared__ uint8_t raw[];
    uint8_t* sm = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
    uint8_t* As = sm;                    // 256 rows x 128 B
    uint8_t* Bs = sm + 256 * 128;        // 64 rows x 128 B
    uint64_t* bar = reinterpret_cast<uint64_t*>(Bs + 64 * 128);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
    const int tid = threadIdx.x;
    for (int i = tid; i < 256 * 8; i += blockDim.x) {
        const int row = i >> 3, ch = i & 7;
        *reinterpret_cast<uint4*>(As + row * 128 + ((ch ^ (row & 7)) << 4)) = *reinterpret_cast<const uint4*>(A_all + row * 64 + ch * 8);
    }
    for (int i = tid; i < 64 * 8; i += blockDim.x) {
        const int row = i >> 3, ch = i & 7;
        *reinterpret_cast<uint4*>(Bs + row * 128 + ((ch ^ (row & 7)) << 4)) = *reinterpret_cast<const uint4*>(B + row * 64 + ch * 8);
    }
    if (tid == 0) { mbar_init(bar, 1); fence_barrier_init(); }
    if (tid < 32) tmem_alloc(slot, 64);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = *slot;
    if (tid == 0) {
        const uint32_t idesc = make_idesc(64);
        const uint32_t a0 = smem_u32(As) + shift * 128, b0 = smem_u32(Bs);
        for (int kk = 0; kk < 4; ++kk) {
            const uint32_t aa = a0 + kk * 32;
            const int bo = mode == 1 ? ((aa >> 7) & 7) : 0;
            umma_f16(tm, make_desc_bo(aa, bo), make_desc(b0 + kk * 32), idesc, kk ? 1u : 0u);
        }
        umma_commit(bar);
    }
    mbar_wait(bar, 0);
    tc_fence_after();
    if (tid < 128) {
        const int warp = tid >> 5, lane = tid & 31;
        for (int c0 = 0; c0 < 64; c0 += 16) {
            float v[16];
            tmem_ld16(tm + ((uint32_t)(warp * 32) << 16) + c0, v);
            for (int i = 0; i < 16; ++i) D[(warp * 32 + lane) * 64 + c0 + i] = v[i];
        }
    }
    tc_fence_before();
    __syncthreads();
    if (tid < 32) tmem_dealloc(tm, 64);
}

int main() {
    std::vector<__half> A(256 * 64), B(64 * 64);
    std::vector<float> Af(256 * 64), Bf(64 * 64);
    srand(1);
    for (int i = 0; i < 256 * 64; ++i) { Af[i] = (float)(rand() % 7 - 3); A[i] = __float2half(Af[i]); }
    for (int i = 0; i < 64 * 64; ++i) { Bf[i] = (float)(rand() % 5 - 2); B[i] = __float2half(Bf[i]); }
    __half *dA, *dB; float* dD;
    cudaMalloc(&dA, A.size() * 2); cudaMalloc(&dB, B.size() * 2); cudaMalloc(&dD, 128 * 64 * 4);
    cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice);
    const int smem = 256 * 128 + 64 * 128 + 64 + 1024;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int shifts[] = {0, 8, 1, 2, 3, 5, 7, 9, 12, 41, 83, 127};
    for (int mode = 0; mode < 2; ++mode)
        for (int s : shifts) {
            cudaMemset(dD, 0, 128 * 64 * 4);
            k<<<1, 128, smem>>>(dA, dB, dD, s, mode);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("mode %d shift %d: CUDA error %s\n", mode, s, cudaGetErrorString(e)); return 1; }
            std::vector<float> D(128 * 64);
            cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
            int bad = 0;
            for (int r = 0; r < 128; ++r)
                for (int n = 0; n < 64; ++n) {
                    float ref = 0;
                    for (int kk = 0; kk < 64; ++kk) ref += Af[(r + s) * 64 + kk] * Bf[n * 64 + kk];
                    if (ref != D[r * 64 + n]) ++bad;
                }
            printf("mode %d (base_offset %s) shift %3d: %s (%d mismatches)\n", mode, mode ? "(addr>>7)&7" : "0", s, bad ? "WRONG" : "ok", bad);
        }
    return 0;
}
