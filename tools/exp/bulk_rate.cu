// Experiment: cp.async.bulk global->shared streaming rate per SM / per chip vs copy size, copies per stage and stages in flight.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../eabnet_b200/csrc/umma.cuh"
using namespace eab::umma;

__global__ void __launch_bounds__(128) k(const uint8_t* src, size_t per_cta, int bytes, int ncopy, int nbuf, int iters, unsigned long long* sink) {
    extern __shared__ uint8_t raw[];
    uint8_t* sm = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
    __shared__ uint64_t full[8];
    if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) mbar_init(&full[i], 1); fence_barrier_init(); }
    __syncthreads();
    const uint8_t* base = src + (size_t)blockIdx.x * per_cta;
    const int stage_bytes = bytes * ncopy;
    if (threadIdx.x == 0) {
        size_t off = 0;
        // prologue: fill nbuf stages
        for (int s = 0; s < nbuf && s < iters; ++s) {
            mbar_arrive_expect_tx(&full[s], (uint32_t)stage_bytes);
            for (int c = 0; c < ncopy; ++c) { bulk_copy_g2s(sm + s * stage_bytes + c * bytes, base + off, bytes, &full[s]); off += bytes; if (off + bytes > per_cta) off = 0; }
        }
        unsigned long long acc = 0;
        for (int it = 0; it < iters; ++it) {
            const int s = it % nbuf;
            mbar_wait(&full[s], (uint32_t)((it / nbuf) & 1));
            acc += sm[s * stage_bytes];
            if (it + nbuf < iters) {
                mbar_arrive_expect_tx(&full[s], (uint32_t)stage_bytes);
                for (int c = 0; c < ncopy; ++c) { bulk_copy_g2s(sm + s * stage_bytes + c * bytes, base + off, bytes, &full[s]); off += bytes; if (off + bytes > per_cta) off = 0; }
            }
        }
        sink[blockIdx.x] = acc;
    }
}

int main() {
    const size_t per_cta = 64ull << 20;
    uint8_t* d; cudaMalloc(&d, per_cta * 148); cudaMemset(d, 1, per_cta * 148);
    unsigned long long* sink; cudaMalloc(&sink, 148 * 8);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    struct Cfg { int bytes, ncopy, nbuf; } cfgs[] = {{19456, 4, 2}, {19456, 4, 1}, {19456, 2, 4}, {19456, 1, 8}, {8192, 1, 8}, {8192, 8, 2}, {4096, 16, 2}, {32768, 2, 3}, {65536, 1, 3}, {16384, 1, 3}};
    for (auto c : cfgs) {
        const int iters = (int)((48ull << 20) / (c.bytes * c.ncopy));
        k<<<148, 128, 220 * 1024>>>(d, per_cta, c.bytes, c.ncopy, c.nbuf, iters, sink);
        cudaEventRecord(e0);
        k<<<148, 128, 220 * 1024>>>(d, per_cta, c.bytes, c.ncopy, c.nbuf, iters, sink);
        cudaEventRecord(e1);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("error %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double gb = 148.0 * iters * c.bytes * c.ncopy / 1e9;
        printf("copy %6d B x %2d per stage, %d stages in flight: %.0f GB/s chip, %.1f B/clk/SM @1.965GHz (%.1f KB in flight per SM)\n",
               c.bytes, c.ncopy, c.nbuf, gb / (ms * 1e-3), gb / (ms * 1e-3) / 148 / 1.965, c.bytes * c.ncopy * c.nbuf / 1024.0);
    }
    return 0;
}
