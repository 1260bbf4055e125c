// Probe: issue cost per warp instruction of MUFU.EX2 / MUFU.RCP / F2FP.PACK_AB / FMNMX on one SM with 4 warps per scheduler
// (the cell-phase occupancy of the LSTM kernels).  Prints cycles per warp instruction per scheduler.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o xu_rate xu_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
template <int OP> __global__ void probe(float* out, long long* cyc, int iters) {
    float v[8];
    for (int i = 0; i < 8; ++i) v[i] = 0.5f + 0.001f * (threadIdx.x + i);
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
            if (OP == 1) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
            if (OP == 2) { uint32_t h; asm volatile("cvt.rn.f16x2.f32 %0, %1, %1;" : "=r"(h) : "f"(v[i])); v[i] = __uint_as_float(h | 0x3f000000u); }
            if (OP == 3) asm volatile("min.ftz.f32 %0, %0, 40.0;" : "+f"(v[i]));
            if (OP == 4) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[i])); asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(v[i])); }
        }
    }
    const long long t1 = clock64();
    __syncthreads();
    float s = 0.f;
    for (int i = 0; i < 8; ++i) s += v[i];
    out[threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
    float* out; long long* cyc;
    cudaMalloc(&out, 4096); cudaMalloc(&cyc, 8);
    const int iters = 2000;
    const char* names[5] = {"MUFU.EX2", "MUFU.RCP", "F2FP.PACK_AB (+LOP)", "FMNMX", "EX2+RCP"};
    for (int warps : {4, 16}) {
        for (int op = 0; op < 5; ++op) {
            for (int rep = 0; rep < 2; ++rep) {
                if (op == 0) probe<0><<<1, warps * 32>>>(out, cyc, iters);
                if (op == 1) probe<1><<<1, warps * 32>>>(out, cyc, iters);
                if (op == 2) probe<2><<<1, warps * 32>>>(out, cyc, iters);
                if (op == 3) probe<3><<<1, warps * 32>>>(out, cyc, iters);
                if (op == 4) probe<4><<<1, warps * 32>>>(out, cyc, iters);
                cudaDeviceSynchronize();
            }
            long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
            const double per_sched = (double)c / ((double)iters * 8 * (warps / 4.0) * (op == 4 ? 2 : 1));
            printf("%2d warps  %-22s %8lld cycles  = %.2f cycles per warp instruction per scheduler\n", warps, names[op], c, per_sched);
        }
    }
    return 0;
}
