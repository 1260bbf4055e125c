// Bandwidth-bound elementwise kernels: residual/skip sums of normalised tensors and the complex
// filter-and-sum over microphones (EaBNet.py:114-117, :386, :104).
#include <cuda_fp16.h>

#include "common.cuh"

namespace eab {

namespace {

// out[b,p,c] = sum_s xform_s(src_s[b,p,c]);  grid (chunks, B); 128-bit accesses when C % 4 == 0
__global__ void __launch_bounds__(256) combine_kernel(const CombineArgs a) {
    extern __shared__ float coef[];                  // [nsrc][3][C]
    const int b = blockIdx.y;
    const int C = a.C;
    for (int i = threadIdx.x; i < a.nsrc * C; i += blockDim.x) {
        const int s = i / C, c = i - s * C;
        float cs, ch, ca;
        xform_coeffs(a.src[s].xf, b, C, c, cs, ch, ca);
        coef[(s * 3 + 0) * C + c] = cs;
        coef[(s * 3 + 1) * C + c] = ch;
        coef[(s * 3 + 2) * C + c] = ca;
    }
    __syncthreads();
    const size_t per_b = (size_t)a.P * C;
    const int step = a.step ? *a.step : 0;
    const size_t base = a.step ? ((size_t)b * a.out_RT + ring_slot(step, a.out_RT)) * per_b : (size_t)b * per_b;
    size_t sbase[3];
    for (int s = 0; s < a.nsrc; ++s)
        sbase[s] = a.step ? ((size_t)b * a.src[s].RT + ring_slot(step, a.src[s].RT)) * per_b : (size_t)b * per_b;
    if ((C & 3) == 0) {
        const size_t n4 = per_b / 4;
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
            const int c = (int)((i * 4) % C);
            float o[4] = {0.f, 0.f, 0.f, 0.f};
            for (int s = 0; s < a.nsrc; ++s) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(a.src[s].x + sbase[s]) + i);
                const float x[4] = {v.x, v.y, v.z, v.w};
                const int pr = a.src[s].xf.prelu;
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    o[q] += xform_apply(x[q], coef[(s * 3 + 0) * C + c + q], coef[(s * 3 + 1) * C + c + q],
                                        coef[(s * 3 + 2) * C + c + q], pr);
            }
            reinterpret_cast<float4*>(a.out + base)[i] = make_float4(o[0], o[1], o[2], o[3]);
        }
    } else {
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < per_b; i += (size_t)gridDim.x * blockDim.x) {
            const int c = (int)(i % C);
            float o = 0.f;
            for (int s = 0; s < a.nsrc; ++s)
                o += xform_apply(__ldg(a.src[s].x + sbase[s] + i), coef[(s * 3 + 0) * C + c], coef[(s * 3 + 1) * C + c],
                                 coef[(s * 3 + 2) * C + c], a.src[s].xf.prelu);
            a.out[base + i] = o;
        }
    }
}

// mimo: y[b,:,t,f] = sum_m w[b,t,f,m] * x[b,t,f,m]  (complex);  one thread per (b,t,f)
__global__ void __launch_bounds__(256) beam_mimo_kernel(const BeamArgs a) {
    const size_t n = (size_t)a.B * a.T * a.F;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    size_t iw = i, ix = i;
    if (a.step) {                                   // T == 1: i = stream * F + f, operands are rings
        const int step = *a.step;
        const size_t s = i / a.F, f = i - s * a.F;
        iw = (s * a.w_RT + ring_slot(step, a.w_RT)) * a.F + f;
        ix = (s * a.inpt_RT + ring_slot(step, a.inpt_RT)) * a.F + f;
    }
    const float2* w = reinterpret_cast<const float2*>(a.w + iw * a.w_ld);
    const float2* x = reinterpret_cast<const float2*>(a.inpt) + ix * a.M;
    float yr = 0.f, yi = 0.f;
    for (int m = 0; m < a.M; ++m) {
        const float2 wv = __ldg(w + m), xv = __ldg(x + m);
        yr += wv.x * xv.x - wv.y * xv.y;
        yi += wv.x * xv.y + wv.y * xv.x;
    }
    const size_t TF = (size_t)a.T * a.F;
    const size_t b = i / TF, p = i - b * TF;
    a.out[(b * 2 + 0) * TF + p] = yr;
    a.out[(b * 2 + 1) * TF + p] = yi;
}

// miso: the reference multiplies by mic 0 and then sums its last axis, which is F (EaBNet.py:123-124),
// so the result is [B,2,T].  One warp per (b,t).
__global__ void __launch_bounds__(256) beam_miso_kernel(const BeamArgs a) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= a.B * a.T) return;
    size_t row = (size_t)warp * a.F, rowx = row;
    if (a.step) {
        const int step = *a.step;
        row = ((size_t)warp * a.w_RT + ring_slot(step, a.w_RT)) * a.F;
        rowx = ((size_t)warp * a.inpt_RT + ring_slot(step, a.inpt_RT)) * a.F;
    }
    float yr = 0.f, yi = 0.f;
    for (int f = lane; f < a.F; f += 32) {
        const float2 wv = __ldg(reinterpret_cast<const float2*>(a.w + (row + f) * a.w_ld));
        const float2 xv = __ldg(reinterpret_cast<const float2*>(a.inpt) + (rowx + f) * a.M);
        yr += wv.x * xv.x - wv.y * xv.y;
        yi += wv.x * xv.y + wv.y * xv.x;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        yr += __shfl_xor_sync(0xffffffffu, yr, o);
        yi += __shfl_xor_sync(0xffffffffu, yi, o);
    }
    if (lane == 0) {
        const int b = warp / a.T, t = warp - b * a.T;
        a.out[((size_t)b * 2 + 0) * a.T + t] = yr;
        a.out[((size_t)b * 2 + 1) * a.T + t] = yi;
    }
}

}  // namespace

int launch_combine(const CombineArgs& a, cudaStream_t st) {
    if (a.nsrc < 1 || a.nsrc > 3) return fail("combine: 1..3 sources");
    if (a.B <= 0 || a.P <= 0) return 0;
    const size_t per_b = (size_t)a.P * a.C;
    size_t work = (a.C & 3) == 0 ? per_b / 4 : per_b;
    int blocks = (int)((work + 256 * 4 - 1) / (256 * 4));
    if (blocks < 1) blocks = 1;
    if (blocks > 148 * 8) blocks = 148 * 8;
    const size_t smem = (size_t)a.nsrc * 3 * a.C * sizeof(float);
    ProfScope ps("combine", (double)a.nsrc * a.B * per_b, 4.0 * (a.nsrc + 1) * a.B * per_b, st);
    EAB_CUDA(launch_k(combine_kernel, dim3(blocks, a.B), dim3(256), smem, st, a));
    EAB_LAUNCH_CHECK("combine_kernel");
    return 0;
}

int launch_beam(const BeamArgs& a, cudaStream_t st) {
    const double npos = (double)a.B * a.T * a.F;
    const int nw = a.miso ? 2 : 2 * a.M;
    ProfScope ps("beam", 4.0 * nw * npos, 4.0 * npos * (nw + 2.0 * a.M + 2), st);
    if (a.miso) {
        const size_t warps = (size_t)a.B * a.T;
        EAB_CUDA(launch_k(beam_miso_kernel, dim3((unsigned)((warps * 32 + 255) / 256)), dim3(256), (size_t)0, st, a));
        EAB_LAUNCH_CHECK("beam_miso_kernel");
    } else {
        const size_t n = (size_t)a.B * a.T * a.F;
        EAB_CUDA(launch_k(beam_mimo_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), (size_t)0, st, a));
        EAB_LAUNCH_CHECK("beam_mimo_kernel");
    }
    return 0;
}

}  // namespace eab
