// Bandwidth-bound kernels of the GaGNet post-filter:
//   gag_pack_kernel : cat(inpt, pre_x) of GaGNet.forward (GaGNet.py:80) as a channels-last [B,T,F,4] tensor for the
//                     first gated conv, plus pre_x as the [B,T,KP] row (channel = ri*F + f, zero-padded to whole
//                     64-channel slabs) that the glance / gaze 1x1 convs concatenate with the encoder feature
//                     (GaGNet.py:189-190, 249-250).
//   gag_crm_kernel  : GlanceGazeModule.forward's coarse filtering + residual (GaGNet.py:128-133):
//                     mag(pre) * acti(gain) * (cos, sin)(phase(pre)) + resi  ==  pre * acti(gain) + resi
//                     (|z| cos(arg z) = re z, |z| sin(arg z) = im z; atan2(0,0) = 0 gives 0 both ways),
//                     written twice: as the next stage's [B,T,KP] row and as the stage's [B,2,T,F] estimate.
#include "common.cuh"

namespace eab {

namespace {

// one thread per (b, t, f)
__global__ void __launch_bounds__(256) gag_pack_kernel(const GagPackArgs a) {
    const size_t n = (size_t)a.B * a.T * a.KP2;              // KP2 = KP / 2 >= F : threads f >= F write the zero padding
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const size_t bt = i / a.KP2;
    const int f = (int)(i - bt * a.KP2);
    const int b = (int)(bt / a.T), t = (int)(bt - (size_t)b * a.T);
    float* prow = a.pre_row + bt * a.KP;
    if (f >= a.F) {
        // padding channels 2F .. KP-1 (two per thread)
        const int c = 2 * a.F + 2 * (f - a.F);
        if (c < a.KP) prow[c] = 0.f;
        if (c + 1 < a.KP) prow[c + 1] = 0.f;
        return;
    }
    const size_t TF = (size_t)a.T * a.F;
    const size_t p = (size_t)t * a.F + f;
    const float pr = __ldg(a.pre + ((size_t)b * 2 + 0) * TF + p);
    const float pi = __ldg(a.pre + ((size_t)b * 2 + 1) * TF + p);
    long long ib = (long long)b * a.sb + (long long)t * a.st + (long long)f * a.sf;
    size_t xrow = bt;
    if (a.step) {                                            // streaming: ring slots of the absolute frame index
        const int n = *a.step;
        ib += (long long)ring_slot(n, a.in_RT) * a.in_slot;
        xrow = bt * a.x_RT + ring_slot(n, a.x_RT);
    }
    const float xr = __ldg(a.inpt + ib), xi = __ldg(a.inpt + ib + a.sc);
    // memory channel order of the first conv: m*2 + ri with m = (inpt, pre_x)  <->  reference channel ri*2 + m after the
    // weight packer's permutation (see Builder::gated / Packer::conv perm_ri)
    reinterpret_cast<float4*>(a.x4)[xrow * a.F + f] = make_float4(xr, pr, xi, pi);
    prow[f] = pr;
    prow[a.F + f] = pi;
}

// one thread per (b, t, f)
__global__ void __launch_bounds__(256) gag_crm_kernel(const GagCrmArgs a) {
    const size_t n = (size_t)a.B * a.T * a.KP2;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const size_t bt = i / a.KP2;
    const int f = (int)(i - bt * a.KP2);
    float* nrow = a.next_row ? a.next_row + bt * a.KP : nullptr;
    if (f >= a.F) {
        if (nrow) {
            const int c = 2 * a.F + 2 * (f - a.F);
            if (c < a.KP) nrow[c] = 0.f;
            if (c + 1 < a.KP) nrow[c + 1] = 0.f;
        }
        return;
    }
    const float* prow = a.pre_row + bt * a.KP;
    float g = __ldg(a.gain + bt * a.ld_g + f);
    if (a.acti == 0) g = sigmoid_f(g);
    else if (a.acti == 1) g = tanh_f(g);
    else g = fmaxf(g, 0.f);
    const float yr = fmaf(__ldg(prow + f), g, __ldg(a.res_r + bt * a.ld_r + f));
    const float yi = fmaf(__ldg(prow + a.F + f), g, __ldg(a.res_i + bt * a.ld_r + f));
    if (nrow) { nrow[f] = yr; nrow[a.F + f] = yi; }
    const int b = (int)(bt / a.T), t = (int)(bt - (size_t)b * a.T);
    const size_t TF = (size_t)a.T * a.F;
    const size_t p = (size_t)t * a.F + f;
    a.out[((size_t)b * 2 + 0) * TF + p] = yr;
    a.out[((size_t)b * 2 + 1) * TF + p] = yi;
}

}  // namespace

int launch_gag_pack(const GagPackArgs& a, cudaStream_t st) {
    if (a.B <= 0 || a.T <= 0) return 0;
    if (a.KP < 2 * a.F || (a.KP & 1) || a.KP2 * 2 != a.KP) return fail("gag_pack: bad row width");
    if (a.step && (a.T != 1 || a.in_RT < 1 || a.x_RT < 1)) return fail("gag_pack: a streaming step is one frame per stream");
    const size_t n = (size_t)a.B * a.T * a.KP2;
    ProfScope ps("gag_elementwise", 0.0, 4.0 * a.B * a.T * (4.0 * a.F + 4.0 * a.F + a.KP), st);
    EAB_CUDA(launch_k(gag_pack_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), (size_t)0, st, a));
    EAB_LAUNCH_CHECK("gag_pack_kernel");
    return 0;
}

int launch_gag_crm(const GagCrmArgs& a, cudaStream_t st) {
    if (a.B <= 0 || a.T <= 0) return 0;
    if (a.KP < 2 * a.F || a.KP2 * 2 != a.KP) return fail("gag_crm: bad row width");
    const size_t n = (size_t)a.B * a.T * a.KP2;
    ProfScope ps("gag_elementwise", 6.0 * a.B * a.T * a.F, 4.0 * a.B * a.T * (5.0 * a.F + 2.0 * a.F + (a.next_row ? a.KP : 0)), st);
    EAB_CUDA(launch_k(gag_crm_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), (size_t)0, st, a));
    EAB_LAUNCH_CHECK("gag_crm_kernel");
    return 0;
}

}  // namespace eab
