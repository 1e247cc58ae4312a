// l_k2.cu -- instantiations of K2 (k2_wave.cuh), the wavefront fill for long pairs.
// K2 is launched COOPERATIVELY (all CTAs co-resident, which its spin waits need) as a persistent grid of
// pair groups: Q consecutive CTAs work on one pair.  (Thread-block clusters would give the same
// guarantee, but clusters of 4 must sit inside one GPC and strand 16 of the B200's 148 SMs:
// 33 resident clusters instead of 37 groups -- measured, see profiles/.)
#include "launch.h"
#include "k2_wave.cuh"
#include "k2f_fine.cuh"

namespace bg {

template <class Kern>
static cudaError_t launch_k2_impl(Kern kern, int n_cta, int wpc, size_t smem, cudaStream_t st, const WaveArgs& a) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)n_cta);          // one CTA per SM (launch bounds: 1 block of 16 warps per SM)
    cfg.blockDim = dim3((unsigned)wpc * 32);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, a);
}

cudaError_t launch_k2(bool local, bool prof4, int C, int n_cta, int wpc, size_t smem, cudaStream_t st, const WaveArgs& a, bool ckpt) {
    if (C == WAVE_C_NARROW) {
        if (ckpt) return cudaErrorInvalidValue;        // the host never plans a bounded-memory launch on narrow bands
        if (local) {
            if (prof4) return launch_k2_impl(k2_wave<WAVE_C_NARROW, true, true>, n_cta, wpc, smem, st, a);
            return launch_k2_impl(k2_wave<WAVE_C_NARROW, true, false>, n_cta, wpc, smem, st, a);
        }
        if (prof4) return launch_k2_impl(k2_wave<WAVE_C_NARROW, false, true>, n_cta, wpc, smem, st, a);
        return launch_k2_impl(k2_wave<WAVE_C_NARROW, false, false>, n_cta, wpc, smem, st, a);
    }
    if (C != WAVE_C) return cudaErrorInvalidValue;
    if (ckpt) {
        if (local) {
            if (prof4) return launch_k2_impl(k2_wave<WAVE_C, true, true, true>, n_cta, wpc, smem, st, a);
            return launch_k2_impl(k2_wave<WAVE_C, true, false, true>, n_cta, wpc, smem, st, a);
        }
        if (prof4) return launch_k2_impl(k2_wave<WAVE_C, false, true, true>, n_cta, wpc, smem, st, a);
        return launch_k2_impl(k2_wave<WAVE_C, false, false, true>, n_cta, wpc, smem, st, a);
    }
    if (local) {
        if (prof4) return launch_k2_impl(k2_wave<WAVE_C, true, true>, n_cta, wpc, smem, st, a);
        return launch_k2_impl(k2_wave<WAVE_C, true, false>, n_cta, wpc, smem, st, a);
    }
    if (prof4) return launch_k2_impl(k2_wave<WAVE_C, false, true>, n_cta, wpc, smem, st, a);
    return launch_k2_impl(k2_wave<WAVE_C, false, false>, n_cta, wpc, smem, st, a);
}

template <class Kern>
static cudaError_t launch_k2f_impl(Kern kern, int n_cta, int warps, size_t smem, cudaStream_t st, const FineArgs& a) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)n_cta);
    cfg.blockDim = dim3((unsigned)warps * 32);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, a);
}

cudaError_t launch_k2f(bool local, bool prof4, int n_cta, int warps_per_cta, size_t smem, cudaStream_t st, const FineArgs& a) {
    if (local) {
        if (prof4) return launch_k2f_impl(k2f_fine<true, true>, n_cta, warps_per_cta, smem, st, a);
        return launch_k2f_impl(k2f_fine<true, false>, n_cta, warps_per_cta, smem, st, a);
    }
    if (prof4) return launch_k2f_impl(k2f_fine<false, true>, n_cta, warps_per_cta, smem, st, a);
    return launch_k2f_impl(k2f_fine<false, false>, n_cta, warps_per_cta, smem, st, a);
}

}  // namespace bg
