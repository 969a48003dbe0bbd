// K5: standalone OU (+ oscillation) generator, one CTA per trial.
// generate_ou_process / generate_ou_with_oscillation
// (reference src/utils/ou_process.jl:46-62, 114-148, 176-218) written to a
// (n_trials x n_t) column-major Float64 matrix, as Julia holds it.
#include "kernels.h"
#include "simulate.cuh"

namespace itjl {

template <bool OSC>
__global__ void __launch_bounds__(kThreads) k_ou_generate(const __grid_constant__ OuGenParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* xs = reinterpret_cast<float*>(smem_raw);
    SimShared<kThreads>* sh = reinterpret_cast<SimShared<kThreads>*>(xs + (size_t)P.chunk * kThreads);
    const int trial = blockIdx.x;
    // the output is the trajectory itself: a growing one may be rescaled when it is standardised anyway,
    // never time-reversed
    const OuConsts oc = make_ou_consts(P.tau, P.dt, P.update, OSC ? ITJL_KIND_OU_OSC : ITJL_KIND_OU,
                                       P.freq, P.coeff, P.n_t, P.mode == kScaleStd, false);
    SimCache sim_cache;
    SimArgs g;
    g.n_t = P.n_t; g.chunk = P.chunk;
    g.inv_n = 1.0 / (double)P.n_t; g.sqrt_nm1 = sqrtf((float)(P.n_t - 1));
    g.keys = &P.keys; g.c3_noise = P.c3_noise; g.c3_init = P.c3_init;
    g.particle = P.particle;
    g.u0 = P.u0; g.noise = P.noise; g.phases = P.phases;
    g.mask_bits = nullptr; g.n_valid = nullptr; g.mask_words = 0;
    simulate_trial<kThreads, false, OSC>(g, oc, trial, xs, sh, P.mode, P.scale, P.shift, sim_cache);
    __syncthreads();
    for (int j = threadIdx.x; j < P.n_t; j += kThreads)
        P.out[(size_t)j * P.n_trials + trial] = (double)xs[j];
}

cudaError_t ou_generate_launch(const OuGenParams& P, cudaStream_t st) {
    const size_t smem = (size_t)P.chunk * kThreads * sizeof(float) + sizeof(SimShared<kThreads>);
    void (*kern)(const OuGenParams) = (P.kind == ITJL_KIND_OU_OSC) ? k_ou_generate<true> : k_ou_generate<false>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<(unsigned)P.n_trials, kThreads, smem, st>>>(P);
    return cudaGetLastError();
}

}  // namespace itjl
