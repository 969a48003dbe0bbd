// K2: fused  simulate OU(+oscillation) trials -> Hamming periodogram -> trial mean ->
// distance, one CTA per particle, with the n2 = 2*nextpow2(n_t) point real FFT held
// entirely in shared memory (n2/2 complex points, in place).
//
// Replaces (reference) generate_data -> generate_ou_with_oscillation
// (src/utils/ou_process.jl:176-218), summary_stats ->
// mean(comp_psd_adfriendly(data, 1/dt)[1], dims=1)[:][freq_idx]
// (src/models/one_timescale_and_osc.jl:293 -> src/stats/summary.jl:204-235) and
// logarithmic_distance / linear_distance (src/stats/distances.jl:29-53).
//
// Real FFT of length n2: pack z[m] = x[2m] + i x[2m+1] (exactly the float layout the
// simulator writes), run an in-place radix-4 DIF complex FFT of N = n2/2 points (output in
// digit-reversed order, which is never undone - bins are read through the permutation) and
// recombine X[k] = E[k] + exp(-2 pi i k/n2) O[k] only for the selected bins.
#include <math.h>

#include "kernels.h"
#include "simulate.cuh"
#include "psd_fft.cuh"

namespace itjl {

constexpr int kPsdThreads = 512;      // fused PSD kernel: 1 CTA/SM (the FFT buffer fills shared memory)

template <bool OSC>
__global__ void __launch_bounds__(kPsdThreads, 1) k_abc_psd(const __grid_constant__ AbcPsdParams P) {
    constexpr int NT = kPsdThreads;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* xs = reinterpret_cast<float*>(smem_raw);                 // xs_len floats: one simulated trial
    float2* buf = reinterpret_cast<float2*>(xs + P.xs_len);         // padded FFT buffer, fft_buf_len(N) float2
    float* accb_smem = reinterpret_cast<float*>(buf + fft_buf_len(P.n2 >> 1));    // n_stats (or nothing: sums in global memory)
    const int n_stats_pad = (P.n_stats + 3) & ~3;
    float* accb = P.accb_global ? P.accb_global + (size_t)blockIdx.x * n_stats_pad : accb_smem;
    SimShared<NT>* sh = reinterpret_cast<SimShared<NT>*>(accb_smem + (P.accb_global ? 0 : n_stats_pad));
    float2* twtab = reinterpret_cast<float2*>(sh + 1);
    __shared__ double red[NT / 32];

    const int tid = threadIdx.x;
    const int64_t particle = blockIdx.x;
    const int N = P.n2 >> 1;

    const double tau = P.theta[particle];
    double freq = 0.0, coeff = 1.0;
    if (OSC) {
        freq = P.theta[particle + P.theta_ld];
        coeff = P.theta[particle + 2 * P.theta_ld];
    }
    const OuConsts oc = make_ou_consts(tau, P.dt, P.update, OSC ? ITJL_KIND_OU_OSC : ITJL_KIND_OU,
                                       freq, coeff, P.n_t, true, true);   // |FFT| of the Hamming-windowed trial is reversal invariant
    SimCache sim_cache;
    SimArgs g;
    g.n_t = P.n_t; g.chunk = P.chunk;
    g.inv_n = 1.0 / (double)P.n_t; g.sqrt_nm1 = sqrtf((float)(P.n_t - 1));
    g.keys = &P.keys; g.c3_noise = P.c3_noise; g.c3_init = P.c3_init;
    g.particle = (uint32_t)(P.particle_offset + particle);
    const size_t pt = (size_t)particle * P.n_trials;
    g.u0 = P.u0 ? P.u0 + pt : nullptr;
    g.noise = P.noise ? P.noise + pt * P.n_t : nullptr;
    g.phases = P.phases ? P.phases + pt : nullptr;
    g.mask_bits = nullptr; g.n_valid = nullptr; g.mask_words = 0;

    for (int b = tid; b < P.n_stats; b += NT) accb[b] = 0.f;
    const Tw2 tw = tw2_build(twtab, P.twiddle, N, tid, NT);          // visible after the first __syncthreads() below

    for (int trial = 0; trial < P.n_trials; ++trial) {
        // standardised series * data_sd; the reference then adds data_mean and
        // comp_psd_adfriendly removes the mean again (summary.jl:209) - skipped both.
        simulate_trial<NT, false, OSC>(g, oc, trial, xs, sh, kScaleStd, (float)P.data_sd, 0.0f, sim_cache);
        __syncthreads();
        // window, pack z[m] = x[2m] + i x[2m+1] into the padded FFT buffer, zero-pad to N points
        // (xs holds zeros from n_t to chunk * NT)
        // data part (window loads from L2, several in flight), then the zero padding (stores only)
        const int m_data = (P.n_t + 1) >> 1;
#pragma unroll 5
        for (int m = tid; m < m_data; m += NT) {
            const float2 v = *reinterpret_cast<const float2*>(xs + 2 * m);
            const float2 w = __ldg(reinterpret_cast<const float2*>(P.window) + m);   // window is padded to an even length
            buf[fft_pad(m)] = make_float2(v.x * w.x, (2 * m + 1 < P.n_t) ? v.y * w.y : 0.f);
        }
        const int m_zero = ((m_data + NT - 1) / NT) * NT;     // first multiple of NT at or above m_data
#pragma unroll 4
        for (int m = m_zero + tid - NT; m < N; m += NT)
            if (m >= m_data) buf[fft_pad(m)] = make_float2(0.f, 0.f);
        __syncthreads();
        fft_inplace_dif<NT>(buf, N, tw);
        // bin table rows come from L2 (16 B each, 100 KB per trial): four in flight
#pragma unroll 4
        for (int b = tid; b < P.n_stats; b += NT)
            accb[b] += real_fft_power_at(buf, __ldg(P.binpos + b));
        __syncthreads();
    }

    const double sc = P.psd_scale / (double)P.n_trials;
    double local = 0.0;
    for (int b = tid; b < P.n_stats; b += NT) {
        const double v = (double)accb[b] * sc;
        const int bi = __ldg(P.binidx + b);                      // accb is in storage order
        if (P.stats_out) P.stats_out[(size_t)particle * P.n_stats + bi] = v;
        const double tgt = (double)P.target[bi];
        const double e = (P.distance == ITJL_DIST_LINEAR) ? (v - tgt) : (log(v) - log(tgt));
        local += e * e;
    }
    local = warp_sum(local);
    if ((tid & 31) == 0) red[tid >> 5] = local;
    __syncthreads();
    if (tid == 0) {
        double d = 0.0;
#pragma unroll
        for (int w = 0; w < NT / 32; ++w) d += red[w];
        P.dist_out[particle] = d / (double)P.n_stats;
    }
}

// comp_psd_adfriendly on stored data: one CTA per slice, all n2/2 - 1 bins out.
__global__ void __launch_bounds__(kThreads, 1) k_psd_data(const double* __restrict__ data, int64_t n_slices,
                                                          int n_t, int n2, const float* __restrict__ window,
                                                          const float2* __restrict__ U, double psd_scale,
                                                          double* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* xs = reinterpret_cast<float*>(smem_raw);
    float2* buf = reinterpret_cast<float2*>(smem_raw);
    __shared__ double red[kWarps];
    const int tid = threadIdx.x;
    const int64_t s = blockIdx.x;
    const int N = n2 >> 1;
    const Tw2 tw = tw2_build(buf + fft_buf_len(N), U, N, tid, kThreads);
    double local = 0.0;
    for (int j = tid; j < n_t; j += kThreads) local += data[s + (size_t)j * n_slices];
    local = warp_sum(local);
    if ((tid & 31) == 0) red[tid >> 5] = local;
    __syncthreads();
    double mean = 0.0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) mean += red[w];
    mean /= (double)n_t;
    for (int j = tid; j < n2; j += kThreads) {
        const float v = (j < n_t) ? (float)((data[s + (size_t)j * n_slices] - mean) * (double)__ldg(window + j)) : 0.f;
        xs[2 * fft_pad(j >> 1) + (j & 1)] = v;
    }
    __syncthreads();
    fft_inplace_dif<kThreads>(buf, N, tw);
    for (int k = 1 + tid; k < N; k += kThreads)
        out[s + (size_t)(k - 1) * n_slices] = (double)real_fft_power(buf, k, N, U) * psd_scale;
}

int psd_fft_pos(int k, int N) { return fft_pad(fft_pos(k, N)); }      // position in the padded buffer
size_t psd_fft_smem_bytes(int n2) { return (size_t)(fft_buf_len(n2 / 2) + tw2_entries(n2 / 2)) * sizeof(float2); }

size_t abc_psd_smem_bytes(int n_t, int n2, int n_stats, int* chunk_out, int* xs_len_out, bool accb_in_smem) {
    int c = (n_t + kPsdThreads - 1) / kPsdThreads;
    c = (c + 3) & ~3;                                   // (<= kQtab: n_t <= 2^24 / ... is bounded by the FFT buffer long before)
    const int xs_len = c * kPsdThreads;                // staging buffer of one simulated trial
    *chunk_out = c;
    *xs_len_out = xs_len;
    return (size_t)xs_len * sizeof(float) + (size_t)fft_buf_len(n2 / 2) * sizeof(float2) +
           (accb_in_smem ? (size_t)((n_stats + 3) & ~3) * sizeof(float) : 0) + sizeof(SimShared<kPsdThreads>) +
           (size_t)tw2_entries(n2 / 2) * sizeof(float2);
}

void psd_make_tables(int n_t, int n2, float* window_host, float2* twiddle_host, double* win_ss) {
    const double PI = 3.14159265358979323846;
    double ss = 0.0;
    for (int i = 0; i < n_t; ++i) {
        const double w = 0.54 - 0.46 * cos(2.0 * PI * (double)i / (double)(n_t - 1));   // summary.jl:214
        window_host[i] = (float)w;
        ss += w * w;
    }
    *win_ss = ss;
    for (int k = 0; k < n2 / 2; ++k) {
        const double ang = -2.0 * PI * (double)k / (double)n2;
        twiddle_host[k] = make_float2((float)cos(ang), (float)sin(ang));
    }
}

cudaError_t abc_psd_launch(const AbcPsdParams& P, bool osc, size_t smem, cudaStream_t st) {
    void (*kern)(const AbcPsdParams) = osc ? k_abc_psd<true> : k_abc_psd<false>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<(unsigned)(P.grid > 0 ? P.grid : P.n_particles), kPsdThreads, smem, st>>>(P);
    return cudaGetLastError();
}

cudaError_t psd_data_launch(const double* data, int64_t n_slices, int n_t, int n2, int log2_n2,
                            const float* window, const float2* twiddle, double psd_scale, double* out,
                            size_t smem, cudaStream_t st) {
    (void)log2_n2;
    cudaError_t e = cudaFuncSetAttribute(k_psd_data, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    k_psd_data<<<(unsigned)n_slices, kThreads, smem, st>>>(data, n_slices, n_t, n2, window, twiddle, psd_scale, out);
    return cudaGetLastError();
}

}  // namespace itjl
