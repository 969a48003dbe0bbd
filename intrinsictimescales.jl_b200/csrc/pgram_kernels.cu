// K6: DSP.jl-convention periodogram (reference src/stats/summary.jl:146-155, `dsp.periodogram(x; fs, window =
// hamming)`), the summary statistic of one_timescale_model(summary_method = :psd)
// (src/models/one_timescale.jl:177-203, :255) and the spectrum behind acw(...; acwtypes = :knee)
// (src/core/acw.jl:244).  Conventions restated in oracle/summary.py::comp_psd_periodogram_vec: no detrending,
// symmetric Hamming window of the signal's length, nfft = nextfastfft(n) = the next 7-smooth integer (5000 for
// 10 s at 500 Hz - not a power of two), one-sided power |X|^2 / (fs sum w^2), doubled except at DC and Nyquist,
// DC bin dropped.
//
//   k_abc_pgram<OSC> : fused simulate -> periodogram -> trial mean -> distance, one CTA per particle.  Two trials
//                      share one nfft-point complex transform (trial A = real part, trial B = imaginary part):
//                      |X_A[k]|^2 + |X_B[k]|^2 = (|Z[k]|^2 + |Z[nfft - k]|^2) / 2, and only the trial SUM is needed.
//                      Mixed-radix Stockham FFT (fft_mixed.cuh) ping-ponging between two shared-memory buffers.
//   k_pgram_data     : the same spectrum of stored data in fp64 (two slices per transform, buffers in a per-CTA
//                      global scratch area that stays in L2), all nfft / 2 bins out.
//   k_lomb_scargle   : generalised Lomb-Scargle periodogram of stored data with NaN gaps (summary.jl:254-356)
#include <math.h>
#include <algorithm>

#include "fft_mixed.cuh"
#include "kernels.h"
#include "simulate.cuh"

namespace itjl {

constexpr int kPgramThreads = 512;

// Two CTAs per SM (the trial buffer lives in the second transform buffer, which is idle while a trial is simulated):
// one CTA's simulate phase runs under the other's transform.
template <bool OSC>
__global__ void __launch_bounds__(kPgramThreads, 2) k_abc_pgram(const __grid_constant__ AbcPgramParams P) {
    constexpr int NT = kPgramThreads;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Cx<float>* bufa = reinterpret_cast<Cx<float>*>(smem_raw);                // nfft complex points
    Cx<float>* bufb = bufa + ((P.plan.n + 1) & ~1);                          // second transform buffer, 16-byte aligned (bufb_len entries) ...
    float* xs = reinterpret_cast<float*>(bufb);                              // ... = chunk * NT floats: one simulated trial
    float* accb = reinterpret_cast<float*>(bufb + P.bufb_len);               // n_stats bin sums (padded to 4)
    SimShared<NT>* sh = reinterpret_cast<SimShared<NT>*>(accb + ((P.n_stats + 3) & ~3));
    Cx<float>* rtab = reinterpret_cast<Cx<float>*>(sh + 1);                  // two-level roots of unity
    __shared__ double red[NT / 32];

    const int tid = threadIdx.x;
    const int64_t particle = blockIdx.x;
    const int nfft = P.plan.n;

    const double tau = P.theta[particle];
    double freq = 0.0, coeff = 1.0;
    if (OSC) {
        freq = P.theta[particle + P.theta_ld];
        coeff = P.theta[particle + 2 * P.theta_ld];
    }
    // not reversible: the window is symmetric but no mean is removed by the periodogram - the standardised trial
    // has mean zero either way, so the reversed exponential is still the same statistic
    const OuConsts oc = make_ou_consts(tau, P.dt, P.update, OSC ? ITJL_KIND_OU_OSC : ITJL_KIND_OU, freq, coeff, P.n_t, true, true);
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

    // generate_ou_process(tau, data_sd, ...) has mean zero (one_timescale.jl:226-228); the oscillation generator adds data_mean
    const float shift = OSC ? (float)P.data_mean : 0.f;
    for (int b = tid; b < P.n_stats; b += NT) accb[b] = 0.f;
    const int n_hi = (nfft + (1 << kRootLoBits) - 1) >> kRootLoBits;
    for (int e = tid; e < n_hi + (1 << kRootLoBits); e += NT) {
        const float2 r = __ldg(P.roots + e);
        rtab[e] = {r.x, r.y};
    }
    const RootsTwoLevel<float> roots{rtab, rtab + n_hi};

    for (int trial = 0; trial < P.n_trials; trial += 2) {
        // trial A -> real parts, trial B -> imaginary parts (zero when n_trials is odd), both windowed, zero padded
        simulate_trial<NT, false, OSC>(g, oc, trial, xs, sh, kScaleStd, (float)P.data_sd, shift, sim_cache);
        __syncthreads();
        for (int m = tid; m < nfft; m += NT) bufa[m].x = m < P.n_t ? xs[m] * __ldg(P.window + m) : 0.f;
        __syncthreads();
        if (trial + 1 < P.n_trials) {
            simulate_trial<NT, false, OSC>(g, oc, trial + 1, xs, sh, kScaleStd, (float)P.data_sd, shift, sim_cache);
            __syncthreads();
            for (int m = tid; m < nfft; m += NT) bufa[m].y = m < P.n_t ? xs[m] * __ldg(P.window + m) : 0.f;
        } else {
            for (int m = tid; m < nfft; m += NT) bufa[m].y = 0.f;
        }
        __syncthreads();
        const Cx<float>* z = fft_mixed_cta<float>(P.plan, bufa, bufb, roots);
        for (int b = tid; b < P.n_stats; b += NT) {
            const int k = __ldg(P.bins + b);
            const Cx<float> zk = z[k], zn = z[nfft - k];
            accb[b] += fmaf(zk.x, zk.x, zk.y * zk.y) + fmaf(zn.x, zn.x, zn.y * zn.y);
        }
        __syncthreads();
    }

    // power = |X|^2 / (fs sum w^2), doubled except at Nyquist (DC is never selected); 0.5: the pair identity above
    const double sc = 0.5 * P.psd_scale / (double)P.n_trials;
    double local = 0.0;
    for (int b = tid; b < P.n_stats; b += NT) {
        const int k = __ldg(P.bins + b);
        const double v = (double)accb[b] * sc * ((2 * k == nfft) ? 1.0 : 2.0);
        if (P.stats_out) P.stats_out[(size_t)particle * P.n_stats + b] = v;
        const double tgt = (double)P.target[b];
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

size_t abc_pgram_smem_bytes(int n_t, int nfft, int n_stats, int* chunk_out, int* bufb_len_out) {
    int c = (n_t + kPgramThreads - 1) / kPgramThreads;
    c = (c + 3) & ~3;
    *chunk_out = c;
    // the second buffer also holds the simulated trial: max(nfft float2, chunk * NT floats)
    const int bufb = (std::max(nfft, (c * kPgramThreads + 1) / 2) + 1) & ~1;
    *bufb_len_out = bufb;
    return (size_t)(((nfft + 1) & ~1) + bufb) * sizeof(float2) +
           (size_t)((n_stats + 3) & ~3) * sizeof(float) + sizeof(SimShared<kPgramThreads>) +
           (size_t)roots_two_level_entries(nfft) * sizeof(float2);
}

void pgram_make_tables(int n_t, int nfft, float* window_f, double* window_d, float2* roots2_f, double2* roots_d, double* win_ss) {
    const double PI = 3.14159265358979323846;
    double ss = 0.0;
    for (int i = 0; i < n_t; ++i) {
        const double w = n_t > 1 ? 0.54 - 0.46 * cos(2.0 * PI * (double)i / (double)(n_t - 1)) : 1.0;   // DSP.hamming(n): symmetric
        if (window_f) window_f[i] = (float)w;
        if (window_d) window_d[i] = w;
        ss += w * w;
    }
    *win_ss = ss;
    if (roots2_f) {
        const int hi = (nfft + (1 << kRootLoBits) - 1) >> kRootLoBits;
        for (int h = 0; h < hi; ++h) {
            const double ang = -2.0 * PI * (double)(h << kRootLoBits) / (double)nfft;
            roots2_f[h] = make_float2((float)cos(ang), (float)sin(ang));
        }
        for (int l = 0; l < (1 << kRootLoBits); ++l) {
            const double ang = -2.0 * PI * (double)l / (double)nfft;
            roots2_f[hi + l] = make_float2((float)cos(ang), (float)sin(ang));
        }
    }
    if (roots_d)
        for (int k = 0; k < nfft; ++k) {
            const double ang = -2.0 * PI * (double)k / (double)nfft;
            roots_d[k] = make_double2(cos(ang), sin(ang));
        }
}

cudaError_t abc_pgram_launch(const AbcPgramParams& P, bool osc, size_t smem, cudaStream_t st) {
    void (*kern)(const AbcPgramParams) = osc ? k_abc_pgram<true> : k_abc_pgram<false>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<(unsigned)P.n_particles, kPgramThreads, smem, st>>>(P);
    return cudaGetLastError();
}

// ---- stored data, fp64 ----------------------------------------------------------------------------------------
constexpr int kPgramDataThreads = 256;

// CTA b transforms the slice pairs b, b + gridDim.x, ...; scratch: 2 nfft double2 per CTA.
// out: (n_slices x nfft/2) column-major, bin k = 1 .. nfft/2.
__global__ void __launch_bounds__(kPgramDataThreads) k_pgram_data(const double* __restrict__ data, int64_t n_slices, int n_t,
                                                                  const __grid_constant__ FftPlan plan,
                                                                  const double* __restrict__ window, const double2* __restrict__ roots_tab,
                                                                  double psd_scale, double2* __restrict__ scratch, double* __restrict__ out) {
    const int tid = threadIdx.x, nfft = plan.n;
    Cx<double>* a = reinterpret_cast<Cx<double>*>(scratch) + (size_t)blockIdx.x * 2 * nfft;
    Cx<double>* b = a + nfft;
    const RootsTable<double> roots{reinterpret_cast<const Cx<double>*>(roots_tab)};
    const int64_t n_pairs = (n_slices + 1) / 2;
    for (int64_t pr = blockIdx.x; pr < n_pairs; pr += gridDim.x) {
        const int64_t s0 = 2 * pr, s1 = s0 + 1;
        for (int m = tid; m < nfft; m += kPgramDataThreads) {
            Cx<double> v = {0.0, 0.0};
            if (m < n_t) {
                const double w = window[m];
                v.x = data[s0 + (size_t)m * n_slices] * w;
                if (s1 < n_slices) v.y = data[s1 + (size_t)m * n_slices] * w;
            }
            a[m] = v;
        }
        __syncthreads();
        const Cx<double>* z = fft_mixed_cta<double>(plan, a, b, roots);
        for (int k = 1 + tid; k <= nfft / 2; k += kPgramDataThreads) {
            const Cx<double> zk = z[k], zn = z[nfft - k];
            // X_A = (Z[k] + conj Z[n-k]) / 2,  X_B = (Z[k] - conj Z[n-k]) / 2i
            const double ar = 0.5 * (zk.x + zn.x), ai = 0.5 * (zk.y - zn.y);
            const double br = 0.5 * (zk.y + zn.y), bi = -0.5 * (zk.x - zn.x);
            const double f = psd_scale * ((2 * k == nfft) ? 1.0 : 2.0);
            out[s0 + (size_t)(k - 1) * n_slices] = (ar * ar + ai * ai) * f;
            if (s1 < n_slices) out[s1 + (size_t)(k - 1) * n_slices] = (br * br + bi * bi) * f;
        }
        __syncthreads();
    }
}

cudaError_t pgram_data_launch(const double* data, int64_t n_slices, int n_t, const FftPlan& plan, const double* window,
                              const double2* roots, double psd_scale, double2* scratch, int grid, double* out, cudaStream_t st) {
    k_pgram_data<<<grid, kPgramDataThreads, 0, st>>>(data, n_slices, n_t, plan, window, roots, psd_scale, scratch, out);
    return cudaGetLastError();
}

// ---- Lomb-Scargle -----------------------------------------------------------------------------------------------
// Generalised Lomb-Scargle periodogram (floating mean, standard normalisation: what LombScargle.jl's
// `lombscargle(plan(times, signal; frequencies))` computes with its defaults fit_mean = true, center_data = true,
// normalization = :standard) of every slice with NaN = missing, on the common grid f_k = f0 + k df.  Samples sit
// on the regular grid t_j = (j + 1) dt, so the phasor of frequency k advances by a fixed rotation per sample
// (fp64 recurrence, re-seeded from sincospi every 256 samples).  Thread = frequency, CTA = (slice, block of
// frequencies); the slice (centred, NaN -> flag) is staged in shared memory in tiles.
constexpr int kLsThreads = 128;
constexpr int kLsTile = 1024;

__global__ void __launch_bounds__(kLsThreads) k_lomb_scargle(const double* __restrict__ data, int64_t n_slices, int n_t, double dt,
                                                             double f0, double df, int n_freq, double* __restrict__ out) {
    __shared__ double ys[kLsTile];
    __shared__ double red[2][kLsThreads / 32];
    const int tid = threadIdx.x;
    const int64_t s = blockIdx.y;
    const int k = blockIdx.x * kLsThreads + tid;
    // mean over the valid samples (center_data)
    double sum = 0.0, cnt = 0.0;
    for (int j = tid; j < n_t; j += kLsThreads) {
        const double v = data[s + (size_t)j * n_slices];
        if (v == v) { sum += v; cnt += 1.0; }
    }
    sum = warp_sum(sum); cnt = warp_sum(cnt);
    if ((tid & 31) == 0) { red[0][tid >> 5] = sum; red[1][tid >> 5] = cnt; }
    __syncthreads();
    sum = 0.0; cnt = 0.0;
#pragma unroll
    for (int w = 0; w < kLsThreads / 32; ++w) { sum += red[0][w]; cnt += red[1][w]; }
    const double mean = sum / cnt;

    const double f = f0 + df * (double)k;
    const double turn = f * dt;                       // turns per sample
    double rs, rc;
    sincospi(2.0 * turn, &rs, &rc);                   // rotation per sample
    double C = 0.0, S = 0.0, YC = 0.0, YS = 0.0, CC = 0.0, CS = 0.0, YY = 0.0;
    for (int t0 = 0; t0 < n_t; t0 += kLsTile) {
        __syncthreads();
        for (int j = tid; j < kLsTile && t0 + j < n_t; j += kLsThreads) ys[j] = data[s + (size_t)(t0 + j) * n_slices] - mean;   // NaN stays NaN
        __syncthreads();
        const int lim = min(kLsTile, n_t - t0);
        double c = 1.0, sn = 0.0;
        for (int j = 0; j < lim; ++j) {
            if ((j & 255) == 0) {
                double ph = turn * (double)(t0 + j + 1);     // t = (j + 1) dt
                ph -= floor(ph);
                sincospi(2.0 * ph, &sn, &c);
            }
            const double y = ys[j];
            if (y == y) {
                C += c; S += sn; YC = fma(y, c, YC); YS = fma(y, sn, YS);
                CC = fma(c, c, CC); CS = fma(c, sn, CS); YY = fma(y, y, YY);
            }
            const double c2 = c * rc - sn * rs;
            sn = fma(sn, rc, c * rs);
            c = c2;
        }
    }
    if (k >= n_freq) return;
    // weights 1 / n_valid; Y = 0 after centring
    const double w = 1.0 / cnt;
    C *= w; S *= w; YC *= w; YS *= w; CC *= w; CS *= w; YY *= w;
    const double SS = (1.0 - CC) - S * S;
    CC -= C * C; CS -= C * S;
    const double D = CC * SS - CS * CS;
    double pw = (SS * YC * YC + CC * YS * YS - 2.0 * CS * YC * YS) / (YY * D);
    // degenerate frequencies (exactly Nyquist on the regular grid: every sin(omega t) vanishes, D = 0 / 0): the model
    // has one basis function left - the power is the projection on it
    if (!(D > 1e-12 * (CC + SS) * (CC + SS))) pw = SS < CC ? YC * YC / (YY * CC) : YS * YS / (YY * SS);
    out[s + (size_t)k * n_slices] = pw;
}

cudaError_t lomb_scargle_launch(const double* data, int64_t n_slices, int n_t, double dt, double f0, double df, int n_freq,
                                double* out, cudaStream_t st) {
    dim3 grid((unsigned)((n_freq + kLsThreads - 1) / kLsThreads), (unsigned)n_slices);
    k_lomb_scargle<<<grid, kLsThreads, 0, st>>>(data, n_slices, n_t, dt, f0, df, n_freq, out);
    return cudaGetLastError();
}


// ---- Lomb-Scargle by chirp-z transforms --------------------------------------------------------------------------
// The samples sit on the regular grid t = (k + 1) dt with gaps, so the five sums of the periodogram are values of
//   A_j = sum_k c_k e^{-i w_j t_k},   B_j = sum_k m_k e^{-i w_j t_k},   C_j = sum_k m_k e^{-2 i w_j t_k}
// (c = centred data with zeros in the gaps, m = the mask of valid samples, w_j = 2 pi (f0 + j df)): three chirp-z
// (Bluestein) transforms of length-n_t sequences to n_freq equispaced frequencies.  With p = k + 1 and
// a = f0 dt, b = df dt (in turns):  phase(p, j) = a p + b j p = a p + b (p^2 + j^2 - (j - p)^2) / 2, so
//   A_j = e^{-2 pi i b j^2 / 2} * sum_p [c_p e^{-2 pi i (a p + b p^2 / 2)}] * e^{+2 pi i b (j - p)^2 / 2}
// - a convolution, done with two FFTs of a 7-smooth length M >= n_t + n_freq per transform (mixed-radix Stockham,
// fft_mixed.cuh, fp64, buffers in a per-CTA global scratch area) against the precomputed transform of the chirp
// filter.  O(M log M) per slice instead of O(n_t n_freq): 2000 x 5000 samples x 12 498 frequencies: 94 -> ~10 ms.
constexpr int kCztThreads = 256;

struct LsCztParams {
    const double* data; int64_t n_slices; int n_t, n_freq;
    double a_turn, b_turn;                // f0 dt, df dt
    FftPlan plan;                         // length M
    const double2* roots;                 // M-th roots of unity e^{-2 pi i k / M}
    double2* chirp_in;                    // [2][n_t]   e^{-2 pi i (a p + b p^2 / 2)} and the same for (2a, 2b), p = k + 1
    double2* filt;                        // [2][M]     FFT of v1[l] = e^{+2 pi i b l^2 / 2}, v2[l] = e^{+2 pi i b l^2} (l = -n_t .. n_freq - 1, circular)
    double2* chirp_out;                   // [2][n_freq] e^{-2 pi i b j^2 / 2}, e^{-2 pi i b j^2}
    double2* scratch;                     // per CTA: 2 M (transform buffers) + 2 n_freq (A, B)
    double* out;                          // (n_slices x n_freq) column-major
};

// e^{2 pi i t} for a phase of t turns (|t| up to ~1e5: reduced before sincospi)
__device__ __forceinline__ Cx<double> cis_turns(double t) {
    t -= floor(t);
    double s, c;
    sincospi(2.0 * t, &s, &c);
    return {c, s};
}

// one CTA: the chirps and the transformed filters
__global__ void __launch_bounds__(kCztThreads) k_lomb_czt_prepare(const __grid_constant__ LsCztParams P) {
    const int tid = threadIdx.x, M = P.plan.n, n = P.n_t, J = P.n_freq;
    const double a = P.a_turn, b = P.b_turn;
    for (int k = tid; k < n; k += kCztThreads) {
        const double p = (double)(k + 1);
        const Cx<double> c1 = cis_turns(-(a * p + 0.5 * b * p * p)), c2 = cis_turns(-(2.0 * a * p + b * p * p));
        P.chirp_in[k] = make_double2(c1.x, c1.y);
        P.chirp_in[n + k] = make_double2(c2.x, c2.y);
    }
    for (int j = tid; j < J; j += kCztThreads) {
        const double q = (double)j;
        const Cx<double> c1 = cis_turns(-0.5 * b * q * q), c2 = cis_turns(-b * q * q);
        P.chirp_out[j] = make_double2(c1.x, c1.y);
        P.chirp_out[J + j] = make_double2(c2.x, c2.y);
    }
    Cx<double>* x = reinterpret_cast<Cx<double>*>(P.scratch);
    Cx<double>* y = x + M;
    const RootsTable<double> roots{reinterpret_cast<const Cx<double>*>(P.roots)};
    for (int which = 0; which < 2; ++which) {
        const double bb = which == 0 ? 0.5 * b : b;
        __syncthreads();
        for (int i = tid; i < M; i += kCztThreads) {
            // circular index i <-> lag l: l = i for i < n_freq, l = i - M for i >= M - n_t, nothing in between
            Cx<double> v = {0.0, 0.0};
            if (i < J) { const double l = (double)i; v = cis_turns(bb * l * l); }
            else if (i >= M - n) { const double l = (double)(i - M); v = cis_turns(bb * l * l); }
            x[i] = v;
        }
        __syncthreads();
        const Cx<double>* z = fft_mixed_cta<double>(P.plan, x, y, roots);
        for (int i = tid; i < M; i += kCztThreads) P.filt[(size_t)which * M + i] = make_double2(z[i].x, z[i].y);
    }
}

__global__ void __launch_bounds__(kCztThreads) k_lomb_czt(const __grid_constant__ LsCztParams P) {
    __shared__ double red[3][kCztThreads / 32];
    const int tid = threadIdx.x, M = P.plan.n, n = P.n_t, J = P.n_freq;
    Cx<double>* buf0 = reinterpret_cast<Cx<double>*>(P.scratch) + (size_t)blockIdx.x * (2 * (size_t)M + 2 * (size_t)J);
    Cx<double>* buf1 = buf0 + M;
    Cx<double>* SA = buf1 + M;
    Cx<double>* SB = SA + J;
    const RootsTable<double> roots{reinterpret_cast<const Cx<double>*>(P.roots)};
    const double inv_m = 1.0 / (double)M;
    for (int64_t s = blockIdx.x; s < P.n_slices; s += gridDim.x) {
        const double* col = P.data + s;
        const size_t ld = (size_t)P.n_slices;
        // mean over the valid samples (center_data), their count, and sum c^2
        double sum = 0.0, cnt = 0.0;
        for (int k = tid; k < n; k += kCztThreads) { const double v = col[(size_t)k * ld]; if (v == v) { sum += v; cnt += 1.0; } }
        sum = warp_sum(sum); cnt = warp_sum(cnt);
        __syncthreads();
        if ((tid & 31) == 0) { red[0][tid >> 5] = sum; red[1][tid >> 5] = cnt; }
        __syncthreads();
        sum = 0.0; cnt = 0.0;
#pragma unroll
        for (int w = 0; w < kCztThreads / 32; ++w) { sum += red[0][w]; cnt += red[1][w]; }
        const double mean = sum / cnt;
        double yy = 0.0;
        for (int k = tid; k < n; k += kCztThreads) { const double v = col[(size_t)k * ld]; if (v == v) yy = fma(v - mean, v - mean, yy); }
        yy = warp_sum(yy);
        if ((tid & 31) == 0) red[2][tid >> 5] = yy;
        __syncthreads();
        yy = 0.0;
#pragma unroll
        for (int w = 0; w < kCztThreads / 32; ++w) yy += red[2][w];

        for (int tr = 0; tr < 3; ++tr) {
            const double2* cin = P.chirp_in + (tr == 2 ? n : 0);
            const double2* flt = P.filt + (tr == 2 ? (size_t)M : 0);
            const double2* cout = P.chirp_out + (tr == 2 ? J : 0);
            // u[p] = value_{p-1} * chirp_in[p-1] at index p = 1 .. n_t, zero elsewhere
            for (int i = tid; i < M; i += kCztThreads) {
                Cx<double> u = {0.0, 0.0};
                if (i >= 1 && i <= n) {
                    const double v = col[(size_t)(i - 1) * ld];
                    if (v == v) {
                        const double val = tr == 0 ? v - mean : 1.0;
                        const double2 c = cin[i - 1];
                        u = {val * c.x, val * c.y};
                    }
                }
                buf0[i] = u;
            }
            __syncthreads();
            Cx<double>* z = fft_mixed_cta<double>(P.plan, buf0, buf1, roots);
            Cx<double>* other = z == buf0 ? buf1 : buf0;
            // inverse transform of z * filt by conjugation: ifft(x) = conj(fft(conj(x))) / M
            for (int i = tid; i < M; i += kCztThreads) {
                const double2 f = flt[i];
                const Cx<double> pz = z[i];
                z[i] = {pz.x * f.x - pz.y * f.y, -(pz.x * f.y + pz.y * f.x)};
            }
            __syncthreads();
            const Cx<double>* w = fft_mixed_cta<double>(P.plan, z, other, roots);
            for (int j = tid; j < J; j += kCztThreads) {
                const double wr = w[j].x * inv_m, wi = -w[j].y * inv_m;
                const double2 c = cout[j];
                const Cx<double> r = {wr * c.x - wi * c.y, wr * c.y + wi * c.x};
                if (tr == 0) SA[j] = r;
                else if (tr == 1) SB[j] = r;
                else {
                    // sums of the periodogram (weights 1 / n_valid, Y = 0 after centring): cos <-> Re, sin <-> -Im
                    const double wgt = 1.0 / cnt;
                    const Cx<double> A = SA[j], B = SB[j];
                    const double C = B.x * wgt, S = -B.y * wgt, YC = A.x * wgt, YS = -A.y * wgt, YY = yy * wgt;
                    double CC = 0.5 * (cnt + r.x) * wgt, CS = -0.5 * r.y * wgt;
                    const double SS = (1.0 - CC) - S * S;
                    CC -= C * C; CS -= C * S;
                    const double D = CC * SS - CS * CS;
                    double pw = (SS * YC * YC + CC * YS * YS - 2.0 * CS * YC * YS) / (YY * D);
                    if (!(D > 1e-12 * (CC + SS) * (CC + SS))) pw = SS < CC ? YC * YC / (YY * CC) : YS * YS / (YY * SS);
                    P.out[s + (size_t)j * ld] = pw;
                }
            }
            __syncthreads();
        }
    }
}

// chirp-z plan: transform length (0 when the direct kernel has to be used), device memory needs in double2 units
int lomb_czt_length(int n_t, int n_freq) {
    if (n_freq < 64 || (long long)n_t + n_freq > 65536) return 0;
    const int M = next_fast_len(n_t + n_freq);
    FftPlan plan;
    return (M <= 65536 && fft_plan_make(M, &plan)) ? M : 0;
}
size_t lomb_czt_table_elems(int n_t, int n_freq, int M) { return (size_t)M + 2 * (size_t)n_t + 2 * (size_t)M + 2 * (size_t)n_freq; }
size_t lomb_czt_scratch_elems(int n_freq, int M, int grid) { return (size_t)grid * (2 * (size_t)M + 2 * (size_t)n_freq); }

// tables: [roots M | chirp_in 2 n_t | filt 2 M | chirp_out 2 n_freq] (roots are written by the caller)
cudaError_t lomb_czt_launch(const double* data, int64_t n_slices, int n_t, double dt, double f0, double df, int n_freq, int M,
                            double2* tables, double2* scratch, int grid, double* out, cudaStream_t st) {
    LsCztParams P{};
    P.data = data; P.n_slices = n_slices; P.n_t = n_t; P.n_freq = n_freq; P.a_turn = f0 * dt; P.b_turn = df * dt;
    if (!fft_plan_make(M, &P.plan)) return cudaErrorInvalidValue;
    P.roots = tables; P.chirp_in = tables + M; P.filt = P.chirp_in + 2 * (size_t)n_t; P.chirp_out = P.filt + 2 * (size_t)M;
    P.scratch = scratch; P.out = out;
    k_lomb_czt_prepare<<<1, kCztThreads, 0, st>>>(P);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    k_lomb_czt<<<grid, kCztThreads, 0, st>>>(P);
    return cudaGetLastError();
}

}  // namespace itjl
