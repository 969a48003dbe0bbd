// K7: the PSD-side reducers, batched - Lorentzian knee (find_knee_frequency), oscillation peak
// (find_oscillation_peak), Gaussian peak fit (fit_gaussian) and the FOOOF-style loop (fooof_fit) of reference
// src/utils/utils.jl:359-445, 479-538, 598-675, 721-838.  One CTA per spectrum (row); rows are contiguous in memory.
//
// The reference's residual functions fill the whole residual vector with mean(sqrt(abs2(model - psd))), so its
// Levenberg-Marquardt solve minimises the MEAN ABSOLUTE deviation.  NonlinearSolve's iterates are not restated
// (third-party, see oracle/knee.py): the kernel returns the argmin of that objective reached from the reference's
// initial guess by a deterministic Nelder-Mead search in (log amplitude, log knee) [(amplitude, centre, log sd) for
// the Gaussian], every objective evaluation a fixed-order block reduction in fp64 - the same answer on every run and
// every GPU.  The knee is confined to (0, max(freqs)]: beyond the observed band it is not identifiable, and the
// reference's own edge-case tests (flat / white spectra -> freqs[end], test_utils.jl:72-82) expect exactly that.
#include <math.h>

#include "kernels.h"

namespace itjl {

// 64 threads per row, 16 CTAs per SM (64 registers): an objective evaluation is a short latency-bound loop + reduction,
// so many small CTAs beat few wide ones - 20 000 spectra x 994 bins, knee only: 119 ms (256 threads, 2 CTAs per SM, one
// partial sum per thread) -> 66 ms (64 threads, four independent partial sums) -> 53 ms (16 CTAs per SM);
// benchmarks/knee_time.py
constexpr int kKneeThreads = 64;

struct KneeShared {
    double red[kKneeThreads / 32];
    double bcast[4];
    int ibcast[4];
};

__device__ __forceinline__ double block_sum(double v, KneeShared* sh) {
    v = warp_sum(v);
    __syncthreads();                                     // previous result consumed
    if ((threadIdx.x & 31) == 0) sh->red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < kKneeThreads / 32; ++w) t += sh->red[w];
    return t;
}
__device__ __forceinline__ double block_min(double v, KneeShared* sh) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh->red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = sh->red[0];
#pragma unroll
    for (int w = 1; w < kKneeThreads / 32; ++w) t = fmin(t, sh->red[w]);
    return t;
}
__device__ __forceinline__ double block_max(double v, KneeShared* sh) { return -block_min(-v, sh); }
// largest / smallest index with a flag set (-1 / INT_MAX when none)
__device__ __forceinline__ int block_max_int(int v, KneeShared* sh) { return (int)block_max((double)v, sh); }
__device__ __forceinline__ int block_min_int(int v, KneeShared* sh) { return (int)block_min((double)v, sh); }

// 1 / d for d >= 1 (finite): hardware seed (~20 bits) + two Newton steps, within ~1 ulp of the correctly rounded
// quotient - the objective loop is bound by the FP64 pipe and a full-precision division costs about twice as many
// DP operations (no special cases can occur here: d = 1 + (f / knee)^2)
__device__ __forceinline__ double rcp_ge1(double d) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
    r = fma(r, fma(-d, r, 1.0), r);
    r = fma(r, fma(-d, r, 1.0), r);
    return r;
}

// mean |model(f_i; x) - y_i| over the row; MODEL 0: Lorentzian x = (log amp, log knee), knee clamped to fmax;
// MODEL 1: Gaussian x = (amp, centre, log sd)
template <int MODEL>
__device__ __forceinline__ double l1_objective(const double* __restrict__ y, const double* __restrict__ f, int n, const double* x,
                                               double f_hi, KneeShared* sh) {
    // four independent partial sums per thread: the divisions (exps) of four elements overlap instead of queueing behind
    // one another's ~100-cycle dependent chains
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    constexpr int T = kKneeThreads;
    int i = threadIdx.x;
    if (MODEL == 0) {
        const double amp = exp(x[0]), ik = 1.0 / fmin(exp(x[1]), f_hi);
        for (; i + 3 * T < n; i += 4 * T) {
            const double t0 = f[i] * ik, t1 = f[i + T] * ik, t2 = f[i + 2 * T] * ik, t3 = f[i + 3 * T] * ik;
            const double y0 = y[i], y1 = y[i + T], y2 = y[i + 2 * T], y3 = y[i + 3 * T];
            a0 += fabs(fma(amp, rcp_ge1(fma(t0, t0, 1.0)), -y0));
            a1 += fabs(fma(amp, rcp_ge1(fma(t1, t1, 1.0)), -y1));
            a2 += fabs(fma(amp, rcp_ge1(fma(t2, t2, 1.0)), -y2));
            a3 += fabs(fma(amp, rcp_ge1(fma(t3, t3, 1.0)), -y3));
        }
        for (; i < n; i += T) {
            const double t = f[i] * ik;
            a0 += fabs(fma(amp, rcp_ge1(fma(t, t, 1.0)), -y[i]));
        }
    } else {
        // Gaussian: beyond 10 sd of the centre the model is < 2e-22 of its amplitude and |model - y| == |y| to the last
        // bit, so only the bins inside that window are visited; the rest is known: f_hi carries sum |y| of the row
        // (freqs ascending - checked by the callers - so the window is an index range found by bisection)
        const double amp = x[0], c = x[1], sd = exp(x[2]), h = -0.5 / (sd * sd), r = 10.0 * sd;
        int lo = 0, hi = n;
        if (r == r && c == c) {
            int a = 0, b = n;                                   // first index with f >= c - r
            while (a < b) { const int m = (a + b) >> 1; if (f[m] < c - r) a = m + 1; else b = m; }
            lo = a; b = n;                                      // first index with f > c + r
            while (a < b) { const int m = (a + b) >> 1; if (f[m] <= c + r) a = m + 1; else b = m; }
            hi = a;
        }
        i = lo + (int)threadIdx.x;
        for (; i + 3 * T < hi; i += 4 * T) {
            const double d0 = f[i] - c, d1 = f[i + T] - c, d2 = f[i + 2 * T] - c, d3 = f[i + 3 * T] - c;
            const double y0 = y[i], y1 = y[i + T], y2 = y[i + 2 * T], y3 = y[i + 3 * T];
            a0 += fabs(amp * exp(h * d0 * d0) - y0) - fabs(y0);
            a1 += fabs(amp * exp(h * d1 * d1) - y1) - fabs(y1);
            a2 += fabs(amp * exp(h * d2 * d2) - y2) - fabs(y2);
            a3 += fabs(amp * exp(h * d3 * d3) - y3) - fabs(y3);
        }
        for (; i < hi; i += T) {
            const double d = f[i] - c, yi = y[i];
            a0 += fabs(amp * exp(h * d * d) - yi) - fabs(yi);
        }
        return (f_hi + block_sum((a0 + a1) + (a2 + a3), sh)) / (double)n;
    }
    return block_sum((a0 + a1) + (a2 + a3), sh) / (double)n;
}

// Nelder-Mead (reflection 1, expansion 2, contraction 1/2, shrink 1/2), run redundantly by every thread of the CTA on
// identical values; five rounds with initial simplex steps step, step / 10, ... step / 10^4 from the best point so far
// (a kinked L1 objective can stall a simplex: restarts from the best point get it moving again).
template <int MODEL, int ND>
__device__ void nelder_mead(const double* y, const double* f, int n, double* x, const double* step, double f_hi, KneeShared* sh) {
    double best[ND], fbest = 0.0;
    for (int d = 0; d < ND; ++d) best[d] = x[d];
    for (int round = 0; round < 5; ++round) {
        const double sc = round == 0 ? 1.0 : (round == 1 ? 0.1 : (round == 2 ? 0.01 : (round == 3 ? 1e-3 : 1e-4)));
        double P[ND + 1][ND], F[ND + 1];
        for (int v = 0; v <= ND; ++v) {
            for (int d = 0; d < ND; ++d) P[v][d] = best[d] + ((v == d + 1) ? sc * step[d] : 0.0);
            F[v] = l1_objective<MODEL>(y, f, n, P[v], f_hi, sh);
        }
        if (round == 0) fbest = F[0];
        for (int it = 0; it < 600; ++it) {
            // order: lo = best, hi = worst, nh = second worst
            int lo = 0, hi = 0;
            for (int v = 1; v <= ND; ++v) { if (F[v] < F[lo]) lo = v; if (F[v] > F[hi]) hi = v; }
            int nh = lo;
            for (int v = 0; v <= ND; ++v) if (v != hi && F[v] > F[nh]) nh = v;
            double size = 0.0;
            for (int v = 0; v <= ND; ++v) for (int d = 0; d < ND; ++d) size = fmax(size, fabs(P[v][d] - P[lo][d]));
            if (!(F[hi] - F[lo] > 1e-15 * fabs(F[lo])) && size < 1e-9) break;
            if (size < 1e-13) break;
            double c[ND], xr[ND];
            for (int d = 0; d < ND; ++d) {
                c[d] = 0.0;
                for (int v = 0; v <= ND; ++v) if (v != hi) c[d] += P[v][d];
                c[d] /= (double)ND;
                xr[d] = c[d] + (c[d] - P[hi][d]);
            }
            const double fr = l1_objective<MODEL>(y, f, n, xr, f_hi, sh);
            if (fr < F[lo]) {
                double xe[ND];
                for (int d = 0; d < ND; ++d) xe[d] = c[d] + 2.0 * (c[d] - P[hi][d]);
                const double fe = l1_objective<MODEL>(y, f, n, xe, f_hi, sh);
                if (fe < fr) { for (int d = 0; d < ND; ++d) P[hi][d] = xe[d]; F[hi] = fe; }
                else { for (int d = 0; d < ND; ++d) P[hi][d] = xr[d]; F[hi] = fr; }
            } else if (fr < F[nh]) {
                for (int d = 0; d < ND; ++d) P[hi][d] = xr[d];
                F[hi] = fr;
            } else {
                double xc[ND];
                const bool outside = fr < F[hi];
                for (int d = 0; d < ND; ++d) xc[d] = outside ? c[d] + 0.5 * (xr[d] - c[d]) : c[d] + 0.5 * (P[hi][d] - c[d]);
                const double fc = l1_objective<MODEL>(y, f, n, xc, f_hi, sh);
                if (fc < (outside ? fr : F[hi])) { for (int d = 0; d < ND; ++d) P[hi][d] = xc[d]; F[hi] = fc; }
                else {
                    for (int v = 0; v <= ND; ++v) {
                        if (v == lo) continue;
                        for (int d = 0; d < ND; ++d) P[v][d] = P[lo][d] + 0.5 * (P[v][d] - P[lo][d]);
                        F[v] = l1_objective<MODEL>(y, f, n, P[v], f_hi, sh);
                    }
                }
            }
        }
        int lo = 0;
        for (int v = 1; v <= ND; ++v) if (F[v] < F[lo]) lo = v;
        if (F[lo] <= fbest) { fbest = F[lo]; for (int d = 0; d < ND; ++d) best[d] = P[lo][d]; }
    }
    for (int d = 0; d < ND; ++d) x[d] = best[d];
}

// find_knee_frequency on y[0..n): out = (amp, knee); NaN when the initial guess is not positive / finite
__device__ void knee_fit_row(const double* y, const double* f, int n, double min_freq, double fmaxv, double* amp_out, double* knee_out,
                             KneeShared* sh, bool fit = true) {
    const int tid = threadIdx.x;
    // lorentzian_initial_guess (utils.jl:423-445): amplitude = mean over freqs <= 2 min_freq; knee = last frequency whose
    // power is at least half of it, else the middle of the band
    double s = 0.0, c = 0.0, bad = 0.0;
    for (int i = tid; i < n; i += kKneeThreads) {
        if (f[i] <= 2.0 * min_freq) { s += y[i]; c += 1.0; }
        if (!isfinite(y[i])) bad = 1.0;
    }
    s = block_sum(s, sh); c = block_sum(c, sh); bad = block_sum(bad, sh);
    const double amp0 = s / c;                                               // NaN when no frequency qualifies, as mean([]) is
    int last = -1;
    for (int i = tid; i < n; i += kKneeThreads) if (y[i] >= 0.5 * amp0) last = i;      // i increases: the thread's largest
    last = block_max_int(last, sh);
    const double knee0 = last >= 0 ? f[last] : 0.5 * (f[0] + f[n - 1]);
    if (!fit) { *amp_out = amp0; *knee_out = knee0; return; }            // lorentzian_initial_guess itself
    if (!isfinite(amp0) || bad > 0.0) { *amp_out = NAN; *knee_out = NAN; return; }
    // a non-positive amplitude guess (white / sign-changing input): no Lorentzian to descend to in (log amp, log knee);
    // the initial guess stands - the reference's own edge-case test expects knee ~ freqs[end] there (test_utils.jl:80-81)
    if (!(amp0 > 0.0) || !(knee0 > 0.0)) { *amp_out = amp0; *knee_out = knee0; return; }
    double x[2] = {log(amp0), log(knee0)};
    const double step[2] = {0.1, 0.1};
    nelder_mead<0, 2>(y, f, n, x, step, fmaxv, sh);
    *amp_out = exp(x[0]);
    *knee_out = fmin(exp(x[1]), fmaxv);
}

// find_oscillation_peak (utils.jl:721-765) on y[0..n) within [min_freq, max_freq]; pre / suf: scratch of n doubles
__device__ double osc_peak_row(const double* y, const double* f, int n, double min_freq, double max_freq, double ratio,
                               double* pre, double* suf, KneeShared* sh) {
    const int tid = threadIdx.x;
    // masked window [a, b): freqs ascending
    int a = n, b = -1;
    for (int i = tid; i < n; i += kKneeThreads) if (f[i] >= min_freq && f[i] <= max_freq) { a = min(a, i); b = max(b, i); }
    a = block_min_int(a, sh); b = block_max_int(b, sh) + 1;
    const int m = b - a;
    if (m < 3) return NAN;
    // one-sided running minima: thread t owns a contiguous chunk
    const int per = (m + kKneeThreads - 1) / kKneeThreads;
    const int lo = min(tid * per, m), hi = min(lo + per, m);
    double cmin = INFINITY, cmax = -INFINITY;
    for (int i = lo; i < hi; ++i) { cmin = fmin(cmin, y[a + i]); cmax = fmax(cmax, y[a + i]); }
    __shared__ double chunk_min[kKneeThreads];
    __syncthreads();
    chunk_min[tid] = cmin;
    const double gmax = block_max(cmax, sh);
    __syncthreads();
    double run = INFINITY;
    for (int t = 0; t < tid; ++t) run = fmin(run, chunk_min[t]);
    for (int i = lo; i < hi; ++i) { run = fmin(run, y[a + i]); pre[i] = run; }
    run = INFINITY;
    for (int t = kKneeThreads - 1; t > tid; --t) run = fmin(run, chunk_min[t]);
    for (int i = hi - 1; i >= lo; --i) { run = fmin(run, y[a + i]); suf[i] = run; }
    __syncthreads();
    // most prominent strict local maximum (first one on ties, as argmax does)
    const double need = gmax * ratio;
    double bp = -INFINITY; int bi = 0x7fffffff;
    for (int i = 1 + tid; i < m - 1; i += kKneeThreads) {
        const double v = y[a + i];
        if (v > y[a + i - 1] && v > y[a + i + 1]) {
            const double p = v - fmax(pre[i], suf[i]);
            if (p >= need && (p > bp || (p == bp && i < bi))) { bp = p; bi = i; }
        }
    }
    const double gp = block_max(bp, sh);
    if (!(gp > -INFINITY)) return NAN;
    const int gi = block_min_int(bp == gp ? bi : 0x7fffffff, sh);
    return f[a + gi];
}

// fit_gaussian (utils.jl:813-838) on y[0..n): g = (amp, centre, sd)
__device__ void gaussian_fit_row(const double* y, const double* f, int n, double peak, double min_freq, double max_freq, double* g,
                                 KneeShared* sh) {
    const int tid = threadIdx.x;
    // argmin |f - peak| (first)
    double bd = INFINITY; int bi = 0x7fffffff;
    for (int i = tid; i < n; i += kKneeThreads) { const double d = fabs(f[i] - peak); if (d < bd) { bd = d; bi = i; } }
    const double gd = block_min(bd, sh);
    const int k = block_min_int(bd == gd ? bi : 0x7fffffff, sh);
    const double amp0 = y[k], half = 0.5 * amp0;
    int left = -1, right = 0x7fffffff;
    for (int i = tid; i < n; i += kKneeThreads) {
        if (i <= k && y[i] <= half) left = max(left, i);
        if (i >= k && y[i] <= half) right = min(right, i);
    }
    left = block_max_int(left, sh); right = block_min_int(right, sh);
    const double sd0 = (left < 0 || right == 0x7fffffff) ? (max_freq - min_freq) / 10.0 : (f[right] - f[left]) / 2.355;
    if (!(sd0 > 0.0) || !isfinite(amp0)) { g[0] = g[1] = g[2] = NAN; return; }
    double x[3] = {amp0, peak, log(sd0)};
    const double step[3] = {0.1 * fabs(amp0) + 1e-300, 0.5 * sd0, 0.1};
    double sabs = 0.0;
    for (int i = tid; i < n; i += kKneeThreads) sabs += fabs(y[i]);
    sabs = block_sum(sabs, sh);
    nelder_mead<1, 3>(y, f, n, x, step, sabs, sh);
    g[0] = x[0]; g[1] = x[1]; g[2] = exp(x[2]);
}

// rows: [n_rows][ld] doubles; the fit uses columns [i0, i0 + n) (fooof_fit's frequency mask; the whole row otherwise).
// mode 0: find_knee_frequency (+ peak_out: find_oscillation_peak of the residual, as the oscillation model's combined
// distance does); mode 1: fooof_fit(oscillation_peak = true, max_peaks) -> knee of the cleaned spectrum; mode 2: as 0
// with lorentzian_initial_guess in place of the fit (the oscillation models' informed prior,
// one_timescale_and_osc.jl:16-28); mode 3: find_oscillation_peak of the row itself.
// work: gridDim.x * 3 * n doubles.
__global__ void __launch_bounds__(kKneeThreads, 16) k_knee_fit(const double* __restrict__ rows, int64_t n_rows, int64_t ld, int i0, int n,
                                                           const double* __restrict__ freqs, double min_freq, double max_freq,
                                                           int mode, int max_peaks, double* __restrict__ knee_out,
                                                           double* __restrict__ amp_out, double* __restrict__ peak_out,
                                                           double* __restrict__ work) {
    __shared__ KneeShared sh;
    const int tid = threadIdx.x;
    const double* f = freqs + i0;
    double* res = work + (size_t)blockIdx.x * 3 * n;
    double* pre = res + n;
    double* suf = pre + n;
    const double fmaxv = f[n - 1];
    for (int64_t r = blockIdx.x; r < n_rows; r += gridDim.x) {
        const double* y = rows + r * ld + i0;
        double amp = NAN, knee = NAN;
        double first_peak = NAN;
        if (mode == 3) {
            first_peak = osc_peak_row(y, f, n, min_freq, max_freq, 0.1, pre, suf, &sh);
        } else {
            knee_fit_row(y, f, n, min_freq, fmaxv, &amp, &knee, &sh, mode != 2);
        }
        if (mode != 3 && (mode == 1 || peak_out) && knee == knee) {
            for (int i = tid; i < n; i += kKneeThreads) { const double t = f[i] / knee; res[i] = y[i] - amp / fma(t, t, 1.0); }
            __syncthreads();
            if (mode != 1) {
                first_peak = osc_peak_row(res, f, n, min_freq, max_freq, 0.1, pre, suf, &sh);
            } else {
                int n_peaks = 0;
                for (int p = 0; p < max_peaks; ++p) {
                    const double pf = osc_peak_row(res, f, n, min_freq, max_freq, 0.1, pre, suf, &sh);
                    if (!(pf == pf)) break;
                    if (p == 0) first_peak = pf;
                    double g[3];
                    gaussian_fit_row(res, f, n, pf, min_freq, max_freq, g, &sh);
                    if (!(g[0] == g[0])) break;
                    ++n_peaks;
                    const double h = -0.5 / (g[2] * g[2]);
                    __syncthreads();
                    for (int i = tid; i < n; i += kKneeThreads) { const double d = f[i] - g[1]; res[i] -= g[0] * exp(h * d * d); }
                    __syncthreads();
                }
                if (n_peaks > 0) {
                    // cleaned = psd - sum of the Gaussians = residual + first Lorentzian; refit from its own initial guess
                    for (int i = tid; i < n; i += kKneeThreads) { const double t = f[i] / knee; res[i] += amp / fma(t, t, 1.0); }
                    __syncthreads();
                    knee_fit_row(res, f, n, min_freq, fmaxv, &amp, &knee, &sh);
                }
            }
        }
        if (tid == 0) {
            knee_out[r] = knee;
            if (amp_out) amp_out[r] = amp;
            if (peak_out) peak_out[r] = first_peak;
        }
        __syncthreads();
    }
}

// CTAs that fill the machine once (the registers per thread bound the resident CTAs per SM)
int knee_fit_grid(int sm_count, int64_t n_rows) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_knee_fit, kKneeThreads, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
    const int64_t g = (int64_t)sm_count * per_sm;
    return (int)(n_rows < g ? n_rows : g);
}
cudaError_t knee_fit_launch(const double* rows, int64_t n_rows, int64_t ld, int i0, int n, const double* freqs, double min_freq,
                            double max_freq, int mode, int max_peaks, double* knee_out, double* amp_out, double* peak_out,
                            double* work, int grid, cudaStream_t st) {
    k_knee_fit<<<grid, kKneeThreads, 0, st>>>(rows, n_rows, ld, i0, n, freqs, min_freq, max_freq, mode, max_peaks, knee_out,
                                              amp_out, peak_out, work);
    return cudaGetLastError();
}

// (n_rows x n_cols) column-major -> row-major (rows contiguous), 32 x 32 tiles through shared memory
__global__ void k_transpose_cm_to_rm(const double* __restrict__ in, int64_t n_rows, int64_t n_cols, double* __restrict__ out) {
    __shared__ double tile[32][33];
    const int64_t r0 = (int64_t)blockIdx.x * 32, c0 = (int64_t)blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int64_t r = r0 + threadIdx.x, c = c0 + j;
        if (r < n_rows && c < n_cols) tile[j][threadIdx.x] = in[r + c * n_rows];
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int64_t r = r0 + j, c = c0 + threadIdx.x;
        if (r < n_rows && c < n_cols) out[r * n_cols + c] = tile[threadIdx.x][j];
    }
}
cudaError_t transpose_launch(const double* in, int64_t n_rows, int64_t n_cols, double* out, cudaStream_t st) {
    dim3 grid((unsigned)((n_rows + 31) / 32), (unsigned)((n_cols + 31) / 32)), block(32, 8);
    k_transpose_cm_to_rm<<<grid, block, 0, st>>>(in, n_rows, n_cols, out);
    return cudaGetLastError();
}

// column means of a column-major (n_rows x n_cols) matrix -> out[n_cols] (average_over_trials of the spectra)
__global__ void k_col_mean(const double* __restrict__ in, int64_t n_rows, int64_t n_cols, double* __restrict__ out) {
    const int64_t c = blockIdx.x;
    double s = 0.0;
    for (int64_t r = threadIdx.x; r < n_rows; r += blockDim.x) s += in[r + c * n_rows];
    s = warp_sum(s);
    __shared__ double red[8];
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) { double t = 0.0; for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w]; out[c] = t / (double)n_rows; }
}
cudaError_t col_mean_launch(const double* in, int64_t n_rows, int64_t n_cols, double* out, cudaStream_t st) {
    k_col_mean<<<(unsigned)n_cols, 256, 0, st>>>(in, n_rows, n_cols, out);
    return cudaGetLastError();
}


// combined_distance (one_timescale.jl:278-290, one_timescale_and_osc.jl:324-347):
// d <- w0 d + w1 (data_tau - tau)^2 [+ w2 (data_osc - peak)^2]; fit = tau (ACF) or the knee frequency (PSD: tau = 1 / 2 pi knee)
__global__ void __launch_bounds__(256) k_combine(double* __restrict__ dist, int64_t n, const double* __restrict__ fit, int is_knee,
                                                 const double* __restrict__ peak, double w0, double w1, double w2, double data_tau,
                                                 double data_osc) {
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const double tau = is_knee ? 1.0 / (6.283185307179586 * fit[i]) : fit[i];
    const double e = data_tau - tau;
    double d = w0 * dist[i] + w1 * (e * e);
    if (peak) { const double o = data_osc - peak[i]; d += w2 * (o * o); }
    dist[i] = d;
}
cudaError_t combine_launch(double* dist, int64_t n, const double* fit, int is_knee, const double* peak, double w0, double w1, double w2,
                           double data_tau, double data_osc, cudaStream_t st) {
    k_combine<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(dist, n, fit, is_knee, peak, w0, w1, w2, data_tau, data_osc);
    return cudaGetLastError();
}


// ---- fit_expdecay_3_parameters (reference src/utils/utils.jl:85-91, 159-179) ------------------------------------------
// A (exp(-t / tau) + B) fitted to acf[2:end] over lags[2:end] (the zero lag is skipped: acw(...; skip_zero_lag = true)),
// residual = mean squared error.  For a given tau the model is linear in (A, A B), so the best pair is the closed-form
// 2 x 2 least-squares solution and the objective becomes a function of log tau alone, minimised around the reference's
// initial guess tau0 = tau_from_acw50(acw50(lags, acf)) (1.0 when the ACF never falls to 0.5) by a 65-point log scan of
// tau0 / 32 .. 32 tau0 and a golden-section search (a noisy ACF gives this function more than one dip: a local 1-D
// simplex and SciPy's Brent stopped in different ones, benchmarks/api_fuzz.py).  Returns the argmin of the reference's objective; its LM may stop elsewhere
// (unpinned) and returns NaN for a negative tau, which the log parametrisation cannot produce.
struct Exp3Sums { double se, see, sy, sey, syy; };

__device__ __forceinline__ double exp3_objective(const double* __restrict__ y, const double* __restrict__ t, int n, double log_tau,
                                                 KneeShared* sh, double* a_out = nullptr, double* c_out = nullptr) {
    const double it = exp(-log_tau);
    double se = 0.0, see = 0.0, sy = 0.0, sey = 0.0, syy = 0.0;
    for (int i = threadIdx.x; i < n; i += kKneeThreads) {
        const double e = exp(-t[i] * it), v = y[i];
        se += e; see = fma(e, e, see); sy += v; sey = fma(e, v, sey); syy = fma(v, v, syy);
    }
    se = block_sum(se, sh); see = block_sum(see, sh); sy = block_sum(sy, sh); sey = block_sum(sey, sh); syy = block_sum(syy, sh);
    // minimise sum (a e_i + c - y_i)^2: [see se; se n] [a; c] = [sey; sy]
    const double nn = (double)n, det = see * nn - se * se;
    double a, c;
    if (fabs(det) > 1e-300 * (see * nn + 1e-300)) { a = (sey * nn - se * sy) / det; c = (see * sy - se * sey) / det; }
    else { a = 0.0; c = sy / nn; }
    if (a_out) { *a_out = a; *c_out = c; }
    const double sse = syy - a * sey - c * sy;
    return fmax(sse, 0.0) / nn;
}

// rows: [n_rows][ld] (lag index fastest); lags: n_lags values; uses entries 1 .. n_lags - 1
__global__ void __launch_bounds__(kKneeThreads) k_expdecay3_fit(const double* __restrict__ rows, int64_t n_rows, int64_t ld, int n_lags,
                                                                const double* __restrict__ lags, double* __restrict__ tau_out) {
    __shared__ KneeShared sh;
    const int tid = threadIdx.x;
    for (int64_t r = blockIdx.x; r < n_rows; r += gridDim.x) {
        const double* acf = rows + r * ld;
        // acw50: first lag with acf <= 0.5 (NaN values never compare true)
        int first = 0x7fffffff;
        double bad = 0.0;
        for (int i = tid; i < n_lags; i += kKneeThreads) {
            if (acf[i] <= 0.5) first = min(first, i);
            if (i >= 1 && !isfinite(acf[i])) bad = 1.0;
        }
        first = block_min_int(first, &sh);
        bad = block_sum(bad, &sh);
        const double tau0 = first == 0x7fffffff ? 1.0 : lags[first] / 0.6931471805599453;       // tau_from_acw50
        double tau = NAN;
        if (bad == 0.0 && n_lags >= 4 && tau0 > 0.0) {
            // the objective in log tau can have more than one dip: scan tau0 / 32 .. 32 tau0 on 65 log-spaced points
            // (first minimum on ties), then golden-section search between the neighbours of the best one
            const double* yy = acf + 1; const double* tt = lags + 1; const int m = n_lags - 1;
            const double x0 = log(tau0), span = log(32.0), dx = span / 32.0;
            int bi = 0; double bf = INFINITY;
            for (int i = 0; i < 65; ++i) {
                const double fv = exp3_objective(yy, tt, m, x0 - span + i * dx, &sh);
                if (fv < bf) { bf = fv; bi = i; }
            }
            double lo = x0 - span + (bi - 1) * dx, hi = x0 - span + (bi + 1) * dx;
            const double gr = 0.6180339887498949;
            double c1 = hi - gr * (hi - lo), c2 = lo + gr * (hi - lo);
            double f1 = exp3_objective(yy, tt, m, c1, &sh), f2 = exp3_objective(yy, tt, m, c2, &sh);
            for (int it = 0; it < 70; ++it) {
                if (f1 < f2) { hi = c2; c2 = c1; f2 = f1; c1 = hi - gr * (hi - lo); f1 = exp3_objective(yy, tt, m, c1, &sh); }
                else { lo = c1; c1 = c2; f1 = f2; c2 = lo + gr * (hi - lo); f2 = exp3_objective(yy, tt, m, c2, &sh); }
            }
            const double x = 0.5 * (lo + hi);
            tau = exp(x);
        }
        if (tid == 0) tau_out[r] = tau;
        __syncthreads();
    }
}

cudaError_t expdecay3_launch(const double* rows, int64_t n_rows, int64_t ld, int n_lags, const double* lags, double* tau_out, int grid,
                             cudaStream_t st) {
    k_expdecay3_fit<<<grid, kKneeThreads, 0, st>>>(rows, n_rows, ld, n_lags, lags, tau_out);
    return cudaGetLastError();
}

}  // namespace itjl
