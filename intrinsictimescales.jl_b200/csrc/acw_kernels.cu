// K4: acw() on stored data - per-slice autocorrelation (fp64, direct lag sum in shared
// memory) with the threshold reducers, plus the Romberg-AUC and exp-decay-fit epilogues.
//
// Replaces the ACF-based part of acw(data, fs; ...) (reference src/core/acw.jl:141-231):
//   comp_ac_fft / comp_ac_time_missing     src/stats/summary.jl:46-63, 455-484
//   acw0 / acw50 / acweuler                src/utils/utils.jl:257-264, 218-226, 294-302
//   acw_romberg                            src/utils/utils.jl:332-334 (Romberg.jl, see oracle/acwred.py)
//   fit_expdecay                           src/utils/utils.jl:113-126 (argmin of the LM objective)
//
// The reference computes all n lags of every slice by FFT and then keeps
// n_lags = ceil(1.1 max k0) of them; here each slice is scanned in growing lag blocks
// (20, 20, 40, 80, ...) only until its own zero crossing, in fp64 so that the crossing
// indices agree with an fp64 reference bit for bit (up to ties at the 1e-15 level).
// Lag-sum tiling as in abc_kernels.cu, with doubles: thread = (tile of 10 lags) x (time
// stream), 100 DFMA per 10 LDS.128.
#include <math.h>

#include <stdlib.h>

#include "kernels.h"

namespace itjl {

constexpr int TKD = 10;   // lags per thread
constexpr int TTD = 10;   // samples per time tile

__device__ __forceinline__ void load5d(const double2* __restrict__ p, double* dst) {
#pragma unroll
    for (int q = 0; q < 5; ++q) {
        const double2 v = p[q];
        dst[2 * q] = v.x; dst[2 * q + 1] = v.y;
    }
}

__device__ __forceinline__ void acf_tiles_f64(const double* __restrict__ xs, double (&acc)[TKD], int k0,
                                              int tile_begin, int tile_end) {
    double b[2 * TTD];
    double a[TTD];
    const double2* pa = reinterpret_cast<const double2*>(xs + tile_begin * TTD);
    const double2* pb = reinterpret_cast<const double2*>(xs + tile_begin * TTD + k0);
    load5d(pb, b);
    pb += 5;
    int n = tile_end - tile_begin;
    while (n > 0) {
        load5d(pb, b + TTD); pb += 5;
        load5d(pa, a); pa += 5;
#pragma unroll
        for (int i = 0; i < TTD; ++i)
#pragma unroll
            for (int j = 0; j < TKD; ++j) acc[j] = fma(a[i], b[i + j], acc[j]);
        if (--n == 0) break;
        load5d(pb, b); pb += 5;
        load5d(pa, a); pa += 5;
#pragma unroll
        for (int i = 0; i < TTD; ++i)
#pragma unroll
            for (int j = 0; j < TKD; ++j) acc[j] = fma(a[i], b[(TTD + i + j) % (2 * TTD)], acc[j]);
        --n;
    }
}

struct AcfSliceParams {
    const double* data;     // (n_slices x n_t) column-major
    int64_t n_slices;
    int n_t;
    int missing;            // NaN-aware centring (comp_ac_time_missing)
    int mode;               // 0 = scan until the zero crossing, 1 = fill lags [n_done, need)
    int max_lags;           // scan limit
    int need;               // fill target
    int xs_len;             // doubles in the staging buffer (>= n_t + max block reach)
    double* acf_work;       // [slice][cap]
    int64_t cap;
    int32_t* n_done;        // [slice] lags available in acf_work
    int32_t* k0; int32_t* k50; int32_t* ke;
    int32_t* nan_flag;      // set to 1 when any slice holds a NaN (may be null)
    const int32_t* slice_list;  // optional: CTA b works on slice slice_list[b]
};

constexpr double kInvE = 0.36787944117144233;   // Julia: 1 / ℯ

__global__ void __launch_bounds__(kThreads, 2) k_acf_slices(const AcfSliceParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* xs = reinterpret_cast<double*>(smem_raw);
    double* part = xs + P.xs_len;                 // [n_streams][lt*TKD] <= kThreads*TKD doubles
    __shared__ double red[kWarps];
    __shared__ int s_first[3];
    __shared__ double s_ss;

    const int tid = threadIdx.x;
    const int64_t s = P.slice_list ? (int64_t)P.slice_list[blockIdx.x] : (int64_t)blockIdx.x;
    const int n_t = P.n_t;

    int start = 0, stop = P.max_lags;
    if (P.mode == 1) {
        // resume on a whole lag tile: a previous fill may have stopped at any lag count (the caller's n_lags),
        // and the tile loop reads the series with 16-byte loads at lag offsets that must stay even
        start = (P.n_done[s] / TKD) * TKD;
        stop = P.need;
        if (P.n_done[s] >= stop) return;          // uniform per CTA
    }

    // ---- load + centre -------------------------------------------------------------
    double lsum = 0.0;
    int lcnt = 0;
    for (int j = tid; j < P.xs_len; j += kThreads) {
        double v = 0.0;
        if (j < n_t) {
            v = P.data[s + (size_t)j * P.n_slices];
            if (P.missing && isnan(v)) v = 0.0; else lcnt++;
            lsum += v;
        }
        xs[j] = v;
    }
    lsum = warp_sum(lsum);
    double dcnt = warp_sum((double)lcnt);
    if ((tid & 31) == 0) { red[tid >> 5] = lsum; part[tid >> 5] = dcnt; }
    __syncthreads();
    double tot = 0.0, cnt = 0.0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) { tot += red[w]; cnt += part[w]; }
    const double mean = tot / cnt;                // cnt == n_t for complete data
    if (tid == 0 && P.nan_flag && cnt < (double)n_t) *P.nan_flag = 1;
    __syncthreads();
    double lss = 0.0;
    for (int j = tid; j < n_t; j += kThreads) {   // summary.jl:471: every entry (masked too) -= mean
        const double v = xs[j] - mean;
        xs[j] = v;
        lss = fma(v, v, lss);
    }
    lss = warp_sum(lss);
    if ((tid & 31) == 0) red[tid >> 5] = lss;
    if (tid < 3) s_first[tid] = 0x7fffffff;
    __syncthreads();
    if (tid == 0) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) t += red[w];
        s_ss = t;
    }
    __syncthreads();
    const double inv_ss = 1.0 / s_ss;
    const int n_tiles = (n_t + TTD - 1) / TTD;

    // ---- lag blocks --------------------------------------------------------------------
    int kb = start;
    int lb = 2 * TKD;
    while (kb < stop) {
        if (P.mode == 0 && kb >= 4 * TKD) lb = kb;            // 20, 20, 40, 80, 160, ...
        if (P.mode == 1) lb = min(stop - kb, kThreads * TKD);
        lb = min(lb, stop - kb);
        lb = min(lb, kThreads * TKD);
        const int lt = (lb + TKD - 1) / TKD;
        int n_streams = kThreads / lt;
        if (n_streams > n_tiles) n_streams = n_tiles;
        const int tps = (n_tiles + n_streams - 1) / n_streams;
        const int tile = tid % lt, stream = tid / lt;
        const int row = lt * TKD;
        if (stream < n_streams) {
            double acc[TKD];
#pragma unroll
            for (int j = 0; j < TKD; ++j) acc[j] = 0.0;
            const int tb = stream * tps, te = min(tb + tps, n_tiles);
            if (tb < te) acf_tiles_f64(xs, acc, kb + tile * TKD, tb, te);
#pragma unroll
            for (int j = 0; j < TKD; ++j) part[stream * row + tile * TKD + j] = acc[j];
        }
        __syncthreads();
        for (int k = tid; k < lb; k += kThreads) {
            double v = 0.0;
            for (int st = 0; st < n_streams; ++st) v += part[st * row + k];
            const int lag = kb + k;
            // lag 0 is acov[0]/acov[0] in the reference: exactly 1 (NaN for a constant slice)
            v = (lag == 0) ? (s_ss / s_ss) : v * inv_ss;
            P.acf_work[(size_t)s * P.cap + lag] = v;
            if (P.mode == 0) {
                if (v <= 0.0) atomicMin(&s_first[0], lag);
                if (v <= 0.5) atomicMin(&s_first[1], lag);
                if (v <= kInvE) atomicMin(&s_first[2], lag);
            }
        }
        __syncthreads();
        kb += lb;
        if (P.mode == 0 && s_first[0] != 0x7fffffff) break;
    }
    if (tid == 0) {
        P.n_done[s] = kb;
        if (P.mode == 0) {
            P.k0[s] = (s_first[0] == 0x7fffffff) ? -1 : s_first[0];
            P.k50[s] = (s_first[1] == 0x7fffffff) ? -1 : s_first[1];
            P.ke[s] = (s_first[2] == 0x7fffffff) ? -1 : s_first[2];
        }
    }
}

// crossing indices of stored ACF rows (acf[row*ld + k], k < n_lags): one warp per row
__global__ void k_acf_cross(const double* __restrict__ acf, int64_t n_rows, int64_t ld, int n_lags,
                            int32_t* k0, int32_t* k50, int32_t* ke) {
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= n_rows) return;
    const int lane = threadIdx.x & 31;
    int f0 = 0x7fffffff, f50 = 0x7fffffff, fe = 0x7fffffff;
    for (int base = 0; base < n_lags && f0 == 0x7fffffff; base += 32) {
        const int k = base + lane;
        const double v = (k < n_lags) ? acf[row * ld + k] : 2.0;
        const unsigned b0 = __ballot_sync(0xffffffffu, v <= 0.0);
        const unsigned b50 = __ballot_sync(0xffffffffu, v <= 0.5);
        const unsigned be = __ballot_sync(0xffffffffu, v <= kInvE);
        if (f50 == 0x7fffffff && b50) f50 = base + __ffs(b50) - 1;
        if (fe == 0x7fffffff && be) fe = base + __ffs(be) - 1;
        if (b0) f0 = base + __ffs(b0) - 1;
    }
    if (lane == 0) {
        k0[row] = (f0 == 0x7fffffff) ? -1 : f0;
        k50[row] = (f50 == 0x7fffffff) ? -1 : f50;
        ke[row] = (fe == 0x7fffffff) ? -1 : fe;
    }
}

// column mean over rows: out[k] = mean_row acf[row*ld + k]
__global__ void k_acf_mean(const double* __restrict__ acf, int64_t n_rows, int64_t ld, int n_lags,
                           double* __restrict__ out) {
    __shared__ double red[kWarps];
    const int k = blockIdx.x;
    double s = 0.0;
    for (int64_t r = threadIdx.x; r < n_rows; r += kThreads) s += acf[r * ld + k];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) t += red[w];
        out[k] = t / (double)n_rows;
    }
}

// (rows x ld) row-major work buffer -> (rows x n_lags) column-major output
__global__ void k_acf_export(const double* __restrict__ acf, int64_t n_rows, int64_t ld, int n_lags,
                             double* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rows * n_lags) return;
    const int64_t row = i % n_rows, k = i / n_rows;
    out[i] = acf[row * ld + k];
}

// ---- Romberg AUC ---------------------------------------------------------------------
// Romberg.jl romberg(dt, y[0..w)) : trapezoid sums at the strides given by the prime
// factorisation of w-1 (host-computed, increasing), Richardson extrapolation with power 2,
// returning the tableau entry with the smallest error estimate (oracle/acwred.py).
constexpr int kMaxRombergLevels = 40;

__global__ void k_acw_auc(const double* __restrict__ acf, int64_t n_rows, int64_t ld, int w, double dt,
                          const int* __restrict__ strides, int n_strides, double* __restrict__ auc) {
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n_rows) return;
    const double* y = acf + row * ld;
    const double endsum = 0.5 * (y[0] + y[w - 1]);
    double vals[kMaxRombergLevels], hp[kMaxRombergLevels];
    // order: coarsest stride first (decreasing h)
    for (int q = 0; q < n_strides; ++q) {
        const int step = strides[n_strides - 1 - q];
        double inner = 0.0;
        for (int i = step; i < w - 1; i += step) inner += y[i];
        const double h = (double)step * dt;
        vals[q] = h * (inner + endsum);
        hp[q] = h * h;
    }
    double f0 = vals[0], err = INFINITY;
    double nev[kMaxRombergLevels];
    nev[0] = vals[0];
    const double rtol = 1.4901161193847656e-08;      // sqrt(eps(Float64))
    for (int j = 1; j < n_strides; ++j) {
        nev[j] = vals[j];
        double minerr = INFINITY;
        for (int i = j - 1; i >= 0; --i) {
            const double old = nev[i];
            nev[i] = nev[i + 1] + (nev[i + 1] - nev[i]) / (hp[i] / hp[j] - 1.0);
            const double e = fabs(nev[i] - old);
            minerr = fmin(minerr, e);
            if (e < err) { f0 = nev[i]; err = e; }
        }
        if (minerr > 2.0 * err || !isfinite(minerr)) break;
        if (err <= rtol * fabs(f0)) break;
    }
    auc[row] = f0;
}

// ---- exp-decay fit -------------------------------------------------------------------
// argmin_tau mean((exp(-lags/tau) - acf)^2), lags = k*dt, k < n_lags: one warp per row.
// Same algorithm as oracle/acwred.py fit_expdecay (bracket walk in log tau + Illinois).
__device__ __forceinline__ double dres_dlogtau(const double* __restrict__ y, int n_lags, double dt, double u,
                                               int lane) {
    const double sr = exp(-u);
    double g = 0.0;
    for (int k = lane; k < n_lags; k += 32) {
        const double t = (double)k * dt;
        const double e = exp(-sr * t);
        g += (e - y[k]) * e * t;
    }
    g = warp_sum(g);
    return g * sr;
}

__global__ void k_acw_tau(const double* __restrict__ acf, int64_t n_rows, int64_t ld, int n_lags, double dt,
                          double* __restrict__ tau) {
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= n_rows) return;
    const int lane = threadIdx.x & 31;
    const double* y = acf + row * ld;
    // u0 = log(acw50 / ln 2) on the truncated ACF (utils.jl:115-120)
    int f50 = 0x7fffffff;
    for (int base = 0; base < n_lags && f50 == 0x7fffffff; base += 32) {
        const int k = base + lane;
        const unsigned b = __ballot_sync(0xffffffffu, (k < n_lags) && (y[k] <= 0.5));
        if (b) f50 = base + __ffs(b) - 1;
    }
    double tau0 = (f50 == 0x7fffffff) ? 1.0 : -((double)f50 * dt) / log(0.5);
    if (!(tau0 > 0.0) || !isfinite(tau0)) tau0 = 1.0;
    const double u0 = log(tau0);
    const double g0 = dres_dlogtau(y, n_lags, dt, u0, lane);
    double result = exp(u0);
    if (g0 != 0.0) {
        const double dir = (g0 > 0.0) ? -1.0 : 1.0;
        double up = u0, gp = g0, lo = u0, glo = g0, hi = u0, ghi = g0;
        bool found = false, exact = false;
        for (int it = 0; it < 24; ++it) {
            const double u = up + dir * 0.5;
            const double g = dres_dlogtau(y, n_lags, dt, u, lane);
            if (g == 0.0) { result = exp(u); exact = true; break; }
            if ((g > 0.0) != (gp > 0.0)) {
                if (dir < 0.0) { lo = u; glo = g; hi = up; ghi = gp; }
                else { lo = up; glo = gp; hi = u; ghi = g; }
                found = true;
                break;
            }
            up = u; gp = g;
        }
        if (!exact) {
            if (!found) result = exp(up);
            else {
                double a = lo, ga = glo, b = hi, gb = ghi;
                int side = 0;
                double u = 0.5 * (a + b), ulast = a;
                for (int it = 0; it < 100; ++it) {
                    u = (a * gb - b * ga) / (gb - ga);
                    if (!(a < u && u < b)) u = 0.5 * (a + b);
                    const double g = dres_dlogtau(y, n_lags, dt, u, lane);
                    if (g == 0.0 || fabs(u - ulast) < 1e-13 * fmax(1.0, fabs(u))) break;
                    ulast = u;
                    if ((g > 0.0) == (gb > 0.0)) { b = u; gb = g; if (side == -1) ga *= 0.5; side = -1; }
                    else { a = u; ga = g; if (side == 1) gb *= 0.5; side = 1; }
                }
                result = exp(u);
            }
        }
    }
    if (lane == 0) tau[row] = result;
}

// ---- streaming first pass (short-memory data: HBM-bound) -----------------------------------
// One pass over the column-major input computes, for every slice, the RAW lagged products
//   P_k = sum_t y_t y_{t-k}, k < SK = 32,   and S = sum y_t,     y = x - ref  (per-slice shift)
// from which k_acw_finish derives the centred autocorrelation algebraically:
//   C_k = P_k - m (2S - head_k - tail_k) + (n - k) m^2,  m = S / n,
// head_k / tail_k = sums of the first / last k shifted samples.
// Precision (north_star: "fp32 accumulate, fp64 check"): the shift is applied in fp64, the
// products are fp32 and flushed to fp64 every 256 samples - ACF error ~1e-8.  Every crossing
// decision whose ACF value lies within 1e-6 of a threshold (0, 0.5, 1/e), every slice containing a
// NaN (it propagates into P_k), every slice that has not crossed zero within SK lags and every
// slice whose shifted mean is still large (n m^2 > C_0) is redone in fp64 by k_acf_slices, so the
// crossing indices stay bit-exact against an fp64 reference.
// Layout: thread = (slice, time segment); consecutive threads own consecutive slices, so every
// time step of a warp is one coalesced 256-byte load straight into registers (16 loads in flight
// per thread).  A thread keeps the 32 accumulators and a 32-deep circular window
// of its slice in registers (static indexing: the time loop is unrolled by 32) - 32 FFMA per
// sample next to one LDG.64, one DADD and one F2F.  A segment warms its window up on the 32
// samples before it.  The shift is the mean of the slice's first 32 samples (same arithmetic in
// k_acw_finish).  Partial sums go out as fp32 [segment][lag][slice] (coalesced).
constexpr int SK = 32;     // lags of the streaming pass
constexpr int SREF = 32;   // samples averaged for the per-slice shift
constexpr int SPART = SK + 1;   // per (segment, slice): P_0..P_31, S
constexpr int STREAM_THREADS = 128;
constexpr int SFLUSH = 256;     // fp32 partial sums are folded into fp64 every SFLUSH samples
constexpr double kStreamTol = 1e-6;

struct AcwStreamParams {
    const double* data;
    int64_t n_slices;
    int n_t;
    int n_seg, seg_len;          // time segments (seg_len multiple of SK)
    float* part;                 // [seg][SPART][slice]
};

// per-slice shift: mean of the first min(SREF, n_t) samples, summed in time order
__device__ __forceinline__ double slice_ref(const double* __restrict__ data, int64_t s, int64_t n_slices, int n_t) {
    const int m = min(SREF, n_t);
    double r = 0.0;
    for (int t = 0; t < m; ++t) r += data[s + (size_t)t * n_slices];
    return r / (double)m;
}

// SKT = lags kept by the pass (8 / 16 / 32): the window, the accumulators and so the registers and the FFMA per
// sample scale with it - 32 lags: 128 registers, 4 CTAs per SM, 38 instructions per sample; 16 lags: 6 CTAs, 22; 8 lags:
// 8 CTAs, 14.  Blocks stay 32 samples long for every SKT (k_acw_finish adds the n_t % 32 tail in fp64).
template <int SKT> struct StreamCfg { static constexpr int kCtas = SKT == 32 ? 4 : (SKT == 16 ? 6 : 8); };
template <int SKT>
__global__ void __launch_bounds__(STREAM_THREADS, StreamCfg<SKT>::kCtas) k_acw_stream(const AcwStreamParams P) {
    const int64_t s = (int64_t)blockIdx.x * STREAM_THREADS + threadIdx.x;
    if (s >= P.n_slices) return;
    const int seg = blockIdx.y;
    const int tb = seg * P.seg_len;
    const int te = min(tb + P.seg_len, P.n_t & ~(SK - 1));
    const double* __restrict__ col = P.data + s;
    const size_t ld = (size_t)P.n_slices;
    const double ref = slice_ref(P.data, s, P.n_slices, P.n_t);

    // fp64 running sums live in shared memory ([lag][thread]: conflict-free), touched every SFLUSH samples
    __shared__ double acc64[SKT + 1][STREAM_THREADS];
    float acc[SKT], C[SKT];
#pragma unroll
    for (int j = 0; j < SKT; ++j) { acc[j] = 0.f; acc64[j][threadIdx.x] = 0.0; }
    acc64[SKT][threadIdx.x] = 0.0;
    float S = 0.f;

    // window warm-up: C[i] = y_{tb - SKT + i} (zero before the series starts)
    constexpr int HB = SKT < 16 ? SKT : 16;
#pragma unroll
    for (int h = 0; h < SKT; h += HB) {
        double x[HB];
#pragma unroll
        for (int i = 0; i < HB; ++i) {
            const int t = tb - SKT + h + i;
            x[i] = (t >= 0) ? col[(size_t)t * ld] : ref;
        }
#pragma unroll
        for (int i = 0; i < HB; ++i) C[h + i] = (float)(x[i] - ref);       // t < 0: ref - ref = 0
        asm volatile("" ::: "memory");
    }

    // whole 32-sample blocks only: the host makes every segment a multiple of 32 samples and
    // k_acw_finish adds the n_t % 32 samples at the end of the series in fp64.
    // (Tried and dropped, B200, SKT = 32: loads issued half a block ahead of the math - 162 registers, 3 CTAs/SM,
    // 116 us - and an extra prefetch.global.L2 96 steps ahead - 122 us - against 111 us for this loop:
    // 38 instructions per sample fill 65 % of the issue slots and 16 resident warps per SM (128 registers per
    // thread) leave DRAM latency exposed - trading registers for load depth loses on both counts.)
    int since_flush = 0;
    const double* __restrict__ p = col + (size_t)tb * ld;
    for (int t0 = tb; t0 + SK <= te; t0 += SK) {
#pragma unroll
        for (int h = 0; h < SK; h += 16) {
            double x[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) { x[i] = *p; p += ld; }
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const float z = (float)(x[i] - ref);
                C[(h + i) & (SKT - 1)] = z;                // (t - tb) mod SKT
                S += z;
#pragma unroll
                for (int j = 0; j < SKT; ++j) acc[j] = fmaf(z, C[(h + i - j) & (SKT - 1)], acc[j]);
            }
            asm volatile("" ::: "memory");                 // keep the next 16 loads behind this half's math
        }
        since_flush += SK;
        if (since_flush >= SFLUSH) {
#pragma unroll
            for (int j = 0; j < SKT; ++j) { acc64[j][threadIdx.x] += (double)acc[j]; acc[j] = 0.f; }
            acc64[SKT][threadIdx.x] += (double)S; S = 0.f;
            since_flush = 0;
        }
    }
    float* out = P.part + (size_t)seg * SPART * ld + s;
#pragma unroll
    for (int j = 0; j < SKT; ++j) out[(size_t)j * ld] = (float)(acc64[j][threadIdx.x] + (double)acc[j]);
    out[(size_t)SK * ld] = (float)(acc64[SKT][threadIdx.x] + (double)S);
}

// Measured and not kept (B200, C2 = randn 10000 x 5000; this kernel: 102 us, 55 % of the measured HBM rate):
//  - packed fp32 pairs (FFMA2, `__ffma2_rn`): the FMA pipe rate is the same (74 TFLOP/s either way,
//    benchmarks/micro/ffma2_rate.cu), only issue slots are saved; the time-pair formulation needs 64 accumulator +
//    32 window registers (142-168 per thread, 8-12 warps per SM): 125-170 us;
//  - a TMA ring (producer warp, cp.async.bulk rows of 1 KB into 3-4 stages of 16 time steps, consumers pull a tile
//    into registers and hand the stage back): 125 us with FFMA, 125-132 us with FFMA2 - and 90 us with the lag
//    products compiled out altogether, i.e. the ring of 1 KB rows itself tops out at 4.4 TB/s, no better than what
//    plain coalesced LDG.64 already delivers here;
//  - round 1's variants (deeper LDG pipelining, L2 prefetch, in-line cp.async.bulk producer): 116-182 us.
// The pass is bounded by how the (slice-fastest, 80 KB per time step) layout can be streamed per slice block, not
// by the FMA pipe (44 us of work) nor by raw HBM bandwidth (62 us).

struct AcwFinishParams {
    const double* data;
    int64_t n_slices;
    int n_t, n_seg;
    const float* part;
    double* acf_work; int64_t cap;
    int32_t* n_done; int32_t* k0; int32_t* k50; int32_t* ke;
    int32_t* redo;               // [n_slices] 1 = needs the fp64 kernel
    int w;                       // lags the streaming pass kept (8 / 16 / 32)
};

// CTA = 32 slices: the segment partials, the first 32 and the last 64 samples of every slice are
// staged coalesced in shared memory ([row][slice] tiles), then one warp per slice (lane = lag)
// derives the ACF and the crossings
constexpr int kFinishWarps = 16;       // 512 threads: all CTAs of C2 (313) in one wave of 4 per SM; 8 warps: 25 us, 32 warps (two waves): 21.6 us
__global__ void __launch_bounds__(32 * kFinishWarps) k_acw_finish(const AcwFinishParams P) {
    __shared__ double tile[SPART][33];
    __shared__ double head[SREF][33];          // samples 0 .. 31
    __shared__ double tail[2 * SK][33];        // samples n_t - 64 .. n_t - 1
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t s0 = (int64_t)blockIdx.x * 32;
    const size_t ld = (size_t)P.n_slices;
    const bool col_ok = s0 + lane < P.n_slices;
    for (int r = warp; r < SPART; r += kFinishWarps) {
        double v = 0.0;
        if (col_ok && (r < P.w || r == SK)) {
            const float* src = P.part + (size_t)r * ld + s0 + lane;
            const size_t gs = (size_t)SPART * ld;
            int g = 0;
            for (; g + 4 <= P.n_seg; g += 4) {              // four independent loads in flight
                const float p0 = src[(size_t)g * gs], p1 = src[(size_t)(g + 1) * gs];
                const float p2 = src[(size_t)(g + 2) * gs], p3 = src[(size_t)(g + 3) * gs];
                v += (double)p0; v += (double)p1; v += (double)p2; v += (double)p3;
            }
            for (; g < P.n_seg; ++g) v += (double)src[(size_t)g * gs];
        }
        tile[r][lane] = v;
    }
    for (int r = warp; r < SREF; r += kFinishWarps) head[r][lane] = (col_ok && r < P.n_t) ? P.data[s0 + lane + (size_t)r * ld] : 0.0;
    const int t_tail = P.n_t - 2 * SK;         // time index of tail row 0 (may be negative)
    for (int r = warp; r < 2 * SK; r += kFinishWarps) {
        const int t = t_tail + r;
        tail[r][lane] = (col_ok && t >= 0) ? P.data[s0 + lane + (size_t)t * ld] : 0.0;
    }
    __syncthreads();
    for (int w = warp; w < 32; w += kFinishWarps) {
        const int64_t s = s0 + w;
        if (s >= P.n_slices) break;
        const int k = lane;
        double pk = tile[k][w], S = tile[SK][w];
        const double n = (double)P.n_t;
        // per-slice shift: same arithmetic as slice_ref()
        double ref = 0.0;
        const int mref = min(SREF, P.n_t);
        for (int t = 0; t < mref; ++t) ref += head[t][w];
        ref /= (double)mref;
        // the n_t % 32 samples after the last whole block of the streaming pass, in fp64
        for (int t = P.n_t & ~(SK - 1); t < P.n_t; ++t) {
            const double yt = tail[t - t_tail][w] - ref;
            if (t - k >= 0) pk += yt * (tail[t - k - t_tail][w] - ref);
            S += yt;
        }
        const double m = S / n;
        // head_k = sum of the first k shifted samples, tail_k = sum of the last k: inclusive warp scans
        double a = 0.0, b = 0.0;
        if (k > 0 && k <= P.n_t) {
            a = head[k - 1][w] - ref;
            b = tail[2 * SK - k][w] - ref;
        }
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const double ua = __shfl_up_sync(0xffffffffu, a, d), ub = __shfl_up_sync(0xffffffffu, b, d);
            if (k >= d) { a += ua; b += ub; }
        }
        const double ck = pk - m * (2.0 * S - a - b) + (n - (double)k) * m * m;
        const double c0 = __shfl_sync(0xffffffffu, ck, 0);
        const bool in_range = k < P.n_t && k < P.w;
        const double v = (k == 0) ? (c0 / c0) : ck / c0;
        if (in_range) P.acf_work[(size_t)s * P.cap + k] = v;
        const unsigned b0 = __ballot_sync(0xffffffffu, in_range && v <= 0.0);
        const unsigned b50 = __ballot_sync(0xffffffffu, in_range && v <= 0.5);
        const unsigned be = __ballot_sync(0xffffffffu, in_range && v <= kInvE);
        const unsigned bad = __ballot_sync(0xffffffffu, in_range && !(v == v));      // NaN anywhere
        const int f0 = b0 ? __ffs(b0) - 1 : -1;
        // fp32 products: a value this close to a threshold, at a lag that matters for a crossing,
        // is decided in fp64 instead
        const bool matters = in_range && k > 0 && (f0 < 0 || k <= f0);
        const bool close = fabs(v) < kStreamTol || fabs(v - 0.5) < kStreamTol || fabs(v - kInvE) < kStreamTol;
        const unsigned unsure = __ballot_sync(0xffffffffu, matters && close);
        if (k == 0) {
            const int kmax = min(P.w, P.n_t);
            const bool ill = !(n * m * m <= c0);             // shifted mean still dominates: cancellation
            const bool redo = ill || bad || unsure || (f0 < 0 && kmax < P.n_t);
            P.redo[s] = redo ? 1 : 0;
            P.n_done[s] = redo ? 0 : kmax;
            P.k0[s] = f0;
            P.k50[s] = b50 ? __ffs(b50) - 1 : -1;
            P.ke[s] = be ? __ffs(be) - 1 : -1;
        }
    }
}

// ---- host launchers ------------------------------------------------------------------
static size_t acf_slices_smem(int n_t, int reach, int* xs_len) {
    int len = n_t + reach + 4 * TTD;
    len = (len + 1) & ~1;
    *xs_len = len;
    return (size_t)len * sizeof(double) + (size_t)kThreads * TKD * sizeof(double);
}

// largest lag reach whose staging buffer still fits in shared memory
int acf_max_reach(int n_t, int max_smem) {
    const long long avail = (long long)max_smem - (long long)kThreads * TKD * 8 - (long long)(n_t + 4 * TTD + 2) * 8;
    long long r = avail / 8;
    if (r > n_t) r = n_t;
    return (int)(r < 0 ? 0 : r);
}

static cudaError_t launch_slices(AcfSliceParams& P, int reach, int max_smem, int64_t n_ctas, cudaStream_t st) {
    int xs_len = 0;
    size_t smem = acf_slices_smem(P.n_t, reach, &xs_len);
    if ((int64_t)smem > (int64_t)max_smem) return cudaErrorInvalidValue;
    P.xs_len = xs_len;
    cudaError_t e = cudaFuncSetAttribute(k_acf_slices, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    k_acf_slices<<<(unsigned)n_ctas, kThreads, smem, st>>>(P);
    return cudaGetLastError();
}

cudaError_t acw_stream_launch(const double* data, int64_t n_slices, int n_t, double* part, int n_seg, int seg_len,
                              double* acf_work, int64_t cap, int32_t* n_done, int32_t* k0, int32_t* k50,
                              int32_t* ke, int32_t* redo, int w, cudaStream_t st) {
    AcwStreamParams S{};
    S.data = data; S.n_slices = n_slices; S.n_t = n_t; S.n_seg = n_seg; S.seg_len = seg_len;
    S.part = reinterpret_cast<float*>(part);
    dim3 grid((unsigned)((n_slices + STREAM_THREADS - 1) / STREAM_THREADS), (unsigned)n_seg);
    cudaError_t e;
    if (w == 8) k_acw_stream<8><<<grid, STREAM_THREADS, 0, st>>>(S);
    else if (w == 16) k_acw_stream<16><<<grid, STREAM_THREADS, 0, st>>>(S);
    else if (w == 32) k_acw_stream<32><<<grid, STREAM_THREADS, 0, st>>>(S);
    else return cudaErrorInvalidValue;
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    AcwFinishParams F{};
    F.data = data; F.n_slices = n_slices; F.n_t = n_t; F.n_seg = n_seg; F.part = reinterpret_cast<const float*>(part);
    F.acf_work = acf_work; F.cap = cap; F.n_done = n_done; F.k0 = k0; F.k50 = k50; F.ke = ke;
    F.redo = redo; F.w = w;
    k_acw_finish<<<(unsigned)((n_slices + 31) / 32), 32 * kFinishWarps, 0, st>>>(F);
    return cudaGetLastError();
}
int acw_stream_lags() { return SK; }
// fp32 partials, counted in doubles for the caller's reserve()
size_t acw_stream_part_doubles(int64_t n_slices, int n_seg) { return ((size_t)n_seg * SPART * (size_t)n_slices + 1) / 2; }
int acw_stream_ctas_per_sm(int w) { return w == 32 ? StreamCfg<32>::kCtas : (w == 16 ? StreamCfg<16>::kCtas : StreamCfg<8>::kCtas); }
void acw_stream_segments(int n_t, int64_t n_slices, int sm_count, int* n_seg, int* seg_len, int force_waves, int w) {
    // The grid is (blocks of 128 slices) x (time segments); 4 CTAs of 128 threads are resident per SM
    // (128 registers per thread).  Pick the segment count whose grid fills whole waves of those slots:
    // cost = waves x (segment length + 32-sample halo); segments are multiples of 32 samples and at least
    // 256 long (each one also writes 33 partial sums per slice).  C2 (10 000 x 5000): 7 segments of 736
    // samples, 553 of 592 slots, one wave (16 segments were 2.13 waves, i.e. three rounds).
    const long long slots = (long long)acw_stream_ctas_per_sm(w) * sm_count;
    const long long blocks = (n_slices + 127) / 128;
    const int max_g = n_t / 256 > 0 ? n_t / 256 : 1;
    long long best_cost = -1;
    int best_len = ((n_t + SK - 1) / SK) * SK;
    for (int waves = 1; waves <= 4; ++waves) {
        if (force_waves && waves != force_waves) continue;
        long long g = waves * slots / blocks;
        if (g < 1) g = 1;
        if (g > max_g) g = max_g;
        int len = (int)((n_t + g - 1) / g);
        len = ((len + SK - 1) / SK) * SK;
        const long long segs = (n_t + len - 1) / len;
        const long long rounds = (segs * blocks + slots - 1) / slots;
        const long long cost = rounds * (len + SK);
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; best_len = len; }
    }
    *seg_len = best_len;
    *n_seg = (n_t + best_len - 1) / best_len;
}

cudaError_t acw_scan_launch(const double* data, int64_t n_slices, int n_t, int missing, int max_lags,
                            double* acf_work, int64_t cap, int32_t* n_done, int32_t* k0, int32_t* k50,
                            int32_t* ke, int32_t* nan_flag, const int32_t* slice_list, int64_t n_list,
                            int max_smem, cudaStream_t st) {
    AcfSliceParams P{};
    P.nan_flag = nan_flag; P.slice_list = slice_list;
    P.data = data; P.n_slices = n_slices; P.n_t = n_t; P.missing = missing; P.mode = 0;
    P.max_lags = max_lags; P.need = 0; P.acf_work = acf_work; P.cap = cap; P.n_done = n_done;
    P.k0 = k0; P.k50 = k50; P.ke = ke;
    return launch_slices(P, max_lags, max_smem, slice_list ? n_list : n_slices, st);
}

cudaError_t acf_fill_launch(const double* data, int64_t n_slices, int n_t, int missing, int need,
                            double* acf_work, int64_t cap, int32_t* n_done, const int32_t* slice_list,
                            int64_t n_list, int max_smem, cudaStream_t st, int32_t* nan_flag) {
    AcfSliceParams P{};
    P.slice_list = slice_list; P.nan_flag = nan_flag;
    P.data = data; P.n_slices = n_slices; P.n_t = n_t; P.missing = missing; P.mode = 1;
    P.max_lags = need; P.need = need; P.acf_work = acf_work; P.cap = cap; P.n_done = n_done;
    return launch_slices(P, need, max_smem, slice_list ? n_list : n_slices, st);
}

cudaError_t acf_cross_launch(const double* acf, int64_t n_rows, int64_t ld, int n_lags, int32_t* k0,
                             int32_t* k50, int32_t* ke, cudaStream_t st) {
    const int wpb = 4;
    k_acf_cross<<<(unsigned)((n_rows + wpb - 1) / wpb), wpb * 32, 0, st>>>(acf, n_rows, ld, n_lags, k0, k50, ke);
    return cudaGetLastError();
}

cudaError_t acf_mean_launch(const double* acf, int64_t n_rows, int64_t ld, int n_lags, double* out,
                            cudaStream_t st) {
    k_acf_mean<<<n_lags, kThreads, 0, st>>>(acf, n_rows, ld, n_lags, out);
    return cudaGetLastError();
}

cudaError_t acf_export_launch(const double* acf, int64_t n_rows, int64_t ld, int n_lags, double* out,
                              cudaStream_t st) {
    const int64_t total = n_rows * n_lags;
    k_acf_export<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(acf, n_rows, ld, n_lags, out);
    return cudaGetLastError();
}

cudaError_t acw_auc_launch(const double* acf, int64_t n_rows, int64_t ld, int w, double dt,
                           const int* strides, int n_strides, double* auc, cudaStream_t st) {
    if (n_strides > kMaxRombergLevels) return cudaErrorInvalidValue;
    k_acw_auc<<<(unsigned)((n_rows + 127) / 128), 128, 0, st>>>(acf, n_rows, ld, w, dt, strides, n_strides, auc);
    return cudaGetLastError();
}

cudaError_t acw_tau_launch(const double* acf, int64_t n_rows, int64_t ld, int n_lags, double dt, double* tau,
                           cudaStream_t st) {
    const int wpb = 4;
    k_acw_tau<<<(unsigned)((n_rows + wpb - 1) / wpb), wpb * 32, 0, st>>>(acf, n_rows, ld, n_lags, dt, tau);
    return cudaGetLastError();
}

// ---- strided input -> kernel layout ---------------------------------------------------------------------------
// out[s + t * n_slices] = raw[s * ss + t * ts].  32 x 32 tiles through shared memory: the tile is read with
// consecutive threads along whichever of (s, t) has the smaller stride in `raw` (ts = 1 for a Julia matrix with
// time down the columns: coalesced 256-byte reads) and written with consecutive threads along s (coalesced).
__global__ void __launch_bounds__(256) k_relayout(const double* __restrict__ raw, int64_t n_slices, int64_t n_t,
                                                  int64_t ss, int64_t ts, double* __restrict__ out) {
    __shared__ double tile[32][33];
    const int64_t s0 = (int64_t)blockIdx.x * 32, t0 = (int64_t)blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;              // 32 x 8
    const bool t_fast = ts <= ss;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int a = ty + 8 * r;
        const int64_t s = s0 + (t_fast ? a : tx), t = t0 + (t_fast ? tx : a);
        if (s < n_slices && t < n_t) tile[t - t0][s - s0] = raw[s * ss + t * ts];
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int64_t t = t0 + ty + 8 * r, s = s0 + tx;
        if (s < n_slices && t < n_t) out[s + t * n_slices] = tile[ty + 8 * r][tx];
    }
}

cudaError_t relayout_launch(const double* raw, int64_t n_slices, int64_t n_t, int64_t ss, int64_t ts, double* out,
                            cudaStream_t st) {
    const dim3 grid((unsigned)((n_slices + 31) / 32), (unsigned)((n_t + 31) / 32));
    if (grid.y > 65535) return cudaErrorInvalidValue;
    k_relayout<<<grid, 256, 0, st>>>(raw, n_slices, n_t, ss, ts, out);
    return cudaGetLastError();
}

}  // namespace itjl
