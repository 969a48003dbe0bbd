// K6: the between-generation step of PMC-ABC on the device (reference src/core/abc.jl:520-567 calc_weights,
// :582-613 weighted_covar): N x N_prev Gaussian-kernel evaluations.  At the default 100 accepted particles it
// is nothing; at BASELINE config C5 (10^6 accepted per generation) it is 10^12 evaluations per generation and
// cannot run on the host at all.
//
//   den_i = sum_j w_j N(theta_i; theta_prev_j, tau^2)
//         = |2 pi tau^2|^-1/2 sum_j w_j exp(-1/2 |z_i - z_j|^2),   z = L^-1 (theta - mu),  tau^2 = L L^T
//
// The host whitens both particle sets in fp64 (so the kernel never sees the covariance and its inputs are
// O(1) numbers centred on the previous population), the kernel does the pair sum:
//   - thread = kRows rows i kept in registers, CTA tile of 256 * kRows rows; the previous particles stream
//     through shared memory in tiles of kTileJ (z_j, w_j), every thread reads them as broadcasts;
//   - the j range is split over blockIdx.y so that small problems still fill the 148 SMs; partial sums land in
//     part[split][row] and are added in a fixed order by k_pmc_finish (bit-deterministic: no atomics);
//   - T = double for up to ~10^8 pairs (the reference's arithmetic, exp() in fp64), T = float above: the
//     exponent argument is an O(1) whitened distance, so fp32 costs ~1e-6 relative on a weight, and the pair
//     rate is the MUFU rate (one ex2 per pair: 148 SMs x 16 / clk) instead of the fp64 exp() rate.
#include <cub/device/device_radix_sort.cuh>

#include "kernels.h"
#include "philox.cuh"

namespace itjl {

constexpr int kPmcThreads = 256;
constexpr int kPmcRows = 4;
constexpr int kPmcTileJ = 1024;

// The host folds the factor of the exponent into the whitened coordinates (pmc_coord_scale): the kernel
// evaluates exp(-|dz|^2) in fp64 and 2^(-|dz|^2) in fp32 (one FADD, one FMUL, one MUFU.EX2, one FFMA per pair).
template <typename T> __device__ __forceinline__ T pmc_kernel_value(T q);
template <> __device__ __forceinline__ double pmc_kernel_value<double>(double q) { return exp(-q); }
template <> __device__ __forceinline__ float pmc_kernel_value<float>(float q) {
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-q));
    return e;
}

// z: [P][n] new particles (whitened), zp: [P][n_prev] previous, wp: [n_prev]; part: [gridDim.y][n]
template <typename T, int P>
__global__ void __launch_bounds__(kPmcThreads) k_pmc_mixture(const T* __restrict__ z, const T* __restrict__ zp,
                                                              const T* __restrict__ wp, int64_t n, int64_t n_prev,
                                                              int64_t j_per_split, double* __restrict__ part) {
    __shared__ T sz[P][kPmcTileJ];
    __shared__ T sw[kPmcTileJ];
    const int tid = threadIdx.x;
    const int64_t row0 = ((int64_t)blockIdx.x * kPmcThreads + tid) * kPmcRows;
    T zi[kPmcRows][P];
    double acc[kPmcRows];
#pragma unroll
    for (int r = 0; r < kPmcRows; ++r) {
        acc[r] = 0.0;
#pragma unroll
        for (int d = 0; d < P; ++d) zi[r][d] = (row0 + r < n) ? z[(int64_t)d * n + row0 + r] : (T)0;
    }
    const int64_t j_lo = (int64_t)blockIdx.y * j_per_split;
    const int64_t j_hi = j_lo + j_per_split < n_prev ? j_lo + j_per_split : n_prev;
    for (int64_t jt = j_lo; jt < j_hi; jt += kPmcTileJ) {
        const int cnt = (int)(j_hi - jt < kPmcTileJ ? j_hi - jt : kPmcTileJ);
        __syncthreads();
        for (int k = tid; k < cnt; k += kPmcThreads) {
#pragma unroll
            for (int d = 0; d < P; ++d) sz[d][k] = zp[(int64_t)d * n_prev + jt + k];
            sw[k] = wp[jt + k];
        }
        __syncthreads();
        T a[kPmcRows];
#pragma unroll
        for (int r = 0; r < kPmcRows; ++r) a[r] = (T)0;
#pragma unroll 4
        for (int k = 0; k < cnt; ++k) {
            const T w = sw[k];
#pragma unroll
            for (int r = 0; r < kPmcRows; ++r) {
                T q = (T)0;
#pragma unroll
                for (int d = 0; d < P; ++d) { const T df = zi[r][d] - sz[d][k]; q = fma(df, df, q); }
                a[r] = fma(w, pmc_kernel_value<T>(q), a[r]);
            }
        }
#pragma unroll
        for (int r = 0; r < kPmcRows; ++r) acc[r] += (double)a[r];      // tile sums (<= 1024 terms) leave the fp32 domain here
    }
#pragma unroll
    for (int r = 0; r < kPmcRows; ++r)
        if (row0 + r < n) part[(int64_t)blockIdx.y * n + row0 + r] = acc[r];
}

// w_i = prior_i / (norm * sum_split part[split][i]); block sums of w for the normalisation
__global__ void __launch_bounds__(256) k_pmc_finish(const double* __restrict__ part, int n_split, int64_t n,
                                                    const double* __restrict__ prior_pdf, double norm,
                                                    double* __restrict__ w_out, double* __restrict__ block_sums) {
    __shared__ double red[8];
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    double w = 0.0;
    if (i < n) {
        double den = 0.0;
        for (int s = 0; s < n_split; ++s) den += part[(int64_t)s * n + i];
        w = prior_pdf[i] / (norm * den);
        w_out[i] = w;
    }
    w = warp_sum(w);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = w;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int k = 0; k < 8; ++k) t += red[k];
        block_sums[blockIdx.x] = t;
    }
}

// One-parameter models with large populations: the mixture density f(z) = sum_j w_j exp(-(z - z_j)^2) is a smooth function
// of ONE whitened coordinate (every term has width 1 in z), so it is evaluated exactly (fp64 pair sums, k_pmc_mixture)
// on a uniform grid of n_grid points spanning the new particles and read off by 4-point Lagrange interpolation:
// error <= (3/128) h^4 max|d^4 f / dz^4| ~ 0.3 h^4 f_max with h = range / n_grid ~ 1e-3, i.e. ~1e-12 of the density's
// scale (below 1e-9 relative out to the tails) - tighter than the fp32 pair sums it replaces, at n_grid x n_prev
// instead of n x n_prev pairs.  dens[m] = sum over the j splits of part[split][m] (fixed order).
__global__ void __launch_bounds__(256) k_pmc_grid_density(const double* __restrict__ part, int n_split, int n_grid,
                                                          double* __restrict__ dens) {
    const int m = blockIdx.x * 256 + threadIdx.x;
    if (m >= n_grid) return;
    double d = 0.0;
    for (int s = 0; s < n_split; ++s) d += part[(int64_t)s * n_grid + m];
    dens[m] = d;
}
// grid point m sits at z_lo + (m - 1) h; w_i = prior_i / (norm * f(z_i)); block sums of w for the normalisation
__global__ void __launch_bounds__(256) k_pmc_grid_weights(const double* __restrict__ dens, int n_grid, double z_lo, double inv_h,
                                                          const double* __restrict__ z, int64_t n,
                                                          const double* __restrict__ prior_pdf, double norm,
                                                          double* __restrict__ w_out, double* __restrict__ block_sums) {
    __shared__ double red[8];
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    double w = 0.0;
    if (i < n) {
        const double u = (z[i] - z_lo) * inv_h + 1.0;
        int m0 = (int)floor(u);
        m0 = m0 < 1 ? 1 : (m0 > n_grid - 3 ? n_grid - 3 : m0);
        const double t = u - (double)m0;                               // in [0, 1] (up to rounding at the ends)
        const double fm = dens[m0 - 1], f0 = dens[m0], f1 = dens[m0 + 1], f2 = dens[m0 + 2];
        // Lagrange basis on nodes -1, 0, 1, 2
        const double lm = -t * (t - 1.0) * (t - 2.0) * (1.0 / 6.0);
        const double l0 = (t + 1.0) * (t - 1.0) * (t - 2.0) * 0.5;
        const double l1 = -(t + 1.0) * t * (t - 2.0) * 0.5;
        const double l2 = (t + 1.0) * t * (t - 1.0) * (1.0 / 6.0);
        const double den = fma(lm, fm, fma(l0, f0, fma(l1, f1, l2 * f2)));
        w = prior_pdf[i] / (norm * den);
        w_out[i] = w;
    }
    w = warp_sum(w);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = w;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int k = 0; k < 8; ++k) t += red[k];
        block_sums[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(256) k_scale(double* __restrict__ w, int64_t n, const double* __restrict__ block_sums, int n_blocks) {
    __shared__ double total;
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int k = 0; k < n_blocks; ++k) t += block_sums[k];          // fixed order
        total = t;
    }
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i < n) w[i] = w[i] / total;
}

template <typename T>
static cudaError_t mixture_launch_t(const void* z, const void* zp, const void* wp, int64_t n, int64_t n_prev, int p,
                                    int n_split, int64_t j_per_split, double* part, cudaStream_t st) {
    const dim3 grid((unsigned)((n + (int64_t)kPmcThreads * kPmcRows - 1) / ((int64_t)kPmcThreads * kPmcRows)), (unsigned)n_split);
    const T* a = (const T*)z; const T* b = (const T*)zp; const T* c = (const T*)wp;
    switch (p) {
        case 1: k_pmc_mixture<T, 1><<<grid, kPmcThreads, 0, st>>>(a, b, c, n, n_prev, j_per_split, part); break;
        case 2: k_pmc_mixture<T, 2><<<grid, kPmcThreads, 0, st>>>(a, b, c, n, n_prev, j_per_split, part); break;
        case 3: k_pmc_mixture<T, 3><<<grid, kPmcThreads, 0, st>>>(a, b, c, n, n_prev, j_per_split, part); break;
        case 4: k_pmc_mixture<T, 4><<<grid, kPmcThreads, 0, st>>>(a, b, c, n, n_prev, j_per_split, part); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

// z = coord_scale * L^-1 (theta - mu): exp(-1/2 |.|^2) = exp(-|z|^2) (fp64) = 2^(-|z|^2) (fp32)
double pmc_coord_scale(bool fp32) { return fp32 ? 0.84932180028801904 /* sqrt(log2(e) / 2) */ : 0.70710678118654752; }

void pmc_mixture_split(int64_t n, int64_t n_prev, int sm_count, int* n_split, int64_t* j_per_split) {
    const int64_t row_ctas = (n + (int64_t)kPmcThreads * kPmcRows - 1) / ((int64_t)kPmcThreads * kPmcRows);
    int64_t want = (2 * (int64_t)sm_count + row_ctas - 1) / row_ctas;              // ~2 CTAs per SM
    const int64_t max_split = (n_prev + kPmcTileJ - 1) / kPmcTileJ;
    if (want > max_split) want = max_split;
    if (want < 1) want = 1;
    if (want > 1024) want = 1024;
    int64_t per = (n_prev + want - 1) / want;
    per = (per + kPmcTileJ - 1) / kPmcTileJ * kPmcTileJ;
    *j_per_split = per;
    *n_split = (int)((n_prev + per - 1) / per);
}

cudaError_t pmc_mixture_launch(bool fp32, const void* z, const void* zp, const void* wp, int64_t n, int64_t n_prev, int p,
                               int n_split, int64_t j_per_split, double* part, cudaStream_t st) {
    return fp32 ? mixture_launch_t<float>(z, zp, wp, n, n_prev, p, n_split, j_per_split, part, st)
                : mixture_launch_t<double>(z, zp, wp, n, n_prev, p, n_split, j_per_split, part, st);
}

cudaError_t pmc_finish_launch(const double* part, int n_split, int64_t n, const double* prior_pdf, double norm,
                              double* w_out, double* block_sums, bool normalize, cudaStream_t st) {
    const int blocks = (int)((n + 255) / 256);
    k_pmc_finish<<<blocks, 256, 0, st>>>(part, n_split, n, prior_pdf, norm, w_out, block_sums);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess || !normalize) return e;
    k_scale<<<blocks, 256, 0, st>>>(w_out, n, block_sums, blocks);
    return cudaGetLastError();
}

cudaError_t pmc_grid_density_launch(const double* part, int n_split, int n_rows, double* dens, cudaStream_t st) {
    k_pmc_grid_density<<<(n_rows + 255) / 256, 256, 0, st>>>(part, n_split, n_rows, dens);
    return cudaGetLastError();
}
cudaError_t pmc_grid_weights_launch(const double* dens, int n_grid, double z_lo, double h, const double* z, int64_t n,
                                    const double* prior_pdf, double norm, double* w_out, double* block_sums, cudaStream_t st) {
    k_pmc_grid_weights<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(dens, n_grid, z_lo, 1.0 / h, z, n, prior_pdf, norm, w_out, block_sums);
    return cudaGetLastError();
}

__global__ void __launch_bounds__(256) k_block_sums(const double* __restrict__ w, int64_t n, double* __restrict__ block_sums) {
    __shared__ double red[8];
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    double v = warp_sum(i < n ? w[i] : 0.0);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int k = 0; k < 8; ++k) t += red[k];
        block_sums[blockIdx.x] = t;
    }
}

// w <- w / sum(w), the sum taken in a fixed order (block sums, then serially)
cudaError_t pmc_normalize_launch(double* w, int64_t n, double* block_sums, cudaStream_t st) {
    const int blocks = (int)((n + 255) / 256);
    k_block_sums<<<blocks, 256, 0, st>>>(w, n, block_sums);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    k_scale<<<blocks, 256, 0, st>>>(w, n, block_sums, blocks);
    return cudaGetLastError();
}

// ---- weighted covariance (abc.jl:582-613): two passes, per-CTA partial sums combined on the host ------------
// pass 0: out[cta][0] = sum w, [1] = sum w^2, [2 + k] = sum w x_k          (x: [p][n] column-major)
// pass 1: out[cta][j * p + k] = sum w (x_j - xbar_j)(x_k - xbar_k)        (w already divided by sum w: `wsum`)
constexpr int kWcovMaxP = 4;
__global__ void __launch_bounds__(256) k_wcov(const double* __restrict__ x, const double* __restrict__ w, int64_t n, int p,
                                              int pass, double wsum, const double* __restrict__ xbar, double* __restrict__ out) {
    __shared__ double red[8][2 + kWcovMaxP * kWcovMaxP];
    const int n_out = pass == 0 ? 2 + p : p * p;
    double acc[2 + kWcovMaxP * kWcovMaxP];
    for (int k = 0; k < n_out; ++k) acc[k] = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
        const double wi = w[i] / wsum;
        if (pass == 0) {
            acc[0] += wi; acc[1] += wi * wi;
            for (int k = 0; k < p; ++k) acc[2 + k] += wi * x[(int64_t)k * n + i];
        } else {
            double c[kWcovMaxP];
            for (int k = 0; k < p; ++k) c[k] = x[(int64_t)k * n + i] - xbar[k];
            for (int j = 0; j < p; ++j)
                for (int k = 0; k < p; ++k) acc[j * p + k] += c[j] * c[k] * wi;
        }
    }
    for (int k = 0; k < n_out; ++k) {
        const double v = warp_sum(acc[k]);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][k] = v;
    }
    __syncthreads();
    if (threadIdx.x < n_out) {
        double t = 0.0;
        for (int wv = 0; wv < 8; ++wv) t += red[wv][threadIdx.x];
        out[(int64_t)blockIdx.x * n_out + threadIdx.x] = t;
    }
}

cudaError_t wcov_launch(const double* x, const double* w, int64_t n, int p, int pass, double wsum, const double* xbar,
                        double* out, int blocks, cudaStream_t st) {
    if (p < 1 || p > kWcovMaxP) return cudaErrorInvalidValue;
    k_wcov<<<blocks, 256, 0, st>>>(x, w, n, p, pass, wsum, xbar, out);
    return cudaGetLastError();
}


// ---- PMC proposals (draw_theta_pmc, reference src/core/abc.jl:101-117), one thread per proposal ----------------
// theta* = theta_prev[sample(pweights(weights))] by inverse-cdf look-up (53-bit uniform against the running sum of
// the weights), theta = theta* + L z with L the Cholesky factor of tau_squared + jitter I and z ~ N(0, I) (Box-Muller
// on Philox block `attempt`), redrawn around the SAME theta* while any component is negative (:112-115).
// Counter-based: proposal i of (seed, generation) does not depend on n, on the launch geometry or on the GPU count.
struct PmcProposeParams {
    const double* theta_prev;    // (n_prev x p) column-major
    const double* cdf;           // running sum of the weights, n_prev
    int64_t n_prev, n;
    int p;
    double L[16];                // lower Cholesky factor, column-major p x p
    PhiloxKeys keys;
    uint32_t c3;
    int64_t first;               // global index of proposal 0 of this launch
    double* out;                 // (n x p) column-major
};

__global__ void __launch_bounds__(256) k_pmc_propose(const __grid_constant__ PmcProposeParams P) {
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= P.n) return;
    const uint32_t gi = (uint32_t)(P.first + i), gh = (uint32_t)((uint64_t)(P.first + i) >> 32);
    uint32_t r[4];
    philox4x32(0u, gh, gi, P.c3, P.keys, r);
    const double u = ((double)(((uint64_t)r[0] << 21) ^ (uint64_t)(r[1] >> 11)) + 0.5) * 1.1102230246251565e-16;   // (0, 1), 53 bits
    const double target = u * P.cdf[P.n_prev - 1];
    int64_t lo = 0, hi = P.n_prev - 1;                   // first index with cdf > target
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (__ldg(P.cdf + mid) > target) hi = mid; else lo = mid + 1;
    }
    double star[4], th[4];
    for (int d = 0; d < P.p; ++d) star[d] = P.theta_prev[lo + (int64_t)d * P.n_prev];
    for (uint32_t attempt = 1u;; ++attempt) {
        philox4x32(attempt, gh, gi, P.c3, P.keys, r);
        float z[4];
        box_muller(r[0], r[1], z[0], z[1]);
        box_muller(r[2], r[3], z[2], z[3]);
        bool neg = false;
        for (int d = 0; d < P.p; ++d) {
            double v = star[d];
            for (int k = 0; k <= d; ++k) v = fma(P.L[d + k * P.p], (double)z[k], v);
            th[d] = v;
            neg |= v < 0.0;
        }
        if (!neg || attempt == 0xFFFFu) break;           // (65535 redraws: theta* itself is far below zero - give up as is)
    }
    for (int d = 0; d < P.p; ++d) P.out[i + (int64_t)d * P.n] = th[d];
}

cudaError_t pmc_propose_launch(const double* theta_prev, const double* cdf, int64_t n_prev, int p, const double* L, int64_t n,
                               uint64_t seed, uint64_t generation, int64_t first, double* out, cudaStream_t st) {
    PmcProposeParams P{};
    P.theta_prev = theta_prev; P.cdf = cdf; P.n_prev = n_prev; P.n = n; P.p = p;
    for (int i = 0; i < p * p; ++i) P.L[i] = L[i];
    P.keys = philox_keys(seed);
    P.c3 = philox_c3(kStreamPropose, generation);
    P.first = first; P.out = out;
    k_pmc_propose<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(P);
    return cudaGetLastError();
}


// ---- order statistics of the distances (select_epsilon, reference src/core/abc.jl:700-760) ------------------------
// values that are NaN or >= upper are moved to +inf (they sort last) and counted out; the sort itself is CUB's radix
// sort - a library call, once per generation on <= 10^6 doubles, not a hot kernel.
__global__ void __launch_bounds__(256) k_mask_invalid(const double* __restrict__ v, int64_t n, double upper, double* __restrict__ out,
                                                      unsigned long long* __restrict__ n_valid) {
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    bool ok = false;
    if (i < n) {
        const double x = v[i];
        ok = (x == x) && x < upper;
        out[i] = ok ? x : INFINITY;
    }
    const unsigned m = __ballot_sync(0xffffffffu, ok);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(n_valid, (unsigned long long)__popc(m));
}
__global__ void k_gather(const double* __restrict__ sorted, const int64_t* __restrict__ idx, int k, double* __restrict__ out) {
    const int i = threadIdx.x;
    if (i < k) out[i] = sorted[idx[i]];
}

size_t sort_temp_bytes(int64_t n) {
    size_t bytes = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, bytes, (const double*)nullptr, (double*)nullptr, (int)n);
    return bytes;
}
cudaError_t mask_invalid_launch(const double* v, int64_t n, double upper, double* out, unsigned long long* n_valid, cudaStream_t st) {
    k_mask_invalid<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(v, n, upper, out, n_valid);
    return cudaGetLastError();
}
cudaError_t sort_doubles_launch(void* temp, size_t temp_bytes, const double* in, double* out, int64_t n, cudaStream_t st) {
    return cub::DeviceRadixSort::SortKeys(temp, temp_bytes, in, out, (int)n, 0, 64, st);
}
cudaError_t gather_launch(const double* sorted, const int64_t* idx, int k, double* out, cudaStream_t st) {
    k_gather<<<1, 64, 0, st>>>(sorted, idx, k, out);
    return cudaGetLastError();
}

}  // namespace itjl
