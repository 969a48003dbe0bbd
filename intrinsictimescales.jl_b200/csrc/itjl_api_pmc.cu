// C-ABI layer, part 4: the between-generation step of PMC-ABC (reference src/core/abc.jl:520-567, 582-613).
#include <math.h>

#include <vector>

#include "api_internal.h"

using namespace itjl;

namespace {

// lower Cholesky factor of a symmetric positive definite p x p matrix (column-major); false when not SPD
bool cholesky(const double* a, int p, double* L) {
    for (int i = 0; i < p * p; ++i) L[i] = 0.0;
    for (int j = 0; j < p; ++j) {
        double d = a[j + j * p];
        for (int k = 0; k < j; ++k) d -= L[j + k * p] * L[j + k * p];
        if (!(d > 0.0) || !isfinite(d)) return false;
        L[j + j * p] = sqrt(d);
        for (int i = j + 1; i < p; ++i) {
            double s = a[i + j * p];
            for (int k = 0; k < j; ++k) s -= L[i + k * p] * L[j + k * p];
            L[i + j * p] = s / L[j + j * p];
        }
    }
    return true;
}

// z[d][i] = scale * (L^-1 (theta[i, :] - mu))[d], stored as T, for rows [lo, hi) of a column-major (n x p) matrix
template <typename T>
void whiten(const double* theta, int64_t n, int p, const double* L, const double* mu, double scale, int64_t lo, int64_t hi, T* z) {
    const int64_t m = hi - lo;
    for (int64_t i = 0; i < m; ++i) {
        double y[8];
        for (int d = 0; d < p; ++d) {
            double s = theta[(lo + i) + (int64_t)d * n] - mu[d];
            for (int k = 0; k < d; ++k) s -= L[d + k * p] * y[k];
            y[d] = s / L[d + d * p];
        }
        for (int d = 0; d < p; ++d) z[(int64_t)d * m + i] = (T)(scale * y[d]);
    }
}

}  // namespace

extern "C" {

int itjl_pmc_weights(itjl_ctx* ctx, const double* theta_prev, int64_t n_prev, const double* theta, int64_t n,
                     int32_t n_theta, const double* tau_squared, const double* weights_prev,
                     const double* prior_pdf, int32_t normalize, double* weights_out) {
    if (!ctx) return fail(nullptr, ITJL_ERR_ARG, "itjl_pmc_weights: null context");
    if (!theta_prev || !theta || !tau_squared || !weights_prev || !prior_pdf || !weights_out)
        return fail(ctx, ITJL_ERR_ARG, "itjl_pmc_weights: null argument");
    if (n < 1 || n_prev < 1 || n_theta < 1 || n_theta > 4) return fail(ctx, ITJL_ERR_ARG, "itjl_pmc_weights: need n, n_prev >= 1 and 1 <= n_theta <= 4");
    ITJL_LOCK(ctx);
    NvtxRange nvtx("itjl_pmc_weights");
    const int p = n_theta;
    double L[16], mu[4] = {0, 0, 0, 0};
    if (!cholesky(tau_squared, p, L)) return fail(ctx, ITJL_ERR_ARG, "itjl_pmc_weights: tau_squared is not positive definite");
    double logdet = 0.0;
    for (int d = 0; d < p; ++d) logdet += 2.0 * log(L[d + d * p]);
    const double norm = exp(-0.5 * (p * log(2.0 * M_PI) + logdet));      // 1 / sqrt((2 pi)^p det tau^2)
    for (int d = 0; d < p; ++d) {                                        // centre on the previous population (fp32 path: O(1) inputs)
        double s = 0.0;
        for (int64_t j = 0; j < n_prev; ++j) s += theta_prev[j + (int64_t)d * n_prev];
        mu[d] = s / (double)n_prev;
    }
    // one parameter and a large population: exact fp64 density on a grid + interpolation (k_pmc_grid_weights)
    const bool grid1d = p == 1 && n >= 16 && (ctx->tuning.pmc_fp32 == 2 || (ctx->tuning.pmc_fp32 == -1 && (double)n * (double)n_prev > 1e8 && n >= 4 * 8192));
    const int n_grid = 8192;
    const bool fp32 = !grid1d && (ctx->tuning.pmc_fp32 == 1 || (ctx->tuning.pmc_fp32 == -1 && (double)n * (double)n_prev > 1e8));
    const double scale = pmc_coord_scale(fp32);
    const size_t el = fp32 ? sizeof(float) : sizeof(double);

    std::vector<ShardMember> mem;
    int world = 1;
    shard_members(ctx, &mem, &world);
    const int64_t width = (n + world - 1) / world;
    // previous population: whitened once on the host, uploaded to every member
    std::vector<unsigned char> zp((size_t)n_prev * p * el), wp((size_t)n_prev * el);
    if (fp32) {
        whiten<float>(theta_prev, n_prev, p, L, mu, scale, 0, n_prev, (float*)zp.data());
        for (int64_t j = 0; j < n_prev; ++j) ((float*)wp.data())[j] = (float)weights_prev[j];
    } else {
        whiten<double>(theta_prev, n_prev, p, L, mu, scale, 0, n_prev, (double*)zp.data());
        memcpy(wp.data(), weights_prev, (size_t)n_prev * sizeof(double));
    }
    std::vector<std::vector<unsigned char>> zs(mem.size());
    // grid path: whitened coordinate of every new particle (range of the grid), the grid itself
    std::vector<double> z_all, grid;
    double z_lo = 0.0, h = 1.0;
    bool use_grid = grid1d;
    if (use_grid) {
        z_all.resize((size_t)n);
        whiten<double>(theta, n, 1, L, mu, scale, 0, n, z_all.data());
        double lo = z_all[0], hi = z_all[0];
        bool finite = true;
        for (int64_t i = 0; i < n; ++i) {
            const double v = z_all[(size_t)i];
            if (!(v == v) || std::isinf(v)) finite = false;
            lo = v < lo ? v : lo; hi = v > hi ? v : hi;
        }
        if (!finite || !(hi > lo)) use_grid = false;                // degenerate population: the exact pair sums handle it
        else {
            z_lo = lo; h = (hi - lo) / (double)(n_grid - 3);        // points lie in [g_1, g_{n_grid - 2}]
            grid.resize((size_t)n_grid);
            for (int m = 0; m < n_grid; ++m) grid[(size_t)m] = z_lo + (double)(m - 1) * h;
        }
    }
    // grid rows are sharded over the members too: each one sums its slice of the density, an in-place all-gather
    // (n_grid doubles) gives every member the whole grid, then each interpolates its own particles
    const int64_t g_width = (n_grid + world - 1) / world;
    const int n_grid_pad = (int)(g_width * world);
    std::vector<double*> dens_bases(mem.size(), nullptr);
    for (size_t i = 0; use_grid && i < mem.size(); ++i) {
        itjl_ctx* c = mem[i].c;
        ITJL_LOCK(c);
        ITJL_CUDA(ctx, cudaSetDevice(c->device));
        const int64_t g_lo = std::min<int64_t>((int64_t)mem[i].rank * g_width, n_grid), g_hi = std::min<int64_t>(g_lo + g_width, n_grid);
        const int64_t ng = g_hi - g_lo;
        // the split of the previous particles is a function of n_prev alone, so every grid point's sum has the same
        // grouping on any number of GPUs (results identical to one GPU's)
        int64_t j_per = 8192;
        if ((n_prev + j_per - 1) / j_per > 1024) j_per = (((n_prev + 1023) / 1024 + 1023) / 1024) * 1024;
        const int n_split = (int)((n_prev + j_per - 1) / j_per);
        // device layout in c->data: [zp | wp | grid slice]; c->part: [density of all grid points (padded) | partial sums]
        const size_t off_wp = (size_t)n_prev * sizeof(double), off_g = off_wp + (size_t)n_prev * sizeof(double);
        ITJL_CUDA(ctx, c->data.reserve(off_g + (size_t)g_width * sizeof(double)));
        ITJL_CUDA(ctx, c->part.reserve(((size_t)n_grid_pad + (size_t)n_split * g_width) * sizeof(double)));
        dens_bases[i] = (double*)c->part.p;
        if (ng <= 0) continue;
        unsigned char* base = (unsigned char*)c->data.p;
        double* part = dens_bases[i] + n_grid_pad;
        ITJL_CUDA(ctx, cudaMemcpyAsync(base, zp.data(), zp.size(), cudaMemcpyHostToDevice, c->stream));
        ITJL_CUDA(ctx, cudaMemcpyAsync(base + off_wp, wp.data(), wp.size(), cudaMemcpyHostToDevice, c->stream));
        ITJL_CUDA(ctx, cudaMemcpyAsync(base + off_g, grid.data() + g_lo, (size_t)ng * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        cudaError_t e = pmc_mixture_launch(false, base + off_g, base, base + off_wp, ng, n_prev, 1, n_split, j_per, part, c->stream);
        if (e != cudaSuccess) return fail(ctx, ITJL_ERR_CUDA, "launch k_pmc_mixture(grid): %s", cudaGetErrorString(e));
        e = pmc_grid_density_launch(part, n_split, (int)ng, dens_bases[i] + g_lo, c->stream);
        if (e != cudaSuccess) return fail(ctx, ITJL_ERR_CUDA, "launch k_pmc_grid_density: %s", cudaGetErrorString(e));
        c->launches += 2;
    }
    if (use_grid) {
        const int rcg = shard_allgather_inplace(ctx, dens_bases, g_width);
        if (rcg != ITJL_OK) return rcg;
    }
    for (size_t i = 0; use_grid && i < mem.size(); ++i) {
        itjl_ctx* c = mem[i].c;
        ITJL_LOCK(c);
        int64_t lo, hi;
        shard_bounds(n, world, mem[i].rank, &lo, &hi);
        ITJL_CUDA(ctx, cudaSetDevice(c->device));
        ITJL_CUDA(ctx, c->gather.reserve((size_t)world * width * sizeof(double)));
        if (hi <= lo) continue;
        const int64_t nl = hi - lo;
        ITJL_CUDA(ctx, c->out.reserve((size_t)nl * sizeof(double)));
        ITJL_CUDA(ctx, c->out2.reserve((size_t)nl * sizeof(double)));
        ITJL_CUDA(ctx, c->out3.reserve((size_t)((nl + 255) / 256) * sizeof(double)));
        ITJL_CUDA(ctx, cudaMemcpyAsync(c->out.p, z_all.data() + lo, (size_t)nl * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        ITJL_CUDA(ctx, cudaMemcpyAsync(c->out2.p, prior_pdf + lo, (size_t)nl * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        cudaError_t e = pmc_grid_weights_launch(dens_bases[i], n_grid, z_lo, h, (const double*)c->out.p, nl, (const double*)c->out2.p, norm,
                                                (double*)c->gather.p + (int64_t)mem[i].rank * width, (double*)c->out3.p, c->stream);
        if (e != cudaSuccess) return fail(ctx, ITJL_ERR_CUDA, "launch k_pmc_grid_weights: %s", cudaGetErrorString(e));
        c->launches++;
    }
    for (size_t i = 0; !use_grid && i < mem.size(); ++i) {
        itjl_ctx* c = mem[i].c;
        ITJL_LOCK(c);
        int64_t lo, hi;
        shard_bounds(n, world, mem[i].rank, &lo, &hi);
        ITJL_CUDA(ctx, cudaSetDevice(c->device));
        ITJL_CUDA(ctx, c->gather.reserve((size_t)world * width * sizeof(double)));
        if (hi <= lo) continue;
        const int64_t nl = hi - lo;
        zs[i].resize((size_t)nl * p * el);
        if (fp32) whiten<float>(theta, n, p, L, mu, scale, lo, hi, (float*)zs[i].data());
        else whiten<double>(theta, n, p, L, mu, scale, lo, hi, (double*)zs[i].data());
        int n_split = 1;
        int64_t j_per = n_prev;
        pmc_mixture_split(nl, n_prev, c->sm_count, &n_split, &j_per);
        // device layout in c->data: [zp | wp | z], partial sums in c->part, prior pdf in c->out2, block sums in c->out3
        const size_t off_wp = (size_t)n_prev * p * el, off_z = off_wp + (((size_t)n_prev * el + 15) & ~(size_t)15);
        ITJL_CUDA(ctx, c->data.reserve(off_z + (size_t)nl * p * el));
        ITJL_CUDA(ctx, c->part.reserve((size_t)n_split * nl * sizeof(double)));
        ITJL_CUDA(ctx, c->out2.reserve((size_t)nl * sizeof(double)));
        ITJL_CUDA(ctx, c->out3.reserve((size_t)((nl + 255) / 256) * sizeof(double)));
        unsigned char* base = (unsigned char*)c->data.p;
        ITJL_CUDA(ctx, cudaMemcpyAsync(base, zp.data(), zp.size(), cudaMemcpyHostToDevice, c->stream));
        ITJL_CUDA(ctx, cudaMemcpyAsync(base + off_wp, wp.data(), wp.size(), cudaMemcpyHostToDevice, c->stream));
        ITJL_CUDA(ctx, cudaMemcpyAsync(base + off_z, zs[i].data(), zs[i].size(), cudaMemcpyHostToDevice, c->stream));
        ITJL_CUDA(ctx, cudaMemcpyAsync(c->out2.p, prior_pdf + lo, (size_t)nl * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        cudaError_t e = pmc_mixture_launch(fp32, base + off_z, base, base + off_wp, nl, n_prev, p, n_split, j_per,
                                           (double*)c->part.p, c->stream);
        if (e != cudaSuccess) return fail(ctx, ITJL_ERR_CUDA, "launch k_pmc_mixture: %s", cudaGetErrorString(e));
        // un-normalised weights of this shard straight into its all-gather slot (normalised below, over all rows)
        e = pmc_finish_launch((const double*)c->part.p, n_split, nl, (const double*)c->out2.p, norm,
                              (double*)c->gather.p + (int64_t)mem[i].rank * width, (double*)c->out3.p, false, c->stream);
        if (e != cudaSuccess) return fail(ctx, ITJL_ERR_CUDA, "launch k_pmc_finish: %s", cudaGetErrorString(e));
        c->launches += 2;
    }
    double* w_dev = nullptr;
    int rc = shard_allgather_compact(ctx, n, &w_dev);
    if (rc != ITJL_OK) return rc;
    itjl_ctx* c0 = primary(ctx);
    ITJL_LOCK(c0);
    if (normalize) {
        ITJL_CUDA(ctx, c0->out3.reserve((size_t)((n + 255) / 256) * sizeof(double)));
        cudaError_t e = pmc_normalize_launch(w_dev, n, (double*)c0->out3.p, c0->stream);
        if (e != cudaSuccess) return fail(ctx, ITJL_ERR_CUDA, "launch k_block_sums / k_scale: %s", cudaGetErrorString(e));
        c0->launches += 2;
    }
    ITJL_CUDA(ctx, cudaMemcpyAsync(weights_out, w_dev, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, c0->stream));
    ITJL_CUDA(ctx, cudaStreamSynchronize(c0->stream));
    for (const ShardMember& sm : mem) {                     // host staging vectors stay alive until every copy is done
        cudaSetDevice(sm.c->device);
        ITJL_CUDA(ctx, cudaStreamSynchronize(sm.c->stream));
    }
    return ITJL_OK;
}

int itjl_weighted_covar(itjl_ctx* ctx, const double* x, int64_t n, int32_t n_theta, const double* w, double* cov_out) {
    if (!ctx) return fail(nullptr, ITJL_ERR_ARG, "itjl_weighted_covar: null context");
    if (!x || !w || !cov_out || n < 1 || n_theta < 1 || n_theta > 4) return fail(ctx, ITJL_ERR_ARG, "itjl_weighted_covar: invalid argument");
    ITJL_LOCK(ctx);
    ctx = primary(ctx);
    ITJL_LOCK(ctx);
    NvtxRange nvtx("itjl_weighted_covar");
    const int p = n_theta;
    ITJL_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc = upload(ctx, ctx->data, x, (size_t)n * p * sizeof(double));
    if (rc == ITJL_OK) rc = upload(ctx, ctx->out2, w, (size_t)n * sizeof(double));
    if (rc != ITJL_OK) return rc;
    const int blocks = (int)std::min<int64_t>((n + 255) / 256, 2 * (int64_t)ctx->sm_count);
    const int n0 = 2 + p, n1 = p * p;
    ITJL_CUDA(ctx, ctx->out3.reserve((size_t)blocks * (n0 > n1 ? n0 : n1) * sizeof(double) + 64));
    ITJL_CUDA(ctx, ctx->scratch.reserve(8 * sizeof(double)));
    std::vector<double> part((size_t)blocks * (n0 > n1 ? n0 : n1));
    // the reference normalises w by its sum first (abc.jl:584): pass 0 with wsum = 1 gives that sum, then both
    // passes run on w / sum(w)
    double raw_sum = 0.0;
    for (int stage = 0; stage < 2; ++stage) {
        cudaError_t e = wcov_launch((const double*)ctx->data.p, (const double*)ctx->out2.p, n, p, 0, stage == 0 ? 1.0 : raw_sum,
                                    nullptr, (double*)ctx->out3.p, blocks, ctx->stream);
        if (e != cudaSuccess) return fail(ctx, ITJL_ERR_CUDA, "launch k_wcov: %s", cudaGetErrorString(e));
        ctx->launches++;
        ITJL_CUDA(ctx, cudaMemcpyAsync(part.data(), ctx->out3.p, (size_t)blocks * n0 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        ITJL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (stage == 0) { raw_sum = 0.0; for (int b = 0; b < blocks; ++b) raw_sum += part[(size_t)b * n0]; }
    }
    double sumw = 0.0, sum2 = 0.0, xbar[4] = {0, 0, 0, 0};
    for (int b = 0; b < blocks; ++b) {
        sumw += part[(size_t)b * n0]; sum2 += part[(size_t)b * n0 + 1];
        for (int k = 0; k < p; ++k) xbar[k] += part[(size_t)b * n0 + 2 + k];
    }
    if (!(fabs(sumw - 1.0) <= 1e-10 * fmax(fabs(sumw), 1.0))) {          // abc.jl:588-591: fallback covariance
        for (int i = 0; i < p * p; ++i) cov_out[i] = 0.0;
        return ITJL_OK;
    }
    ITJL_CUDA(ctx, cudaMemcpyAsync(ctx->scratch.p, xbar, sizeof(xbar), cudaMemcpyHostToDevice, ctx->stream));
    cudaError_t e = wcov_launch((const double*)ctx->data.p, (const double*)ctx->out2.p, n, p, 1, raw_sum,
                                (const double*)ctx->scratch.p, (double*)ctx->out3.p, blocks, ctx->stream);
    if (e != cudaSuccess) return fail(ctx, ITJL_ERR_CUDA, "launch k_wcov: %s", cudaGetErrorString(e));
    ctx->launches++;
    ITJL_CUDA(ctx, cudaMemcpyAsync(part.data(), ctx->out3.p, (size_t)blocks * n1 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    ITJL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const double f = sumw / (sumw * sumw - sum2);
    for (int i = 0; i < n1; ++i) {
        double s = 0.0;
        for (int b = 0; b < blocks; ++b) s += part[(size_t)b * n1 + i];
        cov_out[i] = s * f;
    }
    return ITJL_OK;
}


int itjl_pmc_propose(itjl_ctx* ctx, const double* theta_prev, int64_t n_prev, int32_t n_theta, const double* weights_prev,
                     const double* tau_squared, double jitter, int64_t n, uint64_t seed, uint64_t generation,
                     int64_t first, double* theta_out) {
    if (!ctx) return fail(nullptr, ITJL_ERR_ARG, "itjl_pmc_propose: null context");
    if (!theta_prev || !weights_prev || !tau_squared || !theta_out) return fail(ctx, ITJL_ERR_ARG, "itjl_pmc_propose: null argument");
    if (n < 1 || n_prev < 1 || n_theta < 1 || n_theta > 4 || !(jitter >= 0.0) || first < 0)
        return fail(ctx, ITJL_ERR_ARG, "itjl_pmc_propose: need n, n_prev >= 1, 1 <= n_theta <= 4, jitter >= 0, first >= 0");
    ITJL_LOCK(ctx);
    ctx = primary(ctx);                                  // every rank / device draws the same proposals: nothing to shard
    ITJL_LOCK(ctx);
    NvtxRange nvtx("itjl_pmc_propose");
    const int p = n_theta;
    double cov[16], L[16];
    for (int i = 0; i < p * p; ++i) cov[i] = tau_squared[i];
    for (int d = 0; d < p; ++d) cov[d + d * p] += jitter;                     // abc.jl:106-108
    if (!cholesky(cov, p, L)) return fail(ctx, ITJL_ERR_ARG, "itjl_pmc_propose: tau_squared + jitter I is not positive definite");
    std::vector<double> cdf((size_t)n_prev);
    double run = 0.0;
    for (int64_t j = 0; j < n_prev; ++j) {
        if (!(weights_prev[j] >= 0.0)) return fail(ctx, ITJL_ERR_ARG, "itjl_pmc_propose: weights must be non-negative and finite");
        run += weights_prev[j];
        cdf[(size_t)j] = run;
    }
    if (!(run > 0.0) || !isfinite(run)) return fail(ctx, ITJL_ERR_ARG, "itjl_pmc_propose: weights sum to zero");
    ITJL_CUDA(ctx, cudaSetDevice(ctx->device));
    ITJL_CUDA(ctx, ctx->raw.reserve((size_t)n_prev * p * sizeof(double)));
    ITJL_CUDA(ctx, ctx->scratch.reserve((size_t)n_prev * sizeof(double)));
    ITJL_CUDA(ctx, ctx->theta.reserve((size_t)n * p * sizeof(double)));
    ITJL_CUDA(ctx, cudaMemcpyAsync(ctx->raw.p, theta_prev, (size_t)n_prev * p * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    ITJL_CUDA(ctx, cudaMemcpyAsync(ctx->scratch.p, cdf.data(), (size_t)n_prev * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    cudaError_t e = pmc_propose_launch((const double*)ctx->raw.p, (const double*)ctx->scratch.p, n_prev, p, L, n, seed, generation, first,
                                       (double*)ctx->theta.p, ctx->stream);
    if (e != cudaSuccess) return fail(ctx, ITJL_ERR_CUDA, "launch k_pmc_propose: %s", cudaGetErrorString(e));
    ctx->launches++;
    ITJL_CUDA(ctx, cudaMemcpyAsync(theta_out, ctx->theta.p, (size_t)n * p * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    ITJL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));              // cdf stays alive until here
    return ITJL_OK;
}


int itjl_order_stats(itjl_ctx* ctx, const double* values, int64_t n, double upper, const double* probs, int32_t k,
                     double* out, int64_t* n_valid_out) {
    if (!ctx) return fail(nullptr, ITJL_ERR_ARG, "itjl_order_stats: null context");
    if (!values || !n_valid_out || n < 1 || n > 0x7fffffffll || k < 0 || k > 32 || (k > 0 && (!probs || !out)))
        return fail(ctx, ITJL_ERR_ARG, "itjl_order_stats: invalid argument (1 <= n < 2^31, 0 <= k <= 32)");
    for (int i = 0; i < k; ++i)
        if (!(probs[i] >= 0.0 && probs[i] <= 1.0)) return fail(ctx, ITJL_ERR_ARG, "itjl_order_stats: probabilities must lie in [0, 1]");
    ITJL_LOCK(ctx);
    ctx = primary(ctx);
    ITJL_LOCK(ctx);
    NvtxRange nvtx("itjl_order_stats");
    ITJL_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t bytes = (size_t)n * sizeof(double), tb = sort_temp_bytes(n);
    ITJL_CUDA(ctx, ctx->raw.reserve(bytes));
    ITJL_CUDA(ctx, ctx->out.reserve(bytes));
    ITJL_CUDA(ctx, ctx->out2.reserve(bytes));
    ITJL_CUDA(ctx, ctx->part.reserve(tb));
    ITJL_CUDA(ctx, ctx->scratch.reserve(64 * sizeof(int64_t) + 64 * sizeof(double) + 16));
    unsigned long long* nv = (unsigned long long*)ctx->scratch.p;
    int64_t* idx_dev = (int64_t*)((char*)ctx->scratch.p + 16);
    double* got_dev = (double*)((char*)ctx->scratch.p + 16 + 64 * sizeof(int64_t));
    ITJL_CUDA(ctx, cudaMemcpyAsync(ctx->raw.p, values, bytes, cudaMemcpyHostToDevice, ctx->stream));
    ITJL_CUDA(ctx, cudaMemsetAsync(nv, 0, sizeof(unsigned long long), ctx->stream));
    cudaError_t e = mask_invalid_launch((const double*)ctx->raw.p, n, upper, (double*)ctx->out.p, nv, ctx->stream);
    if (e == cudaSuccess) e = sort_doubles_launch(ctx->part.p, tb, (const double*)ctx->out.p, (double*)ctx->out2.p, n, ctx->stream);
    if (e != cudaSuccess) return fail(ctx, ITJL_ERR_CUDA, "itjl_order_stats: %s", cudaGetErrorString(e));
    ctx->launches += 2;
    unsigned long long n_valid = 0;
    ITJL_CUDA(ctx, cudaMemcpyAsync(&n_valid, nv, sizeof(n_valid), cudaMemcpyDeviceToHost, ctx->stream));
    ITJL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *n_valid_out = (int64_t)n_valid;
    if (k > 0 && n_valid > 0) {
        // quantile type 7: position p (n_valid - 1) between the order statistics floor and floor + 1
        int64_t idx[64];
        for (int i = 0; i < k; ++i) {
            const double pos = probs[i] * (double)(n_valid - 1);
            const int64_t lo = (int64_t)floor(pos);
            idx[2 * i] = lo;
            idx[2 * i + 1] = lo + 1 < (int64_t)n_valid ? lo + 1 : lo;
        }
        ITJL_CUDA(ctx, cudaMemcpyAsync(idx_dev, idx, (size_t)2 * k * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream));
        e = gather_launch((const double*)ctx->out2.p, idx_dev, 2 * k, got_dev, ctx->stream);
        if (e != cudaSuccess) return fail(ctx, ITJL_ERR_CUDA, "launch k_gather: %s", cudaGetErrorString(e));
        ctx->launches++;
        ITJL_CUDA(ctx, cudaMemcpyAsync(out, got_dev, (size_t)2 * k * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        ITJL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return ITJL_OK;
}

}  // extern "C"
