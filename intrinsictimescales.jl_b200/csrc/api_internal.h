// Declarations shared by the translation units of the C-ABI layer (itjl_api*.cu).
#pragma once
#include <vector>

#include "kernels.h"

struct itjl_model {
    itjl_ctx* ctx = nullptr;
    itjl_model_desc d{};
    itjl::DevBuf target, mask_bits, n_valid, bins, binidx, window, twiddle;
    int mask_words = 0;
    double mask_frac = 0.0;        // fraction of masked samples (missing-data models)
    itjl::AbcAcfGeometry geo{};
    itjl::AbcTcGeometry tc{};
    bool use_tc = false;           // lag sums on the tensor cores (abc_tc_kernels.cu)
    double tc_pred_err = 0.0;      // error model's bound for the chosen plan (0 for the CUDA-core kernel)
    // PSD
    int n2 = 0, log2_n2 = 0, psd_chunk = 0, psd_xs_len = 0;
    size_t psd_smem = 0;
    bool psd_accb_global = false;  // PSD bin sums in global memory (too many selected bins for shared memory)
    double psd_scale = 0.0;
    bool psd_v2 = false;           // warp-specialised kernel k_abc_psd2 (4096 < n_t <= 16384 and everything fits)
    int psd_round0_bins = 0;
    // DSP-convention periodogram (ITJL_SUMMARY_PGRAM)
    itjl::FftPlan pg_plan{};
    int pg_chunk = 0, pg_bufb = 0;
    size_t pg_smem = 0;
    // multi-GPU parent model: one model per device of the parent context (itjl_api_multi.cu)
    std::vector<itjl_model*> subs;
};

namespace itjl {

// the context that owns device resources for calls that do not shard: the first device of a multi-GPU context
inline itjl_ctx* primary(itjl_ctx* ctx) { return (ctx && !ctx->subs.empty()) ? ctx->subs[0] : ctx; }
inline itjl_model* primary(itjl_model* m) { return (m && !m->subs.empty()) ? m->subs[0] : m; }
inline bool is_sharded(const itjl_ctx* ctx) { return ctx && (ctx->world > 1 || !ctx->subs.empty()); }

int upload(itjl_ctx* ctx, DevBuf& buf, const void* host, size_t bytes);

// Launch the fused kernel for one batch; every pointer is a device pointer (or null).  theta_dev is
// (n_particles x n_theta) column-major with leading dimension n_particles.
int launch_distances(itjl_model* m, const double* theta_dev, int64_t n_particles, int32_t n_theta,
                     uint64_t seed, uint64_t generation, int64_t particle_offset,
                     const double* u0_dev, const double* noise_dev, const double* phases_dev,
                     double* dist_dev, double* stats_dev, int64_t theta_ld = 0);

// contiguous shard of `rank` when n items are split over `world` ranks (same rule as distributed.py::shard_bounds)
inline void shard_bounds(int64_t n, int world, int rank, int64_t* lo, int64_t* hi) {
    const int64_t base = n / world, rem = n % world;
    *lo = rank * base + (rank < rem ? rank : rem);
    *hi = *lo + base + (rank < rem ? 1 : 0);
}

// itjl_api_multi.cu: score n_particles (host theta, column-major, ld = n_particles) sharded over the context's
// communicator; on return dist_dev_out (on primary(ctx)'s device, n_particles doubles, its ctx->dist buffer)
// holds every particle's distance, ordered by the stream of primary(ctx).
int sharded_distances(itjl_model* m, const double* theta, int64_t n_particles, int32_t n_theta, uint64_t seed,
                      uint64_t generation, int64_t particle_offset, double** dist_dev_out);
// all-gather of per-row doubles computed by `fill(member ctx, rank, lo, hi, dst_dev)` into primary(ctx)->dist
struct ShardMember { itjl_ctx* c; int rank; };
void shard_members(itjl_ctx* ctx, std::vector<ShardMember>* out, int* world);
int shard_allgather_compact(itjl_ctx* ctx, int64_t n, double** out_dev);
int shard_allgather_inplace(itjl_ctx* ctx, const std::vector<double*>& bases, int64_t width);
int multi_model_create(itjl_ctx* ctx, const itjl_model_desc* desc, itjl_model** model_out);
void comm_destroy(itjl_ctx* ctx);

}  // namespace itjl
