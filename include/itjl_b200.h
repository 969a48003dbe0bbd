/*
 * itjl_b200.h - C ABI of libitjl_b200.so, the B200 (sm_100a) implementation of the
 * data-parallel hot path of IntrinsicTimescales.jl v0.7.0.
 *
 * The reference is pure Julia and has no FFI layer; its seam is Julia multiple dispatch
 * on the model protocol.  Each entry point below names the reference function(s) whose
 * body it replaces (paths relative to the reference checkout).  The Julia-side `ccall`
 * shim a maintainer would add is shown in INTEGRATION.md.
 *
 * Conventions
 *   - plain C, no exceptions: every call returns 0 on success or a negative itjl_status;
 *     itjl_last_error(ctx) gives the message (ctx may be NULL for creation errors).
 *   - all `const double*` / `double*` / `uint8_t*` / `int32_t*` parameters are HOST pointers
 *     owned by the caller (Julia `Array`s, column-major) unless the name ends in `_dev`.
 *   - matrices follow Julia layout: `theta` is (n_particles x n_theta) column-major;
 *     `data` for the summary/acw calls is (n_slices x n_t) column-major, i.e. element
 *     (slice s, time t) lives at data[s + t * n_slices] (what `acw(data, fs)` holds for
 *     dims = 2).
 *   - fp64 at the boundary; fp32 arithmetic inside the fused simulate+summary kernels,
 *     fp64 inside the acw kernels (bit-exact crossing indices against an fp64 reference).
 *   - one context = one GPU + one stream (or, from itjl_ctx_create_multi, one per GPU of the box); every call
 *     that takes a context or one of its models holds the context's lock for its duration, so a context can
 *     be shared between host threads: concurrent calls are serialised, not undefined.
 *   - the library never reads the environment: everything that changes which kernel runs or how it rounds is
 *     in itjl_tuning, per context.
 *   - a non-finite simulated trajectory yields distance = NaN (the observable contract of
 *     src/utils/ou_process.jl:56-61 + src/core/abc.jl:185).
 */
#ifndef ITJL_B200_H
#define ITJL_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct itjl_ctx itjl_ctx;
typedef struct itjl_model itjl_model;

typedef enum {
    ITJL_OK = 0,
    ITJL_ERR_ARG = -1,      /* invalid argument / unsupported configuration */
    ITJL_ERR_CUDA = -2,     /* CUDA runtime error (message has the CUDA string) */
    ITJL_ERR_NO_DEVICE = -3,
    ITJL_ERR_ALLOC = -4
} itjl_status;

enum { ITJL_KIND_OU = 0, ITJL_KIND_OU_OSC = 1 };
enum { ITJL_SUMMARY_ACF = 0, ITJL_SUMMARY_ACF_MISSING = 1,
       ITJL_SUMMARY_PSD = 2,      /* comp_psd_adfriendly: demeaned, n2 = 2 nextpow2(n) (the oscillation models) */
       ITJL_SUMMARY_PGRAM = 3 };  /* comp_psd = DSP.periodogram: nfft = nextfastfft(n), one-sided (one_timescale_model, :psd) */
enum { ITJL_DIST_LINEAR = 0, ITJL_DIST_LOG = 1 };
enum { ITJL_UPDATE_EXACT = 0, ITJL_UPDATE_EM = 1 };

/* acw type mask bits (src/core/acw.jl:132, acf_acwtypes) */
enum { ITJL_ACW0 = 1, ITJL_ACW50 = 2, ITJL_ACWEULER = 4, ITJL_AUC = 8, ITJL_TAU = 16,
       ITJL_TAU3 = 32 };   /* tau by fit_expdecay_3_parameters (skip_zero_lag = true, acw.jl:212-216) instead of fit_expdecay */

/* ---- tuning ------------------------------------------------------------------------ */
/* Per-context knobs (itjl_tuning_default gives the values below).  They replace the ITJL_* environment
 * switches of the first version: a model reads them once, in itjl_model_create. */
typedef struct {
    int32_t abc_tc;          /* fused ACF kernel: -1 automatic (tensor-core kernel k_abc_acf_tc when the shape has a plan
                                AND its predicted error is within tc_error_bound, else the CUDA-core kernel k_abc_acf);
                                0 always k_abc_acf (fp32 FFMA lag sums: the exact-parity switch); 1 always k_abc_acf_tc
                                (itjl_model_create fails when the shape has no plan) */
    int32_t tc_terms;        /* 0 automatic (from the error model); 1 / 2 / 3 product terms of the fp16 hi/lo split */
    int32_t tc_tail_max;     /* -1 default (15): lags beyond a tensor-memory plan summed by the FFMA tail (<= 60) */
    int32_t tc_groups;       /* 0 automatic; 1..4 simulate groups */
    int32_t tc_seg_trials;   /* 0 automatic; trials per accumulator segment (drain) */
    int32_t tc_bias;         /* 1 (default): undo the accumulator's round-toward-zero to first order; 0 off */
    int32_t abc_threads;     /* 0 automatic; 128 / 256 threads per CTA of k_abc_acf */
    int32_t abc_bufs;        /* 0 automatic; 1 / 2 trial buffers of k_abc_acf */
    int32_t acw_stream;      /* 1 (default): acw() starts with the one-read streaming pass over 32 lags; 0: shared-memory
                                kernel only; 8 / 16 / 32: lags that pass keeps (slices that do not cross zero inside the
                                window go to the fp64 kernel: a narrower window pays only on short-memory data) */
    int32_t acw_waves;       /* 0 automatic; time segments of the streaming pass sized for this many waves of CTAs */
    int32_t pinned_upload;   /* 1 (default): large host <-> device copies go through the pinned staging ring */
    int32_t copy_threads;    /* 0 automatic (half the hardware threads, at most 8); host threads feeding that ring */
    int32_t pmc_fp32;        /* PMC weights: -1 automatic (fp64 pair sums up to 1e8 pairs, fp32 pair sums above; one-parameter
                                models above 1e8 pairs with >= 32768 new particles: exact fp64 mixture density on an
                                8192-point grid + 4-point interpolation, ~1e-12); 0 always fp64 pair sums; 1 always fp32
                                pair sums; 2 always the grid (n_theta = 1) */
    int32_t psd_v2;          /* bit 0 (default 1): PSD models with 4096 < n_t <= 16384 run the warp-specialised kernel
                                k_abc_psd2 (0: always k_abc_psd); bit 1 (default 0): Lomb-Scargle by the direct
                                O(n_t n_freq) sums instead of the chirp-z transforms */
    double tc_error_bound;   /* 5e-6: bound on the predicted |error| of a trial-mean ACF value of k_abc_acf_tc */
} itjl_tuning;
int itjl_tuning_default(itjl_tuning* out);
int itjl_ctx_set_tuning(itjl_ctx* ctx, const itjl_tuning* t);
int itjl_ctx_get_tuning(const itjl_ctx* ctx, itjl_tuning* out);

/* ---- library / context ------------------------------------------------------------ */
const char* itjl_version(void);
int itjl_device_count(int* n_out);
int itjl_ctx_create(int device, itjl_ctx** ctx_out);
/* One process driving n_dev GPUs (what a Julia host gets): one sub-context and stream per device and an NCCL
 * communicator over them (ncclCommInitAll; libnccl.so.2 is loaded on first use, the library itself does not
 * link it).  itjl_model_create / itjl_abc_distances / itjl_abc_generation / itjl_pmc_weights on such a
 * context shard the particles over the devices as contiguous ranges of the global particle index, all-gather
 * the per-particle results device to device and return them exactly as the single-GPU calls do (the Philox
 * streams are keyed on the global index: results do not depend on n_dev).  Everything else runs on the
 * first device.  dev_ids = NULL means devices 0 .. n_dev-1. */
int itjl_ctx_create_multi(int n_dev, const int* dev_ids, itjl_ctx** ctx_out);
/* One process per GPU (torchrun / MPI): rank 0 calls itjl_comm_unique_id, ships the 128 bytes to every rank by
 * its own means, then every rank attaches its single-device context to the communicator (ncclCommInitRank,
 * collective).  The sharded calls listed above are then COLLECTIVE: every rank passes the same arguments
 * and receives the full result. */
int itjl_comm_unique_id(void* id_out_128_bytes);
int itjl_ctx_attach_comm(itjl_ctx* ctx, int rank, int world, const void* id_128_bytes);
/* rank / world size of the context's communicator (0 / 1 without one) and the number of devices this process drives */
int itjl_ctx_world(const itjl_ctx* ctx, int* rank, int* world, int* n_local);
int itjl_ctx_destroy(itjl_ctx* ctx);
const char* itjl_last_error(const itjl_ctx* ctx);
int itjl_ctx_synchronize(itjl_ctx* ctx);
/* cudaStream_t of the context, as an integer handle (for event timing by the caller) */
uint64_t itjl_ctx_stream(const itjl_ctx* ctx);
/* number of kernels this context has launched so far */
int64_t itjl_ctx_launch_count(const itjl_ctx* ctx);
int itjl_ctx_sm_count(const itjl_ctx* ctx);

/* ---- model: the constants of one AbstractTimescaleModel, uploaded once ------------- */
/* Fields mirror OneTimescaleModel / OneTimescaleWithMissingModel / OneTimescaleAndOscModel
 * (src/models/one_timescale.jl:61-83, one_timescale_with_missing.jl:61-83,
 *  one_timescale_and_osc.jl:74-96). */
typedef struct {
    int32_t kind;          /* ITJL_KIND_*    : generate_data method */
    int32_t summary;       /* ITJL_SUMMARY_* : summary_stats method */
    int32_t distance;      /* ITJL_DIST_*    : distance_function method */
    int32_t update;        /* ITJL_UPDATE_*  : OU transition (exact = what SOSRA converges to) */
    int64_t n_trials;      /* model.numTrials */
    int64_t n_t;           /* length(model.dt:model.dt:model.T) */
    double dt;             /* model.dt */
    double data_mean;      /* model.data_mean (osc model only) */
    double data_sd;        /* model.data_sd */
    int64_t n_stats;       /* length(model.data_sum_stats): n_lags (ACF) or count(freq_idx) (PSD) */
    const double* target_stats;   /* model.data_sum_stats, n_stats */
    const uint8_t* missing_mask;  /* Matrix{UInt8}(model.missing_mask), (n_trials x n_t) col-major, or NULL */
    const int32_t* freq_bins;     /* PSD / PGRAM: findall(model.freq_idx) .- 1 (0-based into the power vector the summary returns,
                                     whose first entry is the bin after DC), n_stats; else NULL */
} itjl_model_desc;

int itjl_model_create(itjl_ctx* ctx, const itjl_model_desc* desc, itjl_model** model_out);
int itjl_model_destroy(itjl_model* model);

/* ---- ABC inner loop --------------------------------------------------------------- */
/* Replaces the body of the `trial_count` loop of basic_abc for a batch of particles:
 *   d[i] = Models.generate_data_and_reduce(model, theta[i, :])
 * (src/core/abc.jl:181 -> src/core/model.jl:146-151 -> generate_data / summary_stats /
 * distance_function of the model).  Particle i uses the Philox stream keyed by
 * (seed, generation, particle_offset + i), so results do not depend on batching or on how
 * particles are sharded over GPUs.  u0 / noise / phases (each may be NULL) inject the
 * per-trial initial state (n_particles x n_trials, row-major), the N(0,1) increments
 * (n_particles x n_trials x n_t, row-major) and the oscillation phases (as u0) for parity
 * tests.  stats_out (may be NULL) receives the per-particle summary statistics
 * (n_particles x n_stats, row-major).  */
int itjl_abc_distances(itjl_model* model, const double* theta, int64_t n_particles, int32_t n_theta,
                       uint64_t seed, uint64_t generation, int64_t particle_offset,
                       const double* u0, const double* noise, const double* phases,
                       double* dist_out, double* stats_out);

/* d[i] = distance_function(model, summary_stats(model, generate_data(model, theta[i, :])), data_sum_stats) with
 * distance_combined = true (one_timescale.jl:306-320 + combined_distance :278-290; one_timescale_and_osc.jl:368-385 +
 * :324-347), the whole batch on the device: fused kernel -> statistics (kept on the device) -> batched fit (ACF:
 * fit_expdecay over lags = (0 : n_stats - 1) dt; PSD: find_knee_frequency of every particle's spectrum over lags_freqs,
 * and for the oscillation model find_oscillation_peak of its residual) -> weights[0] d + weights[1] (data_tau - tau)^2
 * [+ weights[2] (data_osc - peak)^2].  n_weights = 2, or 3 for the PSD summary of the oscillation model.  A particle
 * whose statistics are not finite gets NaN. */
int itjl_abc_distances_combined(itjl_model* model, const double* theta, int64_t n_particles, int32_t n_theta,
                                uint64_t seed, uint64_t generation, int64_t particle_offset, const double* lags_freqs,
                                const double* weights, int32_t n_weights, double data_tau, double data_osc, double* dist_out);

/* Same, fused with the bookkeeping of src/core/abc.jl:182-199: applies `d <= epsilon`
 * (NaN -> rejected) and the sequential stopping rule (stop right after the
 * min_accepted-th acceptance, in particle order).  isaccepted (n_particles bytes) is only
 * meaningful for the first *n_total entries. */
int itjl_abc_generation(itjl_model* model, const double* theta, int64_t n_particles, int32_t n_theta,
                        uint64_t seed, uint64_t generation, int64_t particle_offset,
                        double epsilon, int64_t min_accepted,
                        double* dist_out, uint8_t* isaccepted,
                        int64_t* n_total, int64_t* n_accepted);

/* Accept / stop bookkeeping alone on host distances (src/core/abc.jl:185-199). */
int itjl_abc_accept(itjl_ctx* ctx, const double* dist, int64_t n, double epsilon, int64_t min_accepted,
                    uint8_t* isaccepted, int64_t* n_total, int64_t* n_accepted);

/* Device-resident variant used for kernel-only timing and multi-GPU sharding: theta_dev
 * (n_particles x n_theta col-major, double) and dist_dev (double) are DEVICE pointers; the
 * call is asynchronous on the context stream. */
int itjl_abc_distances_dev(itjl_model* model, const double* theta_dev, int64_t n_particles,
                           int32_t n_theta, uint64_t seed, uint64_t generation,
                           int64_t particle_offset, double* dist_dev);

/* ---- between generations: PMC weights and covariance ---------------------------------- */
/* calc_weights(theta_prev, theta, tau_squared, weights, prior)  (src/core/abc.jl:520-567, the multi-parameter
 * branch :547-566, which is also the intended formula of the 1-parameter branch):
 *     w_i = prior_pdf[i] / sum_j weights_prev[j] N(theta[i, :]; theta_prev[j, :], tau_squared),
 * normalised to sum 1 when `normalize` != 0.  theta_prev (n_prev x n_theta) and theta (n x n_theta) are
 * column-major, tau_squared (n_theta x n_theta, symmetric positive definite), n_theta <= 4.  prior_pdf[i] =
 * prod_k pdf(prior[k], theta[i, k]) is evaluated by the caller (Distributions.pdf on the Julia side): the
 * library knows no distribution families.  n * n_prev Gaussian evaluations on the device (k_pmc_mixture);
 * on a multi-GPU context the rows i are sharded and the result all-gathered. */
int itjl_pmc_weights(itjl_ctx* ctx, const double* theta_prev, int64_t n_prev, const double* theta, int64_t n,
                     int32_t n_theta, const double* tau_squared, const double* weights_prev,
                     const double* prior_pdf, int32_t normalize, double* weights_out);
/* n i.i.d. draws of draw_theta_pmc(model, theta_prev, weights, tau_squared; jitter)  (src/core/abc.jl:101-117):
 * theta* resampled from theta_prev (n_prev x n_theta, column-major) with probabilities weights_prev, then
 * MvNormal(theta*, tau_squared + jitter I), redrawn around the same theta* while any component is negative.
 * Philox stream (seed, generation, first + i): proposal i does not depend on n or on the number of GPUs (every
 * rank of a communicator draws the same matrix).  theta_out: (n x n_theta) column-major.  StatsBase's and
 * Distributions' own streams are not reproduced - the draws are i.i.d. from the same distribution. */
int itjl_pmc_propose(itjl_ctx* ctx, const double* theta_prev, int64_t n_prev, int32_t n_theta, const double* weights_prev,
                     const double* tau_squared, double jitter, int64_t n, uint64_t seed, uint64_t generation,
                     int64_t first, double* theta_out);
/* Order statistics for select_epsilon (src/core/abc.jl:700-760: percentiles of the valid distances).  The values
 * that are not NaN and < upper are sorted ascending on the device; *n_valid_out = their count m, and for every
 * probability probs[i] in [0, 1] out[2i], out[2i+1] = the order statistics of 0-based ranks floor(probs[i] (m - 1)) and
 * that + 1 (clamped): the two numbers quantile type 7 (StatsBase.percentile) interpolates between - the caller does
 * the interpolation with its own rounding.  k <= 32; nothing is written to `out` when m = 0. */
int itjl_order_stats(itjl_ctx* ctx, const double* values, int64_t n, double upper, const double* probs, int32_t k,
                     double* out, int64_t* n_valid_out);
/* weighted_covar(x, w)  (src/core/abc.jl:582-613): x (n x n_theta) column-major, cov_out (n_theta x n_theta). */
int itjl_weighted_covar(itjl_ctx* ctx, const double* x, int64_t n, int32_t n_theta, const double* w, double* cov_out);

/* The same on a communicator of single-device contexts (itjl_ctx_attach_comm; COLLECTIVE): theta_dev holds ALL
 * n_particles on this rank's device, the rank scores its contiguous share, the shares are all-gathered over
 * NVLink and dist_dev (device, n_particles) receives every distance.  Asynchronous on the context stream.
 * Without a communicator it is itjl_abc_distances_dev. */
int itjl_abc_distances_sharded_dev(itjl_model* model, const double* theta_dev, int64_t n_particles,
                                   int32_t n_theta, uint64_t seed, uint64_t generation,
                                   int64_t particle_offset, double* dist_dev);

/* Algorithmic work of one itjl_abc_distances call, for roofline accounting:
 * flops_per_sample = 2*n_lags + 8 for the direct lag-sum kernels. */
int itjl_model_work(const itjl_model* model, double* flops_per_sample, int64_t* samples_per_particle);
/* Which fused kernel itjl_abc_distances launches for this model ("k_abc_acf_tc", "k_abc_acf", "k_abc_psd";
 * `name_out` holds >= 32 bytes) and the tensor-core flops it EXECUTES per OU sample (fp16 hi/lo split,
 * tile and K padding included; 0 for the CUDA-core kernels) - the algorithmic figure is itjl_model_work's. */
int itjl_model_kernel_info(const itjl_model* model, char* name_out, double* tensor_flops_per_sample);
/* Host-only (no GPU needed): the tensor-memory plan k_abc_acf_tc would use for a model shape on a device
 * with `max_smem_optin` bytes of opt-in shared memory per CTA and `sm_count` SMs (B200: 232448, 148).
 * Returns ITJL_OK and fills out[0..15] = {groups, chunk, ksteps, rows_total, used_cols, mixed (plan B),
 * seg_trials, n_terms, tc_lags, tail_start, tail_lt, n_ops, smem_bytes, grid, n_windows (kernel launches
 * per batch: lag windows of 385 lags beyond the first 449), 0}, or ITJL_ERR_ARG when
 * the shape is outside the kernel's plans (the CUDA-core kernel k_abc_acf then runs). */
int itjl_tc_plan(int64_t n_t, int64_t n_lags, int64_t n_trials, int32_t max_smem_optin, int32_t sm_count, int32_t* out);
/* The error model behind abc_tc = -1 (host only): predicted bound on |error| of a trial-mean ACF value of
 * k_abc_acf_tc with `n_terms` product terms for a model of n_trials x n_t samples, `seg_rows` 128-sample rows per
 * accumulator segment and a fraction `mask_frac` of masked samples (missing-data models; `mean_ratio` =
 * data_mean / data_sd of the oscillation + missing-data model, whose masked samples sit at that offset, else 0).  DESIGN.md section 4 derives it;
 * tests/test_gpu_tc.py holds the kernel to it. */
double itjl_tc_predicted_error(int64_t n_t, int64_t n_trials, int32_t n_terms, int32_t seg_rows, double mask_frac, double mean_ratio);

/* ---- OU generator ----------------------------------------------------------------- */
/* generate_ou_process(tau, true_D, dt, duration, num_trials; standardize)
 * (src/utils/ou_process.jl:46-62, 114-148).  out: (n_trials x n_t) column-major.
 * u0 (n_trials) / noise (n_trials x n_t row-major) may be NULL (Philox stream of
 * (seed, generation = 0, particle = 0)). */
int itjl_generate_ou(itjl_ctx* ctx, double tau, double true_D, double dt, int64_t n_t, int64_t n_trials,
                     int32_t standardize, int32_t update, uint64_t seed,
                     const double* u0, const double* noise, double* out);

/* generate_ou_with_oscillation(theta = [tau, freq, coeff], dt, duration, num_trials,
 * data_mean, data_sd)  (src/utils/ou_process.jl:176-218). */
int itjl_generate_ou_osc(itjl_ctx* ctx, double tau, double freq, double coeff, double dt, int64_t n_t,
                         int64_t n_trials, double data_mean, double data_sd, int32_t update,
                         uint64_t seed, const double* u0, const double* noise, const double* phases,
                         double* out);

/* Models.generate_data(::TwoTimescaleModel, [tau1, tau2, coeff])  (src/models/two_timescale.jl:30-51, the
 * generator of the reference's unfinished two-timescale model): sqrt(data_var) * (sqrt(coeff) * ou1 + sqrt(1 - coeff)
 * * ou2) with ou_i = generate_ou_process(tau_i, 1 / tau_i, dt, duration, num_trials).  out: (n_trials x n_t)
 * column-major. */
int itjl_generate_ou_two(itjl_ctx* ctx, double tau1, double tau2, double coeff, double dt, int64_t n_t, int64_t n_trials,
                         double data_var, int32_t update, uint64_t seed, double* out);

/* ---- summary statistics on stored data -------------------------------------------- */
/* comp_ac_fft(data; dims = 2, n_lags)  (src/stats/summary.jl:46-63, 79-89) computed as the
 * mathematically identical time-domain lag sum.  out: (n_slices x n_lags) column-major. */
int itjl_acf(itjl_ctx* ctx, const double* data, int64_t n_slices, int64_t n_t, int64_t n_lags, double* out);
/* Strided variants: element (slice s, time t) lives at data[s * slice_stride + t * time_stride]
 * (strides in elements, > 0).  slice_stride = 1, time_stride = n_slices is the layout of the calls above
 * (`dims = 2` of a Julia matrix); slice_stride = n_t, time_stride = 1 is `dims = 1` (time down the columns):
 * the matrix is uploaded as it lies and transposed on the device, no host copy.  `missing` != 0 selects
 * comp_ac_time_missing. */
int itjl_acf_strided(itjl_ctx* ctx, const double* data, int64_t n_slices, int64_t n_t, int64_t slice_stride,
                     int64_t time_stride, int64_t n_lags, int32_t missing, double* out);
/* comp_ac_time_missing(data; dims = 2, n_lags)  (src/stats/summary.jl:455-496); NaN = missing. */
int itjl_acf_missing(itjl_ctx* ctx, const double* data, int64_t n_slices, int64_t n_t, int64_t n_lags,
                     double* out);
/* comp_psd_adfriendly(data, fs; dims = 2)[1]  (src/stats/summary.jl:183-235).
 * out: (n_slices x (n2/2 - 1)) column-major with n2 = 2 * nextpow2(n_t). */
int itjl_psd_hamming(itjl_ctx* ctx, const double* data, int64_t n_slices, int64_t n_t, double fs, double* out);
/* comp_psd(data, fs; dims = 2, method = "periodogram")[1]  (src/stats/summary.jl:111-155): DSP.periodogram with the
 * Hamming window, nfft = itjl_nextfastfft(n_t), one-sided, DC dropped.  fp64.
 * out: (n_slices x nfft/2) column-major; freqs[k] = (k + 1) fs / nfft. */
int itjl_psd_periodogram(itjl_ctx* ctx, const double* data, int64_t n_slices, int64_t n_t, double fs, double* out);
/* DSP.nextfastfft(n) = nextprod([2, 3, 5, 7], n)  (host only) */
int64_t itjl_nextfastfft(int64_t n);
/* comp_psd_lombscargle(times, data, nanmask, dt; dims = 2)[1]  (src/stats/summary.jl:321-356) for times =
 * dt:dt:n_t*dt: the generalised (floating-mean, standard-normalised) Lomb-Scargle periodogram of the valid
 * samples of every slice (NaN = missing) on the grid f0 + k df, k < n_freq (the caller passes LombScargle.jl's
 * autofrequency grid: df = 1 / (5 (n_t - 1) dt), f0 = df / 2, up to 1 / (2 dt)).  Evaluated exactly;
 * LombScargle.jl's default for more than 200 samples is the Press-Rybicki approximation of the same quantity.
 * out: (n_slices x n_freq) column-major. */
int itjl_psd_lombscargle(itjl_ctx* ctx, const double* data, int64_t n_slices, int64_t n_t, double dt,
                         double f0, double df, int64_t n_freq, double* out);

/* ---- PSD-side reducers: Lorentzian knee, oscillation peak, FOOOF-style loop ------------------------ */
/* find_knee_frequency / fooof_fit / find_oscillation_peak / fit_gaussian of src/utils/utils.jl:479-538, 598-675,
 * 721-765, 813-838 for every row of psd (n_rows x n_freq; col_major = 0: rows contiguous, as itjl_abc_distances'
 * stats_out; 1: Julia (n_rows x n_freq) matrix), freqs ascending.  The reference's residuals make its
 * Levenberg-Marquardt solve a least-ABSOLUTE-deviation fit; the library returns the argmin of mean |model - psd|
 * reached from the reference's initial guess (deterministic Nelder-Mead, fp64), knee confined to (0, max(freqs)].
 *   fooof = 0: [amp, knee] = find_knee_frequency(psd, freqs; min_freq, max_freq); peak_out (may be NULL) =
 *              find_oscillation_peak(psd - lorentzian, freqs; min_freq, max_freq) (NaN: no peak) - the two numbers
 *              distance_function of the PSD models needs (one_timescale.jl:311-316, one_timescale_and_osc.jl:377-384);
 *   fooof = 1: fooof_fit(psd, freqs; min_freq, max_freq, oscillation_peak = true, max_peaks, return_only_knee = true):
 *              frequencies outside [min_freq, max_freq] dropped, up to max_peaks Gaussians removed, knee refitted
 *              (when no peak is found the first knee is returned; the reference throws UndefVarError there);
 *   fooof = 2: as 0 with lorentzian_initial_guess (utils.jl:423-445) in place of the fit - what the informed prior of
 *              the oscillation models uses (one_timescale_and_osc.jl:16-28);
 *   fooof = 3: peak_out = find_oscillation_peak(psd, freqs; min_freq, max_freq) of the rows themselves (knee = NaN).
 * Only the default constrained = false, allow_variable_exponent = false branch.  amp_out / peak_out may be NULL. */
int itjl_fit_knee(itjl_ctx* ctx, const double* psd, int64_t n_rows, int64_t n_freq, int32_t col_major,
                  const double* freqs, double min_freq, double max_freq, int32_t fooof, int32_t max_peaks,
                  double* knee_out, double* amp_out, double* peak_out);

/* The :knee branch of acw(data, fs; acwtypes = :knee, freqlims, average_over_trials, oscillation_peak, max_peaks)
 * (src/core/acw.jl:242-270) in one call: spectrum of every slice on the device (comp_psd periodogram; with NaN in
 * the data comp_psd_lombscargle on LombScargle.autofrequency(dt:dt:n_t*dt; maximum_frequency = fs/2)), optional
 * mean over the slices, then itjl_fit_knee's fooof loop (flags bit 0 = oscillation_peak) or the plain knee.
 * freq_lo / freq_hi: freqlims, NaN = (freqs[1], freqs[end]).  knee_out: n_out knee FREQUENCIES (tau =
 * 1 / (2 pi knee)), n_out = 1 when flags bit 1 (average_over_trials) else n_slices.  psd_out: NULL or
 * (n_out x *n_freq_out) column-major, capacity psd_capacity doubles; freqs_out: NULL or n_freq doubles
 * (itjl_acw_knee_nfreq gives the size beforehand). */
int itjl_acw_knee(itjl_ctx* ctx, const double* data, int64_t n_slices, int64_t n_t, double fs, double freq_lo, double freq_hi,
                  uint32_t flags, int32_t max_peaks, double* knee_out, double* psd_out, int64_t psd_capacity,
                  double* freqs_out, int64_t* n_freq_out);
int64_t itjl_acw_knee_nfreq(int64_t n_t, int32_t has_nan);

/* ---- acw() ------------------------------------------------------------------------ */
/* The ACF-based part of acw(data, fs; acwtypes, n_lags, average_over_trials)
 * (src/core/acw.jl:141-231) with the reducers of src/utils/utils.jl:113-138, 218-345.
 *   k0 / k50 / ke : 0-based first lag index with acf <= 0 / 0.5 / 1/e, -1 if never
 *                   (the caller indexes its own `lags` range; NaN for -1);
 *   auc           : acw_romberg(dt, acf[1:k0_first]) per slice (NaN when k0_first < 2);
 *   tau           : fit_expdecay(lags[1:n_lags], acf[.., 1:n_lags]) per slice (ITJL_TAU), or the tau of
 *                   fit_expdecay_3_parameters - A (exp(-t / tau) + B) over lags[2:end], src/utils/utils.jl:159-179 -
 *                   with ITJL_TAU3 (argmin of its mean-squared-error objective; LM's stopping point is not reproduced);
 *   n_lags_io     : in: user n_lags or 0 for the reference default ceil(1.1*max k0);
 *                   out: the value used;
 *   acf_out       : NULL, or capacity acf_capacity doubles; receives (n_out x n_lags)
 *                   column-major, n_out = 1 when average_over_trials else n_slices.
 * Any of k0/k50/ke/auc/tau may be NULL when not requested in typemask. */
int itjl_acw(itjl_ctx* ctx, const double* data, int64_t n_slices, int64_t n_t, double fs,
             uint32_t typemask, int32_t average_over_trials, int64_t* n_lags_io,
             int32_t* k0, int32_t* k50, int32_t* ke, double* auc, double* tau,
             double* acf_out, int64_t acf_capacity);

/* itjl_acw on a strided matrix (see itjl_acf_strided). */
int itjl_acw_strided(itjl_ctx* ctx, const double* data, int64_t n_slices, int64_t n_t, int64_t slice_stride,
                     int64_t time_stride, double fs, uint32_t typemask, int32_t average_over_trials,
                     int64_t* n_lags_io, int32_t* k0, int32_t* k50, int32_t* ke, double* auc, double* tau,
                     double* acf_out, int64_t acf_capacity);
/* Device-resident data: upload once (itjl_data_upload keeps the matrix on the device, in the context), then
 * any number of itjl_acw_resident calls run without host -> device traffic - what a Julia caller doing repeated
 * acw() calls on the same recording pays.  itjl_data_upload takes the same strides as above. */
int itjl_data_upload(itjl_ctx* ctx, const double* data, int64_t n_slices, int64_t n_t, int64_t slice_stride,
                     int64_t time_stride);
int itjl_acw_resident(itjl_ctx* ctx, double fs, uint32_t typemask, int32_t average_over_trials,
                      int64_t* n_lags_io, int32_t* k0, int32_t* k50, int32_t* ke, double* auc, double* tau,
                      double* acf_out, int64_t acf_capacity);
/* ACF matrix of the most recent itjl_acw / _strided / _resident call on this context (the reference's
 * ACWResults.acf, acw.jl:275-307): n_rows x n_lags, column-major, kept on the device until the next acw call.
 * n_lags is only known after the call (ceil(1.1 max acw0), acw.jl:203-210): pass acf_out = NULL to those calls,
 * read the shape here (acf_out = NULL: shape query only), then fetch into an array of exactly that size - a
 * worst-case n_slices x n_t host buffer costs a first-touch page fault per 4 KB the library writes. */
int itjl_acw_acf(itjl_ctx* ctx, double* acf_out, int64_t acf_capacity, int64_t* n_rows_out, int64_t* n_lags_out);

/* fit_expdecay(lags, acf; dims = 2) with lags = (0:n_lags-1)*dt  (src/utils/utils.jl:113-138):
 * acf is (n_rows x n_lags) column-major; tau_out has n_rows entries. */
int itjl_fit_expdecay(itjl_ctx* ctx, const double* acf, int64_t n_rows, int64_t n_lags, double dt,
                      double* tau_out);
/* fit_expdecay_3_parameters(lags, acf; dims = 2) (reference src/utils/utils.jl:159-192; what
 * acw(...; skip_zero_lag = true) runs, acw.jl:212-216): A (exp(-t / tau) + B) fitted to acf[:, 2:end] over
 * lags[2:end] = dt, 2 dt, ...; same layout as itjl_fit_expdecay; NaN for a row holding a non-finite value. */
int itjl_fit_expdecay_3_parameters(itjl_ctx* ctx, const double* acf, int64_t n_rows, int64_t n_lags, double dt,
                                   double* tau_out);

/* ---- measurement helpers ---------------------------------------------------------- */
/* FP32 FMA issue-rate microbenchmark (the roofline denominator of the fused simulate+ACF
 * kernel): returns achieved TFLOP/s over `iters` launches. */
int itjl_fp32_peak(itjl_ctx* ctx, int32_t iters, double* tflops_out);
/* average device time (ms) of the kernels timed by the context since the last reset,
 * measured with CUDA events on the context stream around each dominant-kernel launch. */
int itjl_ctx_timing_reset(itjl_ctx* ctx);
int itjl_ctx_timing_get(itjl_ctx* ctx, int64_t* n_launches, double* total_ms);
/* stop timing: launches are asynchronous again (while timing is on, every timed launch is followed by an
 * event synchronisation, itjl_abc_distances_dev included) */
int itjl_ctx_timing_stop(itjl_ctx* ctx);
/* Tensor-core self-test: loads `image` (n_bytes of fp16 data, multiple of 16, <= 200 KiB) into shared
 * memory, issues `n_ops` tcgen05.mma (M = 128, kind::f16) described by ops[4*i .. 4*i+3] =
 * {A byte offset, B byte offset, TMEM column of D, accumulate flag} with the given no-swizzle
 * leading / stride byte offsets and instruction descriptor, and returns the 128 x 512 fp32
 * accumulator block (row-major).  Pins the operand layout the fused ACF kernel relies on. */
int itjl_tc_probe(itjl_ctx* ctx, const void* image, int64_t n_bytes, const int32_t* ops, int32_t n_ops,
                  uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t idesc, float* out);

#ifdef __cplusplus
}
#endif
#endif /* ITJL_B200_H */
