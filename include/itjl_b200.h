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
 *   - one context = one GPU + one stream; a context is not re-entrant (serialise calls).
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
enum { ITJL_SUMMARY_ACF = 0, ITJL_SUMMARY_ACF_MISSING = 1, ITJL_SUMMARY_PSD = 2 };
enum { ITJL_DIST_LINEAR = 0, ITJL_DIST_LOG = 1 };
enum { ITJL_UPDATE_EXACT = 0, ITJL_UPDATE_EM = 1 };

/* acw type mask bits (src/core/acw.jl:132, acf_acwtypes) */
enum { ITJL_ACW0 = 1, ITJL_ACW50 = 2, ITJL_ACWEULER = 4, ITJL_AUC = 8, ITJL_TAU = 16 };

/* ---- library / context ------------------------------------------------------------ */
const char* itjl_version(void);
int itjl_device_count(int* n_out);
int itjl_ctx_create(int device, itjl_ctx** ctx_out);
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
    const int32_t* freq_bins;     /* PSD: findall(model.freq_idx) .- 1 (0-based into psd[2:n2/2]), n_stats; else NULL */
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

/* ---- summary statistics on stored data -------------------------------------------- */
/* comp_ac_fft(data; dims = 2, n_lags)  (src/stats/summary.jl:46-63, 79-89) computed as the
 * mathematically identical time-domain lag sum.  out: (n_slices x n_lags) column-major. */
int itjl_acf(itjl_ctx* ctx, const double* data, int64_t n_slices, int64_t n_t, int64_t n_lags, double* out);
/* comp_ac_time_missing(data; dims = 2, n_lags)  (src/stats/summary.jl:455-496); NaN = missing. */
int itjl_acf_missing(itjl_ctx* ctx, const double* data, int64_t n_slices, int64_t n_t, int64_t n_lags,
                     double* out);
/* comp_psd_adfriendly(data, fs; dims = 2)[1]  (src/stats/summary.jl:183-235).
 * out: (n_slices x (n2/2 - 1)) column-major with n2 = 2 * nextpow2(n_t). */
int itjl_psd_hamming(itjl_ctx* ctx, const double* data, int64_t n_slices, int64_t n_t, double fs, double* out);

/* ---- acw() ------------------------------------------------------------------------ */
/* The ACF-based part of acw(data, fs; acwtypes, n_lags, average_over_trials)
 * (src/core/acw.jl:141-231) with the reducers of src/utils/utils.jl:113-138, 218-345.
 *   k0 / k50 / ke : 0-based first lag index with acf <= 0 / 0.5 / 1/e, -1 if never
 *                   (the caller indexes its own `lags` range; NaN for -1);
 *   auc           : acw_romberg(dt, acf[1:k0_first]) per slice (NaN when k0_first < 2);
 *   tau           : fit_expdecay(lags[1:n_lags], acf[.., 1:n_lags]) per slice;
 *   n_lags_io     : in: user n_lags or 0 for the reference default ceil(1.1*max k0);
 *                   out: the value used;
 *   acf_out       : NULL, or capacity acf_capacity doubles; receives (n_out x n_lags)
 *                   column-major, n_out = 1 when average_over_trials else n_slices.
 * Any of k0/k50/ke/auc/tau may be NULL when not requested in typemask. */
int itjl_acw(itjl_ctx* ctx, const double* data, int64_t n_slices, int64_t n_t, double fs,
             uint32_t typemask, int32_t average_over_trials, int64_t* n_lags_io,
             int32_t* k0, int32_t* k50, int32_t* ke, double* auc, double* tau,
             double* acf_out, int64_t acf_capacity);

/* fit_expdecay(lags, acf; dims = 2) with lags = (0:n_lags-1)*dt  (src/utils/utils.jl:113-138):
 * acf is (n_rows x n_lags) column-major; tau_out has n_rows entries. */
int itjl_fit_expdecay(itjl_ctx* ctx, const double* acf, int64_t n_rows, int64_t n_lags, double dt,
                      double* tau_out);

/* ---- measurement helpers ---------------------------------------------------------- */
/* FP32 FMA issue-rate microbenchmark (the roofline denominator of the fused simulate+ACF
 * kernel): returns achieved TFLOP/s over `iters` launches. */
int itjl_fp32_peak(itjl_ctx* ctx, int32_t iters, double* tflops_out);
/* average device time (ms) of the kernels timed by the context since the last reset,
 * measured with CUDA events on the context stream around each dominant-kernel launch. */
int itjl_ctx_timing_reset(itjl_ctx* ctx);
int itjl_ctx_timing_get(itjl_ctx* ctx, int64_t* n_launches, double* total_ms);
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
