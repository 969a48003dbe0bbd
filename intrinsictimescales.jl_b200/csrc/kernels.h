// Internal declarations shared between the kernel translation units and the C-ABI layer.
#pragma once
#include "common.cuh"
#include "philox.cuh"
#include "fft_mixed.cuh"

namespace itjl {

// ---- abc_kernels.cu ------------------------------------------------------------------
struct AbcAcfParams {
    const double* theta;         // (n_particles x n_theta) column-major, leading dimension theta_ld
    int64_t n_particles;
    int64_t theta_ld;
    int n_theta;
    int n_t, n_trials, n_lags;
    int chunk, xs_stride, n_bufs;
    int lt, n_streams, tiles_per_stream, n_tiles;
    double dt, data_mean, data_sd;
    int update, kind, distance;
    PhiloxKeys keys;
    uint32_t c3_noise, c3_init;
    int64_t particle_offset;
    const float* target;
    const uint32_t* mask_bits;
    const int* n_valid;
    int mask_words;
    const double* u0;
    const double* noise;
    const double* phases;
    double* dist_out;
    double* stats_out;
    int debug;                   // tuning only (ITJL_ABC_DEBUG): 1 = skip simulate, 2 = skip lag sum
};
struct AbcAcfGeometry {
    int threads, n_bufs, chunk, xs_stride, lt, n_streams, tiles_per_stream, n_tiles;
    size_t smem_bytes;
};
int abc_acf_geometry(int n_t, int n_lags, int max_smem, AbcAcfGeometry* g, const itjl_tuning* tun = nullptr);
cudaError_t abc_acf_launch(const AbcAcfParams& P, int threads, bool missing, bool osc, size_t smem,
                           cudaStream_t st);
cudaError_t abc_accept_launch(const double* dist, int64_t n, double eps, int64_t min_accepted,
                              uint8_t* isacc, int64_t* out2, cudaStream_t st);

// ---- psd_kernels.cu ------------------------------------------------------------------
struct AbcPsdParams {
    const double* theta;         // (n_particles x n_theta) column-major, leading dimension theta_ld
    int64_t n_particles;
    int64_t theta_ld;
    int n_theta;
    int n_t, n_trials, n_stats;
    int chunk, xs_len;          // simulate chunk; floats in the staging buffer
    int n2, log2_n2;            // zero-padded real FFT length 2*nextpow2(n_t)
    double dt, data_mean, data_sd;
    int update, kind, distance;
    PhiloxKeys keys;
    uint32_t c3_noise, c3_init;
    int64_t particle_offset;
    const float* target;        // n_stats
    const int4* binpos;         // n_stats: {storage pos of Z[k], of Z[N-k], bits of the twiddle exp(-2 pi i k / n2)}, k = bin + 1
    const int32_t* binidx;      // n_stats: statistic index of the i-th entry of binpos (sorted by storage position)
    const float* window;        // n_t Hamming window
    const float2* twiddle;      // n2/2 complex roots exp(-2 pi i k / (n2/2))... see psd_kernels.cu
    double psd_scale;           // 1 / (fs * sum(window^2))
    const double* u0;
    const double* noise;
    const double* phases;
    double* dist_out;
    double* stats_out;
    float* accb_global = nullptr;   // per-particle bin sums in global memory [grid][(n_stats + 3) & ~3] when they do not fit
                                    // in shared memory beside the FFT buffer (many selected bins), else null
    int64_t grid = 0;               // CTAs of this launch (a sub-batch of the particles when accb_global is used)
    int round0_bins = 0;            // k_abc_psd2: entries of binpos that belong to the first round of sub-transforms
};
size_t abc_psd_smem_bytes(int n_t, int n2, int n_stats, int* chunk_out, int* xs_len_out, bool accb_in_smem = true);
int psd_fft_pos(int k, int N);
size_t psd_fft_smem_bytes(int n2);
cudaError_t abc_psd_launch(const AbcPsdParams& P, bool osc, size_t smem, cudaStream_t st);
// psd2_kernels.cu: warp-specialised version for 4096 < n_t <= 16384
bool abc_psd2_supported(int n_t, int n2);
size_t abc_psd2_smem_bytes(int n_t, int n2, int n_stats, int* chunk_out, int* xs_len_out);
void abc_psd2_bin(int k, int N, int* addr, int* round);
cudaError_t abc_psd2_launch(const AbcPsdParams& P, bool osc, size_t smem, cudaStream_t st);
// standalone: comp_psd_adfriendly on stored data (n_slices x n_t col-major double)
cudaError_t psd_data_launch(const double* data, int64_t n_slices, int n_t, int n2, int log2_n2,
                            const float* window, const float2* twiddle, double psd_scale,
                            double* out, size_t smem, cudaStream_t st);
void psd_make_tables(int n_t, int n2, float* window_host, float2* twiddle_host, double* win_ss);

// ---- pgram_kernels.cu ----------------------------------------------------------------
struct AbcPgramParams {
    const double* theta;         // (n_particles x n_theta) column-major, leading dimension theta_ld
    int64_t n_particles;
    int64_t theta_ld;
    int n_theta;
    int n_t, n_trials, n_stats;
    int chunk;                  // simulate chunk (samples per thread)
    int bufb_len;               // float2 entries of the second transform buffer (it also holds the simulated trial)
    FftPlan plan;               // nfft = nextfastfft(n_t) and its radices
    double dt, data_mean, data_sd;
    int update, kind, distance;
    PhiloxKeys keys;
    uint32_t c3_noise, c3_init;
    int64_t particle_offset;
    const float* target;        // n_stats
    const int32_t* bins;        // n_stats: rfft bin k (>= 1) of each statistic
    const float* window;        // n_t symmetric Hamming window
    const float2* roots;        // two-level table of the nfft-th roots of unity (roots_two_level_entries(nfft))
    double psd_scale;           // 1 / (fs * sum(window^2))
    const double* u0;
    const double* noise;
    const double* phases;
    double* dist_out;
    double* stats_out;
};
size_t abc_pgram_smem_bytes(int n_t, int nfft, int n_stats, int* chunk_out, int* bufb_len_out);
void pgram_make_tables(int n_t, int nfft, float* window_f, double* window_d, float2* roots2_f, double2* roots_d, double* win_ss);
cudaError_t abc_pgram_launch(const AbcPgramParams& P, bool osc, size_t smem, cudaStream_t st);
// stored data (n_slices x n_t col-major double), fp64: all nfft/2 bins of every slice; scratch = grid * 2 * nfft double2
cudaError_t pgram_data_launch(const double* data, int64_t n_slices, int n_t, const FftPlan& plan, const double* window,
                              const double2* roots, double psd_scale, double2* scratch, int grid, double* out, cudaStream_t st);
// generalised Lomb-Scargle periodogram of stored data with NaN gaps on the grid f0 + k df, k < n_freq
int lomb_czt_length(int n_t, int n_freq);
size_t lomb_czt_table_elems(int n_t, int n_freq, int M);
size_t lomb_czt_scratch_elems(int n_freq, int M, int grid);
cudaError_t lomb_czt_launch(const double* data, int64_t n_slices, int n_t, double dt, double f0, double df, int n_freq, int M,
                            double2* tables, double2* scratch, int grid, double* out, cudaStream_t st);
cudaError_t lomb_scargle_launch(const double* data, int64_t n_slices, int n_t, double dt, double f0, double df, int n_freq,
                                double* out, cudaStream_t st);

// ---- knee_kernels.cu -----------------------------------------------------------------
int knee_fit_grid(int sm_count, int64_t n_rows);
cudaError_t knee_fit_launch(const double* rows, int64_t n_rows, int64_t ld, int i0, int n, const double* freqs, double min_freq,
                            double max_freq, int mode, int max_peaks, double* knee_out, double* amp_out, double* peak_out,
                            double* work, int grid, cudaStream_t st);
cudaError_t combine_launch(double* dist, int64_t n, const double* fit, int is_knee, const double* peak, double w0, double w1, double w2,
                           double data_tau, double data_osc, cudaStream_t st);
cudaError_t expdecay3_launch(const double* rows, int64_t n_rows, int64_t ld, int n_lags, const double* lags, double* tau_out, int grid,
                             cudaStream_t st);
cudaError_t transpose_launch(const double* in, int64_t n_rows, int64_t n_cols, double* out, cudaStream_t st);
cudaError_t col_mean_launch(const double* in, int64_t n_rows, int64_t n_cols, double* out, cudaStream_t st);

// ---- ou_kernels.cu -------------------------------------------------------------------
struct OuGenParams {
    int n_t, n_trials, chunk;
    double tau, freq, coeff, dt;
    int update, kind, mode;     // mode: kScale*
    float scale, shift;
    PhiloxKeys keys;
    uint32_t c3_noise, c3_init, particle;
    const double* u0;
    const double* noise;
    const double* phases;
    double* out;                // (n_trials x n_t) column-major
};
cudaError_t ou_generate_launch(const OuGenParams& P, cudaStream_t st);

// ---- acw_kernels.cu ------------------------------------------------------------------
// acf_work is a row-per-slice work buffer [slice][cap]; n_done[slice] = lags present.
cudaError_t acw_scan_launch(const double* data, int64_t n_slices, int n_t, int missing, int max_lags,
                            double* acf_work, int64_t cap, int32_t* n_done, int32_t* k0, int32_t* k50,
                            int32_t* ke, int32_t* nan_flag, const int32_t* slice_list, int64_t n_list,
                            int max_smem, cudaStream_t st);
// streaming first pass (raw lagged products of the first acw_stream_lags() lags, one read of the data)
cudaError_t acw_stream_launch(const double* data, int64_t n_slices, int n_t, double* part, int n_seg, int seg_len,
                              double* acf_work, int64_t cap, int32_t* n_done, int32_t* k0, int32_t* k50,
                              int32_t* ke, int32_t* redo, int w, cudaStream_t st);
int acw_stream_lags();
size_t acw_stream_part_doubles(int64_t n_slices, int n_seg);
void acw_stream_segments(int n_t, int64_t n_slices, int sm_count, int* n_seg, int* seg_len, int force_waves = 0, int w = 32);
cudaError_t acf_fill_launch(const double* data, int64_t n_slices, int n_t, int missing, int need,
                            double* acf_work, int64_t cap, int32_t* n_done, const int32_t* slice_list,
                            int64_t n_list, int max_smem, cudaStream_t st, int32_t* nan_flag = nullptr);
cudaError_t acf_cross_launch(const double* acf, int64_t n_rows, int64_t ld, int n_lags, int32_t* k0,
                             int32_t* k50, int32_t* ke, cudaStream_t st);
cudaError_t acf_mean_launch(const double* acf, int64_t n_rows, int64_t ld, int n_lags, double* out,
                            cudaStream_t st);
cudaError_t acf_export_launch(const double* acf, int64_t n_rows, int64_t ld, int n_lags, double* out,
                              cudaStream_t st);
cudaError_t acw_auc_launch(const double* acf, int64_t n_rows, int64_t ld, int w, double dt,
                           const int* strides, int n_strides, double* auc, cudaStream_t st);
cudaError_t acw_tau_launch(const double* acf, int64_t n_rows, int64_t ld, int n_lags, double dt, double* tau,
                           cudaStream_t st);
int acf_max_reach(int n_t, int max_smem);
cudaError_t relayout_launch(const double* raw, int64_t n_slices, int64_t n_t, int64_t ss, int64_t ts, double* out,
                            cudaStream_t st);

// ---- abc_tc_kernels.cu --------------------------------------------------------------
constexpr int kTcMaxOps = 6;
constexpr int kTcMaxWindows = 4;                   // lag windows (launches) of the tensor-core kernel: n_lags <= 1537
struct AbcTcGeometry {
    int groups, chunk, epi_offset;
    int rows_total, ksteps, sbo_bytes;
    int used_cols, mixed, seg_trials, n_terms;
    float bias_a, bias_b;        // first-order correction of the accumulator's round-toward-zero (see abc_tc_geometry)
    double pred_err;             // error model's bound for this plan (abc_tc_predicted_error)
    int tc_lags, tail_start, tail_lt, tail_steps, tail_steps_per_stream;
    // MMA ops of one K step: A / B operand start (8-row MN block, K-row shift), shape, TMEM address
    int n_ops;
    int op_a_blk[kTcMaxOps], op_b_row[kTcMaxOps], op_b_blk[kTcMaxOps];
    uint32_t op_idesc[kTcMaxOps], op_taddr[kTcMaxOps];
    // further lag windows (n_lags 465 .. 1537): the same kernel launched again, window w with the B operand
    // win1_s0 w = 3 w rows further down (plan-A tiles), accumulator column v <-> lag 384 w + v - a; it scores
    // the lags the windows before it have not (from 449, resp. 384 w + 1) and adds its part of the distance
    int n_windows, win1_s0, win1_used_cols;      // win1_used_cols: accumulator columns of the last window
    float tc_scale, tc_fix;
    int grid;
    size_t smem_bytes;
};
struct AbcTcParams {
    const double* theta;         // (n_particles x n_theta) column-major, leading dimension theta_ld
    int64_t n_particles;
    int64_t theta_ld;
    int n_theta;
    int n_t, n_trials, n_lags;
    double dt;
    int update, kind, distance;
    PhiloxKeys keys;
    uint32_t c3_noise, c3_init;
    int64_t particle_offset;
    const float* target;
    const uint32_t* mask_bits;
    const int* n_valid;
    int mask_words;
    const double* u0;
    const double* noise;
    const double* phases;
    double* dist_out;
    double* stats_out;
    // geometry (AbcTcGeometry)
    int groups, chunk, epi_offset;
    int ksteps, sbo_bytes;
    int used_cols, seg_trials, n_terms;
    int tc_lags, tail_start, tail_lt, tail_steps, tail_steps_per_stream;
    int k_lo;                    // lag window of this launch: accumulator lags [k_lo, tc_lags) relative to lag_base
    int lag_base;                // absolute lag of relative lag 0 (384 w for lag window w)
    int accumulate;              // 1: add this window's share of the distance to dist_out (second window)
    int n_ops;
    uint32_t op_a_off[kTcMaxOps], op_b_off[kTcMaxOps];      // descriptor start offsets, 16 B units
    uint32_t op_idesc[kTcMaxOps], op_taddr[kTcMaxOps];
    float tc_scale, tc_unscale, tc_fix, tc_unfix;
    float bias_a, bias_b;
    float miss_mean_ratio;       // data_mean / data_sd for the oscillation + missing-data model, else 0
    size_t smem_bytes;
    int debug;                   // tuning only (ITJL_ABC_DEBUG): 1 = no MMAs, 2 = no simulation after the first particle
};
// `tun` (nullable: defaults) carries the overrides and the error bound; mask_frac / mean_ratio describe the
// missing-data models (fraction of masked samples, data_mean / data_sd of the oscillation variant).  Returns 0 when
// the shape has a plan whose predicted error is within tun->tc_error_bound (or the kernel is forced), -4 when
// only the error bound fails, < 0 otherwise.
int abc_tc_geometry(int n_t, int n_lags, int n_trials, int max_smem, int sm_count, AbcTcGeometry* g,
                    const itjl_tuning* tun = nullptr, double mask_frac = 0.0, double mean_ratio = 0.0, bool missing = false);
double abc_tc_predicted_error(int64_t n_t, int64_t n_trials, int n_terms, int seg_rows, double mask_frac, double mean_ratio);
cudaError_t abc_tc_launch(const AbcTcParams& P, bool missing, bool osc, bool mixed, size_t smem, int grid,
                          cudaStream_t st);
cudaError_t tc_probe_launch(const void* image_dev, int n_bytes, const int4* ops_dev, int n_ops, uint32_t lbo,
                            uint32_t sbo, uint32_t idesc, float* out_dev, cudaStream_t st);

// ---- pmc_kernels.cu ------------------------------------------------------------------
double pmc_coord_scale(bool fp32);
void pmc_mixture_split(int64_t n, int64_t n_prev, int sm_count, int* n_split, int64_t* j_per_split);
cudaError_t pmc_mixture_launch(bool fp32, const void* z, const void* zp, const void* wp, int64_t n, int64_t n_prev, int p,
                               int n_split, int64_t j_per_split, double* part, cudaStream_t st);
cudaError_t pmc_grid_density_launch(const double* part, int n_split, int n_rows, double* dens, cudaStream_t st);
cudaError_t pmc_grid_weights_launch(const double* dens, int n_grid, double z_lo, double h, const double* z, int64_t n,
                                    const double* prior_pdf, double norm, double* w_out, double* block_sums, cudaStream_t st);
cudaError_t pmc_finish_launch(const double* part, int n_split, int64_t n, const double* prior_pdf, double norm,
                              double* w_out, double* block_sums, bool normalize, cudaStream_t st);
cudaError_t pmc_normalize_launch(double* w, int64_t n, double* block_sums, cudaStream_t st);
size_t sort_temp_bytes(int64_t n);
cudaError_t mask_invalid_launch(const double* v, int64_t n, double upper, double* out, unsigned long long* n_valid, cudaStream_t st);
cudaError_t sort_doubles_launch(void* temp, size_t temp_bytes, const double* in, double* out, int64_t n, cudaStream_t st);
cudaError_t gather_launch(const double* sorted, const int64_t* idx, int k, double* out, cudaStream_t st);
cudaError_t pmc_propose_launch(const double* theta_prev, const double* cdf, int64_t n_prev, int p, const double* L, int64_t n,
                               uint64_t seed, uint64_t generation, int64_t first, double* out, cudaStream_t st);
cudaError_t wcov_launch(const double* x, const double* w, int64_t n, int p, int pass, double wsum, const double* xbar,
                        double* out, int blocks, cudaStream_t st);

// ---- peak_kernels.cu -----------------------------------------------------------------
cudaError_t fp32_peak_launch(float* sink, int blocks, int iters, cudaStream_t st);
double fp32_peak_flops_per_launch(int blocks, int iters);

}  // namespace itjl
