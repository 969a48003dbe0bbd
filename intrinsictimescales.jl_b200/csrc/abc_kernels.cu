// K1 / K1m: fused  simulate OU trials -> lag-windowed autocorrelation -> trial mean ->
// distance, one CTA per particle; the trajectories live only in shared memory.
// K3: accept / sequential-stop bookkeeping.
//
// Replaces, for a batch of particles, the body of the serial loop of basic_abc
// (reference src/core/abc.jl:174-195):
//   generate_data      src/models/one_timescale.jl:226-228, one_timescale_with_missing.jl:244-248,
//                      one_timescale_and_osc.jl:263-266 (ACF branch)
//   summary_stats      mean(comp_ac_fft(data; n_lags), dims=1)            (summary.jl:46-63)
//                      mean(comp_ac_time_missing(data; n_lags), dims=1)   (summary.jl:455-484)
//   distance_function  linear_distance / logarithmic_distance            (distances.jl:29-53)
//
// The ACF is evaluated as the direct lag sum  ac[k] = sum_t x'_t x'_{t+k}  on the centred,
// unit-norm series x' (so ac[0] = 1 and no per-trial normalisation is needed afterwards):
// mathematically the same quantity as the reference's zero-padded FFT route.
//
// Lag-sum tiling: thread = (lag tile of TK = 20 lags) x (time stream).  Per step a thread
// holds a[20] = x'[t0 .. t0+19] (warp-broadcast LDS.128) and a sliding window b[40] of
// x'[t0+k0 ..], and issues 400 FFMAs for 10 LDS.128.  The window is 2*TT wide so the slide
// is a two-phase register renaming (no moves).  Lane stride of 20 floats = 5 x 16 B makes
// the b loads bank-conflict free.
#include <stdlib.h>

#include "kernels.h"
#include "lagsum.cuh"
#include "simulate.cuh"

namespace itjl {

constexpr int kFlushTiles = 32;       // level-one sums of k_abc_acf cover at most 32 time tiles (640 samples)

template <int NT, bool MISSING, bool OSC>
__global__ void __launch_bounds__(NT, NT == 128 ? 5 : 2) k_abc_acf(const __grid_constant__ AbcAcfParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int buf_floats = P.xs_stride;
    float* xs0 = reinterpret_cast<float*>(smem_raw);
    float* xs1 = (P.n_bufs == 2) ? xs0 + buf_floats : xs0;
    SimShared<NT>* sh = reinterpret_cast<SimShared<NT>*>(xs0 + P.n_bufs * buf_floats);
    __shared__ double red[NT / 32];

    const int tid = threadIdx.x;
    const int64_t particle = blockIdx.x;

    const double tau = P.theta[particle];
    double freq = 0.0, coeff = 1.0;
    if (OSC) {
        freq = P.theta[particle + P.theta_ld];
        coeff = P.theta[particle + 2 * P.theta_ld];
    }
    const OuConsts oc = make_ou_consts(tau, P.dt, P.update, OSC ? ITJL_KIND_OU_OSC : ITJL_KIND_OU,
                                       freq, coeff, P.n_t, true, !MISSING);

    SimArgs g;
    g.n_t = P.n_t; g.chunk = P.chunk;
    g.inv_n = 1.0 / (double)P.n_t; g.sqrt_nm1 = sqrtf((float)(P.n_t - 1));
    g.keys = &P.keys; g.c3_noise = P.c3_noise; g.c3_init = P.c3_init;
    g.particle = (uint32_t)(P.particle_offset + particle);
    const size_t pt = (size_t)particle * P.n_trials;
    g.u0 = P.u0 ? P.u0 + pt : nullptr;
    g.noise = P.noise ? P.noise + pt * P.n_t : nullptr;
    g.phases = P.phases ? P.phases + pt : nullptr;
    g.mask_bits = P.mask_bits; g.n_valid = P.n_valid; g.mask_words = P.mask_words;
    g.miss_mean_ratio = (OSC && P.data_sd != 0.0) ? (float)(P.data_mean / P.data_sd) : 0.f;

    // zero everything once: the pads at or beyond chunk*NT (and the shifted copies' ends) stay zero
    for (int i = tid; i < P.n_bufs * buf_floats; i += NT) xs0[i] = 0.f;
    __syncthreads();

    // lag-sum work assignment
    const int tile = tid % P.lt;
    const int stream = tid / P.lt;
    const bool active = stream < P.n_streams;
    const int k0 = tile * TK;
    const int tile_begin = stream * P.tiles_per_stream;
    const int tile_end = min(tile_begin + P.tiles_per_stream, P.n_tiles);

    // two-level fp32 accumulation: acc[] collects one trial (a unit-norm series: sums <= 1, addends ~1/n_t),
    // tot[] the trials.  One accumulator for the whole particle stagnates - at 64 trials x 7722 samples and two
    // time streams it had lost 5e-5 at lag 0 (and 1e-4 in the worst particle) against the fp64 oracle
    float acc[TK], tot[TK];
#pragma unroll
    for (int j = 0; j < TK; ++j) { acc[j] = 0.f; tot[j] = 0.f; }
    SimCache sim_cache;

    for (int trial = 0; trial < P.n_trials; ++trial) {
        float* xs = (trial & 1) ? xs1 : xs0;
        if (!(P.debug & 1) || trial < 2)                      // debug bit 0: time the lag sum alone
            simulate_trial<NT, MISSING, OSC>(g, oc, trial, xs, sh, kScaleUnitNorm, 1.0f, 0.0f, sim_cache);
        __syncthreads();
        if (P.debug & 2) continue;                            // debug bit 1: time the simulator alone
        // (and at most kFlushTiles x 20 samples per level-one sum: with few time streams - long lag windows -
        // one thread owns most of a trial, 1e-5 .. 2e-5 off on 12 000-sample trials of the missing-data model)
        if (active) {
            for (int tb = tile_begin; tb < tile_end; tb += kFlushTiles) {
                acf_tiles(xs, acc, k0, tb, min(tb + kFlushTiles, tile_end));
#pragma unroll
                for (int j = 0; j < TK; ++j) { tot[j] += acc[j]; acc[j] = 0.f; }
            }
        }
        if (P.n_bufs == 1) __syncthreads();              // single buffer: lag sum done before it is rewritten
    }

    // ---- epilogue: reduce over time streams, trial mean, distance -----------------------
    __syncthreads();
    float* part = xs0;                               // [n_streams][lt*TK], fits (host-checked)
    const int row = P.lt * TK;
    if (active) {
#pragma unroll
        for (int j = 0; j < TK; ++j) part[stream * row + k0 + j] = tot[j];
    }
    __syncthreads();
    const double inv_trials = 1.0 / (double)P.n_trials;
    double local = 0.0;
    for (int k = tid; k < P.n_lags; k += NT) {
        float s = 0.f;
        for (int st = 0; st < P.n_streams; ++st) s += part[st * row + k];
        double v = __dmul_rn((double)s, inv_trials);         // no FMA contraction: same bits with and without stats_out
        if (k == 0 && isfinite(v)) v = 1.0;                  // ac[0] = acov[0] / acov[0] is exactly 1 in the reference (summary.jl:58, 479)
        if (P.stats_out) P.stats_out[(size_t)particle * P.n_lags + k] = v;
        const double tgt = (double)P.target[k];
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
        P.dist_out[particle] = d / (double)P.n_lags;
    }
}


// ---------------------------------------------------------------------------------------
// K3: accept mask + sequential stopping rule (abc.jl:185-199) for one batch, single CTA.
// n_total = 1 + index of the min_accepted-th acceptance in particle order (or n);
// n_accepted = acceptances among the first n_total distances.
__global__ void __launch_bounds__(1024) k_abc_accept(const double* __restrict__ dist, int64_t n,
                                                     double epsilon, int64_t min_accepted,
                                                     uint8_t* __restrict__ isaccepted,
                                                     int64_t* __restrict__ out /* [n_total, n_accepted] */) {
    __shared__ int warp_cnt[32];
    __shared__ long long s_running;
    __shared__ long long s_stop;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) { s_running = 0; s_stop = -1; }
    __syncthreads();
    for (int64_t base = 0; base < n; base += 1024) {
        const int64_t i = base + tid;
        const bool acc = (i < n) && (dist[i] <= epsilon);          // NaN -> false
        const unsigned bal = __ballot_sync(0xffffffffu, acc);
        if (lane == 0) warp_cnt[warp] = __popc(bal);
        __syncthreads();
        int before = 0, total = 0;
        for (int w = 0; w < 32; ++w) { const int c = warp_cnt[w]; if (w < warp) before += c; total += c; }
        const long long running = s_running;
        const long long rank = running + before + __popc(bal & ((1u << lane) - 1u)) + (acc ? 1 : 0);
        if (i < n) isaccepted[i] = acc ? 1 : 0;
        if (acc && rank == min_accepted) s_stop = i;
        __syncthreads();
        if (tid == 0) s_running = running + total;
        __syncthreads();
        if (s_stop >= 0) break;
    }
    if (tid == 0) {
        if (s_stop >= 0) { out[0] = s_stop + 1; out[1] = min_accepted; }
        else { out[0] = n; out[1] = s_running; }
    }
}

// ---------------------------------------------------------------------------------------
// host-side launch geometry

static int abc_acf_geometry_cfg(int n_t, int n_lags, int max_smem, int nt, int n_bufs, AbcAcfGeometry* g) {
    if (n_t < 2 || n_lags < 1 || n_lags > n_t) return -1;
    g->threads = nt; g->n_bufs = n_bufs;
    g->lt = (n_lags + TK - 1) / TK;
    if (g->lt > nt) return -2;
    g->n_streams = nt / g->lt;
    g->n_tiles = (n_t + TT - 1) / TT;
    if (g->n_streams > g->n_tiles) g->n_streams = g->n_tiles;
    g->tiles_per_stream = (g->n_tiles + g->n_streams - 1) / g->n_streams;
    g->n_streams = (g->n_tiles + g->tiles_per_stream - 1) / g->tiles_per_stream;
    int c = (n_t + nt - 1) / nt;
    g->chunk = (c + 3) & ~3;
    if (g->chunk > kQtab) return -3;
    int need = (g->n_tiles + g->lt + 2) * TT;               // furthest window read
    int stride = g->chunk * nt;
    if (stride < need) stride = need;
    int red = g->n_streams * g->lt * TK;                     // epilogue partials live in xs0
    if (stride < red) stride = red;
    g->xs_stride = (stride + 3) & ~3;
    g->smem_bytes = (size_t)n_bufs * g->xs_stride * sizeof(float) + sizeof(SimShared<256>);
    if ((int64_t)g->smem_bytes > (int64_t)max_smem) return -3;
    return 0;
}

// Default: 128-thread CTAs (up to 5 per SM: independent simulate / lag-sum phases interleave on
// every SM sub-partition), one series buffer per CTA; 256 threads when there are more than 128
// lag tiles.  itjl_tuning::abc_threads / abc_bufs override for tuning.
int abc_acf_geometry(int n_t, int n_lags, int max_smem, AbcAcfGeometry* g, const itjl_tuning* tun) {
    int nt = 128, n_bufs = 1;
    if (tun && tun->abc_threads) nt = (tun->abc_threads == 256) ? 256 : 128;
    if (tun && tun->abc_bufs) n_bufs = (tun->abc_bufs == 2) ? 2 : 1;
    if ((n_lags + TK - 1) / TK > nt) nt = 256;
    int rc = abc_acf_geometry_cfg(n_t, n_lags, max_smem, nt, n_bufs, g);
    if (rc == -3 && n_bufs == 2) rc = abc_acf_geometry_cfg(n_t, n_lags, max_smem, nt, 1, g);
    return rc;
}

template <int NT>
static cudaError_t abc_acf_launch_nt(const AbcAcfParams& P, bool missing, bool osc, size_t smem, cudaStream_t st) {
    void (*kern)(const AbcAcfParams) =
        missing ? (osc ? k_abc_acf<NT, true, true> : k_abc_acf<NT, true, false>)
                : (osc ? k_abc_acf<NT, false, true> : k_abc_acf<NT, false, false>);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<(unsigned)P.n_particles, NT, smem, st>>>(P);
    return cudaGetLastError();
}

cudaError_t abc_acf_launch(const AbcAcfParams& P, int threads, bool missing, bool osc, size_t smem,
                           cudaStream_t st) {
    return threads == 128 ? abc_acf_launch_nt<128>(P, missing, osc, smem, st)
                          : abc_acf_launch_nt<256>(P, missing, osc, smem, st);
}

cudaError_t abc_accept_launch(const double* dist, int64_t n, double eps, int64_t min_accepted,
                              uint8_t* isacc, int64_t* out2, cudaStream_t st) {
    k_abc_accept<<<1, 1024, 0, st>>>(dist, n, eps, min_accepted, isacc, out2);
    return cudaGetLastError();
}

}  // namespace itjl
