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
#include "kernels.h"
#include "simulate.cuh"

namespace itjl {

constexpr int TK = 20;   // lags per thread
constexpr int TT = 20;   // samples per time tile


__device__ __forceinline__ void load5(const float4* __restrict__ p, float* dst) {
#pragma unroll
    for (int q = 0; q < 5; ++q) {
        const float4 v = p[q];
        dst[4 * q + 0] = v.x; dst[4 * q + 1] = v.y; dst[4 * q + 2] = v.z; dst[4 * q + 3] = v.w;
    }
}

// acc[j] += sum over tiles [tile_begin, tile_end) of sum_i x[t0+i] * x[t0+i+k0+j]
__device__ __forceinline__ void acf_tiles(const float* __restrict__ xs, float (&acc)[TK], int k0,
                                          int tile_begin, int tile_end) {
    float b[2 * TT];
    float a[TT];
    const float4* pa = reinterpret_cast<const float4*>(xs + tile_begin * TT);
    const float4* pb = reinterpret_cast<const float4*>(xs + tile_begin * TT + k0);
    load5(pb, b);
    pb += 5;
    int n = tile_end - tile_begin;
    while (n > 0) {
        // phase 0: window = b[0..39]
        load5(pb, b + TT); pb += 5;
        load5(pa, a); pa += 5;
#pragma unroll
        for (int i = 0; i < TT; ++i)
#pragma unroll
            for (int j = 0; j < TK; ++j) acc[j] = fmaf(a[i], b[i + j], acc[j]);
        if (--n == 0) break;
        // phase 1: window = b[20..39], b[0..19]
        load5(pb, b); pb += 5;
        load5(pa, a); pa += 5;
#pragma unroll
        for (int i = 0; i < TT; ++i)
#pragma unroll
            for (int j = 0; j < TK; ++j) acc[j] = fmaf(a[i], b[(TT + i + j) % (2 * TT)], acc[j]);
        --n;
    }
}

template <bool MISSING, bool OSC>
__global__ void __launch_bounds__(kThreads, 2) k_abc_acf(const AbcAcfParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* xs0 = reinterpret_cast<float*>(smem_raw);
    float* xs1 = xs0 + P.xs_stride;
    SimShared* sh = reinterpret_cast<SimShared*>(xs1 + P.xs_stride);
    __shared__ double red[kWarps];

    const int tid = threadIdx.x;
    const int64_t particle = blockIdx.x;

    const double tau = P.theta[particle];
    double freq = 0.0, coeff = 1.0;
    if (OSC) {
        freq = P.theta[particle + P.n_particles];
        coeff = P.theta[particle + 2 * P.n_particles];
    }
    const OuConsts oc = make_ou_consts(tau, P.dt, P.update, OSC ? ITJL_KIND_OU_OSC : ITJL_KIND_OU,
                                       freq, coeff);

    SimArgs g;
    g.n_t = P.n_t; g.chunk = P.chunk;
    g.k0 = P.k0; g.k1 = P.k1; g.c3_noise = P.c3_noise; g.c3_init = P.c3_init;
    g.particle = (uint32_t)(P.particle_offset + particle);
    const size_t pt = (size_t)particle * P.n_trials;
    g.u0 = P.u0 ? P.u0 + pt : nullptr;
    g.noise = P.noise ? P.noise + pt * P.n_t : nullptr;
    g.phases = P.phases ? P.phases + pt : nullptr;
    g.mask_bits = P.mask_bits; g.n_valid = P.n_valid; g.mask_words = P.mask_words;

    // zero the tail pads once: everything at or beyond chunk*kThreads stays zero
    for (int i = P.chunk * kThreads + tid; i < P.xs_stride; i += kThreads) { xs0[i] = 0.f; xs1[i] = 0.f; }

    // lag-sum work assignment
    const int tile = tid % P.lt;
    const int stream = tid / P.lt;
    const bool active = stream < P.n_streams;
    const int k0 = tile * TK;
    const int tile_begin = stream * P.tiles_per_stream;
    const int tile_end = min(tile_begin + P.tiles_per_stream, P.n_tiles);

    float acc[TK];
#pragma unroll
    for (int j = 0; j < TK; ++j) acc[j] = 0.f;

    for (int trial = 0; trial < P.n_trials; ++trial) {
        float* xs = (trial & 1) ? xs1 : xs0;
        simulate_trial<MISSING, OSC>(g, oc, trial, xs, sh, kScaleUnitNorm, 1.0f, 0.0f);
        __syncthreads();
        if (active && tile_begin < tile_end) acf_tiles(xs, acc, k0, tile_begin, tile_end);
    }

    // ---- epilogue: reduce over time streams, trial mean, distance -----------------------
    __syncthreads();
    float* part = xs0;                               // [n_streams][lt*TK], fits (host-checked)
    const int row = P.lt * TK;
    if (active) {
#pragma unroll
        for (int j = 0; j < TK; ++j) part[stream * row + k0 + j] = acc[j];
    }
    __syncthreads();
    const double inv_trials = 1.0 / (double)P.n_trials;
    double local = 0.0;
    for (int k = tid; k < P.n_lags; k += kThreads) {
        float s = 0.f;
        for (int st = 0; st < P.n_streams; ++st) s += part[st * row + k];
        const double v = (double)s * inv_trials;
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
        for (int w = 0; w < kWarps; ++w) d += red[w];
        P.dist_out[particle] = d / (double)P.n_lags;
    }
}

template __global__ void k_abc_acf<false, false>(const AbcAcfParams);
template __global__ void k_abc_acf<true, false>(const AbcAcfParams);
template __global__ void k_abc_acf<false, true>(const AbcAcfParams);
template __global__ void k_abc_acf<true, true>(const AbcAcfParams);

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

int abc_acf_geometry(int n_t, int n_lags, int max_smem, AbcAcfGeometry* g) {
    if (n_t < 2 || n_lags < 1 || n_lags > n_t) return -1;
    g->lt = (n_lags + TK - 1) / TK;
    if (g->lt > kThreads) return -2;
    g->n_streams = kThreads / g->lt;
    g->n_tiles = (n_t + TT - 1) / TT;
    if (g->n_streams > g->n_tiles) g->n_streams = g->n_tiles;
    g->tiles_per_stream = (g->n_tiles + g->n_streams - 1) / g->n_streams;
    g->n_streams = (g->n_tiles + g->tiles_per_stream - 1) / g->tiles_per_stream;
    int c = (n_t + kThreads - 1) / kThreads;
    g->chunk = (c + 3) & ~3;
    int need = (g->n_tiles + g->lt + 2) * TT;               // furthest window read
    int stride = g->chunk * kThreads;
    if (stride < need) stride = need;
    int red = g->n_streams * g->lt * TK;                     // epilogue partials live in xs0
    if (stride < red) stride = red;
    g->xs_stride = (stride + 3) & ~3;
    g->smem_bytes = (size_t)2 * g->xs_stride * sizeof(float) + sizeof(SimShared);
    if ((int64_t)g->smem_bytes > (int64_t)max_smem) return -3;
    return 0;
}

cudaError_t abc_acf_launch(const AbcAcfParams& P, bool missing, bool osc, size_t smem, cudaStream_t st) {
    void (*kern)(const AbcAcfParams) =
        missing ? (osc ? k_abc_acf<true, true> : k_abc_acf<true, false>)
                : (osc ? k_abc_acf<false, true> : k_abc_acf<false, false>);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<(unsigned)P.n_particles, kThreads, smem, st>>>(P);
    return cudaGetLastError();
}

cudaError_t abc_accept_launch(const double* dist, int64_t n, double eps, int64_t min_accepted,
                              uint8_t* isacc, int64_t* out2, cudaStream_t st) {
    k_abc_accept<<<1, 1024, 0, st>>>(dist, n, eps, min_accepted, isacc, out2);
    return cudaGetLastError();
}

}  // namespace itjl
