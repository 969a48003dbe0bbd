// Tensor-core (tcgen05 / TMEM) variant of the fused aABC ACF kernel, plus the self-test probe
// that pins the shared-memory descriptor conventions it relies on.
#include <stdlib.h>

#include "kernels.h"
#include "lagsum.cuh"
#include "simulate.cuh"
#include "tc05.cuh"

namespace itjl {

// ---------------------------------------------------------------------------------------
// K1-TC: fused simulate -> ACF -> trial mean -> distance with the lag sums on the tensor cores.
//
// Replaces the same reference code as k_abc_acf (abc_kernels.cu): the body of basic_abc's loop
// (src/core/abc.jl:174-195) = generate_data + summary_stats (comp_ac_fft / comp_ac_time_missing,
// src/stats/summary.jl:46-63, 455-484) + distance_function (src/stats/distances.jl:29-53).
//
// Lag sums as a Gram product.  With the centred unit-norm series cut into rows of T = 128
// samples, X[i][a] = x'[128 i + a],
//     D_s[a][b] = sum_i X[i][a] * X[i+s][b]   =  sum over t = 128 i + a of x'_t x'_{t + 128 s + b - a},
// so ac[l] = sum over {(s, a, b): 128 s + b - a = l} of D_s[a][b].  Each D_s is one
// tcgen05.mma chain (M = 128 rows a, N = 128 columns b, K = rows i of every trial of the
// particle) accumulating in fp32 in tensor memory; 4 tiles = all 512 TMEM columns cover lags
// 0..384.  Both operands are the SAME shared-memory buffer in the MN-major no-swizzle layout
// (K rows 16 B apart), the B operand of D_s simply starts s rows later.  x' is fed as an fp16
// hi/lo pair (hi*hi + hi*lo + lo*hi: ~22-bit operands, fp32 accumulate) - 3 MMAs per tile.
// Lags beyond the TMEM capacity (n_lags > 385) are summed by the register-tiled FFMA loop of
// lagsum.cuh on the same fp32 series.
//
// CTA = G simulate groups of 128 threads (one trial each at a time, Philox + blocked scan as in
// simulate.cuh, named barriers) + one issuer warp.  Persistent: one CTA per SM loops over
// particles; the issuer's single thread waits for a group's "operands ready" mbarrier, issues
// that trial's MMAs and commits to the group's "operands free" mbarrier, so the tensor pipe
// runs under the simulation of the following trials.  Per particle the accumulators are drained
// once (tcgen05.ld), summed along diagonals through a skewed shared-memory transpose, and the
// distance is formed in fp64.
constexpr int kTcGroupThreads = 128;
constexpr int kTcMaxGroups = 4;
constexpr int kTcMaxThreads = kTcGroupThreads * kTcMaxGroups + 32;
constexpr int kTcZFloats = 33 * 32;                 // per-warp skew scratch
constexpr int kTcBarAll = 8;                        // named barrier over all simulate threads

template <bool MISSING, bool OSC>
__global__ void __launch_bounds__(kTcMaxThreads, 1) k_abc_acf_tc(const AbcTcParams P) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bar_full[kTcMaxGroups];
    __shared__ __align__(8) uint64_t bar_empty[kTcMaxGroups];
    __shared__ __align__(8) uint64_t bar_acc_full, bar_acc_empty;
    __shared__ uint32_t tmem_slot;
    __shared__ SimShared<kTcGroupThreads> shs[kTcMaxGroups];
    __shared__ double red[kTcMaxGroups * 4];

    const int G = P.groups;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_sim = kTcGroupThreads * G;
    const int op_bytes = 16 * P.sbo_bytes;                       // one fp16 operand plane (hi or lo)
    float* stag = reinterpret_cast<float*>(smem_raw);            // G staging buffers of xs_stride floats
    unsigned char* opnd = smem_raw + (size_t)P.opnd_offset;      // G x {hi, lo} planes

    // ---- one-time setup ------------------------------------------------------------------
    {
        uint4* z = reinterpret_cast<uint4*>(smem_raw);
        const int n16 = (P.opnd_offset + G * 2 * op_bytes) / 16;
        for (int i = tid; i < n16; i += blockDim.x) z[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    if (tid == 0) {
        for (int g = 0; g < G; ++g) { tc::mbar_init(&bar_full[g], kTcGroupThreads); tc::mbar_init(&bar_empty[g], 1); }
        tc::mbar_init(&bar_acc_full, 1);
        tc::mbar_init(&bar_acc_empty, (uint32_t)n_sim);
        tc::fence_mbar_init();
    }
    if (warp == 4 * G) tc::tmem_alloc(&tmem_slot, 512);
    tc::fence_proxy_async_smem();                                // zeroed operand rows -> async proxy
    tc::fence_before_thread_sync();
    __syncthreads();
    tc::fence_after_thread_sync();
    const uint32_t tmem_base = tmem_slot;

    if (warp == 4 * G) {
        // ================= MMA issuer (one thread) ==========================================
        if (lane == 0) {
            const uint32_t idesc_full = tc::make_idesc_f16_mn(128, 128);
            const uint32_t idesc_last = tc::make_idesc_f16_mn(128, P.n_last);
            const uint32_t opnd_addr = tc::smem_u32(opnd);
            uint32_t it = 0;
            for (int64_t particle = blockIdx.x; particle < P.n_particles; particle += gridDim.x, ++it) {
                tc::mbar_wait(&bar_acc_empty, (it & 1u) ^ 1u);   // previous particle drained
                tc::fence_after_thread_sync();
                for (int trial = 0; trial < P.n_trials; ++trial) {
                    const int g = trial % G;
                    const uint32_t n_g = (uint32_t)((P.n_trials - g + G - 1) / G);
                    const uint32_t k = it * n_g + (uint32_t)(trial / G);
                    tc::mbar_wait(&bar_full[g], k & 1u);
                    tc::fence_after_thread_sync();
                    const uint32_t hi_addr = opnd_addr + (uint32_t)(g * 2 * op_bytes);
                    const uint64_t d_hi = tc::make_smem_desc(hi_addr, 128u, (uint32_t)P.sbo_bytes);
                    const uint64_t d_lo = tc::make_smem_desc(hi_addr + (uint32_t)op_bytes, 128u, (uint32_t)P.sbo_bytes);
                    for (int ks = 0; ks < P.ksteps; ++ks) {
                        const uint64_t a_hi = d_hi + (uint64_t)(ks * 16);          // 16 rows x 16 B, in 16 B units
                        const uint64_t a_lo = d_lo + (uint64_t)(ks * 16);
                        const uint32_t first = (trial == 0 && ks == 0) ? 0u : 1u;
                        for (int s = 0; s < P.n_tiles_n; ++s) {
                            const uint32_t idesc = (s == P.n_tiles_n - 1) ? idesc_last : idesc_full;
                            const uint32_t d_t = tmem_base + (uint32_t)(128 * s);
                            const uint64_t b_hi = a_hi + (uint64_t)s, b_lo = a_lo + (uint64_t)s;
                            tc::mma_f16_ss(d_t, a_hi, b_hi, idesc, first);
                            tc::mma_f16_ss(d_t, a_hi, b_lo, idesc, 1u);
                            tc::mma_f16_ss(d_t, a_lo, b_hi, idesc, 1u);
                        }
                    }
                    tc::mma_commit(&bar_empty[g]);                // operands of this trial are free again
                }
                tc::mma_commit(&bar_acc_full);                    // accumulators of this particle complete
            }
        }
        __syncwarp();
    } else {
        // ================= simulate groups ===================================================
        const int g = warp >> 2;
        const int tid_g = tid & (kTcGroupThreads - 1);
        const int lq = warp & 3;                                  // TMEM lane quarter of this warp
        float* xs = stag + (size_t)g * P.xs_stride;
        SimShared<kTcGroupThreads>* sh = &shs[g];
        const uint32_t n_g = (uint32_t)((P.n_trials - g + G - 1) / G);

        TcSink sink;
        sink.hi = opnd + (size_t)g * 2 * op_bytes;
        sink.lo = sink.hi + op_bytes;
        sink.sbo = P.sbo_bytes;
        sink.scale = P.tc_scale;
        sink.wait_bar = &bar_empty[g];
        sink.write_back = P.tail_lt > 0;

        // FFMA tail lags [tc_lags, n_lags)
        const int tile = P.tail_lt > 0 ? tid_g % P.tail_lt : 0;
        const int stream = P.tail_lt > 0 ? tid_g / P.tail_lt : 0;
        const bool tail_active = P.tail_lt > 0 && stream < P.tail_streams;
        const int tail_k0 = P.tail_start + tile * TK;          // multiple of 4: float4 window loads
        const int tile_begin = stream * P.tail_tiles_per_stream;
        const int tile_end = min(tile_begin + P.tail_tiles_per_stream, P.n_time_tiles);
        const int tail_row = P.tail_lt * TK;

        SimArgs sa;
        sa.n_t = P.n_t; sa.chunk = P.chunk;
        sa.inv_n = 1.0 / (double)P.n_t; sa.sqrt_nm1 = sqrtf((float)(P.n_t - 1));
        sa.k0 = P.k0; sa.k1 = P.k1; sa.c3_noise = P.c3_noise; sa.c3_init = P.c3_init;
        sa.mask_bits = P.mask_bits; sa.n_valid = P.n_valid; sa.mask_words = P.mask_words;

        const int n_cb = (P.used_cols + 31) / 32;                 // 32-column accumulator blocks
        float* zs = xs + lq * kTcZFloats;                         // this warp's skew scratch
        float* tab = xs + 4 * kTcZFloats;                         // this group's diagonal-sum table

        uint32_t it = 0;
        for (int64_t particle = blockIdx.x; particle < P.n_particles; particle += gridDim.x, ++it) {
            const double tau = P.theta[particle];
            double freq = 0.0, coeff = 1.0;
            if (OSC) {
                freq = P.theta[particle + P.n_particles];
                coeff = P.theta[particle + 2 * P.n_particles];
            }
            const OuConsts oc = make_ou_consts(tau, P.dt, P.update, OSC ? ITJL_KIND_OU_OSC : ITJL_KIND_OU, freq, coeff);
            sa.particle = (uint32_t)(P.particle_offset + particle);
            const size_t pt = (size_t)particle * P.n_trials;
            sa.u0 = P.u0 ? P.u0 + pt : nullptr;
            sa.noise = P.noise ? P.noise + pt * P.n_t : nullptr;
            sa.phases = P.phases ? P.phases + pt : nullptr;

            float acc[TK];
#pragma unroll
            for (int j = 0; j < TK; ++j) acc[j] = 0.f;
            SimCache sim_cache;

            uint32_t mine = 0;
            for (int trial = g; trial < P.n_trials; trial += G, ++mine) {
                const uint32_t k = it * n_g + mine;               // this group's trial counter
                sink.wait_parity = (k & 1u) ^ 1u;                 // MMAs of trial k-1 retired
                simulate_trial<kTcGroupThreads, MISSING, OSC, true>(sa, oc, trial, xs, sh, kScaleUnitNorm, 1.0f, 0.0f,
                                                                    sim_cache, tid_g, 1 + g, &sink);
                tc::fence_proxy_async_smem();                     // my operand stores -> async proxy
                tc::mbar_arrive(&bar_full[g]);
                if (P.tail_lt > 0) {
                    tc::named_bar_sync(1 + g, kTcGroupThreads);
                    if (tail_active && tile_begin < tile_end) acf_tiles(xs, acc, tail_k0, tile_begin, tile_end);
                    tc::named_bar_sync(1 + g, kTcGroupThreads);   // series is rewritten by the next trial
                }
            }

            // ---- epilogue: drain the accumulators, diagonal sums, distance ---------------------
            tc::mbar_wait(&bar_acc_full, it & 1u);
            tc::fence_after_thread_sync();
            for (int c = g, slot = 0; c < n_cb; c += G, ++slot) {
                float v[32];
                tc::tmem_ld_32x32(tmem_base + ((uint32_t)(32 * lq) << 16) + (uint32_t)(32 * c), v);
#pragma unroll
                for (int i = 0; i < 32; ++i) zs[i * 33 + lane] = v[i];
                __syncwarp();
                float s1 = 0.f, s2 = 0.f;
#pragma unroll
                for (int l = 0; l < 32; ++l) {
                    const float val = zs[((l + lane) & 31) * 33 + l];
                    if (l + lane < 32) s1 += val; else s2 += val;
                }
                float* t = tab + (slot * 4 + lq) * 64;
                t[lane] = s1; t[32 + lane] = s2;
                __syncwarp();
            }
            tc::fence_before_thread_sync();
            tc::mbar_arrive(&bar_acc_empty);                      // TMEM may be overwritten
            tc::named_bar_sync(kTcBarAll, n_sim);                 // tables complete

            // FFMA tail partials go below the tables (the skew scratch is dead after the barrier)
            if (tail_active) {
#pragma unroll
                for (int j = 0; j < TK; ++j) xs[stream * tail_row + tile * TK + j] = acc[j];
            }
            const double inv_trials = 1.0 / (double)P.n_trials;
            double local = 0.0;
            for (int k = tid; k < P.tc_lags; k += n_sim) {
                // lag k = 32 u + r collects block diagonals (c - lq = u, offset r) and (c - lq = u + 1, r - 32)
                const int u = k >> 5, r = k & 31;
                float s = 0.f;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int c1 = u + q;
                    if (c1 < n_cb) s += stag[(size_t)(c1 % G) * P.xs_stride + 4 * kTcZFloats + ((c1 / G) * 4 + q) * 64 + r];
                    const int c2 = u + 1 + q;
                    if (r > 0 && c2 < n_cb) s += stag[(size_t)(c2 % G) * P.xs_stride + 4 * kTcZFloats + ((c2 / G) * 4 + q) * 64 + 32 + r];
                }
                const double v = (double)(s * P.tc_unscale) * inv_trials;
                if (P.stats_out) P.stats_out[(size_t)particle * P.n_lags + k] = v;
                const double tgt = (double)P.target[k];
                const double e = (P.distance == ITJL_DIST_LINEAR) ? (v - tgt) : (log(v) - log(tgt));
                local += e * e;
            }
            if (P.tail_lt > 0) {
                tc::named_bar_sync(kTcBarAll, n_sim);             // partials of every group written
                for (int k = P.tc_lags + tid; k < P.n_lags; k += n_sim) {
                    const int kk = k - P.tail_start;
                    float s = 0.f;
                    for (int gg = 0; gg < G; ++gg)
                        for (int st = 0; st < P.tail_streams; ++st) s += stag[(size_t)gg * P.xs_stride + st * tail_row + kk];
                    const double v = (double)s * inv_trials;
                    if (P.stats_out) P.stats_out[(size_t)particle * P.n_lags + k] = v;
                    const double tgt = (double)P.target[k];
                    const double e = (P.distance == ITJL_DIST_LINEAR) ? (v - tgt) : (log(v) - log(tgt));
                    local += e * e;
                }
            }
            local = warp_sum(local);
            if (lane == 0) red[warp] = local;
            tc::named_bar_sync(kTcBarAll, n_sim);
            if (tid == 0) {
                double d = 0.0;
                for (int w = 0; w < 4 * G; ++w) d += red[w];
                P.dist_out[particle] = d / (double)P.n_lags;
            }
            // the scratch / tables / partials may have touched the zero pad behind the series
            for (int i = P.chunk * kTcGroupThreads + tid_g; i < P.xs_stride; i += kTcGroupThreads) xs[i] = 0.f;
            tc::named_bar_sync(kTcBarAll, n_sim);
        }
    }

    // ---- teardown -----------------------------------------------------------------------------
    tc::fence_before_thread_sync();
    __syncthreads();
    if (warp == 4 * G) tc::tmem_dealloc(tmem_base, 512);
}

// Geometry of the tensor-core kernel; returns 0 when the configuration is supported.
int abc_tc_geometry(int n_t, int n_lags, int n_trials, int max_smem, int sm_count, AbcTcGeometry* g) {
    if (n_t < 16 || n_lags < 1 || n_lags > n_t) return -1;
    const int max_tail = 60;                               // more FFMA tail lags than this: not worth it
    int used_cols = ((n_lags + 127) + 15) & ~15;
    if (used_cols > 512) used_cols = 512;
    if (used_cols < 32) used_cols = 32;
    g->used_cols = used_cols;
    g->tc_lags = (n_lags < used_cols - 127) ? n_lags : used_cols - 127;
    g->tail_start = g->tc_lags & ~3;                       // FFMA tail window starts 16 B aligned
    const int tail = (n_lags > g->tc_lags) ? n_lags - g->tail_start : 0;
    if (tail > max_tail) return -2;
    g->n_tiles_n = (used_cols + 127) / 128;
    g->n_last = used_cols - 128 * (g->n_tiles_n - 1);
    const int rows_data = (n_t + 127) / 128;
    g->ksteps = (rows_data + 15) / 16;
    g->rows_total = 16 * g->ksteps + g->n_tiles_n - 1;
    if ((g->rows_total & 1) == 0) g->rows_total++;           // odd row count: 16 B blocks spread over all banks
    g->sbo_bytes = g->rows_total * 16;
    int c = (n_t + kTcGroupThreads - 1) / kTcGroupThreads;
    g->chunk = (c + 7) & ~7;
    // FFMA tail
    g->n_time_tiles = (n_t + TT - 1) / TT;
    g->tail_lt = (tail + TK - 1) / TK;
    g->tail_streams = 0; g->tail_tiles_per_stream = 0;
    int need = g->chunk * kTcGroupThreads;
    if (g->tail_lt > 0) {
        g->tail_streams = kTcGroupThreads / g->tail_lt;
        if (g->tail_streams > g->n_time_tiles) g->tail_streams = g->n_time_tiles;
        g->tail_tiles_per_stream = (g->n_time_tiles + g->tail_streams - 1) / g->tail_streams;
        g->tail_streams = (g->n_time_tiles + g->tail_tiles_per_stream - 1) / g->tail_tiles_per_stream;
        const int reach = (g->n_time_tiles + 2) * TT + g->tail_start + g->tail_lt * TK + TT;   // furthest window read
        if (need < reach) need = reach;
        const int part = g->tail_streams * g->tail_lt * TK;
        if (need < part) need = part;
    }
    const int n_cb = (used_cols + 31) / 32;
    const size_t opnd = (size_t)2 * 16 * g->sbo_bytes;
    const size_t fixed = 1024;                              // alignment slack
    int groups = n_trials < kTcMaxGroups ? (n_trials > 0 ? n_trials : 1) : kTcMaxGroups;   // no idle groups
    for (; groups >= 1; --groups) {
        const int slots = (n_cb + groups - 1) / groups;
        int stride = need;
        const int scratch = 4 * kTcZFloats + slots * 4 * 64;
        if (stride < scratch) stride = scratch;
        stride = (stride + 31) & ~31;
        const size_t total = (size_t)groups * ((size_t)stride * sizeof(float) + opnd) + fixed;
        if ((int64_t)total <= (int64_t)max_smem - 6 * 1024) {   // static shared (barriers, scan) stays below 6 KB
            g->groups = groups; g->xs_stride = stride;
            g->opnd_offset = (int)(((size_t)groups * stride * sizeof(float) + 127) & ~(size_t)127);
            g->smem_bytes = (size_t)g->opnd_offset + (size_t)groups * opnd;
            break;
        }
    }
    if (groups < 1) return -3;
    int p2 = 1; while (p2 * p2 < n_t) p2 <<= 1;                 // power of two near sqrt(n_t): series ~ N(0,1)
    g->tc_scale = (float)p2;
    g->grid = sm_count;
    return 0;
}

cudaError_t abc_tc_launch(const AbcTcParams& P, bool missing, bool osc, size_t smem, int grid, cudaStream_t st) {
    void (*kern)(const AbcTcParams) =
        missing ? (osc ? k_abc_acf_tc<true, true> : k_abc_acf_tc<true, false>)
                : (osc ? k_abc_acf_tc<false, true> : k_abc_acf_tc<false, false>);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    if ((int64_t)grid > P.n_particles) grid = (int)P.n_particles;
    kern<<<grid, kTcGroupThreads * P.groups + 32, smem, st>>>(P);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------
// Probe: copy a caller-built fp16 shared-memory image, issue the listed tcgen05.mma ops
// (M = 128, one thread), drain the whole 128 x 512 fp32 accumulator block to `out`.
// Used by tests/test_gpu_tc_probe.py to check the MN-major no-swizzle layout, the row-shifted
// B operand and the instruction descriptor against exact integer Gram products.
__global__ void __launch_bounds__(128, 1) k_tc_probe(const uint4* __restrict__ image, int n_vec16,
                                                     const int4* __restrict__ ops, int n_ops, uint32_t lbo,
                                                     uint32_t sbo, uint32_t idesc, float* __restrict__ out) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    uint4* img = reinterpret_cast<uint4*>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    for (int i = tid; i < n_vec16; i += 128) img[i] = image[i];
    tc::fence_proxy_async_smem();
    if (tid == 0) { tc::mbar_init(&bar, 1); tc::fence_mbar_init(); }
    if (warp == 0) tc::tmem_alloc(&tmem_base_s, 512);
    tc::fence_before_thread_sync();
    __syncthreads();
    tc::fence_after_thread_sync();
    const uint32_t tmem_base = tmem_base_s;

    if (tid == 0) {
        const uint32_t base = tc::smem_u32(img);
        for (int i = 0; i < n_ops; ++i) {
            const int4 op = ops[i];
            tc::mma_f16_ss(tmem_base + (uint32_t)op.z, tc::make_smem_desc(base + (uint32_t)op.x, lbo, sbo),
                           tc::make_smem_desc(base + (uint32_t)op.y, lbo, sbo), idesc, (uint32_t)op.w);
        }
        tc::mma_commit(&bar);
    }
    tc::mbar_wait(&bar, 0);
    tc::fence_after_thread_sync();

    float v[32];
    for (int c = 0; c < 16; ++c) {
        tc::tmem_ld_32x32(tmem_base + ((uint32_t)(32 * warp) << 16) + (uint32_t)(32 * c), v);
#pragma unroll
        for (int i = 0; i < 32; ++i) out[(size_t)(32 * warp + lane) * 512 + 32 * c + i] = v[i];
    }
    tc::fence_before_thread_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_base, 512);
}

cudaError_t tc_probe_launch(const void* image_dev, int n_bytes, const int4* ops_dev, int n_ops, uint32_t lbo,
                            uint32_t sbo, uint32_t idesc, float* out_dev, cudaStream_t st) {
    cudaError_t e = cudaFuncSetAttribute(k_tc_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, n_bytes);
    if (e != cudaSuccess) return e;
    k_tc_probe<<<1, 128, n_bytes, st>>>((const uint4*)image_dev, n_bytes / 16, ops_dev, n_ops, lbo, sbo, idesc, out_dev);
    return cudaGetLastError();
}

}  // namespace itjl
