// Tensor-core (tcgen05 / TMEM) variant of the fused aABC ACF kernel, plus the self-test probe
// that pins the shared-memory descriptor conventions it relies on.
#include <stdlib.h>

#include "kernels.h"
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
// Lag sums as a Gram product.  With the centred unit-norm series cut into rows of 128 samples,
// X[i][a] = x'[128 i + a],
//     D_s[a][b] = sum_i X[i][a] * X[i+s][b]   =  sum over t = 128 i + a of x'_t x'_{t + 128 s + b - a},
// so with the "virtual column" v = 128 s + b every accumulator cell (a, v) holds products of lag
// v - a, and ac[l] = sum of the cells on the diagonal v - a = l.  Each D_s is a tcgen05.mma chain
// (K = the rows i of the trials) accumulating in fp32 in tensor memory.  Both operands are the
// SAME shared-memory buffer in the MN-major no-swizzle layout (K rows 16 B apart); the B operand
// of D_s simply starts s rows later.  x' is fed as an fp16 hi/lo pair; the number of product terms
// follows the particle size (abc_tc_geometry): hi*hi + hi*lo + lo*hi (~22-bit operands) below 10^5
// samples, hi*hi + hi*lo up to 2 10^5, hi*hi alone above - a dropped cross term is a sum of independent
// zero-mean fp16 rounding residuals, ~1e-4 / sqrt(n_trials n_t) in ACF units.  The missing-data models keep
// cross terms whatever their size: their masked samples share one value, hence one residual.
//
// Tensor-memory plans (128 lanes x 512 columns of fp32):
//   A  (n_lags <= 385)   tiles s = 0 .. 3, M = 128 (cell row a = TMEM lane), N = 128: lags 0 .. 384.
//   B  (n_lags <= 449)   the lower-left quadrant of D_0 (a >= 64, b < 64) only holds negative lags
//      (D_0 is symmetric), so D_0 is issued as two M = 64 MMAs - rows a < 64 x all 128 columns on
//      the lower 16 lanes of every 32-lane quarter, rows a >= 64 x columns 64 .. 127 on the upper
//      16 lanes (D address lane offset 16) - and the freed quadrant receives D_4 (rows a >= 64,
//      b < 64, v = 512 + b): lags up to 448 without touching the CUDA cores.
//   Up to 16 further lags (n_lags <= 464) are summed by an FFMA loop that reads the trial back from its
//   operand planes (the loop handles 60, but beyond one tile a second launch is faster).
//   Lag windows (n_lags <= 1537): the kernel is launched again with the B operand 3 w rows further down
//   (plan-A tiles, column v <-> lag 384 w + v - a) for the lags the launches before it have not scored;
//   longer windows run on the CUDA-core kernel k_abc_acf.
//
// Accumulator rounding.  tcgen05.mma aligns every product to the accumulator and truncates toward
// zero (measured, benchmarks/tc_probe2.py), so a long positive chain shrinks: 1.1e-5 relative at lags
// with ac ~ 1 over 50 trials x 5000 samples.  Because every trial adds the same sum of squares the
// chain grows linearly and the loss is, to first order, a known factor on the drained sum; the
// epilogue multiplies it back (abc_tc_geometry: bias_a, bias_b), leaving ~1e-6.  The accumulators are
// drained every `seg_trials` trials (<= ~2048 K rows, equal segments: once per particle at C5) and the
// segment sums are added to a fixed-point lag table.
//
// Warp roles (one persistent CTA per SM):
//   G simulate groups of 128 threads - one trial each at a time: Philox + blocked scan as in
//     simulate.cuh on named barriers.  A group owns two operand buffers (fp16 hi/lo planes) and
//     alternates between them; pass A parks its fp32 values inside the buffer and pass B converts
//     them in place, so there is no staging buffer and the group only ever waits for the MMAs of
//     the trial BEFORE the last one - drains and MMA latency are hidden behind a whole trial.
//   4 epilogue warps (one per TMEM lane quarter).  Each also issues a quarter of the MMA ops: it
//     waits for a buffer's "operands ready" mbarrier, issues its MMAs of that trial (one elected
//     lane, descriptors warp-uniform) and commits to the buffer's "operands free" mbarrier; after
//     the last trial of a segment it commits to "accumulators ready".  All four then tcgen05.ld
//     the accumulators, sum them along diagonals by lane rotation (shuffles), add the sums to the
//     lag table in shared memory, release the tensor memory, and after the particle's last segment
//     form the trial mean and the distance in fp64.
// (20 warps: the register file is allocated in units of 4 warps, 21 would cap at 80 registers.)
constexpr int kTcGroupThreads = 128;
constexpr int kTcMaxGroups = 4;
constexpr int kTcEpiThreads = 128;
constexpr int kTcMaxThreads = kTcGroupThreads * kTcMaxGroups + kTcEpiThreads;   // 20 warps x 96 registers
constexpr int kTcBarEpi = 8;                        // named barrier over the epilogue warps
constexpr int kTcLagTab = 512;                      // lag table entries (tensor-core lags <= 449)

// Diagonal sums of 16 accumulator columns (col0 .. col0 + 15 of a 32-column block) held one TMEM lane
// per thread: lane t fetches column i from lane (t + i) mod 32.  Plain blocks: a0 / b0 collect the
// un-wrapped cells (two chains for ILP), a1 / b1 the wrapped ones.  Mixed blocks (plan B): a = lower
// 16-lane strip, b = upper strip, index 0 / 1 = un-wrapped / wrapped.
// lag table entry += v, as 32-bit fixed point: integer atomics commute, so the result does not depend
// on the order of the four warps (and ATOMS.ADD is native; a 64-bit add would be a CAS loop).  `fix` is
// the power of two that maps the largest possible sum (lag 0: n_trials * scale^2) to 2^30, i.e. a
// rounding of ~5e-10 of full scale per add.
__device__ __forceinline__ void tc_lag_add(int* lagtab, int k, float v, int lo, int n, float fix, int* bad) {
    if (k >= lo && k < n) {
        const float f = v * fix;
        if (!(fabsf(f) < 2.0e9f)) *bad = 1;               // NaN / Inf: the particle's statistics become NaN
        else atomicAdd(&lagtab[k], __float2int_rn(f));
    }
}

template <bool MIXED, int COL0>
__device__ __forceinline__ void tc_drain_half(const uint32_t (&v)[16], int lane, bool mixed_block,
                                              float& a0, float& a1, float& b0, float& b1) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int src = lane + COL0 + i;
        const float x = __uint_as_float(__shfl_sync(0xffffffffu, v[i], src));
        const bool wrap = src >= 32;
        if (MIXED && mixed_block) {
            const bool upper = (src & 16) != 0;
            a0 += (!upper && !wrap) ? x : 0.f;
            a1 += (!upper && wrap) ? x : 0.f;
            b0 += (upper && !wrap) ? x : 0.f;
            b1 += (upper && wrap) ? x : 0.f;
        } else if (i & 1) {
            if (wrap) b1 += x; else b0 += x;
        } else {
            if (wrap) a1 += x; else a0 += x;
        }
    }
}

// ---- FFMA tail: the lags beyond the tensor-memory plans (449 .. 508), summed on the CUDA cores from
// the operand planes themselves (x' = hi + lo, exact in fp32).  Thread = (tile of 16 lags) x (time
// stream); a step multiplies 16 samples a[] with a 32-wide window b[] (256 FFMA per 8 LDS.128).
constexpr int kTailK = 16;                          // lags per thread
constexpr int kTailT = 16;                          // samples per step

// the 8 samples j .. j + 7 (j % 8 == 0) of a trial, as fp32
__device__ __forceinline__ void tc_load8(const unsigned char* hi, const unsigned char* lo, int sbo, int j, float* dst) {
    const int off = ((j & 127) >> 3) * sbo + (j >> 7) * 16;
    const uint4 h = *reinterpret_cast<const uint4*>(hi + off);
    const uint4 l = *reinterpret_cast<const uint4*>(lo + off);
    const uint32_t hw[4] = {h.x, h.y, h.z, h.w}, lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&hw[p]));
        const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&lw[p]));
        dst[2 * p] = a.x + b.x; dst[2 * p + 1] = a.y + b.y;
    }
}

// acc[j] += sum over steps [step_begin, step_end) of sum_i x[t0 + i] * x[t0 + i + k0 + j],  t0 = 16 step
__device__ __forceinline__ void tc_tail_steps(const unsigned char* hi, const unsigned char* lo, int sbo,
                                              float (&acc)[kTailK], int k0, int step_begin, int step_end) {
    float a[kTailT], b[2 * kTailT];
    int t0 = step_begin * kTailT;
    tc_load8(hi, lo, sbo, t0 + k0, b); tc_load8(hi, lo, sbo, t0 + k0 + 8, b + 8);
    for (int st = step_begin; st < step_end; st += 2, t0 += 2 * kTailT) {
        // phase 0: window = b[0 .. 31]
        tc_load8(hi, lo, sbo, t0 + k0 + 16, b + 16); tc_load8(hi, lo, sbo, t0 + k0 + 24, b + 24);
        tc_load8(hi, lo, sbo, t0, a); tc_load8(hi, lo, sbo, t0 + 8, a + 8);
#pragma unroll
        for (int i = 0; i < kTailT; ++i)
#pragma unroll
            for (int j = 0; j < kTailK; ++j) acc[j] = fmaf(a[i], b[i + j], acc[j]);
        if (st + 1 >= step_end) break;
        // phase 1: window = b[16 .. 31], b[0 .. 15]
        tc_load8(hi, lo, sbo, t0 + k0 + 32, b); tc_load8(hi, lo, sbo, t0 + k0 + 40, b + 8);
        tc_load8(hi, lo, sbo, t0 + 16, a); tc_load8(hi, lo, sbo, t0 + 24, a + 8);
#pragma unroll
        for (int i = 0; i < kTailT; ++i)
#pragma unroll
            for (int j = 0; j < kTailK; ++j) acc[j] = fmaf(a[i], b[(kTailT + i + j) % (2 * kTailT)], acc[j]);
    }
}

template <bool MISSING, bool OSC, bool MIXED, bool TAIL>
__global__ void __launch_bounds__(kTcMaxThreads, 1) k_abc_acf_tc(const __grid_constant__ AbcTcParams P) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bar_full[kTcMaxGroups][2];      // [group][buffer] operands ready
    __shared__ __align__(8) uint64_t bar_empty[kTcMaxGroups][2];     // [group][buffer] operands consumed
    __shared__ __align__(8) uint64_t bar_acc_full, bar_acc_empty, bar_tail_full, bar_tail_free;
    __shared__ float tailred[kTcMaxGroups][4][4 * kTailK];           // [group][warp][tail lag] per-particle sums
    __shared__ uint32_t tmem_slot;
    __shared__ int nonfinite;                                         // a drained sum of the current particle was NaN / Inf
    __shared__ SimShared<kTcGroupThreads> shs[kTcMaxGroups];
    __shared__ double red[4];

    const int G = P.groups;
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);      // provably warp-uniform: role branches stay convergent
    const int op_bytes = 16 * P.sbo_bytes;                       // one fp16 operand plane (hi or lo)
    unsigned char* opnd = smem_raw;                              // [G][2 buffers][hi, lo] planes
    int* lagtab = reinterpret_cast<int*>(smem_raw + (size_t)P.epi_offset);   // [kTcLagTab] fixed point

    // ---- one-time setup ------------------------------------------------------------------
    {
        uint4* z = reinterpret_cast<uint4*>(smem_raw);
        const int n16 = (int)(P.smem_bytes / 16);
        for (int i = tid; i < n16; i += blockDim.x) z[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    if (tid == 0) {
        nonfinite = 0;
        for (int g = 0; g < G; ++g)
            for (int b = 0; b < 2; ++b) { tc::mbar_init(&bar_full[g][b], kTcGroupThreads / 32); tc::mbar_init(&bar_empty[g][b], 4); }
        tc::mbar_init(&bar_acc_full, 4);
        tc::mbar_init(&bar_acc_empty, 4);
        tc::mbar_init(&bar_tail_full, (uint32_t)(4 * G));
        tc::mbar_init(&bar_tail_free, 1);
        tc::fence_mbar_init();
    }
    if (warp == 4 * G) tc::tmem_alloc(&tmem_slot, 512);
    tc::fence_proxy_async_smem();                                // zeroed operand rows -> async proxy
    tc::fence_before_thread_sync();
    __syncthreads();
    tc::fence_after_thread_sync();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, tmem_slot, 0);

    if (warp < 4 * G) {
        // ================= simulate groups ===================================================
        const int g = warp >> 2;
        const int tid_g = tid & (kTcGroupThreads - 1);
        SimShared<kTcGroupThreads>* sh = &shs[g];
        const uint32_t n_g = (uint32_t)((P.n_trials - g + G - 1) / G);
        unsigned char* my_opnd = opnd + (size_t)g * 4 * op_bytes;

        TcSink sink;
        sink.sbo = P.sbo_bytes;
        sink.scale = P.tc_scale;
        sink.write_lo = TAIL || P.n_terms > 1;            // the lo plane is an MMA operand only with cross terms (and the FFMA tail reads it)

        SimArgs sa;
        sa.n_t = P.n_t; sa.chunk = P.chunk;
        sa.inv_n = 1.0 / (double)P.n_t; sa.sqrt_nm1 = sqrtf((float)(P.n_t - 1));
        sa.keys = &P.keys; sa.c3_noise = P.c3_noise; sa.c3_init = P.c3_init;
        sa.mask_bits = P.mask_bits; sa.n_valid = P.n_valid; sa.mask_words = P.mask_words;
        sa.miss_mean_ratio = P.miss_mean_ratio;

        uint32_t it = 0;
        for (int64_t particle = blockIdx.x; particle < P.n_particles; particle += gridDim.x, ++it) {
            const double tau = P.theta[particle];
            double freq = 0.0, coeff = 1.0;
            if (OSC) {
                freq = P.theta[particle + P.theta_ld];
                coeff = P.theta[particle + 2 * P.theta_ld];
            }
            const OuConsts oc = make_ou_consts(tau, P.dt, P.update, OSC ? ITJL_KIND_OU_OSC : ITJL_KIND_OU, freq, coeff, P.n_t, true, !MISSING);
            sa.particle = (uint32_t)(P.particle_offset + particle);
            const size_t pt = (size_t)particle * P.n_trials;
            sa.u0 = P.u0 ? P.u0 + pt : nullptr;
            sa.noise = P.noise ? P.noise + pt * P.n_t : nullptr;
            sa.phases = P.phases ? P.phases + pt : nullptr;
            SimCache sim_cache;
            float tacc[TAIL ? kTailK : 1];
#pragma unroll
            for (int j = 0; j < (TAIL ? kTailK : 1); ++j) tacc[j] = 0.f;

            uint32_t mine = 0;
            for (int trial = g; trial < P.n_trials; trial += G, ++mine) {
                const uint32_t k = it * n_g + mine;               // this group's trial counter
                const uint32_t buf = k & 1u;                      // operand buffers alternate
                sink.hi = my_opnd + (size_t)buf * 2 * op_bytes;
                sink.lo = sink.hi + op_bytes;
                sink.wait_bar = &bar_empty[g][buf];
                sink.wait_parity = ((k >> 1) & 1u) ^ 1u;          // MMAs of trial k - 2 (same buffer) retired
                if (!(P.debug & 2) || it == 0)                    // debug bit 1: time the tensor-core chain alone
                    simulate_trial<kTcGroupThreads, MISSING, OSC, true>(sa, oc, trial, nullptr, sh, kScaleUnitNorm, 1.0f, 0.0f,
                                                                        sim_cache, tid_g, 1 + g, &sink);
                else
                    tc::mbar_wait(sink.wait_bar, sink.wait_parity);
                tc::fence_proxy_async_smem();                     // my operand stores -> async proxy
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&bar_full[g][buf]);    // one arrival per warp
                if (TAIL) {
                    // lags beyond the tensor-memory plan: read the trial back from its operand planes (the
                    // MMAs only read them too; the buffer is rewritten two trials from now)
                    tc::named_bar_sync(1 + g, kTcGroupThreads);
                    const int tile = tid_g & (P.tail_lt - 1), stream = tid_g / P.tail_lt;
                    const int s_begin = stream * P.tail_steps_per_stream;
                    const int s_end = min(s_begin + P.tail_steps_per_stream, P.tail_steps);
                    if (s_begin < s_end)
                        tc_tail_steps(sink.hi, sink.lo, P.sbo_bytes, reinterpret_cast<float (&)[kTailK]>(tacc),
                                      P.tail_start + tile * kTailK, s_begin, s_end);
                }
            }
            if (TAIL) {
                // reduce over the time streams of a warp (lanes with equal lane % tail_lt), hand the sums over
#pragma unroll
                for (int j = 0; j < (TAIL ? kTailK : 1); ++j) {
                    float v = tacc[j];
                    for (int o = 16; o >= P.tail_lt; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                    tacc[j] = v;
                }
                tc::mbar_wait(&bar_tail_free, (it & 1u) ^ 1u);    // previous particle's sums consumed
                if (lane < P.tail_lt) {
#pragma unroll
                    for (int j = 0; j < (TAIL ? kTailK : 1); ++j) tailred[g][warp & 3][lane * kTailK + j] = tacc[j];
                }
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&bar_tail_full);
            }
        }
    } else {
        // ================= epilogue warps (each also issues a quarter of the MMA ops) ============
        const int q = warp & 3;                                   // TMEM lane quarter this warp may read
        const int e_tid = tid - 4 * G * 32;                       // 0 .. 127
        const int e_warp = warp - 4 * G;                          // issues the MMA ops o = e_warp, e_warp + 4, ..
        const uint32_t opnd_addr = tc::smem_u32(opnd);
        const int n_cb = (P.used_cols + 31) / 32;                 // 32-column accumulator blocks
        const double inv_trials = 1.0 / (double)P.n_trials;

        uint32_t it = 0, seg_it = 0;
        for (int64_t particle = blockIdx.x; particle < P.n_particles; particle += gridDim.x, ++it) {
            for (int seg0 = 0; seg0 < P.n_trials; seg0 += P.seg_trials, ++seg_it) {
                {
                    // ---- issue this warp's share of the segment's MMAs as the groups deliver their trials
                    // (whole warp, convergent: one elected lane issues, see tc05.cuh) ----------------
                    tc::mbar_wait_warp(&bar_acc_empty, (seg_it & 1u) ^ 1u);   // previous segment drained
                    tc::fence_after_thread_sync();
                    const int seg1 = min(seg0 + P.seg_trials, P.n_trials);
#ifdef ITJL_V_HOIST
                    // this warp's ops and the per-group trial counters, kept in registers across the segment: the
                    // issue loop is 8 % of the kernel's instructions when it re-derives them per MMA
                    uint32_t my_a[2], my_b[2], my_t[2], my_i[2];
                    int n_my = 0;
                    for (int o = e_warp; o < P.n_ops && n_my < 2; o += 4, ++n_my) {
                        my_a[n_my] = P.op_a_off[o]; my_b[n_my] = P.op_b_off[o];
                        my_t[n_my] = tmem_base + P.op_taddr[o]; my_i[n_my] = P.op_idesc[o];
                    }
                    int g = seg0 % G;
                    uint32_t kq = (uint32_t)(seg0 / G);
                    for (int trial = seg0; trial < seg1; ++trial) {
                        const uint32_t n_g = (uint32_t)((P.n_trials - g + G - 1) / G);
                        const uint32_t k = it * n_g + kq;
                        const uint32_t buf = k & 1u;
                        tc::mbar_wait_warp(&bar_full[g][buf], (k >> 1) & 1u);
                        tc::fence_after_thread_sync();
                        const uint32_t hi_addr = opnd_addr + (uint32_t)((g * 2 + (int)buf) * 2 * op_bytes);
                        const uint64_t d_hi = tc::make_smem_desc(hi_addr, 128u, (uint32_t)P.sbo_bytes);
                        const uint64_t d_lo = tc::make_smem_desc(hi_addr + (uint32_t)op_bytes, 128u, (uint32_t)P.sbo_bytes);
                        const int n_ks = (P.debug & 1) ? 0 : P.ksteps;
                        for (int ks = 0; ks < n_ks; ++ks) {
                            const uint32_t first = (trial == seg0 && ks == 0) ? 0u : 1u;
#pragma unroll
                            for (int m = 0; m < 2; ++m) {
                                if (m < n_my) {
                                    const uint64_t a_off = (uint64_t)(ks * 16 + my_a[m]);
                                    const uint64_t b_off = (uint64_t)(ks * 16 + my_b[m]);
                                    tc::mma_f16_ss_elect(my_t[m], d_hi + a_off, d_hi + b_off, my_i[m], first);
                                    if (P.n_terms >= 2) tc::mma_f16_ss_elect(my_t[m], d_hi + a_off, d_lo + b_off, my_i[m], 1u);
                                    if (P.n_terms == 3) tc::mma_f16_ss_elect(my_t[m], d_lo + a_off, d_hi + b_off, my_i[m], 1u);
                                }
                            }
                        }
                        tc::mma_commit_elect(&bar_empty[g][buf]); // this warp's reads of the buffer have retired
                        if (++g == G) { g = 0; ++kq; }
                    }
#else
                    for (int trial = seg0; trial < seg1; ++trial) {
                        const int g = trial % G;
                        const uint32_t n_g = (uint32_t)((P.n_trials - g + G - 1) / G);
                        const uint32_t k = it * n_g + (uint32_t)(trial / G);
                        const uint32_t buf = k & 1u;
                        tc::mbar_wait_warp(&bar_full[g][buf], (k >> 1) & 1u);
                        tc::fence_after_thread_sync();
                        const uint32_t hi_addr = opnd_addr + (uint32_t)((g * 2 + (int)buf) * 2 * op_bytes);
                        const uint64_t d_hi = tc::make_smem_desc(hi_addr, 128u, (uint32_t)P.sbo_bytes);
                        const uint64_t d_lo = tc::make_smem_desc(hi_addr + (uint32_t)op_bytes, 128u, (uint32_t)P.sbo_bytes);
                        for (int ks = 0; ks < ((P.debug & 1) ? 0 : P.ksteps); ++ks) {      // debug bit 0: no MMAs (simulate alone)
                            const uint32_t first = (trial == seg0 && ks == 0) ? 0u : 1u;
                            for (int o = e_warp; o < P.n_ops; o += 4) {
                                // descriptor start addresses are in 16 B units: K rows are 16 B apart
                                const uint64_t a_off = (uint64_t)(ks * 16 + P.op_a_off[o]);
                                const uint64_t b_off = (uint64_t)(ks * 16 + P.op_b_off[o]);
                                const uint32_t d_t = tmem_base + P.op_taddr[o];
                                const uint32_t idesc = P.op_idesc[o];
                                tc::mma_f16_ss_elect(d_t, d_hi + a_off, d_hi + b_off, idesc, first);
                                if (P.n_terms >= 2) tc::mma_f16_ss_elect(d_t, d_hi + a_off, d_lo + b_off, idesc, 1u);
                                if (P.n_terms == 3) tc::mma_f16_ss_elect(d_t, d_lo + a_off, d_hi + b_off, idesc, 1u);
                            }
                        }
                        tc::mma_commit_elect(&bar_empty[g][buf]); // this warp's reads of the buffer have retired
                    }
#endif
                    tc::mma_commit_elect(&bar_acc_full);          // this warp's accumulators of the segment complete
                }
                __syncwarp();
                tc::mbar_wait_warp(&bar_acc_full, seg_it & 1u);
                tc::fence_after_thread_sync();
                // ---- drain: diagonal sums by lane rotation.  Lane t receives column i of TMEM lane
                // (t + i) mod 32, i.e. the cell of diagonal (column - row) = -t, or 32 - t once t + i wraps.
                // 16 columns at a time, the next tcgen05.ld in flight under the shuffles of the current one.
                // The loop over blocks is NOT unrolled (the unrolled drain was instruction-fetch bound:
                // 7 us per drain) and the sums go straight into the lag table as 32-bit fixed point
                // (native 32-bit integer atomics: the order of the four warps does not matter - deterministic) ----
                if (!(P.debug & 8)) {                             // debug bit 3: segment handshake without the drain
                    const uint32_t t_q = tmem_base + ((uint32_t)(32 * q) << 16);
                    uint32_t va[16], vb[16];
                    tc::tmem_ld_32x16_issue(t_q, va);
#pragma unroll 1
                    for (int c = 0; c < n_cb; ++c) {
                        float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;      // mixed: lower no-wrap / wrap, upper no-wrap / wrap
                        const bool mixed_block = MIXED && c < 4;
                        tc::tmem_ld_wait(va);
                        tc::tmem_ld_32x16_issue(t_q + (uint32_t)(32 * c + 16), vb);
                        tc_drain_half<MIXED, 0>(va, lane, mixed_block, a0, a1, b0, b1);
                        tc::tmem_ld_wait(vb);
                        if (c + 1 < n_cb) tc::tmem_ld_32x16_issue(t_q + (uint32_t)(32 * (c + 1)), va);
                        tc_drain_half<MIXED, 16>(vb, lane, mixed_block, a0, a1, b0, b1);
                        if (mixed_block) {
                            // lower strip: lag = 32 c - 16 q - lane (+ 32 after the wrap);
                            // upper strip: rows a = 48 + 16 q + lane, columns of D_4 (c < 2) start at v = 512
                            const int kl = 32 * c - 16 * q - lane;
                            const int ku = 32 * c - 48 - 16 * q - lane + (c < 2 ? 512 : 0);
                            tc_lag_add(lagtab, kl, a0, P.k_lo, P.tc_lags, P.tc_fix, &nonfinite); tc_lag_add(lagtab, kl + 32, a1, P.k_lo, P.tc_lags, P.tc_fix, &nonfinite);
                            tc_lag_add(lagtab, ku, b0, P.k_lo, P.tc_lags, P.tc_fix, &nonfinite); tc_lag_add(lagtab, ku + 32, b1, P.k_lo, P.tc_lags, P.tc_fix, &nonfinite);
                        } else {
                            const int k = 32 * (c - q) - lane;
                            tc_lag_add(lagtab, k, a0 + b0, P.k_lo, P.tc_lags, P.tc_fix, &nonfinite); tc_lag_add(lagtab, k + 32, a1 + b1, P.k_lo, P.tc_lags, P.tc_fix, &nonfinite);
                        }
                    }
                }
                tc::fence_before_thread_sync();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&bar_acc_empty);   // this quarter of the tensor memory is free
            }

            tc::named_bar_sync(kTcBarEpi, kTcEpiThreads);           // every quarter has added its last segment

            // ---- trial mean, distance ---------------------------------------------------------------
            if (TAIL) tc::mbar_wait(&bar_tail_full, it & 1u);
            const bool bad_particle = nonfinite != 0;
            double local = 0.0;
            // this launch's lag window: relative lags [k_lo, k_hi) of the table, absolute lag = lag_base + k
            const int k_hi = TAIL ? P.n_lags : min(P.tc_lags, P.n_lags - P.lag_base);
            for (int k = P.k_lo + e_tid; k < k_hi; k += kTcEpiThreads) {
                float sk;
                if (bad_particle) {
                    sk = __int_as_float(0x7fc00000);
                } else if (!TAIL || k < P.tc_lags) {
                    sk = (float)lagtab[k] * P.tc_unfix;
                    // undo the accumulator's round-toward-zero to first order (abc_tc_geometry)
                    const float acn = fminf(fabsf(sk) * P.tc_unscale * (float)inv_trials, 1.0f);
                    sk = fmaf(sk, fmaf(P.bias_b * 0.63661977f, asinf(acn), P.bias_a), sk);
                } else {
                    sk = 0.f;
                    for (int gg = 0; gg < G; ++gg)
                        for (int w = 0; w < 4; ++w) sk += tailred[gg][w][k - P.tail_start];
                }
                double v = __dmul_rn((double)(sk * P.tc_unscale), inv_trials);         // no FMA contraction with the
                                                                                      // subtraction below: same bits with and without stats_out
                const int ka = P.lag_base + k;
                if (ka == 0 && isfinite(v)) v = 1.0;              // ac[0] = acov[0] / acov[0] is exactly 1 in the reference (summary.jl:58, 479)
                if (P.stats_out) P.stats_out[(size_t)particle * P.n_lags + ka] = v;
                const double tgt = (double)P.target[ka];
                const double e = (P.distance == ITJL_DIST_LINEAR) ? (v - tgt) : (log(v) - log(tgt));
                local += e * e;
            }
            local = warp_sum(local);
            if (lane == 0) red[q] = local;
            tc::named_bar_sync(kTcBarEpi, kTcEpiThreads);
            if (e_tid == 0) {
                const double part = (red[0] + red[1] + red[2] + red[3]) / (double)P.n_lags;
                P.dist_out[particle] = P.accumulate ? P.dist_out[particle] + part : part;   // second window: launches are stream ordered
                if (TAIL) tc::mbar_arrive(&bar_tail_free);
            }
            for (int k = e_tid; k < kTcLagTab; k += kTcEpiThreads) lagtab[k] = 0;
            if (e_tid == 0) nonfinite = 0;
            tc::named_bar_sync(kTcBarEpi, kTcEpiThreads);
        }
    }

    // ---- teardown -----------------------------------------------------------------------------
    tc::fence_before_thread_sync();
    __syncthreads();
    if (warp == 4 * G) tc::tmem_dealloc(tmem_base, 512);
}

// Bank conflicts (excess 128-byte wavefronts) of the simulate groups' 16-byte operand accesses for a
// given MN-block stride: lane l of quarter `quarter` of warp w owns chunk 16 (l % 8) + 4 w + quarter
// (simulate.cuh) and touches sample chunk_idx * chunk + q.  (Zero for 40-sample chunks whatever the stride.)
static int tc_store_conflicts(int sbo_bytes, int chunk) {
    int excess = 0;
    for (int q = 0; q < chunk; q += 8) {
        for (int w = 0; w < kTcGroupThreads / 32; ++w) {
            for (int quarter = 0; quarter < 4; ++quarter) {           // 8 lanes x 16 B per wavefront
                int used[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                for (int l = 0; l < 8; ++l) {
                    const int j = (16 * l + 4 * w + quarter) * chunk + q;
                    const int off = ((j & 127) >> 3) * sbo_bytes + (j >> 7) * 16;
                    used[(off >> 4) & 7]++;
                }
                int mx = 0;
                for (int b = 0; b < 8; ++b) mx = used[b] > mx ? used[b] : mx;
                excess += mx - 1;
            }
        }
    }
    return excess;
}

// Geometry of the tensor-core kernel; returns 0 when the configuration is supported.
// ---- error model of the tensor-core lag sums -----------------------------------------------------------------
// Bound on what the tensor-core lag sums ADD to the error of one trial-mean ACF value (values in [-1, 1]) over
// the fp32 FFMA sums of the same simulated series, for a particle of N = n_trials * n_t samples.  (Both fused
// kernels share the fp32 simulation, centring and conversion: <= 1.3e-6 against the fp64 reference from
// tau = T / 100 to tau = 50 T, 1 to 200 trials - measured, tests/test_gpu_tc.py; the total stays well within the
// 1e-5 of the contract when this bound is within the default 5e-6.)
//   E_a  dropped cross terms of the fp16 hi/lo split.  x' = hi + lo with |lo| <= 2^-12 |x'|; a dropped term
//        sum lo_t x'_{t+l} is a sum of N independent zero-mean residuals, in units of sum x'^2 = 1 per trial:
//        rms kA sqrt(dropped terms) / sqrt(N), kA = 3.2e-4 (calibrated at C5, 1184 particles x 414 lags:
//        rms 9.0e-7 with both cross terms dropped, maximum 4.7e-6).
//   E_c  fp32 -> fp16 x 2 conversion and the fixed-point lag table: rms 2.5e-7.
//   E_b  what the first-order correction of the accumulator's round-toward-zero leaves: <= 6e-10 per 128-sample
//        row of a segment (1.2e-6 at 2000 rows, measured).
//   E_m  missing-data models: the masked samples all carry ONE value v, so their rounding residual is common
//        and every masked-masked product is off by the same relative 2^-11 (hi*hi only) or 2^-12 (one cross term
//        kept); they contribute mask_frac * rho^2 to ac, rho^2 = n v^2 = (mean offset / std)^2 ~ 0.05 for the plain
//        model, + (data_mean / data_sd)^2 for the oscillation model (whose masked samples sit at -data_mean).
// bound = 4 sqrt(E_a^2 + E_c^2) + E_b + E_m   (4 sigma: the maximum over the ~1e5 .. 1e6 values of a test run).
double abc_tc_predicted_error(int64_t n_t, int64_t n_trials, int n_terms, int seg_rows, double mask_frac, double mean_ratio) {
    const double N = (double)n_t * (double)(n_trials > 0 ? n_trials : 1);
    const int dropped = n_terms >= 3 ? 0 : 3 - n_terms;
    const double e_a = 3.2e-4 * sqrt((double)dropped) / sqrt(N);
    const double e_c = 2.5e-7;
    const double e_b = 6e-10 * (double)seg_rows;
    const double rho2 = 0.05 + mean_ratio * mean_ratio;
    const double e_m = mask_frac * rho2 * (dropped == 2 ? 4.9e-4 : dropped == 1 ? 2.45e-4 : 1e-7);
    return 4.0 * sqrt(e_a * e_a + e_c * e_c) + e_b + e_m;
}

int abc_tc_geometry(int n_t, int n_lags, int n_trials, int max_smem, int sm_count, AbcTcGeometry* g,
                    const itjl_tuning* tun, double mask_frac, double mean_ratio, bool missing) {
    if (n_t < 16 || n_lags < 1 || n_lags > n_t) return -1;
    // FFMA tail lags beyond the tensor-memory plans.  The tail code handles up to 60 (four tiles of 16), but
    // only one tile pays: at C5 shape 2.8e11 samples/s with a 16-lag tail against 2.4e11 for a second lag
    // window, 2.3e11 / 1.7e11 with two / four tail tiles
    int max_tail = 15;
    if (tun && tun->tc_tail_max >= 0 && tun->tc_tail_max <= 60) max_tail = tun->tc_tail_max;   // tuning only
    // n_lags 465 .. 1537: up to four launches (lag windows), each with the B operand 3 more rows down
    // window w >= 1 moves the B operand 3 w rows down and covers the lags [384 w, 384 w + 385)
    g->n_windows = n_lags > 449 + max_tail ? 1 + (n_lags - 385 + 383) / 384 : 1;
    g->win1_s0 = 3; g->win1_used_cols = 0;
    if (g->n_windows > kTcMaxWindows) return -2;
    const int rows_data = (n_t + 127) / 128;
    g->ksteps = (rows_data + 15) / 16;
    int max_shift;
    g->n_ops = 0;
    auto add_op = [&](int a_blk, int b_row, int b_blk, int m, int n, int lane, int col) {
        const int o = g->n_ops++;
        g->op_a_blk[o] = a_blk; g->op_b_row[o] = b_row; g->op_b_blk[o] = b_blk;
        g->op_idesc[o] = tc::make_idesc_f16_mn(m, n);
        g->op_taddr[o] = ((uint32_t)lane << 16) | (uint32_t)col;
    };
    if (n_lags <= 385) {
        // plan A
        int used_cols = ((n_lags + 127) + 15) & ~15;
        if (used_cols < 32) used_cols = 32;
        g->used_cols = used_cols;
        g->mixed = 0;
        const int n_tiles = (used_cols + 127) / 128;
        for (int s = 0; s < n_tiles; ++s) add_op(0, s, 0, 128, s == n_tiles - 1 ? used_cols - 128 * s : 128, 0, 128 * s);
        max_shift = n_tiles - 1;
    } else {
        // plan B
        int n4 = ((n_lags - 385) + 15) & ~15;
        if (n4 > 64) n4 = 64;
        g->used_cols = 512;
        g->mixed = 1;
        add_op(0, 0, 0, 64, 128, 0, 0);                    // D_0 rows a < 64
        add_op(8, 0, 8, 64, 64, 16, 64);                   // D_0 rows a >= 64, columns 64 .. 127
        for (int s = 1; s < 4; ++s) add_op(0, s, 0, 128, 128, 0, 128 * s);
        add_op(8, 4, 0, 64, n4, 16, 0);                    // D_4 rows a >= 64, columns < n4
        max_shift = 4;
    }
    g->tc_lags = n_lags < 449 ? n_lags : 449;
    if (n_lags <= 385) g->tc_lags = n_lags;
    if (g->n_windows >= 2) {
        // accumulator columns of the LAST window (the ones before it use all 512)
        const int base = 128 * g->win1_s0 * (g->n_windows - 1);
        g->win1_used_cols = (((n_lags - base) + 127) + 15) & ~15;      // <= 512
        max_shift = g->win1_s0 * (g->n_windows - 1) + (g->win1_used_cols + 127) / 128 - 1;
    }
    // FFMA tail: tiles of 16 lags from tail_start (a multiple of 8: 16-byte groups of the planes)
    g->tail_start = g->tc_lags & ~7;
    const int tail = (n_lags > g->tc_lags && g->n_windows == 1) ? n_lags - g->tail_start : 0;
    g->tail_lt = tail == 0 ? 0 : (tail <= kTailK ? 1 : (tail <= 2 * kTailK ? 2 : 4));
    if (tail > 4 * kTailK) return -2;
    g->tail_steps = (n_t + kTailT - 1) / kTailT;
    g->tail_steps_per_stream = 0;
    if (g->tail_lt > 0) {
        const int streams = kTcGroupThreads / g->tail_lt;
        g->tail_steps_per_stream = (g->tail_steps + streams - 1) / streams;
    }
    int c = (n_t + kTcGroupThreads - 1) / kTcGroupThreads;
    g->chunk = (c + 7) & ~7;
    if (g->chunk > kQtab) return -3;
    // K rows per plane: data + shifted-B reach; among the strides that keep the group count, take the
    // one with the fewest bank conflicts
    int rows_min = 16 * g->ksteps + max_shift;
    if (g->tail_lt > 0) {
        // furthest sample the tail window reads (zero rows behind the data)
        const int reach = (g->tail_steps + 2) * kTailT + g->tail_start + g->tail_lt * kTailK + 2 * kTailT;
        const int rows_tail = (reach + 127) / 128;
        if (rows_tail > rows_min) rows_min = rows_tail;
    }
    const size_t epi = (size_t)kTcLagTab * sizeof(int);
    int want_groups = n_trials < kTcMaxGroups ? (n_trials > 0 ? n_trials : 1) : kTcMaxGroups;   // no idle groups
    if (tun && tun->tc_groups >= 1 && tun->tc_groups < want_groups) want_groups = tun->tc_groups;   // tuning only
    auto fit = [&](int rows, int groups, size_t* epi_off, size_t* total) {
        const size_t buffer = (size_t)2 * 16 * rows * 16;   // hi + lo planes of one trial
        *epi_off = ((size_t)groups * 2 * buffer + 127) & ~(size_t)127;
        *total = (*epi_off + epi + 15) & ~(size_t)15;
        // static shared memory: barriers, the groups' scan tables and carry-power tables (~6.6 KB), tail sums (+4 KB)
        return (int64_t)*total + 1024 <= (int64_t)max_smem - (g->tail_lt > 0 ? 11 : 7) * 1024;
    };
    int groups = want_groups;
    size_t epi_off = 0, total = 0;
    while (groups >= 1 && !fit(rows_min, groups, &epi_off, &total)) --groups;
    if (groups < 1) return -3;
    int best_rows = rows_min, best = 1 << 30;
    for (int r = rows_min; r < rows_min + 8 && fit(r, groups, &epi_off, &total); ++r) {
        const int cf = tc_store_conflicts(r * 16, g->chunk);
        if (cf < best) { best = cf; best_rows = r; }
    }
    fit(best_rows, groups, &epi_off, &total);
    g->rows_total = best_rows;
    g->sbo_bytes = best_rows * 16;
    g->groups = groups; g->epi_offset = (int)epi_off; g->smem_bytes = total;
    // trials per accumulator segment: up to ~2048 K rows of samples (the whole particle at C5), split evenly
    // over the particle's trials.  The accumulator's round-toward-zero costs ~5.5e-9 per K row at lags
    // with ac ~ 1 (measured at C5: 2.5e-6 / 5.3e-6 / 1.1e-5 for 520 / 1000 / 2000 rows); the epilogue undoes
    // it to first order (bias_a / bias_b below), which leaves < 1.2e-6 at 2000 rows
    const int rows_trial = 16 * g->ksteps;
    int seg = 2048 / rows_data;
    if (seg < 1) seg = 1;
    if (n_trials > 0) {                                    // equal segments: 100 trials x 40 rows -> 50, 50
        const int n_seg = (n_trials + seg - 1) / seg;
        seg = (n_trials + n_seg - 1) / n_seg;
    }
    (void)rows_trial;
    // Product terms of the fp16 hi/lo split x' = hi + lo: hi*hi (+ hi*lo (+ lo*hi)), and trials per segment: the
    // cheapest combination whose predicted error (abc_tc_predicted_error) is within the bound.  At C5 (2.5e5
    // samples per particle) that is hi*hi alone and one segment; the missing-data models keep the cross terms
    // they need through E_m (their masked samples share one rounding residual).  Fewer trials per segment
    // (more drains) are tried when even three terms miss the bound; when nothing fits the caller falls back to
    // the CUDA-core kernel.
    const double bound = (tun && tun->tc_error_bound > 0.0) ? tun->tc_error_bound : 5e-6;
    const int forced_terms = (tun && tun->tc_terms >= 1 && tun->tc_terms <= 3) ? tun->tc_terms : 0;
    const int forced_seg = (tun && tun->tc_seg_trials > 0) ? tun->tc_seg_trials : 0;
    const bool forced_kernel = tun && tun->abc_tc == 1;
    const int min_terms = missing ? 2 : 1;             // see E_m: hi*hi alone is never worth it with a mask
    auto even_seg = [&](int sg) {                      // equal segments: 100 trials with <= 51 per segment -> 50, 50
        if (n_trials <= 0) return sg;
        const int n_seg = (n_trials + sg - 1) / sg;
        return (n_trials + n_seg - 1) / n_seg;
    };
    auto predicted = [&](int terms, int sg) {
        return abc_tc_predicted_error(n_t, n_trials, terms, even_seg(sg) * rows_data, mask_frac, mean_ratio);
    };
    bool within = false;
    int pick_terms = forced_terms ? forced_terms : 3, pick_seg = forced_seg ? forced_seg : 1;   // best effort
    for (int sg = forced_seg ? forced_seg : seg; !within; sg = (sg + 1) / 2) {
        for (int terms = forced_terms ? forced_terms : min_terms; terms <= 3 && !within; ++terms) {
            if (predicted(terms, sg) <= bound) { within = true; pick_terms = terms; pick_seg = sg; }
            if (forced_terms) break;
        }
        if (forced_seg || sg <= 1) break;
    }
    seg = even_seg(pick_seg);
    g->seg_trials = seg;
    g->n_terms = pick_terms;
    g->pred_err = predicted(pick_terms, pick_seg);
    const bool error_ok = within || forced_kernel;
    // The tensor-core accumulator rounds toward zero (tests/test_gpu_tc.py, probe): a cell that grows
    // linearly to c over R accumulate steps loses sum_i ulp(c i / R) / 2 = R 2^-24 phi(c) c with
    // phi in [1/3, 3/8] whatever the binade of c (measured: 2.1e-8 = 0.35 2^-24 per hi*hi tcgen05.mma, 1.6e-8
    // per cross-term MMA), and each
    // product is truncated on alignment to the accumulator (measured 4.7e-9 per K row, times the sign
    // imbalance (2/pi) asin(ac) of the products; constants fitted at 2000-row segments, 1184 particles,
    // benchmarks/tc_error.py).  Every trial contributes the same sum of squares, so the
    // growth is linear and the loss is a factor on the drained sums:
    //     S_true = S (1 + bias_a + bias_b (2/pi) asin|ac|),
    // bias_a = (2.1e-8 + 1.6e-8 (n_terms - 1)) x (K steps per segment), bias_b = 4.7e-9 x (K rows per segment).
    {
        const int n_seg = n_trials > 0 ? (n_trials + seg - 1) / seg : 1;
        const double trials_seg = n_trials > 0 ? (double)n_trials / n_seg : 1.0;
        g->bias_a = (float)((2.1e-8 + 1.6e-8 * (g->n_terms - 1)) * trials_seg * g->ksteps);
        g->bias_b = (float)(4.7e-9 * trials_seg * rows_data);
        if (tun && tun->tc_bias == 0) g->bias_a = g->bias_b = 0.f;
    }
    int p2 = 1; while (p2 * p2 < n_t) p2 <<= 1;                 // power of two near sqrt(n_t): series ~ N(0,1)
    g->tc_scale = (float)p2;
    // fixed-point unit of the lag table: n_trials * scale^2 * fix <= 2^30
    double fix = 1073741824.0 / ((double)(n_trials > 0 ? n_trials : 1) * (double)p2 * (double)p2);
    int e2 = 0; while (fix >= 2.0 && e2 < 30) { fix *= 0.5; ++e2; }
    g->tc_fix = (float)(1u << e2);
    g->grid = sm_count;
    return error_ok ? 0 : -4;
}

template <bool MIXED, bool TAIL>
static void (*tc_pick(bool missing, bool osc))(const AbcTcParams) {
    return missing ? (osc ? k_abc_acf_tc<true, true, MIXED, TAIL> : k_abc_acf_tc<true, false, MIXED, TAIL>)
                   : (osc ? k_abc_acf_tc<false, true, MIXED, TAIL> : k_abc_acf_tc<false, false, MIXED, TAIL>);
}

cudaError_t abc_tc_launch(const AbcTcParams& P, bool missing, bool osc, bool mixed, size_t smem, int grid,
                          cudaStream_t st) {
    void (*kern)(const AbcTcParams) =
        !mixed ? tc_pick<false, false>(missing, osc)
               : (P.tail_lt > 0 ? tc_pick<true, true>(missing, osc) : tc_pick<true, false>(missing, osc));
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    if ((int64_t)grid > P.n_particles) grid = (int)P.n_particles;
    kern<<<grid, kTcGroupThreads * P.groups + kTcEpiThreads, smem, st>>>(P);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------
// Probe: copy a caller-built fp16 shared-memory image, issue the listed tcgen05.mma ops
// (one thread), drain the whole 128 x 512 fp32 accumulator block to `out`.
// Used by tests/test_gpu_tc_probe.py to check the MN-major no-swizzle layout, the row-shifted
// B operand, the M = 64 lane mapping and the instruction descriptor against exact integer Gram
// products.  ops[i] = {A byte offset, B byte offset, TMEM address offset (lane << 16 | column), accumulate}.
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
