// K2 (second version): fused  simulate OU(+oscillation) -> Hamming periodogram (comp_psd_adfriendly) -> trial
// mean -> distance for 4096 < n_t <= 16384, i.e. a packed complex transform of N = 8192 or 16384 points
// (n2 = 2 N = 2 nextpow2(n_t) real points).  Same reference functions as psd_kernels.cu
// (src/models/one_timescale_and_osc.jl:293 -> src/stats/summary.jl:204-235, src/stats/distances.jl:29-53).
//
// What changed against k_abc_psd (measured there: simulate 24 %, window / pack 10 %, FFT + bins 66 % of a trial,
// serialised by __syncthreads, 59 % issue slots):
//   * warp specialisation: a simulate group (12 warps) produces trial t + 1 - already multiplied by the window -
//     into one of two trial buffers while the transform group (16 warps) works on trial t; full / empty
//     mbarriers per buffer, named barriers inside the groups, no CTA-wide barrier in the trial loop.
//     Measured at C4 (clock64 around the waits, make EXTRA=-DITJL_TUNING; ITJL_PSD_DEBUG=4 / 8 runs one group alone):
//     transform group alone 17.0 k cycles per trial at 80 registers (18.3 k at 72), simulate group alone 13.6 k with
//     4 or 8 warps (it is latency bound: Philox / Box-Muller chains), 10.6 k with 12 warps, four counter blocks per
//     iteration and the fp32 phase increments; together 30.7 k (4 warps: the transform group waits 39 % of the
//     time), 24.1 k (8 warps), 21.6 k (12 warps, 72 registers per thread: both groups wait < 3 %).  Which warp ids
//     the simulate group gets and suspend-hint / nanosleep waits made no difference.
//   * the zero padding is never stored or transformed.  The first DIF stage of radix R0 = N / 4096 splits the
//     transform into R0 independent 4096-point transforms of y_J[p] = w_N^(J p) sum_m z[p + 4096 m] W_R0^(J m);
//     only m = 0 and (for p + 4096 < n_t / 2) m = 1 are non-zero, so y_J is formed in registers straight from the
//     trial buffer and fused with the first radix-16 pass of its 4096-point transform: w_N^(J p) with p = u + 256 e
//     factors into a per-thread constant (folded into the pass twiddles) and 16 literals W_64^(J e).
//   * each 4096-point transform is three radix-16 register passes by 256 threads in its own padded 34 KB buffer
//     (two transforms run side by side: J = 0, 2 then 1, 3 for R0 = 4; J = 0, 1 for R0 = 2), synchronised by
//     256-thread named barriers.  Bins k and N - k of the real-FFT recombination live in sub-transforms J and
//     (R0 - J) % R0, which are always in the same round.
#include <math.h>

#include "kernels.h"
#include "simulate.cuh"
#include "psd_fft.cuh"

namespace itjl {

#ifndef ITJL_PSD2_KPRE
#define ITJL_PSD2_KPRE 1   // bin-table rows fetched under the last pass (release builds measured: 1: 1.29e11, 2: 1.24e11, 3: 1.19e11, 4: 1.21e11 samples/s at C4)
#endif
constexpr int kPsd2SimThreads = 384;
constexpr int kPsd2FftThreads = 512;
constexpr int kPsd2Threads = kPsd2SimThreads + kPsd2FftThreads;
constexpr int kSub = 4096;                        // points of one sub-transform
constexpr int kSubBuf = 4096 + 256 + 2;           // fft_buf_len(4096) float2

// named barrier ids (0 = __syncthreads)
enum : int { kBarSim = 1, kBarFft = 2, kBarHalf0 = 3 /* , kBarHalf1 = 4 */ };

// cos / sin of 2 pi m / 64, m = 0 .. 16 (the rest by symmetry); W_64^m = exp(-2 pi i m / 64)
__device__ __forceinline__ float2 w64(int m) {
    constexpr float C[17] = {1.0f, 0.99518472667219693f, 0.98078528040323043f, 0.95694033573220882f, 0.92387953251128674f,
                             0.88192126434835505f, 0.83146961230254524f, 0.77301045336273699f, 0.70710678118654757f,
                             0.63439328416364549f, 0.55557023301960218f, 0.47139673682599764f, 0.38268343236508978f,
                             0.29028467725446233f, 0.19509032201612825f, 0.098017140329560604f, 0.0f};
    m &= 63;
    const int qd = m >> 4, r = m & 15;
    // angle = (16 qd + r) * 2 pi / 64: cos(a + qd pi/2), sin(a + qd pi/2)
    const float c = C[r], s = C[16 - r];
    float cr, sr;
    switch (qd) {
        case 0: cr = c; sr = s; break;
        case 1: cr = -s; sr = c; break;
        case 2: cr = -c; sr = -s; break;
        default: cr = s; sr = -c; break;
    }
    return make_float2(cr, -sr);
}

// the 16-point register butterfly of fft_radix16_pass with the two twiddle sets passed in:
// after level 1, element e0 + 4 r is multiplied by a_r (r = 0 .. 3; a_0 only when A0), after level 2 element
// 4 b + r' by s^r' (when TW2)
template <bool A0, bool TW1, bool TW2>
__device__ __forceinline__ void r16_butterfly(float2 (&x)[16], float2 a0, float2 a1, float2 a2, float2 a3, float2 s1) {
    constexpr float C1 = 0.92387953251128674f, S1 = 0.38268343236508977f, C2 = 0.70710678118654752f;
    bfly4(x[0], x[4], x[8], x[12]);
    bfly4(x[1], x[5], x[9], x[13]);
    x[5] = cmul(x[5], make_float2(C1, -S1)); x[9] = cmul(x[9], make_float2(C2, -C2)); x[13] = cmul(x[13], make_float2(S1, -C1));
    bfly4(x[2], x[6], x[10], x[14]);
    x[6] = cmul(x[6], make_float2(C2, -C2)); x[10] = make_float2(x[10].y, -x[10].x); x[14] = cmul(x[14], make_float2(-C2, -C2));
    bfly4(x[3], x[7], x[11], x[15]);
    x[7] = cmul(x[7], make_float2(S1, -C1)); x[11] = cmul(x[11], make_float2(-C2, -C2)); x[15] = cmul(x[15], make_float2(-C1, S1));
    if (TW1) {
#pragma unroll
        for (int e0 = 0; e0 < 4; ++e0) {
            if (A0) x[e0] = cmul(x[e0], a0);
            x[e0 + 4] = cmul(x[e0 + 4], a1); x[e0 + 8] = cmul(x[e0 + 8], a2); x[e0 + 12] = cmul(x[e0 + 12], a3);
        }
    }
#pragma unroll
    for (int b4 = 0; b4 < 16; b4 += 4) bfly4(x[b4], x[b4 + 1], x[b4 + 2], x[b4 + 3]);
    if (TW2) {
        const float2 s2 = cmul(s1, s1);
        const float2 s3 = cmul(s2, s1);
#pragma unroll
        for (int b4 = 0; b4 < 16; b4 += 4) {
            x[b4 + 1] = cmul(x[b4 + 1], s1); x[b4 + 2] = cmul(x[b4 + 2], s2); x[b4 + 3] = cmul(x[b4 + 3], s3);
        }
    }
}

// First pass of sub-transform J (of R0): thread u < 256 owns points p = u + 256 e.  z = the windowed trial as
// float2 pairs, z_len = number of pairs stored (zeros beyond the data are stored up to there, nothing beyond).
// t1 = W_4096^u, s1 = W_1024^u (the pass twiddles of thread u), c = w_N^(J u): constant over the trials, kept in registers
template <int R0, int J>
__device__ __forceinline__ void psd2_pass1(const float2* __restrict__ z, int z_len, float2* __restrict__ buf, int u,
                                           const float2 t1, const float2 s1, const float2 c) {
    float2 x[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) {
        const int p = u + 256 * e;
        float2 v = p < z_len ? z[p] : make_float2(0.f, 0.f);
        if (R0 == 4 && p + kSub < z_len) {                 // second non-zero input of the radix-4 stage, times (-i)^J
            const float2 q = z[p + kSub];
            if (J == 0) { v.x += q.x; v.y += q.y; }
            else if (J == 1) { v.x += q.y; v.y -= q.x; }
            else if (J == 2) { v.x -= q.x; v.y -= q.y; }
            else { v.x -= q.y; v.y += q.x; }
        }
        // w_N^(J 256 e): N = 16384 -> W_64^(J e); N = 8192 -> W_32^(J e) = W_64^(2 J e)
        if (J != 0 && e != 0) v = cmul(v, w64((R0 == 4 ? 1 : 2) * J * e));
        x[e] = v;
    }
    // pass twiddles of the 4096-point transform (span 4096, q = 256, index u); the per-thread factor w_N^(J u)
    // multiplies every output
    const float2 t2 = cmul(t1, t1);
    const float2 t3 = cmul(t2, t1);
    if (J == 0) r16_butterfly<false, true, true>(x, t1, t1, t2, t3, s1);
    else r16_butterfly<true, true, true>(x, c, cmul(t1, c), cmul(t2, c), cmul(t3, c), s1);
#pragma unroll
    for (int e = 0; e < 16; ++e) buf[fft_pad(u + 256 * e)] = x[e];
}

// passes 2 (span 256) and 3 (span 16, twiddle free) of a 4096-point transform, thread u < 256
__device__ __forceinline__ void psd2_pass2(float2* __restrict__ buf, int u, const float2 t1, const float2 s1) {
    const int j = u & 15;
    const int base = ((u >> 4) << 8) + j;
    float2 x[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) x[e] = buf[fft_pad(base + 16 * e)];
    {
        const float2 t2 = cmul(t1, t1);
        const float2 t3 = cmul(t2, t1);
        r16_butterfly<false, true, true>(x, t1, t1, t2, t3, s1);
    }
#pragma unroll
    for (int e = 0; e < 16; ++e) buf[fft_pad(base + 16 * e)] = x[e];
}
__device__ __forceinline__ void psd2_pass3(float2* __restrict__ buf, int u) {
    float2 x[16];
    const int base = 17 * u;                               // fft_pad(16 u + e) = 17 u + e
#pragma unroll
    for (int e = 0; e < 16; ++e) x[e] = buf[base + e];
    r16_butterfly<false, false, false>(x, x[0], x[0], x[0], x[0], x[0]);
#pragma unroll
    for (int e = 0; e < 16; ++e) buf[base + e] = x[e];
}

template <int R0>
__device__ __forceinline__ void psd2_pass1_any(int J, const float2* z, int z_len, float2* buf, int u, float2 t1, float2 s1, float2 c) {
    if (R0 == 4) {
        switch (J) {
            case 0: psd2_pass1<4, 0>(z, z_len, buf, u, t1, s1, c); break;
            case 1: psd2_pass1<4, 1>(z, z_len, buf, u, t1, s1, c); break;
            case 2: psd2_pass1<4, 2>(z, z_len, buf, u, t1, s1, c); break;
            default: psd2_pass1<4, 3>(z, z_len, buf, u, t1, s1, c); break;
        }
    } else {
        if (J == 0) psd2_pass1<2, 0>(z, z_len, buf, u, t1, s1, c);
        else psd2_pass1<2, 1>(z, z_len, buf, u, t1, s1, c);
    }
}

template <bool OSC, int R0>
__global__ void __launch_bounds__(kPsd2Threads, 1) k_abc_psd2(const __grid_constant__ AbcPsdParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* xs0 = reinterpret_cast<float*>(smem_raw);                         // two trial buffers of xs_len floats
    float2* bufs = reinterpret_cast<float2*>(xs0 + 2 * P.xs_len);            // two sub-transform buffers of kSubBuf float2
    float* accb = reinterpret_cast<float*>(bufs + 2 * kSubBuf);              // n_stats bin sums, in table order
    const int n_stats_pad = (P.n_stats + 3) & ~3;
    SimShared<kPsd2SimThreads>* sh = reinterpret_cast<SimShared<kPsd2SimThreads>*>(accb + n_stats_pad);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sh + 1);                    // full[2], empty[2]
    __shared__ double red[kPsd2FftThreads / 32];

#ifdef ITJL_PSD2_SIM_FIRST
    const int tid0 = threadIdx.x;
    const bool is_sim = tid0 < kPsd2SimThreads;                              // simulate group = the LOWEST warp ids
    const int tid = is_sim ? tid0 + kPsd2FftThreads : tid0 - kPsd2SimThreads;   // role-local numbering as below
#else
    const int tid = threadIdx.x;                                             // transform group first, simulate group = the highest warp ids
#endif
    const int64_t particle = blockIdx.x;
    const int N = P.n2 >> 1;

    for (int b = threadIdx.x; b < P.n_stats; b += kPsd2Threads) accb[b] = 0.f;
    if (threadIdx.x == 0) {
        tc::mbar_init(&bars[0], kPsd2SimThreads); tc::mbar_init(&bars[1], kPsd2SimThreads);
        tc::mbar_init(&bars[2], kPsd2FftThreads); tc::mbar_init(&bars[3], kPsd2FftThreads);
        tc::fence_mbar_init();
    }
    __syncthreads();

    if (tid >= kPsd2FftThreads) {
        // ---------------- simulate group: trial t -> xs[t & 1], windowed ----------------
        const int stid = tid - kPsd2FftThreads;
        const double tau = P.theta[particle];
        double freq = 0.0, coeff = 1.0;
        if (OSC) {
            freq = P.theta[particle + P.theta_ld];
            coeff = P.theta[particle + 2 * P.theta_ld];
        }
        const OuConsts oc = make_ou_consts(tau, P.dt, P.update, OSC ? ITJL_KIND_OU_OSC : ITJL_KIND_OU,
                                           freq, coeff, P.n_t, true, true);   // |FFT| of the Hamming-windowed trial is reversal invariant
        SimCache sim_cache;
        SimArgs g;
        g.n_t = P.n_t; g.chunk = P.chunk;
        g.inv_n = 1.0 / (double)P.n_t; g.sqrt_nm1 = sqrtf((float)(P.n_t - 1));
        g.keys = &P.keys; g.c3_noise = P.c3_noise; g.c3_init = P.c3_init;
        g.particle = (uint32_t)(P.particle_offset + particle);
        const size_t pt = (size_t)particle * P.n_trials;
        g.u0 = P.u0 ? P.u0 + pt : nullptr;
        g.noise = P.noise ? P.noise + pt * P.n_t : nullptr;
        g.phases = P.phases ? P.phases + pt : nullptr;
        g.mask_bits = nullptr; g.n_valid = nullptr; g.mask_words = 0;
        g.win = P.window;
#ifdef ITJL_TUNING
        long long s_wait = 0; const long long s_t0 = clock64();
#endif
        for (int trial = 0; trial < P.n_trials; ++trial) {
            const int b = trial & 1;
#ifdef ITJL_TUNING
            const long long w0 = clock64();
#endif
            if (trial >= 2) tc::mbar_wait(&bars[2 + b], ((trial >> 1) - 1) & 1);     // the transform group has read trial - 2
#ifdef ITJL_TUNING
            s_wait += clock64() - w0;
#endif
            // standardised series * data_sd (the reference adds data_mean and comp_psd_adfriendly removes the mean
            // again, summary.jl:209 - both skipped), times the Hamming window
#ifdef ITJL_TUNING
            if (!(P.grid & 8))
#endif
            simulate_trial<kPsd2SimThreads, false, OSC, false, true>(g, oc, trial, xs0 + b * P.xs_len, sh, kScaleStd, (float)P.data_sd, 0.0f,
                                                        sim_cache, stid, kBarSim);
            tc::mbar_arrive(&bars[b]);
        }
#ifdef ITJL_TUNING
        if (blockIdx.x == 0 && stid == 0) printf("sim group: total %lld clk, waiting for a free buffer %lld clk\n", clock64() - s_t0, s_wait);
#endif
        return;
    }

    // ---------------- transform group ----------------
    const int half = tid >> 8, u = tid & 255;
    float2* mybuf = bufs + half * kSubBuf;
    const int z_len = P.xs_len >> 1;
    constexpr int kRounds = R0 == 4 ? 2 : 1;
    // this thread's twiddles (the same for every trial): exp(-2 pi i idx / 2N) from the global table, once
    const float2 p1_t1 = tw_lookup(P.twiddle, u * ((2 * N) >> 12), N), p1_s1 = tw_lookup(P.twiddle, 4 * u * ((2 * N) >> 12), N);
    const float2 p2_t1 = tw_lookup(P.twiddle, (u & 15) * ((2 * N) >> 8), N), p2_s1 = tw_lookup(P.twiddle, 4 * (u & 15) * ((2 * N) >> 8), N);
    const float2 c_r0 = tw_lookup(P.twiddle, 2 * (R0 == 4 ? 2 * half : half) * u, N);        // w_N^(J u) of round 0 ...
    const float2 c_r1 = tw_lookup(P.twiddle, 2 * (1 + 2 * half) * u, N);                      // ... and round 1
#ifdef ITJL_TUNING
    long long f_wait = 0; const long long f_t0 = clock64();
#endif
    for (int trial = 0; trial < P.n_trials; ++trial) {
        const int b = trial & 1;
        const float2* z = reinterpret_cast<const float2*>(xs0 + b * P.xs_len);
#ifdef ITJL_TUNING
        const long long w0 = clock64();
#endif
        tc::mbar_wait(&bars[b], (trial >> 1) & 1);
#ifdef ITJL_TUNING
        f_wait += clock64() - w0;
#endif
        int bin_lo = 0;
#ifdef ITJL_TUNING
        if (P.grid & 4) { tc::mbar_arrive(&bars[2 + b]); continue; }
#endif
#pragma unroll 1
        for (int round = 0; round < kRounds; ++round) {
            // R0 = 4: round 0 transforms J = 0 (half 0) and 2 (half 1), round 1 J = 1 and 3; R0 = 2: J = half
            const int J = R0 == 4 ? round + 2 * half : half;
            psd2_pass1_any<R0>(J, z, z_len, mybuf, u, p1_t1, p1_s1, round == 0 ? c_r0 : c_r1);
            if (round == kRounds - 1) tc::mbar_arrive(&bars[2 + b]);         // last read of the trial buffer
            tc::named_bar_sync(kBarHalf0 + half, 256);
            psd2_pass2(mybuf, u, p2_t1, p2_s1);
            tc::named_bar_sync(kBarHalf0 + half, 256);
            // bins of this round (table sorted by round, then storage position): the first rows are fetched from L2
            // under the last pass
            const int bin_hi = round == 0 ? P.round0_bins : P.n_stats;
            constexpr int kPre = ITJL_PSD2_KPRE;
            int4 pre[kPre];
#pragma unroll
            for (int i = 0; i < kPre; ++i) {
                const int bb = bin_lo + tid + i * kPsd2FftThreads;
                pre[i] = bb < bin_hi ? __ldg(P.binpos + bb) : make_int4(0, 0, 0, 0);
            }
            psd2_pass3(mybuf, u);
            tc::named_bar_sync(kBarFft, kPsd2FftThreads);
#pragma unroll
            for (int i = 0; i < kPre; ++i) {
                const int bb = bin_lo + tid + i * kPsd2FftThreads;
                if (bb < bin_hi) accb[bb] += real_fft_power_at(bufs, pre[i]);
            }
#pragma unroll 4
            for (int bb = bin_lo + tid + kPre * kPsd2FftThreads; bb < bin_hi; bb += kPsd2FftThreads)
                accb[bb] += real_fft_power_at(bufs, __ldg(P.binpos + bb));
            bin_lo = bin_hi;
            tc::named_bar_sync(kBarFft, kPsd2FftThreads);
        }
    }

#ifdef ITJL_TUNING
    if (blockIdx.x == 0 && tid == 0) printf("transform group: total %lld clk, waiting for a trial %lld clk\n", clock64() - f_t0, f_wait);
#endif
    const double sc = P.psd_scale / (double)P.n_trials;
    double local = 0.0;
    for (int bb = tid; bb < P.n_stats; bb += kPsd2FftThreads) {
        const double v = (double)accb[bb] * sc;
        const int bi = __ldg(P.binidx + bb);                     // accb is in table order
        if (P.stats_out) P.stats_out[(size_t)particle * P.n_stats + bi] = v;
        const double tgt = (double)P.target[bi];
        const double e = (P.distance == ITJL_DIST_LINEAR) ? (v - tgt) : (log(v) - log(tgt));
        local += e * e;
    }
    local = warp_sum(local);
    if ((tid & 31) == 0) red[tid >> 5] = local;
    tc::named_bar_sync(kBarFft, kPsd2FftThreads);
    if (tid == 0) {
        double d = 0.0;
#pragma unroll
        for (int w = 0; w < kPsd2FftThreads / 32; ++w) d += red[w];
        P.dist_out[particle] = d / (double)P.n_stats;
    }
}

// ---- host side ---------------------------------------------------------------------------------------------
bool abc_psd2_supported(int n_t, int n2) { return n2 == 16384 || n2 == 32768 ? n_t > 4 : false; }

size_t abc_psd2_smem_bytes(int n_t, int n2, int n_stats, int* chunk_out, int* xs_len_out) {
    // samples per simulate thread: a multiple of 4 with chunk / 4 odd, so that the 16-byte accesses of a warp at a
    // stride of `chunk` floats spread over all banks (chunk = 80: 16 wavefronts per access instead of 4)
    int c = (n_t + kPsd2SimThreads - 1) / kPsd2SimThreads;
    c = (c + 3) & ~3;
    if (((c >> 2) & 1) == 0) c += 4;
    *chunk_out = c;
    *xs_len_out = c * kPsd2SimThreads;
    return (size_t)2 * c * kPsd2SimThreads * sizeof(float) + (size_t)2 * kSubBuf * sizeof(float2) +
           (size_t)((n_stats + 3) & ~3) * sizeof(float) + sizeof(SimShared<kPsd2SimThreads>) +
           4 * sizeof(uint64_t);
}

// position of bin k (1 <= k < N) of the packed transform in the two sub-transform buffers, and its round
void abc_psd2_bin(int k, int N, int* addr, int* round) {
    const int R0 = N / kSub;
    const int J = k % R0, r = k / R0;
    int half, rnd;
    if (R0 == 4) { half = J >> 1; rnd = J & 1; } else { half = J; rnd = 0; }
    *addr = half * kSubBuf + fft_pad(fft_pos(r, kSub));
    *round = rnd;
}

cudaError_t abc_psd2_launch(const AbcPsdParams& P, bool osc, size_t smem, cudaStream_t st) {
    const int R0 = (P.n2 >> 1) / kSub;
    void (*kern)(const AbcPsdParams) = R0 == 4 ? (osc ? k_abc_psd2<true, 4> : k_abc_psd2<false, 4>)
                                               : (osc ? k_abc_psd2<true, 2> : k_abc_psd2<false, 2>);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<(unsigned)P.n_particles, kPsd2Threads, smem, st>>>(P);
    return cudaGetLastError();
}

}  // namespace itjl
