// K2 (second version): fused  simulate OU(+oscillation) -> Hamming periodogram (comp_psd_adfriendly) -> trial
// mean -> distance for 4096 < n_t <= 16384, i.e. a packed complex transform of N = 8192 or 16384 points
// (n2 = 2 N = 2 nextpow2(n_t) real points).  Same reference functions as psd_kernels.cu
// (src/models/one_timescale_and_osc.jl:293 -> src/stats/summary.jl:204-235, src/stats/distances.jl:29-53).
//
// What changed against k_abc_psd (measured there: simulate 24 %, window / pack 10 %, FFT + bins 66 % of a trial,
// serialised by __syncthreads, 59 % issue slots):
//   * warp specialisation: a simulate group (4 warps) produces trial t + 1 - already multiplied by the window -
//     into one of two trial buffers while the transform group (16 warps) works on trial t; full / empty
//     mbarriers per buffer, named barriers inside the groups, no CTA-wide barrier in the trial loop.
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

constexpr int kPsd2SimThreads = 128;
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
template <int R0, int J>
__device__ __forceinline__ void psd2_pass1(const float2* __restrict__ z, int z_len, float2* __restrict__ buf, int u, int Ntab,
                                           const Tw2& tw) {
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
    // pass twiddles of the 4096-point transform (span 4096, q = 256, index u): W_4096^u = tw(u * 2 Ntab / 4096);
    // the per-thread factor w_N^(J u) = tw(2 J u) multiplies every output
    const int tstep = (2 * Ntab) >> 12;
    if (J == 0) {
        if (u != 0) {
            const float2 t1 = tw(u * tstep);
            const float2 t2 = cmul(t1, t1);
            const float2 t3 = cmul(t2, t1);
            r16_butterfly<false, true, true>(x, t1, t1, t2, t3, tw(4 * u * tstep));
        } else {
            r16_butterfly<false, false, false>(x, x[0], x[0], x[0], x[0], x[0]);
        }
    } else {
        const float2 c = tw(2 * J * u);
        const float2 t1 = tw(u * tstep);
        const float2 t2 = cmul(t1, t1);
        const float2 t3 = cmul(t2, t1);
        r16_butterfly<true, true, true>(x, c, cmul(t1, c), cmul(t2, c), cmul(t3, c), tw(4 * u * tstep));
    }
#pragma unroll
    for (int e = 0; e < 16; ++e) buf[fft_pad(u + 256 * e)] = x[e];
}

// passes 2 (span 256) and 3 (span 16, twiddle free) of a 4096-point transform, thread u < 256
__device__ __forceinline__ void psd2_pass2(float2* __restrict__ buf, int u, int Ntab, const Tw2& tw) {
    const int j = u & 15;
    const int base = ((u >> 4) << 8) + j;
    const int tstep = (2 * Ntab) >> 8;
    float2 x[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) x[e] = buf[fft_pad(base + 16 * e)];
    if (j != 0) {
        const float2 t1 = tw(j * tstep);
        const float2 t2 = cmul(t1, t1);
        const float2 t3 = cmul(t2, t1);
        r16_butterfly<false, true, true>(x, t1, t1, t2, t3, tw(4 * j * tstep));
    } else {
        r16_butterfly<false, false, false>(x, x[0], x[0], x[0], x[0], x[0]);
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
__device__ __forceinline__ void psd2_pass1_any(int J, const float2* z, int z_len, float2* buf, int u, int Ntab, const Tw2& tw) {
    if (R0 == 4) {
        switch (J) {
            case 0: psd2_pass1<4, 0>(z, z_len, buf, u, Ntab, tw); break;
            case 1: psd2_pass1<4, 1>(z, z_len, buf, u, Ntab, tw); break;
            case 2: psd2_pass1<4, 2>(z, z_len, buf, u, Ntab, tw); break;
            default: psd2_pass1<4, 3>(z, z_len, buf, u, Ntab, tw); break;
        }
    } else {
        if (J == 0) psd2_pass1<2, 0>(z, z_len, buf, u, Ntab, tw);
        else psd2_pass1<2, 1>(z, z_len, buf, u, Ntab, tw);
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
    float2* twtab = reinterpret_cast<float2*>(sh + 1);
    uint64_t* bars = reinterpret_cast<uint64_t*>(twtab + tw2_entries(P.n2 >> 1));      // full[2], empty[2]
    __shared__ double red[kPsd2FftThreads / 32];

    const int tid = threadIdx.x;
    const int64_t particle = blockIdx.x;
    const int N = P.n2 >> 1;

    for (int b = tid; b < P.n_stats; b += kPsd2Threads) accb[b] = 0.f;
    const Tw2 tw = tw2_build(twtab, P.twiddle, N, tid, kPsd2Threads);
    if (tid == 0) {
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
        for (int trial = 0; trial < P.n_trials; ++trial) {
            const int b = trial & 1;
            if (trial >= 2) tc::mbar_wait(&bars[2 + b], ((trial >> 1) - 1) & 1);     // the transform group has read trial - 2
            // standardised series * data_sd (the reference adds data_mean and comp_psd_adfriendly removes the mean
            // again, summary.jl:209 - both skipped), times the Hamming window
            simulate_trial<kPsd2SimThreads, false, OSC>(g, oc, trial, xs0 + b * P.xs_len, sh, kScaleStd, (float)P.data_sd, 0.0f,
                                                        sim_cache, stid, kBarSim);
            tc::mbar_arrive(&bars[b]);
        }
        return;
    }

    // ---------------- transform group ----------------
    const int half = tid >> 8, u = tid & 255;
    float2* mybuf = bufs + half * kSubBuf;
    const int z_len = P.xs_len >> 1;
    constexpr int kRounds = R0 == 4 ? 2 : 1;
    for (int trial = 0; trial < P.n_trials; ++trial) {
        const int b = trial & 1;
        const float2* z = reinterpret_cast<const float2*>(xs0 + b * P.xs_len);
        tc::mbar_wait(&bars[b], (trial >> 1) & 1);
        int bin_lo = 0;
#pragma unroll 1
        for (int round = 0; round < kRounds; ++round) {
            // R0 = 4: round 0 transforms J = 0 (half 0) and 2 (half 1), round 1 J = 1 and 3; R0 = 2: J = half
            const int J = R0 == 4 ? round + 2 * half : half;
            psd2_pass1_any<R0>(J, z, z_len, mybuf, u, N, tw);
            if (round == kRounds - 1) tc::mbar_arrive(&bars[2 + b]);         // last read of the trial buffer
            tc::named_bar_sync(kBarHalf0 + half, 256);
            psd2_pass2(mybuf, u, N, tw);
            tc::named_bar_sync(kBarHalf0 + half, 256);
            psd2_pass3(mybuf, u);
            tc::named_bar_sync(kBarFft, kPsd2FftThreads);
            // bins of this round (table sorted by round, then storage position): rows come from L2, four in flight
            const int bin_hi = round == 0 ? P.round0_bins : P.n_stats;
#pragma unroll 4
            for (int bb = bin_lo + tid; bb < bin_hi; bb += kPsd2FftThreads)
                accb[bb] += real_fft_power_at(bufs, __ldg(P.binpos + bb));
            bin_lo = bin_hi;
            tc::named_bar_sync(kBarFft, kPsd2FftThreads);
        }
    }

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
    int c = (n_t + kPsd2SimThreads - 1) / kPsd2SimThreads;
    c = (c + 7) & ~7;
    *chunk_out = c;
    *xs_len_out = c * kPsd2SimThreads;
    return (size_t)2 * c * kPsd2SimThreads * sizeof(float) + (size_t)2 * kSubBuf * sizeof(float2) +
           (size_t)((n_stats + 3) & ~3) * sizeof(float) + sizeof(SimShared<kPsd2SimThreads>) +
           (size_t)tw2_entries(n2 / 2) * sizeof(float2) + 4 * sizeof(uint64_t);
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
