// One OU (+ oscillation) trial simulated by a whole CTA straight into shared memory.
//
// Replaces, per trial, generate_ou_process / generate_ou_with_oscillation
// (reference src/utils/ou_process.jl:114-148, 176-218) followed by the centring /
// scaling that the summary statistic applies first (summary.jl:48, 466-471, 209-213).
//
// The linear recurrence x_{k+1} = a x_k + s xi_k is serial in k, so it is evaluated as a
// blocked scan: every thread integrates its own chunk of `chunk` consecutive samples from
// a zero state (Philox noise generated in registers), the chunk-end affine maps
// (a^chunk, y_end) are combined by a warp-shuffle + shared-memory scan, and the carry-in
// is added back as p_j * carry with p_j = a^(j+1).  Mean and sum of squares follow
// analytically from per-thread partial sums (S_y, S_yy, S_yu, S_p, S_pp), so the trial
// needs exactly two passes over its chunk and three __syncthreads().
//
// The powers enter as q_j = 1 - a^(j+1), a table of `chunk` floats per particle computed in fp64
// (SimShared::qtab) and read four at a time: x_j = y_j + c - c q_j.  Carrying p_j itself by the fp32
// recurrence p <- p - eps p rounds the SAME way at every step when eps is small (p ~ 1, the decrement a
// fixed number of ulps): a systematic error of the decay rate of 0.5 ulp / (eps / 2^-24), i.e. 1e-3 at
// tau = 500 dt-steps x 1000, which put 1e-5 .. 4e-5 on the ACF of near-unit-root trials (|tau| >> T: most of
// what the README's Normal(tau_hat, 20) prior proposes).  The affine maps are composed in the same form,
// (Q, B) with Q = 1 - A.
#pragma once
#include "common.cuh"
#include "philox.cuh"
#include "tc05.cuh"

namespace itjl {

struct OuConsts {
    float eps;    // 1 - a, a = exp(-dt/tau) (exact) or 1 - dt/tau (Euler-Maruyama)
    float s;      // innovation scale sqrt(tau/2 (1-a^2)) or sqrt(dt)
    float sc;     // sqrt(coeff)            (1 for the plain OU model)
    float so;     // sqrt(1-coeff)*sqrt(2)  (0 for the plain OU model)
    double fdt;   // freq * dt, in turns per sample
    float fdt_f;  // the same in fp32, reduced to [-0.5, 0.5]
    float x0_mul; // factor on the drawn initial state (1; e^-G for a growing trajectory, see below)
    float x0_add; // added to it (0; 1 with x0_mul = 0 for the noise-free limit)
    double lna;   // ln a (NaN when a <= 0: Euler-Maruyama with dt > tau), a64 = a: for the table of 1 - a^(j+1)
    double a64;
};

// q_j = 1 - a^(j+1) in fp64 (one call per table entry and particle)
__device__ __forceinline__ float ou_q(const OuConsts& oc, int j) {
    const double k = (double)(j + 1);
    return (float)(oc.lna == oc.lna ? -expm1(k * oc.lna) : 1.0 - pow(oc.a64, k));
}

// Growing trajectories (tau < 0: a > 1).  The reference integrates them like any other draw of its prior - the
// README's Normal(tau_hat, 20) proposes a negative tau half of the time - and its fp64 state stays finite up to
// a log-growth G = n_t ln a of ~354.  fp32 sums of squares overflow from G ~ 40, so such a trial is simulated
// in a scaled form; every statistic downstream is scale free (the trial is standardised first):
//   G <= 30        : as is.
//   30 < G <= 80   : initial state and innovations times e^-G (the trial then ends at O(1)); the oscillation
//                    term, < 1e-13 of the trial's scale, is dropped.
//   80 < G <= 354  : the innovations are < e^-70 of the samples that carry any weight, so the trial IS the
//                    exponential C a^t; its statistics equal those of its time reversal a^-j, which is
//                    what is generated (x0 = 1, no noise, coefficient 1/a).  Not `reversible` (a missing-data
//                    mask is tied to the time axis): NaN - a documented deviation, the reference is finite there.
//   G > 354        : fp64 overflows as well (x^2 > 1.8e308): NaN, like the reference.
// `scalable` is false for standardize = false (the raw values are the output).
__device__ __forceinline__ OuConsts make_ou_consts(double tau, double dt, int update, int kind,
                                                   double freq, double coeff, int n_t = 0, bool scalable = true,
                                                   bool reversible = true) {
    OuConsts c;
    double eps, s;
    if (update == ITJL_UPDATE_EXACT) {
        const double x = -dt / tau;
        eps = -expm1(x);                       // 1 - a without cancellation
        const double a = 1.0 - eps;
        s = sqrt(tau * 0.5 * (1.0 - a * a));
    } else {
        eps = dt / tau;
        s = sqrt(dt);
    }
    c.sc = 1.0f; c.so = 0.0f; c.fdt = 0.0; c.fdt_f = 0.0f;
    c.x0_mul = 1.0f; c.x0_add = 0.0f;
    if (kind == ITJL_KIND_OU_OSC) {
        // ou_process.jl:189-196: only values outside [0, 1] are moved
        if (coeff < 0.0) coeff = 1e-4; else if (coeff > 1.0) coeff = 1.0 - 1e-4;
        c.sc = (float)sqrt(coeff);
        c.so = (float)(sqrt(1.0 - coeff) * 1.4142135623730951);
        c.fdt = freq * dt;
        c.fdt_f = (float)(c.fdt - rint(c.fdt));
    }
    if (eps < 0.0 && scalable && n_t > 0) {
        const double G = (double)n_t * log1p(-eps);          // ln a^n
        if (G > 354.0 || (G > 80.0 && !reversible)) {
            eps = NAN;
        } else if (G > 80.0) {
            eps = 1.0 - 1.0 / (1.0 - eps);                   // coefficient 1/a
            s = 0.0; c.so = 0.0f; c.x0_mul = 0.0f; c.x0_add = 1.0f;
        } else if (G > 30.0) {
            const double g = exp(-G);
            s *= g; c.so = 0.0f; c.x0_mul = (float)g;
        }
    }
    c.eps = (float)eps;
    c.s = (float)s;
    c.a64 = 1.0 - eps;
    c.lna = (update == ITJL_UPDATE_EXACT && c.x0_add == 0.0f) ? -dt / tau : log1p(-eps);   // NaN for a <= 0
    return c;
}

template <int NT>
struct SimShared {
    float scanA[NT / 32 < 32 ? 32 : NT / 32];      // Q = 1 - A of the chunk-end maps (32 entries: the tensor-core sink scans 8 x 4 blocks)
    float scanB[NT / 32 < 32 ? 32 : NT / 32];
    double sums[NT / 32][4];
    __align__(16) float qtab[kQtab];               // q_j = 1 - a^(j+1), j < chunk <= kQtab, of the current particle
    double sp_full, spp_full;                      // sum_j (1 - q_j), sum_j (1 - q_j)^2 over a whole chunk, in fp64
    float x0;                                      // the trial's initial state, drawn by warp 0 (plain OU model)
};

enum : int { kScaleUnitNorm = 0, kScaleStd = 1, kScaleRaw = 2 };

struct SimArgs {
    int n_t;                  // samples per trial
    double inv_n;             // 1 / n_t
    float sqrt_nm1;           // sqrt(n_t - 1)
    int chunk;                // samples per thread (multiple of 4, chunk * NT >= n_t)
    const PhiloxKeys* keys;   // Philox round keys (seed), in the kernel parameters
    uint32_t c3_noise, c3_init;
    uint32_t particle;        // global particle index (low 32 bits)
    const double* u0;         // [trial] injected initial states of this particle, or null
    const double* noise;      // [trial][n_t] injected N(0,1) increments, or null
    const double* phases;     // [trial] injected phases (radians), or null
    float miss_mean_ratio = 0.f;  // MISSING + OSC: data_mean / data_sd (the oscillation model adds data_mean
                                  // before the mask is applied, so masked samples sit at -(mean of valid) - data_mean)
    const uint32_t* mask_bits;  // [trial][mask_words] missing-sample bitmap, or null
    const int* n_valid;         // [trial] number of valid samples
    int mask_words;
    const float* win = nullptr;   // optional window (chunk * NT floats, 16-byte aligned): the output is multiplied by it (plain sink only)
};

// Per-thread values that do not change from trial to trial (filled during the first trial).
struct SimCache {
    double Sp = 0.0, Spp = 0.0;    // sum p_j, sum p_j^2 over this thread's samples (fp64: see simulate_trial)
    bool valid = false;
};

// Tensor-core sink (TC = true): the trial is emitted, scaled by the power of two `scale`, as an fp16
// hi/lo pair (x = hi + lo to ~22 bits) in the MN-major no-swizzle operand layout of tcgen05.mma
// (tc05.cuh): the 8 samples j .. j+7 (j = 128 i + a, a % 8 == 0) occupy the 16 bytes at
// (a/8)*sbo + 16 i of the hi plane and of the lo plane.  There is no separate staging buffer: pass A
// parks its fp32 values in the very same 32 bytes (samples j..j+3 in the hi slot, j+4..j+7 in the lo
// slot) and pass B converts them in place, each thread touching only its own slots.
// `wait_bar` guards the buffer: it completes when the MMAs that read its previous contents have
// retired (the caller alternates between two buffers, so that is the trial before the last one).
struct TcSink {
    unsigned char* hi;
    unsigned char* lo;
    int sbo;
    float scale;              // power of two
    bool write_lo;            // false: only the hi plane is consumed (single-term products), skip the residuals
    uint64_t* wait_bar;
    uint32_t wait_parity;
};

__device__ __forceinline__ void sim_sync(int bar_id, int nt) {
    if (bar_id == 0) __syncthreads();
    else tc::named_bar_sync(bar_id, nt);
}

// Simulate trial `trial` into xs[0 .. chunk*NT) with NT threads (the whole CTA, or - with
// `bar_id` != 0 - one NT-thread group of it synchronising on that named barrier; `tid` is the
// thread's index within the group).  On return (after the
// caller's __syncthreads()) xs holds
//   kScaleUnitNorm : (x - mean) / sqrt(sum (x-mean)^2)                (ACF kernels)
//                    with MISSING: the comp_ac_time_missing centring, masked samples = -m2
//   kScaleStd      : (x - mean) / std_{n-1} * scale + shift            (PSD kernel, generator)
//   kScaleRaw      : x                                                 (standardize = false)
// and zeros for j >= n_t.
// WIDE: four Philox counter blocks (16 samples) per iteration of the interior loop instead of two - for callers that
// run few simulate warps (latency bound: k_abc_psd2's simulate group); with 16+ warps it measured no faster.
template <int NT, bool MISSING, bool OSC, bool TC = false, bool WIDE = false>
__device__ __forceinline__ void simulate_trial(const SimArgs& g, const OuConsts& oc, int trial,
                                               float* __restrict__ xs, SimShared<NT>* sh, int mode,
                                               float scale, float shift, SimCache& cache,
                                               int tid = threadIdx.x, int bar_id = 0,
                                               const TcSink* sink = nullptr) {
    constexpr int NW = NT / 32;
    const int lane = tid & 31, warp = tid >> 5;
    // Chunk owned by this thread.  Plain: chunk = tid.  Tensor-core sink (NT = 128): chunk = 16 m + 4 warp + qt
    // with m = lane % 8, qt = lane / 8 - the eight lanes of a quarter warp then own chunks 16 apart, whose
    // operand slots (five per 40-sample chunk) fall into the SAME 8-sample block 80 slots = 5 K rows apart:
    // eight different 16-byte bank groups, so every 128-bit plane access is conflict free (chunk = tid is a
    // 2-way conflict whatever the row count).
    const int chunk_idx = TC ? 16 * (lane & 7) + 4 * warp + (lane >> 3) : tid;
    const int j0 = chunk_idx * g.chunk;
    const float eps = oc.eps, s = oc.s;

    // per-trial initial state and phase (stream 1, block 0)
    // (plain OU model: only warp 0 draws it and hands it over through shared memory - it is first needed
    // after the scan's barrier; the oscillation model needs the phase in pass A, every thread draws)
    float x0 = 0.f;
    double phase_turns = 0.0;
    if (OSC || warp == 0) {
        uint32_t r[4];
        philox4x32(0u, (uint32_t)trial, g.particle, g.c3_init, *g.keys, r);
        float z1;
        box_muller(r[0], r[1], x0, z1);
        phase_turns = (double)u01(r[2]);
        if (g.u0) x0 = (float)g.u0[trial];
        x0 = fmaf(x0, oc.x0_mul, oc.x0_add);
        if (OSC && g.phases) phase_turns = g.phases[trial] * 0.15915494309189535;
        if (!OSC && lane == 0) sh->x0 = x0;
    }

    // ---- table of q_j = 1 - a^(j+1) (first trial of the particle in this group) -------
    // Their sums over a chunk, sum p_j and sum p_j^2 (p = 1 - q), enter every thread's share of the trial's sum of
    // squares multiplied by carry^2.  Summed per thread in fp32 they come out with the SAME rounding error in every
    // thread (same table, same order): a systematic relative error of ~2e-7 on the dominant term when the trial's
    // mean is large against its spread - a near-unit-root trial, tau >= T - which the subtraction of n mean^2 amplifies
    // into 2e-6 .. 1e-5 on the normalisation, i.e. on every ACF value of a single-trial particle.  They are
    // therefore summed once per particle in fp64 (warp 0) and combined with the carry in fp64.
    if (!cache.valid) {
        sim_sync(bar_id, NT);                            // nobody still reads the previous particle's table
        for (int j = tid; j < g.chunk; j += NT) sh->qtab[j] = ou_q(oc, j);
        sim_sync(bar_id, NT);
        if (warp == 0) {
            double a1 = 0.0, a2 = 0.0;
            for (int j = lane; j < g.chunk; j += 32) { const double pe = 1.0 - (double)sh->qtab[j]; a1 += pe; a2 += pe * pe; }
            a1 = warp_sum(a1); a2 = warp_sum(a2);
            if (lane == 0) { sh->sp_full = a1; sh->spp_full = a2; }
        }
        sim_sync(bar_id, NT);
        const int n_th = min(max(g.n_t - j0, 0), g.chunk);               // samples this thread owns
        if (n_th == g.chunk) { cache.Sp = sh->sp_full; cache.Spp = sh->spp_full; }
        else {
            double a1 = 0.0, a2 = 0.0;
            for (int j = 0; j < n_th; ++j) { const double pe = 1.0 - (double)sh->qtab[j]; a1 += pe; a2 += pe * pe; }
            cache.Sp = a1; cache.Spp = a2;
        }
    }

    // ---- pass A: local recurrence from a zero state + partial sums -----------------
    // x_j = v_j + cs p_j, p_j = 1 - q_j, cs = (sqrt(coeff) x) carry: v is everything that does not depend on
    // the carry (the local recurrence, plus the oscillation term)
    float y = 0.f;
    float Sy = 0.f, Syy = 0.f, Syu = 0.f;
    float Vy = 0.f, Vyy = 0.f, Vyu = 0.f, Mp = 0.f, Mpp = 0.f;   // valid-only sums; sums of p, p^2 over the MASKED samples (MISSING)
    const double* nz = g.noise ? g.noise + (size_t)trial * g.n_t : nullptr;
    const uint32_t* mrow = MISSING ? g.mask_bits + (size_t)trial * g.mask_words : nullptr;
    const float* qt = sh->qtab;
#ifdef ITJL_V_SIMSLEEP
    if (TC) { while (!tc::mbar_try_wait(sink->wait_bar, sink->wait_parity)) __nanosleep(ITJL_V_SIMSLEEP); }
#else
    if (TC) tc::mbar_wait(sink->wait_bar, sink->wait_parity);
#endif
    if (j0 < g.n_t) {
        // recurrence + partial sums + store of the four samples j .. j+3 driven by z[0..3]; u = q of these samples
        auto advance4 = [&](const int j, const float (&z)[4], const float4 u4, const bool full, const uint32_t mbits) {
            const float u[4] = {u4.x, u4.y, u4.z, u4.w};
            float v[4];
            // oscillation phase of the four samples: the first one reduced to [-0.5, 0.5] turns in fp64, the others
            // by fp32 increments of freq * dt turns, each re-reduced by the 1.5 * 2^23 rounding trick (two FADDs on the
            // FMA pipe; the per-sample fp64 fma / rint / conversions kept the XU pipe busy): <= 1.5e-7 turns off,
            // below MUFU.SIN's own 5e-7
            float turns0 = 0.f;
            if (OSC) {
                double t = fma(oc.fdt, (double)(j + 1), phase_turns);
                t -= rint(t);
                turns0 = (float)t;
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                y = fmaf(-eps, y, y);
                y = fmaf(s, z[i], y);
                v[i] = y;
                if (OSC) {
                    float turns = fmaf((float)i, oc.fdt_f, turns0);
                    turns -= (turns + 12582912.0f) - 12582912.0f;
                    v[i] = fmaf(oc.so, __sinf(6.283185307179586f * turns), oc.sc * y);   // turns in [-0.5, 0.5]: MUFU.SIN at its best (abs error 5e-7)
                }
            }
            if (full && !MISSING) {                      // common case: no per-sample predicates
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    Sy += v[i]; Syy = fmaf(v[i], v[i], Syy); Syu = fmaf(v[i], u[i], Syu);
                }
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    if (j + i < g.n_t) {
                        Sy += v[i]; Syy = fmaf(v[i], v[i], Syy); Syu = fmaf(v[i], u[i], Syu);
                        if (MISSING) {
                            if ((mbits >> i) & 1u) { const float pe = 1.0f - u[i]; Mp += pe; Mpp = fmaf(pe, pe, Mpp); }
                            else { Vy += v[i]; Vyy = fmaf(v[i], v[i], Vyy); Vyu = fmaf(v[i], u[i], Vyu); }
                        }
                    }
                }
            }
            if (TC) {
                unsigned char* pl = (j & 4) ? sink->lo : sink->hi;
                *reinterpret_cast<float4*>(pl + ((j & 127) >> 3) * sink->sbo + (j >> 7) * 16) =
                    make_float4(v[0], v[1], v[2], v[3]);
            } else {
                *reinterpret_cast<float4*>(xs + j) = make_float4(v[0], v[1], v[2], v[3]);
            }
        };
        int q = 0;
        if (WIDE && !MISSING && !nz) {
            for (; q + 16 <= g.chunk && j0 + q + 15 < g.n_t; q += 16) {
                const int j = j0 + q;
                uint32_t r[4][4];
                float z[4][4];
#pragma unroll
                for (int b = 0; b < 4; ++b)
                    philox4x32((uint32_t)(j >> 2) + (uint32_t)b, (uint32_t)trial, g.particle, g.c3_noise, *g.keys, r[b]);
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    box_muller(r[b][0], r[b][1], z[b][0], z[b][1]);
                    box_muller(r[b][2], r[b][3], z[b][2], z[b][3]);
                }
#pragma unroll
                for (int b = 0; b < 4; ++b)
                    advance4(j + 4 * b, z[b], *reinterpret_cast<const float4*>(qt + q + 4 * b), true, 0u);
            }
        }
        if (!MISSING && !nz) {
            // interior of the trial, Philox noise: two counter blocks (eight samples) per iteration in one
            // basic block, so the two seven-round chains and their Box-Muller transforms interleave (four
            // blocks per iteration measured no faster)
            for (; q + 8 <= g.chunk && j0 + q + 7 < g.n_t; q += 8) {
                const int j = j0 + q;
                uint32_t ra[4], rb[4];
                philox4x32((uint32_t)(j >> 2), (uint32_t)trial, g.particle, g.c3_noise, *g.keys, ra);
                philox4x32((uint32_t)(j >> 2) + 1u, (uint32_t)trial, g.particle, g.c3_noise, *g.keys, rb);
                float za[4], zb[4];
                box_muller(ra[0], ra[1], za[0], za[1]);
                box_muller(ra[2], ra[3], za[2], za[3]);
                box_muller(rb[0], rb[1], zb[0], zb[1]);
                box_muller(rb[2], rb[3], zb[2], zb[3]);
                advance4(j, za, *reinterpret_cast<const float4*>(qt + q), true, 0u);
                advance4(j + 4, zb, *reinterpret_cast<const float4*>(qt + q + 4), true, 0u);
            }
        }
        for (; q < g.chunk; q += 4) {
            const int j = j0 + q;
            float z[4] = {0.f, 0.f, 0.f, 0.f};
            uint32_t mbits = 0u;
            const bool full = (j + 3 < g.n_t);
            if (j < g.n_t) {
                if (nz) {
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        if (j + i < g.n_t) z[i] = (float)nz[j + i];
                } else {
                    uint32_t r[4];
                    philox4x32((uint32_t)(j >> 2), (uint32_t)trial, g.particle, g.c3_noise, *g.keys, r);
                    box_muller(r[0], r[1], z[0], z[1]);
                    box_muller(r[2], r[3], z[2], z[3]);
                }
                if (MISSING) mbits = (mrow[j >> 5] >> (j & 31)) & 0xFu;
            }
            advance4(j, z, *reinterpret_cast<const float4*>(qt + q), full, mbits);
        }
    }
    cache.valid = true;
    const double Sp = cache.Sp, Spp = cache.Spp;

    // ---- scan of the chunk-end affine maps x -> (1 - Q) x + B ------------------------
    // composition "earlier (Qp, Bp), then (Q, B)":  B <- Bp - Q Bp + B,  Q <- Q + Qp - Q Qp
    float Q = (j0 < g.n_t) ? qt[g.chunk - 1] : 0.f, B = y;
    float c;                                         // carry into this chunk
    if (!TC) {
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const float Qp = __shfl_up_sync(0xffffffffu, Q, d);
            const float Bp = __shfl_up_sync(0xffffffffu, B, d);
            if (lane >= d) { B = fmaf(-Q, Bp, Bp + B); Q = fmaf(-Q, Qp, Q + Qp); }
        }
        if (lane == 31) { sh->scanA[warp] = Q; sh->scanB[warp] = B; }
        sim_sync(bar_id, NT);
        if (!OSC) x0 = sh->x0;
        float xw = x0;                                   // state entering this warp's first chunk
        for (int w = 0; w < warp; ++w) xw = fmaf(-sh->scanA[w], xw, xw + sh->scanB[w]);
        const float Qprev = __shfl_up_sync(0xffffffffu, Q, 1);
        const float Bprev = __shfl_up_sync(0xffffffffu, B, 1);
        c = (lane == 0) ? xw : fmaf(-Qprev, xw, xw + Bprev);
    } else {
        // chunk order = (m, warp, qt): scan the four quarters of this warp for each m (lanes m, m + 8, m + 16,
        // m + 24), publish the 8 x 4 block maps in chunk order, and let every warp scan that 32-entry table
#pragma unroll
        for (int d = 8; d < 32; d <<= 1) {
            const float Qp = __shfl_up_sync(0xffffffffu, Q, d);
            const float Bp = __shfl_up_sync(0xffffffffu, B, d);
            if (lane >= d) { B = fmaf(-Q, Bp, Bp + B); Q = fmaf(-Q, Qp, Q + Qp); }
        }
        if (lane >= 24) { sh->scanA[4 * (lane - 24) + warp] = Q; sh->scanB[4 * (lane - 24) + warp] = B; }
        sim_sync(bar_id, NT);
        if (!OSC) x0 = sh->x0;
        float TQ = sh->scanA[lane], TB = sh->scanB[lane];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const float Qp = __shfl_up_sync(0xffffffffu, TQ, d);
            const float Bp = __shfl_up_sync(0xffffffffu, TB, d);
            if (lane >= d) { TB = fmaf(-TQ, Bp, Bp + TB); TQ = fmaf(-TQ, Qp, TQ + Qp); }
        }
        const int e = 4 * (lane & 7) + warp;             // my block; state entering it = blocks 0 .. e - 1 applied to x0
        const float PQ = __shfl_sync(0xffffffffu, TQ, (e + 31) & 31);
        const float PB = __shfl_sync(0xffffffffu, TB, (e + 31) & 31);
        const float xw = (e == 0) ? x0 : fmaf(-PQ, x0, x0 + PB);
        const float Qprev = __shfl_up_sync(0xffffffffu, Q, 8);
        const float Bprev = __shfl_up_sync(0xffffffffu, B, 8);
        c = (lane < 8) ? xw : fmaf(-Qprev, xw, xw + Bprev);
    }
    const float cs = OSC ? oc.sc * c : c;            // the oscillation model mixes sqrt(coeff) x OU

    // ---- mean / sum of squares from the partial sums --------------------------------
    const double csd = (double)cs;
    double r0 = fma(csd, Sp, (double)Sy);
    double r1 = fma(csd * csd, Spp, (double)fmaf(2.0f * cs, Sy - Syu, Syy));
    double r2 = 0.0, r3 = 0.0;
    if (MISSING) {
        r2 = fma(csd, Sp - (double)Mp, (double)Vy);
        r3 = fma(csd * csd, Spp - (double)Mpp, (double)fmaf(2.0f * cs, Vy - Vyu, Vyy));
    }
    r0 = warp_sum(r0); r1 = warp_sum(r1);
    if (MISSING) { r2 = warp_sum(r2); r3 = warp_sum(r3); }
    if (lane == 0) {
        sh->sums[warp][0] = r0; sh->sums[warp][1] = r1;
        sh->sums[warp][2] = r2; sh->sums[warp][3] = r3;
    }
    sim_sync(bar_id, NT);
    double SX = 0.0, SXX = 0.0, VX = 0.0, VXX = 0.0;
#pragma unroll
    for (int w = 0; w < NW; ++w) {
        SX += sh->sums[w][0]; SXX += sh->sums[w][1];
        if (MISSING) { VX += sh->sums[w][2]; VXX += sh->sums[w][3]; }
    }
    const double n = (double)g.n_t;
    double m = SX * g.inv_n;
    float rf, miss_val = 0.f;
    double mc;                                       // what is subtracted from every sample
    if (!MISSING) {
        // sums are combined in fp64 (cancellation in SXX - n m^2); the final reciprocal square
        // root is an fp32 rsqrt refined by one Newton step (relative error < 1e-7)
        const float ssf = (float)(SXX - n * m * m);
        float r = rsqrtf(ssf);
        r = r * fmaf(-0.5f * ssf * r, r, 1.5f);
        if (mode == kScaleStd) r *= scale * g.sqrt_nm1;
        else if (mode == kScaleRaw) { r = 1.0f; m = 0.0; }
        mc = m; rf = r;
    } else {
        // comp_ac_time_missing (summary.jl:460-471) applied to the standardised, masked
        // trial: valid z = x - m_all, masked z = 0; m2 = mean of valid z; every entry -= m2.
        const double nv = (double)g.n_valid[trial];
        const double m2 = (double)((float)(VX - nv * m) / (float)nv);
        const double mt = m + m2;
        // masked samples: -m2 in units of the raw trial; the oscillation model (OSC) standardises the
        // trial, scales by data_sd and ADDS data_mean before masking (one_timescale_and_osc_with_missing.jl
        // :292-297 -> ou_process.jl:214), so there they sit at -(m2 + data_mean s_x / data_sd), s_x = the
        // trial's std (n - 1)
        double m2e = m2;
        if (OSC) m2e += (double)g.miss_mean_ratio * sqrt((SXX - n * m * m) / (n - 1.0));
        const float ssf = (float)(VXX - 2.0 * mt * VX + nv * mt * mt + (n - nv) * m2e * m2e);
        float r = rsqrtf(ssf);
        r = r * fmaf(-0.5f * ssf * r, r, 1.5f);
        mc = mt; rf = r; miss_val = (float)(-m2e) * r;
    }

    // ---- pass B: out = (v + cs - cs q - m) r + shift = v r - q (cs r) + ((cs - m) r + shift) ----------
    if (TC) {
        // the operands carry the power-of-two factor `scale`: fold it into the affine map
        const float sc = sink->scale;
        rf *= sc; miss_val *= sc; shift *= sc;
    }
    if (j0 < g.n_t) {
        const float ncrf = -cs * rf;
        const float k0 = fmaf((float)((double)cs - mc), rf, shift);
        constexpr int STEP = TC ? 8 : 4;                 // TC: chunk is a multiple of 8
        for (int q = 0; q < g.chunk; q += STEP) {
            const int j = j0 + q;
            float v[STEP], u[STEP];
            const int off = TC ? ((j & 127) >> 3) * sink->sbo + (j >> 7) * 16 : 0;
            if (TC) {
                const float4 a4 = *reinterpret_cast<const float4*>(sink->hi + off);
                const float4 b4 = *reinterpret_cast<const float4*>(sink->lo + off);
                v[0] = a4.x; v[1] = a4.y; v[2] = a4.z; v[3] = a4.w;
                v[STEP - 4] = b4.x; v[STEP - 3] = b4.y; v[STEP - 2] = b4.z; v[STEP - 1] = b4.w;
            } else {
                const float4 v4 = *reinterpret_cast<const float4*>(xs + j);
                v[0] = v4.x; v[1] = v4.y; v[2] = v4.z; v[3] = v4.w;
            }
#pragma unroll
            for (int h = 0; h < STEP; h += 4) {
                const float4 u4 = *reinterpret_cast<const float4*>(qt + q + h);
                u[h] = u4.x; u[h + 1] = u4.y; u[h + 2] = u4.z; u[h + 3] = u4.w;
            }
#pragma unroll
            for (int i = 0; i < STEP; ++i) v[i] = fmaf(v[i], rf, fmaf(u[i], ncrf, k0));
            if (MISSING || j + STEP - 1 >= g.n_t) {
                uint32_t mbits = 0u;
                if (MISSING && j < g.n_t) mbits = (mrow[j >> 5] >> (j & 31)) & (STEP == 8 ? 0xFFu : 0xFu);
#pragma unroll
                for (int i = 0; i < STEP; ++i) {
                    if (MISSING && ((mbits >> i) & 1u)) v[i] = miss_val;
                    if (j + i >= g.n_t) v[i] = 0.f;
                }
            }
            if (!TC) {
                if (g.win) {
                    const float4 w4 = __ldg(reinterpret_cast<const float4*>(g.win + j));
                    v[0] *= w4.x; v[1] *= w4.y; v[2] *= w4.z; v[3] *= w4.w;
                }
#pragma unroll
                for (int h = 0; h < STEP; h += 4)
                    *reinterpret_cast<float4*>(xs + j + h) = make_float4(v[h], v[h + 1], v[h + 2], v[h + 3]);
            }
            if (TC) {
                uint32_t hp[4], lp[4];
                if (sink->write_lo) {
#pragma unroll
                    for (int h = 0; h < 4; ++h) {
                        const __half2 hh = __floats2half2_rn(v[2 * h], v[2 * h + 1]);
                        const float2 hf = __half22float2(hh);
                        const __half2 ll = __floats2half2_rn(v[2 * h] - hf.x, v[2 * h + 1] - hf.y);
                        hp[h] = *reinterpret_cast<const uint32_t*>(&hh);
                        lp[h] = *reinterpret_cast<const uint32_t*>(&ll);
                    }
                    *reinterpret_cast<uint4*>(sink->hi + off) = make_uint4(hp[0], hp[1], hp[2], hp[3]);
                    *reinterpret_cast<uint4*>(sink->lo + off) = make_uint4(lp[0], lp[1], lp[2], lp[3]);
                } else {
#pragma unroll
                    for (int h = 0; h < 4; ++h) {
                        const __half2 hh = __floats2half2_rn(v[2 * h], v[2 * h + 1]);
                        hp[h] = *reinterpret_cast<const uint32_t*>(&hh);
                    }
                    *reinterpret_cast<uint4*>(sink->hi + off) = make_uint4(hp[0], hp[1], hp[2], hp[3]);
                }
            }
        }
    } else if (!TC) {
        for (int q = 0; q < g.chunk; q += 4)
            *reinterpret_cast<float4*>(xs + j0 + q) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

}  // namespace itjl
