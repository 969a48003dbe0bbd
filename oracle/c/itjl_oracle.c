/*
 * CPU oracle (C, fp64) for the IntrinsicTimescales.jl aABC hot path.
 *
 * TEST INFRASTRUCTURE ONLY - used by tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py.  Never linked or called by the
 * product library (libitjl_b200.so).
 *
 * It restates, in the reference's own arithmetic (fp64, FFT-based ACF; the FFT itself is a vectorised
 * radix-2 with two real trials per complex transform, where the reference hands FFTW one complex transform
 * per real trial - the baseline should not be slower than what it stands for), the body
 * of the serial loop in /root/reference/src/core/abc.jl:174-195:
 *   generate_data_and_reduce          src/core/model.jl:146-151
 *   generate_ou_process(_sciml)       src/utils/ou_process.jl:46-62, 114-148
 *   generate_ou_with_oscillation      src/utils/ou_process.jl:176-218
 *   comp_ac_fft                       src/stats/summary.jl:46-63
 *   comp_ac_time_missing              src/stats/summary.jl:455-484
 *   comp_psd_adfriendly               src/stats/summary.jl:204-235
 *   linear_/logarithmic_distance      src/stats/distances.jl:29-31, 51-53
 * The SDE solve is restated as the exact OU transition (see oracle/ou.py header);
 * the noise stream is the Philox4x32-7 stream specified in oracle/philox.py.
 * PARITY: unpinned against Julia itself (not installable here); pinned against the
 * NumPy oracle, which is pinned against the reference's KATs (tests/).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <stdatomic.h>
#include <unistd.h>

#define PI 3.14159265358979323846

/* ------------------------------------------------------------ Philox ---- */
#define PHILOX_ROUNDS 7     /* oracle/philox.py: ROUNDS */
static void philox4x32(uint32_t c[4], uint32_t k0, uint32_t k1) {
    for (int r = 0; r < PHILOX_ROUNDS; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
        c[0] = n0; c[1] = (uint32_t)p1; c[2] = n2; c[3] = (uint32_t)p0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}
static inline double u01(uint32_t r) { return ((double)(r >> 9) + 0.5) * (1.0 / 8388608.0); }
static inline void box_muller(uint32_t ra, uint32_t rb, double* z0, double* z1) {
    double rad = sqrt(-2.0 * log(u01(ra)));
    double ang = 2.0 * PI * u01(rb);
    *z0 = rad * cos(ang); *z1 = rad * sin(ang);
}
static inline uint32_t c3_word(uint32_t stream, uint64_t gen) {
    return ((stream & 0xFu) << 28) | (uint32_t)(gen & 0x0FFFFFFFu);
}

/* --------------------------------------------------------------- FFT ---- */
/* In-place radix-2 decimation-in-time complex FFT, written so that the CPU baseline is not a straw man:
 * the twiddles of every stage lie contiguously (tw + half - 1), so the butterfly loop of a stage is
 * stride-1 in all six arrays and gcc vectorises it (AVX2 / AVX-512 with -march=native); the first two
 * stages (twiddles 1 and -i) are fused into one multiplication-free radix-4 pass.  Real input is exploited
 * by the callers: two trials ride in one complex transform (see pair_acf / pair_psd). */
typedef struct { int n; double* wr; double* wi; int* rev; } fft_plan;

static fft_plan* fft_plan_create(int n) {
    fft_plan* p = (fft_plan*)malloc(sizeof(fft_plan));
    p->n = n;
    p->wr = (double*)malloc(sizeof(double) * (size_t)(n > 1 ? n : 2));
    p->wi = (double*)malloc(sizeof(double) * (size_t)(n > 1 ? n : 2));
    p->rev = (int*)malloc(sizeof(int) * (size_t)n);
    for (int half = 1; half < n; half <<= 1)              /* stage with butterflies of span `half` */
        for (int j = 0; j < half; ++j) {
            p->wr[half - 1 + j] = cos(-PI * (double)j / (double)half);
            p->wi[half - 1 + j] = sin(-PI * (double)j / (double)half);
        }
    int bits = 0; while ((1 << bits) < n) ++bits;
    for (int i = 0; i < n; ++i) {
        int r = 0;
        for (int b = 0; b < bits; ++b) if (i & (1 << b)) r |= 1 << (bits - 1 - b);
        p->rev[i] = r;
    }
    return p;
}
static void fft_plan_destroy(fft_plan* p) { free(p->wr); free(p->wi); free(p->rev); free(p); }

/* in-place forward complex FFT (inverse = conj trick by caller) */
static void fft_exec(const fft_plan* p, double* restrict re, double* restrict im) {
    const int n = p->n;
    for (int i = 0; i < n; ++i) {
        int r = p->rev[i];
        if (r > i) { double t = re[i]; re[i] = re[r]; re[r] = t; t = im[i]; im[i] = im[r]; im[r] = t; }
    }
    int len = 2;
    if (n >= 4) {
        /* stages len = 2 and len = 4 in one pass: twiddles 1, 1, 1, -i */
        for (int s = 0; s < n; s += 4) {
            double ar = re[s] + re[s + 1], ai = im[s] + im[s + 1];
            double br = re[s] - re[s + 1], bi = im[s] - im[s + 1];
            double cr = re[s + 2] + re[s + 3], ci = im[s + 2] + im[s + 3];
            double dr = re[s + 2] - re[s + 3], di = im[s + 2] - im[s + 3];
            re[s] = ar + cr; im[s] = ai + ci;
            re[s + 2] = ar - cr; im[s + 2] = ai - ci;
            re[s + 1] = br + di; im[s + 1] = bi - dr;       /* (dr + i di) * (-i) = di - i dr */
            re[s + 3] = br - di; im[s + 3] = bi + dr;
        }
        len = 8;
    }
    for (; len <= n; len <<= 1) {
        const int half = len >> 1;
        const double* restrict wr = p->wr + (half - 1);
        const double* restrict wi = p->wi + (half - 1);
        for (int s = 0; s < n; s += len) {
            double* restrict ra = re + s; double* restrict ia = im + s;
            double* restrict rb = re + s + half; double* restrict ib = im + s + half;
            for (int j = 0; j < half; ++j) {
                double xr = rb[j] * wr[j] - ib[j] * wi[j], xi = rb[j] * wi[j] + ib[j] * wr[j];
                rb[j] = ra[j] - xr; ib[j] = ia[j] - xi;
                ra[j] += xr; ia[j] += xi;
            }
        }
    }
}

static int nextpow2(int n) { int p = 1; while (p < n) p <<= 1; return p; }

/* ------------------------------------------------------------ model ----- */
typedef struct {
    int32_t kind;        /* 0 plain OU, 1 OU + oscillation */
    int32_t summary;     /* 0 ACF (comp_ac_fft), 1 ACF missing, 2 PSD hamming */
    int32_t distance;    /* 0 linear, 1 logarithmic */
    int32_t update;      /* 0 exact, 1 Euler-Maruyama */
    int64_t n_trials, n_t;
    double dt, data_mean, data_sd;
    int64_t n_stats;
    const double* target_stats;
    const uint8_t* missing_mask;   /* column-major (trials x n_t), 1 = missing */
    const int32_t* freq_bins;      /* 0-based indices into psd[2:n2/2] */
} oracle_model_desc;

typedef struct {
    double *x, *re, *im, *acc, *win;
    fft_plan* plan;
    int n_fft;
    double win_ss;
    int pending;          /* 1: the real part of (re, im) holds a trial waiting for its partner */
} workspace;

static workspace* ws_create(const oracle_model_desc* m) {
    workspace* w = (workspace*)calloc(1, sizeof(workspace));
    int n = (int)m->n_t;
    w->n_fft = (m->summary == 2) ? 2 * nextpow2(n) : nextpow2(2 * n - 1);
    w->plan = fft_plan_create(w->n_fft);
    w->x = (double*)malloc(sizeof(double) * n);
    w->re = (double*)malloc(sizeof(double) * w->n_fft);
    w->im = (double*)malloc(sizeof(double) * w->n_fft);
    int nacc = (m->summary == 2) ? w->n_fft / 2 : (int)m->n_stats;
    w->acc = (double*)malloc(sizeof(double) * (nacc > 0 ? nacc : 1));
    w->win = (double*)malloc(sizeof(double) * n);
    w->win_ss = 0.0;
    for (int i = 0; i < n; ++i) {
        w->win[i] = 0.54 - 0.46 * cos(2.0 * PI * i / (n - 1));
        w->win_ss += w->win[i] * w->win[i];
    }
    return w;
}
static void ws_destroy(workspace* w) {
    fft_plan_destroy(w->plan); free(w->x); free(w->re); free(w->im); free(w->acc); free(w->win); free(w);
}

/* Two real sequences x1 (in re) and x2 (in im), zero padded to n_fft, transformed at once:
 *   Z = FFT(x1 + i x2),  X1[k] = (Z[k] + conj Z[N-k]) / 2,  X2[k] = (Z[k] - conj Z[N-k]) / (2 i).
 * `count` = 1 when im is all zero (a last unpaired trial). */
static void pair_acf(workspace* w, int n_stats, int count) {
    const int N = w->n_fft;
    double* restrict re = w->re; double* restrict im = w->im;
    fft_exec(w->plan, re, im);
    /* power spectra P1, P2 (real, even in k) packed as W = P1 + i P2; FFT(W) = N (acov1 + i acov2): the
       forward transform of a real even sequence equals N times its inverse (summary.jl:56-58) */
    for (int k = 0; k <= N / 2; ++k) {
        const int nk = (N - k) & (N - 1);
        const double zr = re[k], zi = im[k], yr = re[nk], yi = im[nk];
        const double p1 = 0.25 * ((zr + yr) * (zr + yr) + (zi - yi) * (zi - yi));
        const double p2 = 0.25 * ((zi + yi) * (zi + yi) + (yr - zr) * (yr - zr));
        re[k] = p1; im[k] = p2; re[nk] = p1; im[nk] = p2;
    }
    fft_exec(w->plan, re, im);
    const double inv1 = 1.0 / re[0];
    for (int k = 0; k < n_stats; ++k) w->acc[k] += re[k] * inv1;
    if (count == 2) {
        const double inv2 = 1.0 / im[0];
        for (int k = 0; k < n_stats; ++k) w->acc[k] += im[k] * inv2;
    }
}

static void pair_psd(workspace* w, double scl, int count) {
    const int N = w->n_fft;
    double* restrict re = w->re; double* restrict im = w->im;
    fft_exec(w->plan, re, im);
    for (int k = 0; k < N / 2; ++k) {
        const int nk = (N - k) & (N - 1);
        const double zr = re[k], zi = im[k], yr = re[nk], yi = im[nk];
        const double p1 = 0.25 * ((zr + yr) * (zr + yr) + (zi - yi) * (zi - yi));
        const double p2 = 0.25 * ((zi + yi) * (zi + yi) + (yr - zr) * (yr - zr));
        w->acc[k] += (count == 2 ? p1 + p2 : p1) * scl;
    }
}

/* one particle: returns the distance; stats_out (optional, n_stats) receives the trial-mean
 * summary statistic the distance was computed from */
static double particle_distance(const oracle_model_desc* m, workspace* w, const double* theta,
                                uint64_t seed, uint64_t gen, uint64_t particle,
                                const double* u0_in, const double* noise_in, const double* phase_in,
                                double* stats_out) {
    const int n = (int)m->n_t, ntr = (int)m->n_trials;
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    const double tau = theta[0];
    double a, s;
    if (m->update == 0) { a = exp(-m->dt / tau); s = sqrt(tau / 2.0 * (1.0 - a * a)); }
    else { a = 1.0 - m->dt / tau; s = sqrt(m->dt); }
    double freq = 0.0, coeff = 1.0;
    if (m->kind == 1) {
        freq = theta[1]; coeff = theta[2];
        if (coeff < 0.0) coeff = 1e-4; else if (coeff > 1.0) coeff = 1.0 - 1e-4;
    }
    const int nacc = (m->summary == 2) ? w->n_fft / 2 : (int)m->n_stats;
    for (int i = 0; i < nacc; ++i) w->acc[i] = 0.0;
    int bad = 0;

    for (int tr = 0; tr < ntr; ++tr) {
        double u0, phase;
        {
            uint32_t c[4] = {0u, (uint32_t)tr, (uint32_t)particle, c3_word(1, gen)};
            philox4x32(c, k0, k1);
            double z1;
            box_muller(c[0], c[1], &u0, &z1);
            phase = 2.0 * PI * u01(c[2]);
        }
        if (u0_in) u0 = u0_in[tr];
        if (phase_in) phase = phase_in[tr];
        double prev = u0;
        for (int j = 0; j < n; j += 4) {
            double z[4] = {0.0, 0.0, 0.0, 0.0};
            if (noise_in) {
                for (int q = 0; q < 4 && j + q < n; ++q) z[q] = noise_in[(size_t)tr * n + j + q];
            } else {
                uint32_t c[4] = {(uint32_t)(j >> 2), (uint32_t)tr, (uint32_t)particle, c3_word(0, gen)};
                philox4x32(c, k0, k1);
                box_muller(c[0], c[1], &z[0], &z[1]);
                box_muller(c[2], c[3], &z[2], &z[3]);
            }
            for (int q = 0; q < 4 && j + q < n; ++q) { prev = a * prev + s * z[q]; w->x[j + q] = prev; }
        }
        if (m->kind == 1) {
            double sc = sqrt(coeff), so = sqrt(1.0 - coeff) * sqrt(2.0);
            for (int j = 0; j < n; ++j) {
                double t = m->dt * (double)(j + 1);
                w->x[j] = so * sin(phase + 2.0 * PI * freq * t) + sc * w->x[j];
            }
        }
        /* standardise: (x - mean) / std(n-1) * scale (+ mean for the osc model) */
        double mean = 0.0;
        for (int j = 0; j < n; ++j) mean += w->x[j];
        mean /= n;
        double ss = 0.0;
        for (int j = 0; j < n; ++j) { double d = w->x[j] - mean; ss += d * d; }
        double sd = sqrt(ss / (n - 1));
        double scale = m->data_sd / sd;
        double shift = (m->kind == 1) ? m->data_mean : 0.0;
        for (int j = 0; j < n; ++j) w->x[j] = (w->x[j] - mean) * scale + shift;
        for (int j = 0; j < n; ++j) if (!isfinite(w->x[j])) { bad = 1; break; }
        if (bad) break;

        if (m->summary == 0) {
            /* comp_ac_fft, two trials per complex transform: trial `pending` rides in the real part, this
               one in the imaginary part; a last unpaired trial is flushed after the loop */
            double m2 = 0.0;
            for (int j = 0; j < n; ++j) m2 += w->x[j];
            m2 /= n;
            double* dst = w->pending ? w->im : w->re;
            for (int j = 0; j < n; ++j) dst[j] = w->x[j] - m2;
            for (int j = n; j < w->n_fft; ++j) dst[j] = 0.0;
            if (w->pending) { pair_acf(w, (int)m->n_stats, 2); w->pending = 0; }
            else w->pending = 1;
        } else if (m->summary == 1) {
            /* comp_ac_time_missing with the model's mask applied (NaN -> 0, then
               subtract the valid-sample mean from every entry) */
            double sum = 0.0; int nv = 0;
            for (int j = 0; j < n; ++j) {
                int miss = m->missing_mask ? m->missing_mask[(size_t)j * ntr + tr] : 0;
                if (miss) w->x[j] = 0.0; else { sum += w->x[j]; ++nv; }
            }
            double xm = sum / nv;
            double ssq = 0.0;
            for (int j = 0; j < n; ++j) { w->x[j] -= xm; ssq += w->x[j] * w->x[j]; }
            for (int k = 0; k < (int)m->n_stats; ++k) {
                double acc = 0.0;
                for (int j = 0; j + k < n; ++j) acc += w->x[j] * w->x[j + k];
                w->acc[k] += acc / ssq;
            }
        } else {
            /* comp_psd_adfriendly, two trials per complex transform */
            double m2 = 0.0;
            for (int j = 0; j < n; ++j) m2 += w->x[j];
            m2 /= n;
            double* dst = w->pending ? w->im : w->re;
            for (int j = 0; j < n; ++j) dst[j] = (w->x[j] - m2) * w->win[j];
            for (int j = n; j < w->n_fft; ++j) dst[j] = 0.0;
            const double scl = 1.0 / ((1.0 / m->dt) * w->win_ss);
            if (w->pending) { pair_psd(w, scl, 2); w->pending = 0; }
            else w->pending = 1;
        }
    }
    if (!bad && w->pending) {                  /* odd number of trials: the last one alone (imaginary part zero) */
        for (int j = 0; j < w->n_fft; ++j) w->im[j] = 0.0;
        if (m->summary == 0) pair_acf(w, (int)m->n_stats, 1);
        else pair_psd(w, 1.0 / ((1.0 / m->dt) * w->win_ss), 1);
    }
    w->pending = 0;
    if (bad) {
        if (stats_out) for (int k = 0; k < (int)m->n_stats; ++k) stats_out[k] = NAN;
        return NAN;
    }
    /* trial means in place (freq_bins is increasing, so acc[bins[k] + 1] is read before acc[k] is
       overwritten); the distance loop below is the same code with and without stats_out */
    for (int k = 0; k < (int)m->n_stats; ++k)
        w->acc[k] = ((m->summary == 2) ? w->acc[m->freq_bins[k] + 1] : w->acc[k]) / ntr;   /* +1: DC dropped */
    if (stats_out) memcpy(stats_out, w->acc, sizeof(double) * (size_t)m->n_stats);
    double d = 0.0;
    for (int k = 0; k < (int)m->n_stats; ++k) {
        double v = w->acc[k];
        double e = (m->distance == 0) ? (v - m->target_stats[k]) : (log(v) - log(m->target_stats[k]));
        d += e * e;
    }
    return d / (double)m->n_stats;
}

/* theta: column-major (n_particles x p) like a Julia Matrix.  u0/phases: (particle, trial)
 * row-major; noise: (particle, trial, n_t) row-major; any of them may be NULL (Philox).
 * Particles are independent; they are spread over n_threads pthreads (dynamic chunks of 1). */
typedef struct {
    const oracle_model_desc* m; const double* theta; int64_t n_particles; int32_t n_theta;
    uint64_t seed, generation; int64_t particle_offset;
    const double *u0, *noise, *phases; double* dist_out; atomic_llong* next;
    double* stats_out;                       /* optional (n_particles, n_stats) row-major */
} job_t;

static void* worker(void* arg) {
    job_t* j = (job_t*)arg;
    const oracle_model_desc* m = j->m;
    workspace* w = ws_create(m);
    for (;;) {
        int64_t i = (int64_t)atomic_fetch_add(j->next, 1);
        if (i >= j->n_particles) break;
        double th[8] = {0.0};
        for (int q = 0; q < j->n_theta && q < 8; ++q) th[q] = j->theta[(size_t)q * j->n_particles + i];
        size_t pt = (size_t)i * m->n_trials;
        j->dist_out[i] = particle_distance(m, w, th, j->seed, j->generation,
                                           (uint64_t)(j->particle_offset + i),
                                           j->u0 ? j->u0 + pt : NULL,
                                           j->noise ? j->noise + pt * m->n_t : NULL,
                                           j->phases ? j->phases + pt : NULL,
                                           j->stats_out ? j->stats_out + (size_t)i * m->n_stats : NULL);
    }
    ws_destroy(w);
    return NULL;
}

int itjl_oracle_max_threads(void) {
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

int itjl_oracle_abc_stats(const oracle_model_desc* m, const double* theta, int64_t n_particles,
                          int32_t n_theta, uint64_t seed, uint64_t generation,
                          int64_t particle_offset, const double* u0, const double* noise,
                          const double* phases, double* dist_out, double* stats_out, int32_t n_threads) {
    if (n_threads <= 0) n_threads = itjl_oracle_max_threads();
    if (n_threads > 256) n_threads = 256;
    if ((int64_t)n_threads > n_particles) n_threads = (int32_t)(n_particles > 0 ? n_particles : 1);
    atomic_llong next = 0;
    job_t job = {m, theta, n_particles, n_theta, seed, generation, particle_offset,
                 u0, noise, phases, dist_out, &next, stats_out};
    if (n_threads == 1) { worker(&job); return 0; }
    pthread_t tid[256];
    for (int t = 0; t < n_threads; ++t)
        if (pthread_create(&tid[t], NULL, worker, &job) != 0) return -1;
    for (int t = 0; t < n_threads; ++t) pthread_join(tid[t], NULL);
    return 0;
}

int itjl_oracle_abc_distances(const oracle_model_desc* m, const double* theta, int64_t n_particles,
                              int32_t n_theta, uint64_t seed, uint64_t generation,
                              int64_t particle_offset, const double* u0, const double* noise,
                              const double* phases, double* dist_out, int32_t n_threads) {
    return itjl_oracle_abc_stats(m, theta, n_particles, n_theta, seed, generation, particle_offset,
                                 u0, noise, phases, dist_out, NULL, n_threads);
}
