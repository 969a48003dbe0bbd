// Mixed-radix (2, 3, 4, 5, 7) complex FFT of a 7-smooth length in shared memory: the transform behind
// DSP.jl's periodogram (reference src/stats/summary.jl:146-155), whose nfft = nextfastfft(n) = the next
// 7-smooth integer (5000 for the README's 10 s at 500 Hz), not a power of two.
//
// Stockham autosort, decimation in frequency: a stage with sub-transform length `len` = r m and stride s
// (s len = n) maps
//     y[q + s (r p + j)] = w_len^(p j) * sum_k x[q + s (p + m k)] W_r^(j k),   p < m, q < s, j < r,
// ping-ponging between two buffers; after the last stage (len = 1) the result is in natural order.
// Twiddles w_len^(p j) = exp(-2 pi i p j s / n) come from one table of the n-th roots of unity.
// The same code runs on the host (`__host__ __device__`): benchmarks/micro/fft_mixed_host_test.cu checks it
// against a direct DFT for every 7-smooth length up to 12 000.
#pragma once
#include <cuda_runtime.h>

#ifdef __CUDA_ARCH__
#define ITJL_UNROLL _Pragma("unroll")
#else
#define ITJL_UNROLL
#endif

namespace itjl {

constexpr int kFftMaxStages = 24;

struct FftPlan {
    int n;
    int n_stages;
    int radix[kFftMaxStages];
    unsigned magic[kFftMaxStages];      // ceil(2^32 / s) of the stage's stride s (idx / s by one multiply; n <= 65536)
};

// 7-smooth factorisation into radices 4, 2, 3, 5, 7 (returns false when n has another prime factor)
inline bool fft_plan_make(int n, FftPlan* plan) {
    plan->n = n; plan->n_stages = 0;
    int m = n;
    const int primes[4] = {2, 3, 5, 7};
    int count[4] = {0, 0, 0, 0};
    for (int i = 0; i < 4; ++i) while (m % primes[i] == 0) { m /= primes[i]; ++count[i]; }
    if (m != 1 || n < 1) return false;
    for (; count[0] >= 2; count[0] -= 2) plan->radix[plan->n_stages++] = 4;
    if (count[0]) plan->radix[plan->n_stages++] = 2;
    for (int i = 1; i < 4; ++i) for (int c = 0; c < count[i]; ++c) plan->radix[plan->n_stages++] = primes[i];
    if (plan->n_stages > kFftMaxStages || n > 65536) return false;
    long long st = 1;
    for (int k = 0; k < plan->n_stages; ++k) {
        plan->magic[k] = st == 1 ? 0u : (unsigned)(((1ll << 32) + st - 1) / st);
        st *= plan->radix[k];
    }
    return true;
}

inline int next_fast_len(int n) {            // DSP.nextfastfft(n) = nextprod([2, 3, 5, 7], n)
    for (int m = n < 1 ? 1 : n;; ++m) {
        int r = m;
        for (int p : {2, 3, 5, 7}) while (r % p == 0) r /= p;
        if (r == 1) return m;
    }
}

template <typename T> struct Cx { T x, y; };
template <typename T> __host__ __device__ __forceinline__ Cx<T> cx_mul(Cx<T> a, Cx<T> b) { return {a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x}; }
template <typename T> __host__ __device__ __forceinline__ Cx<T> cx_add(Cx<T> a, Cx<T> b) { return {a.x + b.x, a.y + b.y}; }
template <typename T> __host__ __device__ __forceinline__ Cx<T> cx_sub(Cx<T> a, Cx<T> b) { return {a.x - b.x, a.y - b.y}; }

// r-point DFT in place (forward, exp(-2 pi i j k / r)); r in {2, 3, 4, 5, 7}
template <typename T, int R>
__host__ __device__ __forceinline__ void small_dft(Cx<T> (&a)[R]) {
    if constexpr (R == 2) {
        const Cx<T> t = a[1];
        a[1] = cx_sub(a[0], t); a[0] = cx_add(a[0], t);
    } else if constexpr (R == 4) {
        const Cx<T> s02 = cx_add(a[0], a[2]), d02 = cx_sub(a[0], a[2]), s13 = cx_add(a[1], a[3]), d13 = cx_sub(a[1], a[3]);
        a[0] = cx_add(s02, s13); a[2] = cx_sub(s02, s13);
        a[1] = {d02.x + d13.y, d02.y - d13.x};              // d02 - i d13
        a[3] = {d02.x - d13.y, d02.y + d13.x};              // d02 + i d13
    } else {
        // odd prime: symmetric pairs, cos / sin of 2 pi k / R as compile-time literals
        constexpr int H = (R - 1) / 2;
        // cos / sin of 2 pi k / R, k = 1 .. H (only the entries of this R are used; indices are compile-time after unrolling)
        const T c[3] = {(T)(R == 3 ? -0.5 : R == 5 ? 0.30901699437494742410 : 0.62348980185873353053),
                        (T)(R == 5 ? -0.80901699437494742410 : -0.22252093395631440429), (T)-0.90096886790241912624};
        const T s[3] = {(T)(R == 3 ? 0.86602540378443864676 : R == 5 ? 0.95105651629515357212 : 0.78183148246802980871),
                        (T)(R == 5 ? 0.58778525229247312917 : 0.97492791218182360702), (T)0.43388373911755812048};
        Cx<T> p[H], m[H];
ITJL_UNROLL
        for (int k = 0; k < H; ++k) { p[k] = cx_add(a[k + 1], a[R - 1 - k]); m[k] = cx_sub(a[k + 1], a[R - 1 - k]); }
        Cx<T> out[R];
        out[0] = a[0];
ITJL_UNROLL
        for (int k = 0; k < H; ++k) out[0] = cx_add(out[0], p[k]);
ITJL_UNROLL
        for (int j = 1; j <= H; ++j) {
            T re = a[0].x, im = a[0].y, sr = 0, si = 0;
ITJL_UNROLL
            for (int k = 1; k <= H; ++k) {
                const int idx = (j * k) % R;                 // cos is even, sin odd in idx -> R - idx
                const T ck = idx <= H ? c[idx - 1] : c[R - idx - 1];
                const T sk = idx <= H ? s[idx - 1] : -s[R - idx - 1];
                re += ck * p[k - 1].x; im += ck * p[k - 1].y;
                sr += sk * m[k - 1].y; si += sk * m[k - 1].x;
            }
            // X_j = sum_k a_k (cos - i sin)(2 pi j k / R):  real: re + sr, imag: im - si;  X_{R-j} is the conjugate combination
            out[j] = {re + sr, im - si};
            out[R - j] = {re - sr, im + si};
        }
ITJL_UNROLL
        for (int j = 0; j < R; ++j) a[j] = out[j];
    }
}

// the n-th roots of unity exp(-2 pi i t / n) as a plain table ...
template <typename T> struct RootsTable {
    const Cx<T>* tab;
    __host__ __device__ __forceinline__ Cx<T> operator()(int t) const { return tab[t]; }
};
// ... or as a two-level table root(t) = A[t >> 6] * B[t & 63] (ceil(n / 64) + 64 entries: small enough for shared
// memory beside the two transform buffers)
constexpr int kRootLoBits = 6;
template <typename T> struct RootsTwoLevel {
    const Cx<T>* A;
    const Cx<T>* B;
    __host__ __device__ __forceinline__ Cx<T> operator()(int t) const { return cx_mul(A[t >> kRootLoBits], B[t & ((1 << kRootLoBits) - 1)]); }
};
__host__ __device__ __forceinline__ int roots_two_level_entries(int n) { return ((n + (1 << kRootLoBits) - 1) >> kRootLoBits) + (1 << kRootLoBits); }

// idx / s for idx, s < 65536 with magic = ceil(2^32 / s) (0 for s = 1)
__host__ __device__ __forceinline__ int fft_div(int idx, unsigned magic) {
#ifdef __CUDA_ARCH__
    return magic ? (int)__umulhi((unsigned)idx, magic) : idx;
#else
    return magic ? (int)(((unsigned long long)(unsigned)idx * magic) >> 32) : idx;
#endif
}

// one Stockham stage of radix R executed by `nt` cooperating threads (thread `tid`); caller synchronises after it
template <typename T, int R, typename Roots>
__host__ __device__ __forceinline__ void fft_stage(const Cx<T>* __restrict__ x, Cx<T>* __restrict__ y, int n, int len, int s,
                                                   unsigned magic, const Roots& roots, int tid, int nt) {
    const int m = len / R;
    for (int idx = tid; idx < n / R; idx += nt) {
        const int p = fft_div(idx, magic), q = idx - p * s;
        Cx<T> a[R];
ITJL_UNROLL
        for (int k = 0; k < R; ++k) a[k] = x[q + s * (p + m * k)];
        small_dft<T, R>(a);
        y[q + s * (R * p)] = a[0];
        const int step = p * s;                              // w_len^p = roots[p s]
        int t = 0;
ITJL_UNROLL
        for (int j = 1; j < R; ++j) {
            t += step; if (t >= n) t -= n;
            y[q + s * (R * p + j)] = p == 0 ? a[j] : cx_mul(a[j], roots(t));
        }
    }
}

template <typename T, typename Roots>
__host__ __device__ __forceinline__ void fft_stage_any(int radix, const Cx<T>* x, Cx<T>* y, int n, int len, int s,
                                                       unsigned magic, const Roots& roots, int tid, int nt) {
    switch (radix) {
        case 2: fft_stage<T, 2>(x, y, n, len, s, magic, roots, tid, nt); break;
        case 3: fft_stage<T, 3>(x, y, n, len, s, magic, roots, tid, nt); break;
        case 4: fft_stage<T, 4>(x, y, n, len, s, magic, roots, tid, nt); break;
        case 5: fft_stage<T, 5>(x, y, n, len, s, magic, roots, tid, nt); break;
        default: fft_stage<T, 7>(x, y, n, len, s, magic, roots, tid, nt); break;
    }
}

#ifdef __CUDACC__
// whole transform by one CTA: input in `a`, scratch `b`; returns the buffer that holds the result (natural order)
template <typename T, typename Roots>
__device__ __forceinline__ Cx<T>* fft_mixed_cta(const FftPlan& plan, Cx<T>* a, Cx<T>* b, const Roots& roots) {
    int len = plan.n, s = 1;
    Cx<T>* x = a; Cx<T>* y = b;
    for (int st = 0; st < plan.n_stages; ++st) {
        const int r = plan.radix[st];
        fft_stage_any<T>(r, x, y, plan.n, len, s, plan.magic[st], roots, threadIdx.x, blockDim.x);
        __syncthreads();
        len /= r; s *= r;
        Cx<T>* t = x; x = y; y = t;
    }
    return x;
}
#endif

}  // namespace itjl
