// Shared-memory power-of-two FFT building blocks of the fused PSD kernels (psd_kernels.cu, psd2_kernels.cu):
// two-level twiddle table, padded in-place DIF layout, radix-4 / radix-16 register butterflies, the storage
// permutation and the real-FFT recombination of one bin.
#pragma once
#include <cuda_runtime.h>

namespace itjl {

// U[k] = exp(-2 pi i k / n2), k in [0, n2/2)
__device__ __forceinline__ float2 tw_lookup(const float2* __restrict__ U, int idx, int N) {
    // idx in [0, 2N): exp(-2 pi i idx / (2N)); second half by symmetry
    if (idx >= N) { const float2 w = __ldg(U + (idx - N)); return make_float2(-w.x, -w.y); }
    return __ldg(U + idx);
}
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}

// Two-level twiddle table in shared memory: exp(-2 pi i idx / 2N) = A[idx >> 7] * B[idx & 127]
// (2N/128 + 128 entries, 3 KB at N = 16384).  The butterflies of the large stages read twiddles all over
// the 2N-entry table; from global memory those loads were the kernel's top stall (L2 latency - the
// FFT buffer leaves no L1).
constexpr int kTwLoBits = 7;
__host__ __device__ __forceinline__ int tw2_entries(int N) {
    const int hi = (2 * N) >> kTwLoBits;
    return (hi > 0 ? hi : 1) + (1 << kTwLoBits);
}
struct Tw2 {
    const float2* A;
    const float2* B;
    __device__ __forceinline__ float2 operator()(int idx) const {
        return cmul(A[idx >> kTwLoBits], B[idx & ((1 << kTwLoBits) - 1)]);
    }
};
// fills tab[0 .. tw2_entries(N)) from the global table U (all threads of the CTA; caller syncs)
__device__ __forceinline__ Tw2 tw2_build(float2* tab, const float2* __restrict__ U, int N, int tid, int nt) {
    const int hi = ((2 * N) >> kTwLoBits) > 0 ? ((2 * N) >> kTwLoBits) : 1;
    for (int h = tid; h < hi; h += nt) tab[h] = tw_lookup(U, h << kTwLoBits, N);
    for (int l = tid; l < (1 << kTwLoBits); l += nt) tab[hi + l] = (l < 2 * N) ? tw_lookup(U, l, N) : make_float2(1.f, 0.f);
    Tw2 t; t.A = tab; t.B = tab + hi;
    return t;
}

// Shared-memory layout of the FFT buffer: one float2 of padding after every 16 points.  The stages
// with quarter-span q >= 16 stay conflict free (32 consecutive points = two 128-byte wavefronts) and
// the fused radix-16 tail below reads element e of 32 different 16-point groups at a stride of 17
// float2 = 34 words - also two wavefronts, where the unpadded strides of 16 / 4 float2 cost 8.
__host__ __device__ __forceinline__ int fft_pad(int i) { return i + (i >> 4); }
__host__ __device__ __forceinline__ int fft_buf_len(int N) { return (N + (N >> 4) + 2) & ~1; }     // float2 entries (even: what follows stays 16-byte aligned)

__device__ __forceinline__ void bfly4(float2& x0, float2& x1, float2& x2, float2& x3) {
    const float2 a0 = make_float2(x0.x + x2.x, x0.y + x2.y);
    const float2 a1 = make_float2(x0.x - x2.x, x0.y - x2.y);
    const float2 a2 = make_float2(x1.x + x3.x, x1.y + x3.y);
    const float2 a3 = make_float2(x1.x - x3.x, x1.y - x3.y);
    x0 = make_float2(a0.x + a2.x, a0.y + a2.y);
    x1 = make_float2(a1.x + a3.y, a1.y - a3.x);      // a1 - i a3
    x2 = make_float2(a0.x - a2.x, a0.y - a2.y);
    x3 = make_float2(a1.x - a3.y, a1.y + a3.x);      // a1 + i a3
}

// One radix-16 DIF pass (= two radix-4 stages kept in registers) over groups of span S = 16 q:
// butterfly u takes the 16 points base + e q, e = 0 .. 15.  Level 1 works on (e0, e0 + 4, e0 + 8, e0 + 12)
// with twiddles W_S^((j + e0 q) r) = W_S^(j r) * W_16^(e0 r); level 2 on (4 r .. 4 r + 3) with
// W_(S/4)^(j r') = (W_S^(4 j))^r'.  Results go back in place, so the digit-reversed order is that of
// two consecutive radix-4 stages.  q = 1 (S = 16) is the twiddle-free tail of the transform.
template <int NT>
__device__ __forceinline__ void fft_radix16_pass(float2* __restrict__ buf, int N, int span, const Tw2& tw) {
    constexpr float C1 = 0.92387953251128674f, S1 = 0.38268343236508977f, C2 = 0.70710678118654752f;
    const int q = span >> 4;
    const int log2q = 31 - __clz(q);
    const int tstep = (2 * N) >> (log2q + 4);      // W_span^j = exp(-2 pi i j tstep / 2N)
    for (int u = threadIdx.x; u < (N >> 4); u += NT) {
        const int j = u & (q - 1);
        const int base = ((u >> log2q) << (log2q + 4)) + j;
        float2 x[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) x[e] = buf[fft_pad(base + e * q)];
        bfly4(x[0], x[4], x[8], x[12]);
        bfly4(x[1], x[5], x[9], x[13]);
        x[5] = cmul(x[5], make_float2(C1, -S1)); x[9] = cmul(x[9], make_float2(C2, -C2)); x[13] = cmul(x[13], make_float2(S1, -C1));
        bfly4(x[2], x[6], x[10], x[14]);
        x[6] = cmul(x[6], make_float2(C2, -C2)); x[10] = make_float2(x[10].y, -x[10].x); x[14] = cmul(x[14], make_float2(-C2, -C2));
        bfly4(x[3], x[7], x[11], x[15]);
        x[7] = cmul(x[7], make_float2(S1, -C1)); x[11] = cmul(x[11], make_float2(-C2, -C2)); x[15] = cmul(x[15], make_float2(-C1, S1));
        if (j != 0) {
            const float2 t1 = tw(j * tstep);
            const float2 t2 = cmul(t1, t1);
            const float2 t3 = cmul(t2, t1);
#pragma unroll
            for (int e0 = 0; e0 < 4; ++e0) {
                x[e0 + 4] = cmul(x[e0 + 4], t1); x[e0 + 8] = cmul(x[e0 + 8], t2); x[e0 + 12] = cmul(x[e0 + 12], t3);
            }
        }
#pragma unroll
        for (int b4 = 0; b4 < 16; b4 += 4) bfly4(x[b4], x[b4 + 1], x[b4 + 2], x[b4 + 3]);
        if (j != 0) {
            const float2 s1 = tw(4 * j * tstep);
            const float2 s2 = cmul(s1, s1);
            const float2 s3 = cmul(s2, s1);
#pragma unroll
            for (int b4 = 0; b4 < 16; b4 += 4) {
                x[b4 + 1] = cmul(x[b4 + 1], s1); x[b4 + 2] = cmul(x[b4 + 2], s2); x[b4 + 3] = cmul(x[b4 + 3], s3);
            }
        }
#pragma unroll
        for (int e = 0; e < 16; ++e) buf[fft_pad(base + e * q)] = x[e];
    }
    __syncthreads();
}

// In-place DIF FFT of N = 2^m >= 16 complex points in padded shared memory (NT threads): one radix-2
// stage first when m is odd, one radix-4 stage when the remaining log2 is not a multiple of 4, then
// radix-16 passes (two radix-4 stages per shared-memory round trip) down to the twiddle-free tail.
// Output in digit-reversed order (fft_pos).
template <int NT>
__device__ void fft_inplace_dif(float2* __restrict__ buf, int N, const Tw2 tw) {
    const int tid = threadIdx.x;
    int span = N;
    int log2n = 0; while ((1 << log2n) < N) ++log2n;
    if (log2n & 1) {
        const int h = N >> 1;
        for (int u = tid; u < h; u += NT) {
            const float2 x0 = buf[fft_pad(u)], x1 = buf[fft_pad(u + h)];
            buf[fft_pad(u)] = make_float2(x0.x + x1.x, x0.y + x1.y);
            float2 y1 = make_float2(x0.x - x1.x, x0.y - x1.y);
            if (u != 0) y1 = cmul(y1, tw(2 * u));                        // W_N^u = exp(-2 pi i 2u / 2N)
            buf[fft_pad(u + h)] = y1;
        }
        __syncthreads();
        span = h;
        --log2n;
    }
    if (log2n & 2) {                                   // log2(span) = 2 mod 4: one radix-4 stage
        const int q = span >> 2;
        int log2q = 0; while ((1 << log2q) < q) ++log2q;
        const int tstep = (2 * N) / span;
        for (int u = tid; u < (N >> 2); u += NT) {
            const int j = u & (q - 1);
            const int base = ((u >> log2q) << (log2q + 2)) + j;
            const int p0 = fft_pad(base), p1 = fft_pad(base + q), p2 = fft_pad(base + 2 * q), p3 = fft_pad(base + 3 * q);
            float2 x0 = buf[p0], x1 = buf[p1], x2 = buf[p2], x3 = buf[p3];
            bfly4(x0, x1, x2, x3);
            if (j != 0) {
                const float2 w1 = tw(j * tstep);
                const float2 w2 = cmul(w1, w1);
                const float2 w3 = cmul(w2, w1);
                x1 = cmul(x1, w1);
                x2 = cmul(x2, w2);
                x3 = cmul(x3, w3);
            }
            buf[p0] = x0; buf[p1] = x1; buf[p2] = x2; buf[p3] = x3;
        }
        __syncthreads();
        span = q;
    }
    for (; span >= 16; span >>= 4) fft_radix16_pass<NT>(buf, N, span, tw);
}

// storage position (unpadded index) of frequency k after fft_inplace_dif
__host__ __device__ __forceinline__ int fft_pos(int k, int N) {
    int p = 0, size = N;
    int log2n = 0; while ((1 << log2n) < N) ++log2n;
    if (log2n & 1) { size >>= 1; p += (k & 1) * size; k >>= 1; }
    while (size >= 4) { size >>= 2; p += (k & 3) * size; k >>= 2; }
    return p;
}

// |X[k]|^2 of the length-2N real FFT from the packed complex FFT, 1 <= k <= N-1
__device__ __forceinline__ float real_fft_power(const float2* __restrict__ buf, int k, int N,
                                                const float2* __restrict__ U) {
    const float2 zk = buf[fft_pad(fft_pos(k, N))];
    const float2 zn = buf[fft_pad(fft_pos(N - k, N))];
    const float er = 0.5f * (zk.x + zn.x), ei = 0.5f * (zk.y - zn.y);        // E = (Zk + conj Zn)/2
    const float dr = 0.5f * (zk.x - zn.x), di = 0.5f * (zk.y + zn.y);        // D = (Zk - conj Zn)/2
    const float or_ = di, oi = -dr;                                          // O = -i D
    const float2 w = __ldg(U + k);
    const float xr = er + (w.x * or_ - w.y * oi);
    const float xi = ei + (w.x * oi + w.y * or_);
    return xr * xr + xi * xi;
}

// same with the two storage positions and the bin's twiddle precomputed on the host
// (AbcPsdParams::binpos = {pos(k), pos(N - k), bits of cos, bits of -sin})
__device__ __forceinline__ float real_fft_power_at(const float2* __restrict__ buf, int4 bp) {
    const float2 zk = buf[bp.x];
    const float2 zn = buf[bp.y];
    const float er = 0.5f * (zk.x + zn.x), ei = 0.5f * (zk.y - zn.y);
    const float dr = 0.5f * (zk.x - zn.x), di = 0.5f * (zk.y + zn.y);
    const float or_ = di, oi = -dr;
    const float wx = __int_as_float(bp.z), wy = __int_as_float(bp.w);
    const float xr = er + (wx * or_ - wy * oi);
    const float xi = ei + (wx * oi + wy * or_);
    return xr * xr + xi * xi;
}

}  // namespace itjl
