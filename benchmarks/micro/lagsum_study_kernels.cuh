// Study kernels that used to live in csrc/peak_kernels.cu (issue-rate microbenchmarks of the lag-sum inner loop:
// FFMA vs FFMA2, register-tiled variants).  Not part of libitjl_b200.so; kept for reference only.
#pragma once
template <int MODE>
__global__ void __launch_bounds__(kThreads) k_fp32_micro(float* sink, int iters, float x, float y) {
    float2 acc[16];
    unsigned cnt[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) { acc[j] = make_float2((threadIdx.x + j) * 1e-3f, j * 1e-3f); cnt[j] = threadIdx.x + j; }
    const float2 xx = make_float2(x, x), yy = make_float2(y, y * 0.5f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            if (MODE == 1 || MODE == 2) { float2 t = acc[j]; ffma2_f2(t, xx, yy); acc[j] = make_float2(t.x, t.y); }
            if (MODE == 3) { acc[j].x = fmaf(acc[j].x, x, y); acc[j].y = fmaf(acc[j].y, x, y); }
            if (MODE == 2) cnt[j] = (cnt[j] ^ (cnt[j] >> 3)) + 0x9E3779B9u;
            if (MODE == 3) { cnt[j] = (cnt[j] ^ (cnt[j] >> 3)) + 0x9E3779B9u; }
        }
    }
    float s = 0.f; unsigned c = 0;
#pragma unroll
    for (int j = 0; j < 16; ++j) { s += acc[j].x + acc[j].y; c ^= cnt[j]; }
    if (s == 123456.789f || c == 0x12345u) sink[blockIdx.x * kThreads + threadIdx.x] = s + c;
}

// packed-FFMA2 lag-sum phase (tuning only, see lagsum.cuh): A_p = (a[2p], a[2p+1]),
// A'_p = (a[2p-1], a[2p]), B_q = (b[2q], b[2q+1]); A_p*B_{p+r} -> lag 2r, A'_p*B_{p+r} -> lag 2r+1.
typedef unsigned long long f32x2;
__device__ __forceinline__ void ffma2(f32x2& d, f32x2 a, f32x2 b) {
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b));
}
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
template <int TT, int PH>
__device__ __forceinline__ void acf_phase2(f32x2 (&acc2)[TT], const f32x2 (&A)[TT / 2],
                                           const f32x2 (&Ap)[TT / 2], const f32x2 (&B)[TT]) {
    constexpr int H = TT / 2;
#pragma unroll
    for (int p = 0; p < H; ++p)
#pragma unroll
        for (int r = 0; r < H; ++r) {
            const int phys = (p + r + H * PH) % TT;
            ffma2(acc2[2 * r], A[p], B[phys]);
            ffma2(acc2[2 * r + 1], Ap[p], B[phys]);
        }
}
// modes 10 / 12: the lag-sum inner loops of abc_kernels.cu (scalar FFMA 20x20 / FFMA2 20x20) on
// register-resident operands - no shared-memory loads, same FMA instruction stream.
template <int NT>
__global__ void __launch_bounds__(NT, 512 / NT) k_lagsum_ffma(float* sink, int iters, const float* __restrict__ src) {
    float acc[20], a[20], b[40];
#pragma unroll
    for (int j = 0; j < 20; ++j) { acc[j] = 0.f; a[j] = src[(threadIdx.x + j) & 1023]; }
#pragma unroll
    for (int j = 0; j < 40; ++j) b[j] = src[(threadIdx.x * 3 + j) & 1023];
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 20; ++i)
#pragma unroll
            for (int j = 0; j < 20; ++j) acc[j] = fmaf(a[i], b[i + j], acc[j]);
#pragma unroll
        for (int i = 0; i < 20; ++i)
#pragma unroll
            for (int j = 0; j < 20; ++j) acc[j] = fmaf(a[i], b[(20 + i + j) % 40], acc[j]);
    }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 20; ++j) s += acc[j];
    if (s == 123456.789f) sink[blockIdx.x * NT + threadIdx.x] = s;
}
template <int NT>
__global__ void __launch_bounds__(NT, 512 / NT) k_lagsum_ffma2(float* sink, int iters, const f32x2* __restrict__ src) {
    f32x2 acc2[20], A[10], Ap[10], B[20];
#pragma unroll
    for (int j = 0; j < 20; ++j) { acc2[j] = 0ull; B[j] = src[(threadIdx.x * 3 + j) & 511]; }
#pragma unroll
    for (int j = 0; j < 10; ++j) { A[j] = src[(threadIdx.x + j) & 511]; Ap[j] = src[(threadIdx.x + j + 7) & 511]; }
    for (int it = 0; it < iters; ++it) {
        acf_phase2<20, 0>(acc2, A, Ap, B);
        acf_phase2<20, 1>(acc2, A, Ap, B);
    }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 20; ++j) { float lo, hi; unpack2(acc2[j], lo, hi); s += lo + hi; }
    if (s == 123456.789f) sink[blockIdx.x * NT + threadIdx.x] = s;
}

cudaError_t fp32_micro_launch(float* sink, int blocks, int iters, int mode, cudaStream_t st) {
    // lag-sum loops: 800 FMAs per thread per iteration; scale iters so the flop count matches
    // fp32_peak_flops_per_launch(blocks, iters) = 2 * blocks * 256 * iters * 32
    if (mode == 10 || mode == 11) {
        const int nt = (mode == 10) ? 128 : 256;
        const int it2 = (int)((long long)iters * 32 * 256 / (800LL * nt));
        if (nt == 128) k_lagsum_ffma<128><<<blocks, 128, 0, st>>>(sink, it2, sink);
        else k_lagsum_ffma<256><<<blocks, 256, 0, st>>>(sink, it2, sink);
        return cudaGetLastError();
    }
    if (mode == 12 || mode == 13) {
        const int nt = (mode == 12) ? 128 : 256;
        const int it2 = (int)((long long)iters * 32 * 256 / (800LL * nt));
        if (nt == 128) k_lagsum_ffma2<128><<<blocks, 128, 0, st>>>(sink, it2, reinterpret_cast<const f32x2*>(sink));
        else k_lagsum_ffma2<256><<<blocks, 256, 0, st>>>(sink, it2, reinterpret_cast<const f32x2*>(sink));
        return cudaGetLastError();
    }
    if (mode == 1) k_fp32_micro<1><<<blocks, kThreads, 0, st>>>(sink, iters, 0.999999f, 1e-7f);
    else if (mode == 2) k_fp32_micro<2><<<blocks, kThreads, 0, st>>>(sink, iters, 0.999999f, 1e-7f);
    else k_fp32_micro<3><<<blocks, kThreads, 0, st>>>(sink, iters, 0.999999f, 1e-7f);
    return cudaGetLastError();
}

