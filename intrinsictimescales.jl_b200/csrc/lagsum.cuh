// Register-tiled lag-sum inner loops shared by the fused aABC kernel (abc_kernels.cu) and the
// tuning microbenchmarks (peak_kernels.cu).
#pragma once
#include "common.cuh"

namespace itjl {

// Lag-sum tile: TK = 20 lags x TT = 20 samples per thread step (400 FFMA per 10 LDS.128).
//
// Tuning note (round 1, B200): a packed-FFMA2 (fma.rn.f32x2) version of this loop - 12x12 and
// 20x20 tiles, operand pairs from a one-sample-delayed copy of the series - was measured at
// 0.67-0.69 of the FP32 peak end to end against 0.75 for the scalar loop: with two non-reused
// 64-bit operands per FFMA2 the register-file read ports, not the issue slots, are the limit
// (register-resident microbenchmarks k_lagsum_ffma / k_lagsum_ffma2 in peak_kernels.cu:
// 60.7-66.6 vs 57.6-58.9 TFLOP/s of 70.7).  The scalar loop is therefore the product path.
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

}  // namespace itjl
