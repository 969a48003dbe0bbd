// FP32 FMA issue-rate microbenchmark: the roofline denominator of the fused
// simulate+ACF kernel (MEASURED_PEAKS.json only carries HBM and bf16 tensor peaks).
#include "kernels.h"

namespace itjl {

constexpr int kPeakAcc = 32;

__global__ void __launch_bounds__(kThreads) k_fp32_peak(float* sink, int iters, float x, float y) {
    float acc[kPeakAcc];
#pragma unroll
    for (int j = 0; j < kPeakAcc; ++j) acc[j] = (float)(threadIdx.x + j) * 1e-3f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < kPeakAcc; ++j) acc[j] = fmaf(acc[j], x, y);
    }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < kPeakAcc; ++j) s += acc[j];
    if (s == 123456.789f) sink[blockIdx.x * kThreads + threadIdx.x] = s;
}

cudaError_t fp32_peak_launch(float* sink, int blocks, int iters, cudaStream_t st) {
    k_fp32_peak<<<blocks, kThreads, 0, st>>>(sink, iters, 0.999999f, 1e-7f);
    return cudaGetLastError();
}

double fp32_peak_flops_per_launch(int blocks, int iters) {
    return 2.0 * (double)blocks * kThreads * (double)iters * kPeakAcc;
}

}  // namespace itjl
