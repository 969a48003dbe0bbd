// Microbenchmark: issue rate of FFMA vs FFMA2 (fma.rn.f32x2) on sm_100a.  nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cuda_runtime.h>
#include <cstdio>
template <int MODE>
__global__ void k(float* out, int iters) {
    float a[16], b = 1.0001f + threadIdx.x * 1e-7f, c = 0.5f;
    unsigned long long pa[8], pb, pc;
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = threadIdx.x + i;
    { float2 t = make_float2(b, b); pb = *reinterpret_cast<unsigned long long*>(&t); t = make_float2(c, c); pc = *reinterpret_cast<unsigned long long*>(&t); }
#pragma unroll
    for (int i = 0; i < 8; ++i) { float2 t = make_float2(a[2 * i], a[2 * i + 1]); pa[i] = *reinterpret_cast<unsigned long long*>(&t); }
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], b, c);
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(pa[i]) : "l"(pb), "l"(pc));
        }
    }
    float s = 0.f;
    if (MODE == 0) { for (int i = 0; i < 16; ++i) s += a[i]; }
    else { for (int i = 0; i < 8; ++i) { float2 t = *reinterpret_cast<float2*>(&pa[i]); s += t.x + t.y; } }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 1 << 16;
    for (int mode = 0; mode < 2; ++mode) {
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) k<0><<<148 * 8, 256>>>(out, iters); else k<1><<<148 * 8, 256>>>(out, iters);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            double flops = 2.0 * 16 * iters * 148.0 * 8 * 256;
            printf("%s: %.3f ms  %.1f TFLOP/s\n", mode ? "FFMA2" : "FFMA ", ms, flops / ms / 1e9);
        }
    }
    return 0;
}
