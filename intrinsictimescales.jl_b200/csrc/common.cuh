// Shared host/device helpers of libitjl_b200 (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/itjl_b200.h"

namespace itjl {

constexpr int kThreads = 256;          // CTA size of the fused simulate+summary kernels
constexpr int kWarps = kThreads / 32;
// samples per thread of the blocked-scan simulation (simulate.cuh keeps a table of that many carry powers per
// particle): 256 threads x 221 floats already fill the shared memory of one CTA
constexpr int kQtab = 224;

// Grow-only device scratch buffer.
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e == cudaSuccess) cap = bytes;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

}  // namespace itjl

struct itjl_ctx {
    int device = 0;
    int sm_count = 0;
    int max_smem_optin = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    char err[512] = {0};
    int64_t launches = 0;
    // timing of the dominant kernel (events on `stream`)
    bool timing = false;
    int64_t timed_launches = 0;
    double timed_ms = 0.0;
    itjl::DevBuf theta, dist, stats, u0, noise, phases, flags, scratch, data, out, out2, out3, part;
    // pinned staging ring of the large host -> device uploads (itjl_api_acw.cu)
    void* pin[3] = {nullptr, nullptr, nullptr};
    cudaEvent_t pin_ev[3] = {nullptr, nullptr, nullptr};
};

namespace itjl {

extern char g_err[512];

inline int fail(itjl_ctx* ctx, int code, const char* fmt, const char* a = "", const char* b = "") {
    char* dst = ctx ? ctx->err : g_err;
    snprintf(dst, 512, fmt, a, b);
    return code;
}

#define ITJL_CUDA(ctx, call)                                                              \
    do {                                                                                  \
        cudaError_t _e = (call);                                                          \
        if (_e != cudaSuccess)                                                            \
            return itjl::fail((ctx), ITJL_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(_e)); \
    } while (0)

#define ITJL_CHECK_LAUNCH(ctx, name)                                                      \
    do {                                                                                  \
        cudaError_t _e = cudaGetLastError();                                              \
        if (_e != cudaSuccess)                                                            \
            return itjl::fail((ctx), ITJL_ERR_CUDA, "launch %s: %s", name, cudaGetErrorString(_e)); \
        (ctx)->launches++;                                                                \
    } while (0)

// ---- block-level helpers (kThreads threads) ------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace itjl
