// Shared host/device helpers of libitjl_b200 (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <mutex>
#include <vector>

#include "../../include/itjl_b200.h"

namespace itjl {

constexpr int kThreads = 256;          // CTA size of the fused simulate+summary kernels
constexpr int kWarps = kThreads / 32;
// samples per thread of the blocked-scan simulation (simulate.cuh keeps a table of that many carry powers per
// particle): 256 threads x 221 floats already fill the shared memory of one CTA
constexpr int kQtab = 224;

// Grow-only device scratch buffer (freed with its owner: every error exit that deletes a context or a model
// releases what was already allocated).
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { release(); }
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

// NVTX range of one C-ABI call (nsys / ncu --nvtx show the phases of a generation by name)
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

}  // namespace itjl

struct itjl_ctx {
    // one call at a time per context: every entry point that takes a context (or a model of it) holds this
    // lock for its whole duration, so a context may be shared between host threads (Julia tasks)
    std::recursive_mutex mu;
    itjl_tuning tuning{};
    // multi-GPU (itjl_api_multi.cu): a parent context owns one sub-context per device of this process
    // (itjl_ctx_create_multi), or a single-device context joins a communicator of several processes
    // (itjl_ctx_attach_comm); `comm` is the ncclComm_t of THIS device, rank / world its place in it
    std::vector<itjl_ctx*> subs;
    itjl_ctx* parent = nullptr;
    void* comm = nullptr;
    int rank = 0, world = 1;
    itjl::DevBuf gather;                // [world][width] all-gather buffer of per-particle values
    int device = 0;
    int sm_count = 0;
    int max_smem_optin = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    char err[512] = {0};
    int64_t launches = 0;
    // timing of the dominant kernel (events on `stream`)
    bool timing = false;
    bool timing_sync = true;            // false: events are only recorded around the last timed launch
    int64_t timed_launches = 0;
    double timed_ms = 0.0;
    itjl::DevBuf theta, dist, stats, u0, noise, phases, flags, scratch, data, out, out2, out3, part, raw;
    int64_t res_slices = 0, res_nt = 0;   // shape of the matrix itjl_data_upload left in `data` (0: none)
    itjl::DevBuf acf_keep;                // ACF matrix (rows x lags, column-major) of the last acw call (itjl_acw_acf)
    int64_t keep_rows = 0, keep_lags = 0;
    // pinned staging ring of the large host -> device uploads (itjl_api_acw.cu)
    void* pin[3] = {nullptr, nullptr, nullptr};
    cudaEvent_t pin_ev[3] = {nullptr, nullptr, nullptr};
    // small device -> host results staged through pinned memory (itjl_api_acw.cu: d2h_queue / d2h_flush)
    struct PendingCopy { void* dst; size_t off, bytes; };
    void* pin_small = nullptr;
    size_t pin_small_used = 0;
    std::vector<PendingCopy> pending;
};

namespace itjl {

extern char g_err[512];

inline int fail(itjl_ctx* ctx, int code, const char* fmt, const char* a = "", const char* b = "") {
    char* dst = ctx ? ctx->err : g_err;
    snprintf(dst, 512, fmt, a, b);
    // a failed runtime call stays recorded as the thread's "last error" until it is fetched: clear it here, or the next
    // successful launch's cudaGetLastError() reports it (found by benchmarks/psd_fuzz.py: a refused configuration made
    // the following, unrelated call fail with "invalid argument")
    if (code == ITJL_ERR_CUDA) (void)cudaGetLastError();
    return code;
}

#define ITJL_CAT2(a, b) a##b
#define ITJL_CAT(a, b) ITJL_CAT2(a, b)
#define ITJL_LOCK(ctx) std::lock_guard<std::recursive_mutex> ITJL_CAT(_itjl_lock_, __COUNTER__)((ctx)->mu)

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
