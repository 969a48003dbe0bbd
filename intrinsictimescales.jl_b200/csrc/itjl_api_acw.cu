// C-ABI layer, part 2: summary statistics on stored data and acw()
// (reference src/stats/summary.jl:46-89, 183-235, 455-496 and src/core/acw.jl:141-231).
#include <math.h>
#include <algorithm>
#include <stdlib.h>
#include <condition_variable>
#include <pthread.h>
#include <mutex>
#include <thread>
#include <vector>

#include "api_internal.h"

using namespace itjl;

// Host -> device upload of the caller's (pageable) matrix.  Large inputs go through a ring of three
// pinned 16 MB staging buffers: copy_threads() host threads copy chunk c into a buffer while the DMA engine is
// still draining chunk c - 1 (a plain cudaMemcpy from pageable memory runs at ~11 GB/s on this box, the staged
// pipeline with per-chunk std::threads at ~25 GB/s, with the persistent pool below at ~41 GB/s; a whole C2 acw()
// call from a 400 MB host array: 36 ms -> 17 ms -> 9.8 ms).
static const size_t kPinChunk = (size_t)16 << 20;
// host threads that copy between the caller's pageable array and the pinned staging ring
// (itjl_tuning::copy_threads, default: half the hardware threads, at most 8)
constexpr int kMaxCopyThreads = 16;
static int copy_threads(const itjl_ctx* ctx) {
    const int t = ctx->tuning.copy_threads;
    int n = t > 0 ? t : (int)(std::thread::hardware_concurrency() / 2);
    if (t <= 0 && n > 8) n = 8;
    return n < 1 ? 1 : (n > kMaxCopyThreads ? kMaxCopyThreads : n);
}

// Persistent fork-join pool for those copies: starting std::threads per 16 MB chunk cost as much as a third of the
// chunk's copy time (7 thread starts ~ 0.1-0.15 ms against ~0.25 ms of memcpy).  One copy at a time (run_mu); workers
// are detached threads that sleep on a condition variable between jobs.  The pool lives on the heap and is never
// destroyed (nothing to join when the process exits); a forked child starts with a fresh, empty pool (the parent's
// workers do not exist there).
namespace {
class CopyPool {
    std::mutex run_mu, mu;
    std::condition_variable cv_start, cv_done;
    int n_workers = 0;
    uint64_t gen = 0;
    int pending = 0, n_parts = 0;
    char* dst = nullptr; const char* src = nullptr; size_t n = 0, part = 0;

    void copy_part(int t) const {
        const size_t lo = (size_t)t * part < n ? (size_t)t * part : n;
        const size_t hi = (size_t)(t + 1) * part < n ? (size_t)(t + 1) * part : n;
        if (hi > lo) memcpy(dst + lo, src + lo, hi - lo);
    }
    void worker(int t, uint64_t seen) {
        for (;;) {
            std::unique_lock<std::mutex> lk(mu);
            cv_start.wait(lk, [&] { return gen != seen; });
            seen = gen;
            const bool mine = t < n_parts;
            lk.unlock();
            if (mine) copy_part(t);
            lk.lock();
            if (--pending == 0) cv_done.notify_one();
        }
    }

 public:
    void copy(void* d, const void* s, size_t bytes, int n_threads) {
        if (n_threads <= 1 || bytes < ((size_t)1 << 20)) { memcpy(d, s, bytes); return; }
        std::lock_guard<std::mutex> run(run_mu);
        {
            std::lock_guard<std::mutex> lk(mu);
            while (n_workers < n_threads - 1) {
                const int t = ++n_workers;
                const uint64_t seen = gen;
                std::thread([this, t, seen] { worker(t, seen); }).detach();
            }
            dst = reinterpret_cast<char*>(d); src = reinterpret_cast<const char*>(s); n = bytes;
            part = ((bytes / n_threads) + 4095) & ~(size_t)4095;
            n_parts = n_threads;
            pending = n_workers;
            ++gen;
        }
        cv_start.notify_all();
        copy_part(0);
        std::unique_lock<std::mutex> lk(mu);
        cv_done.wait(lk, [&] { return pending == 0; });
    }
};
CopyPool* g_copy_pool_ptr = nullptr;
std::once_flag g_copy_pool_once;
void copy_pool_after_fork() { g_copy_pool_ptr = new CopyPool(); }      // the parent's pool (and its locks) are left behind
CopyPool& copy_pool() {
    std::call_once(g_copy_pool_once, [] {
        g_copy_pool_ptr = new CopyPool();
        pthread_atfork(nullptr, nullptr, copy_pool_after_fork);
    });
    return *g_copy_pool_ptr;
}
}  // namespace

// `bytes` from the caller's array to the device buffer `dst_dev`
static int upload_bytes(itjl_ctx* ctx, void* dst_dev, const void* data, size_t bytes) {
    if (bytes < 4 * kPinChunk || ctx->tuning.pinned_upload == 0) {
        ITJL_CUDA(ctx, cudaMemcpyAsync(dst_dev, data, bytes, cudaMemcpyHostToDevice, ctx->stream));
        return ITJL_OK;
    }
    for (int i = 0; i < 3; ++i) {
        if (!ctx->pin[i]) ITJL_CUDA(ctx, cudaHostAlloc(&ctx->pin[i], kPinChunk, cudaHostAllocDefault));
        if (!ctx->pin_ev[i]) ITJL_CUDA(ctx, cudaEventCreateWithFlags(&ctx->pin_ev[i], cudaEventDisableTiming));
    }
    const char* src = reinterpret_cast<const char*>(data);
    char* dst = reinterpret_cast<char*>(dst_dev);
    const int n_threads = copy_threads(ctx);
    size_t off = 0;
    for (int c = 0; off < bytes; ++c, off += kPinChunk) {
        const int b = c % 3;
        const size_t n = bytes - off < kPinChunk ? bytes - off : kPinChunk;
        if (c >= 3) ITJL_CUDA(ctx, cudaEventSynchronize(ctx->pin_ev[b]));       // the DMA out of this buffer is done
        char* stage = reinterpret_cast<char*>(ctx->pin[b]);
        copy_pool().copy(stage, src + off, n, n_threads);
        ITJL_CUDA(ctx, cudaMemcpyAsync(dst + off, stage, n, cudaMemcpyHostToDevice, ctx->stream));
        ITJL_CUDA(ctx, cudaEventRecord(ctx->pin_ev[b], ctx->stream));
    }
    return ITJL_OK;
}

// The caller's matrix, element (s, t) at data[s * ss + t * ts], into ctx->data in the kernels' layout
// (slice index fastest).  That layout is uploaded as it is; anything else is uploaded as it lies in host memory
// (its whole span) and re-laid on the device by k_relayout - no host-side transpose.
static int upload_data(itjl_ctx* ctx, const double* data, int64_t n_slices, int64_t n_t, int64_t ss = 1, int64_t ts = 0) {
    if (ts == 0) ts = n_slices;
    const size_t bytes = (size_t)n_slices * n_t * sizeof(double);
    ctx->res_slices = ctx->res_nt = 0;                       // whatever was resident is overwritten
    ITJL_CUDA(ctx, ctx->data.reserve(bytes));
    if (ss == 1 && ts == n_slices) return upload_bytes(ctx, ctx->data.p, data, bytes);
    if (ss < 1 || ts < 1) return fail(ctx, ITJL_ERR_ARG, "strides must be positive");
    const size_t span = (size_t)((n_slices - 1) * ss + (n_t - 1) * ts + 1);
    if (span > (size_t)4 * n_slices * n_t) return fail(ctx, ITJL_ERR_ARG, "strided view spans more than 4x its elements: copy it to a dense array first");
    ITJL_CUDA(ctx, ctx->raw.reserve(span * sizeof(double)));
    const int rc = upload_bytes(ctx, ctx->raw.p, data, span * sizeof(double));
    if (rc != ITJL_OK) return rc;
    const cudaError_t e = relayout_launch((const double*)ctx->raw.p, n_slices, n_t, ss, ts, (double*)ctx->data.p, ctx->stream);
    if (e != cudaSuccess) return fail(ctx, ITJL_ERR_CUDA, "launch k_relayout: %s", cudaGetErrorString(e));
    ctx->launches++;
    return ITJL_OK;
}

// Device -> host download of a large result (the exported ACF matrix) through the same pinned ring:
// DMA of chunk c + 1 runs while the host threads copy chunk c out to the caller's pageable array.
static int download_large(itjl_ctx* ctx, void* host_dst, const void* dev_src, size_t bytes) {
    // below 256 KB the driver's own pageable path is as quick; above it a pageable cudaMemcpy crawls (the 1.3 MB ACF of
    // acw() on 10 000 slices: 0.6 ms against 0.1 ms through the pinned ring + one memcpy)
    if (bytes < ((size_t)256 << 10) || ctx->tuning.pinned_upload == 0) {
        ITJL_CUDA(ctx, cudaMemcpyAsync(host_dst, dev_src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
        return ITJL_OK;
    }
    for (int i = 0; i < 3; ++i) {
        if (!ctx->pin[i]) ITJL_CUDA(ctx, cudaHostAlloc(&ctx->pin[i], kPinChunk, cudaHostAllocDefault));
        if (!ctx->pin_ev[i]) ITJL_CUDA(ctx, cudaEventCreateWithFlags(&ctx->pin_ev[i], cudaEventDisableTiming));
    }
    ITJL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));          // earlier users of the ring are done
    char* dst = reinterpret_cast<char*>(host_dst);
    const char* src = reinterpret_cast<const char*>(dev_src);
    const size_t n_chunks = (bytes + kPinChunk - 1) / kPinChunk;
    auto chunk_len = [&](size_t c) { return bytes - c * kPinChunk < kPinChunk ? bytes - c * kPinChunk : kPinChunk; };
    auto issue = [&](size_t c) -> cudaError_t {
        cudaError_t e = cudaMemcpyAsync(ctx->pin[c % 3], src + c * kPinChunk, chunk_len(c), cudaMemcpyDeviceToHost, ctx->stream);
        return e != cudaSuccess ? e : cudaEventRecord(ctx->pin_ev[c % 3], ctx->stream);
    };
    for (size_t c = 0; c < 2 && c < n_chunks; ++c) ITJL_CUDA(ctx, issue(c));
    const int n_threads = bytes < ((size_t)8 << 20) ? 1 : copy_threads(ctx);      // thread start-up costs more than a small copy
    for (size_t c = 0; c < n_chunks; ++c) {
        if (c + 2 < n_chunks) ITJL_CUDA(ctx, issue(c + 2));      // buffer (c + 2) % 3 was emptied at chunk c - 1
        ITJL_CUDA(ctx, cudaEventSynchronize(ctx->pin_ev[c % 3]));
        const char* stage = reinterpret_cast<const char*>(ctx->pin[c % 3]);
        const size_t n = chunk_len(c), off = c * kPinChunk;
        copy_pool().copy(dst + off, stage, n, n_threads);
    }
    return ITJL_OK;
}

// Small results (crossing indices, flags, per-slice values): a cudaMemcpyAsync into pageable memory blocks the host
// for every copy (10-20 us each; acw() makes seven).  d2h_queue copies into a pinned scratch area instead - truly
// asynchronous - and d2h_flush synchronises the stream once and moves the bytes to their destinations.
static const size_t kPinSmall = (size_t)2 << 20;
static int d2h_queue(itjl_ctx* ctx, void* host_dst, const void* dev_src, size_t bytes) {
    if (!ctx->pin_small) ITJL_CUDA(ctx, cudaHostAlloc(&ctx->pin_small, kPinSmall, cudaHostAllocDefault));
    const size_t off = (ctx->pin_small_used + 63) & ~(size_t)63;
    if (off + bytes > kPinSmall) {
        ITJL_CUDA(ctx, cudaMemcpyAsync(host_dst, dev_src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
        return ITJL_OK;
    }
    ITJL_CUDA(ctx, cudaMemcpyAsync(reinterpret_cast<char*>(ctx->pin_small) + off, dev_src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    ctx->pending.push_back({host_dst, off, bytes});
    ctx->pin_small_used = off + bytes;
    return ITJL_OK;
}
static int d2h_flush(itjl_ctx* ctx) {
    ITJL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (const auto& c : ctx->pending) memcpy(c.dst, reinterpret_cast<const char*>(ctx->pin_small) + c.off, c.bytes);
    ctx->pending.clear();
    ctx->pin_small_used = 0;
    return ITJL_OK;
}

static int check_data_args(itjl_ctx* ctx, const double* data, int64_t n_slices, int64_t n_t, const char* who) {
    if (!ctx) return fail(nullptr, ITJL_ERR_ARG, "%s: null context", who);
    if (!data) return fail(ctx, ITJL_ERR_ARG, "%s: null data", who);
    if (n_slices < 1 || n_slices > 0x7fffffffll) return fail(ctx, ITJL_ERR_ARG, "%s: n_slices out of range", who);
    if (n_t < 2 || n_t > (1 << 24)) return fail(ctx, ITJL_ERR_ARG, "%s: need 2 <= n_t <= 2^24", who);
    return ITJL_OK;
}

// Device-side ACF of every slice up to n_lags into ctx->out (row-per-slice, ld = cap).
static int acf_all(itjl_ctx* ctx, int64_t n_slices, int n_t, int n_lags, int missing, int64_t cap, bool reset,
                   int32_t* nan_flag = nullptr) {
    ITJL_CUDA(ctx, ctx->out.reserve((size_t)n_slices * cap * sizeof(double)));
    ITJL_CUDA(ctx, ctx->flags.reserve(((size_t)n_slices * 6 + 4) * sizeof(int32_t)));
    int32_t* n_done = (int32_t*)ctx->flags.p;
    if (reset) ITJL_CUDA(ctx, cudaMemsetAsync(n_done, 0, (size_t)n_slices * sizeof(int32_t), ctx->stream));
    cudaError_t e = acf_fill_launch((const double*)ctx->data.p, n_slices, n_t, missing, n_lags,
                                    (double*)ctx->out.p, cap, n_done, nullptr, 0, ctx->max_smem_optin, ctx->stream, nan_flag);
    if (e != cudaSuccess) return fail(ctx, ITJL_ERR_CUDA, "launch k_acf_slices(fill): %s", cudaGetErrorString(e));
    ctx->launches++;
    return ITJL_OK;
}

static int acf_entry(itjl_ctx* ctx, const double* data, int64_t n_slices, int64_t n_t, int64_t n_lags,
                     int missing, double* out, const char* who, int64_t ss = 1, int64_t ts = 0) {
    int rc = check_data_args(ctx, data, n_slices, n_t, who);
    if (rc != ITJL_OK) return rc;
    if (!out) return fail(ctx, ITJL_ERR_ARG, "%s: null output", who);
    if (n_lags < 1 || n_lags > n_t) return fail(ctx, ITJL_ERR_ARG, "%s: need 1 <= n_lags <= n_t", who);
    ITJL_LOCK(ctx);
    ctx = primary(ctx);
    ITJL_LOCK(ctx);
    NvtxRange nvtx(who);
    ITJL_CUDA(ctx, cudaSetDevice(ctx->device));
    if (n_lags > acf_max_reach((int)n_t, ctx->max_smem_optin))
        return fail(ctx, ITJL_ERR_ARG, "%s: n_t + n_lags does not fit in shared memory", who);
    rc = upload_data(ctx, data, n_slices, n_t, ss, ts);
    if (rc != ITJL_OK) return rc;
    rc = acf_all(ctx, n_slices, (int)n_t, (int)n_lags, missing, n_lags, true);
    if (rc != ITJL_OK) return rc;
    const size_t obytes = (size_t)n_slices * n_lags * sizeof(double);
    ITJL_CUDA(ctx, ctx->out2.reserve(obytes));
    cudaError_t e = acf_export_launch((const double*)ctx->out.p, n_slices, n_lags, (int)n_lags, (double*)ctx->out2.p, ctx->stream);
    if (e != cudaSuccess) return fail(ctx, ITJL_ERR_CUDA, "launch k_acf_export: %s", cudaGetErrorString(e));
    ctx->launches++;
    ITJL_CUDA(ctx, cudaMemcpyAsync(out, ctx->out2.p, obytes, cudaMemcpyDeviceToHost, ctx->stream));
    ITJL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ITJL_OK;
}

static void prime_strides(int w, std::vector<int>& strides) {
    // Romberg.jl: strides 1, f1, f1 f2, ..., w-1 with f_i the prime factors of w-1 (increasing)
    strides.clear();
    strides.push_back(1);
    int m = w - 1;
    for (int p = 2; (long long)p * p <= m; ++p)
        while (m % p == 0) { strides.push_back(strides.back() * p); m /= p; }
    if (m > 1) strides.push_back(strides.back() * m);
}

extern "C" {

int itjl_acf(itjl_ctx* ctx, const double* data, int64_t n_slices, int64_t n_t, int64_t n_lags, double* out) {
    return acf_entry(ctx, data, n_slices, n_t, n_lags, 0, out, "itjl_acf");
}

int itjl_acf_missing(itjl_ctx* ctx, const double* data, int64_t n_slices, int64_t n_t, int64_t n_lags, double* out) {
    return acf_entry(ctx, data, n_slices, n_t, n_lags, 1, out, "itjl_acf_missing");
}

int itjl_acf_strided(itjl_ctx* ctx, const double* data, int64_t n_slices, int64_t n_t, int64_t slice_stride,
                     int64_t time_stride, int64_t n_lags, int32_t missing, double* out) {
    if (slice_stride < 1 || time_stride < 1) return fail(ctx, ITJL_ERR_ARG, "itjl_acf_strided: strides must be positive");
    return acf_entry(ctx, data, n_slices, n_t, n_lags, missing ? 1 : 0, out, "itjl_acf_strided", slice_stride, time_stride);
}

int itjl_psd_hamming(itjl_ctx* ctx, const double* data, int64_t n_slices, int64_t n_t, double fs, double* out) {
    int rc = check_data_args(ctx, data, n_slices, n_t, "itjl_psd_hamming");
    if (rc != ITJL_OK) return rc;
    if (!out || !(fs > 0.0)) return fail(ctx, ITJL_ERR_ARG, "itjl_psd_hamming: null output or fs <= 0");
    ITJL_LOCK(ctx);
    ctx = primary(ctx);
    ITJL_LOCK(ctx);
    NvtxRange nvtx("itjl_psd_hamming");
    ITJL_CUDA(ctx, cudaSetDevice(ctx->device));
    int n2 = 2; while (n2 < n_t) n2 <<= 1; n2 <<= 1;             // 2 * nextpow2(n_t)
    int log2_n2 = 0; while ((1 << log2_n2) < n2) log2_n2++;
    const size_t smem = psd_fft_smem_bytes(n2);
    if ((int64_t)smem > ctx->max_smem_optin) return fail(ctx, ITJL_ERR_ARG, "itjl_psd_hamming: FFT does not fit in shared memory");
    std::vector<float> win((size_t)n_t);
    std::vector<float2> tw((size_t)n2 / 2);
    double win_ss = 0.0;
    psd_make_tables((int)n_t, n2, win.data(), tw.data(), &win_ss);
    const double scale = 1.0 / (fs * win_ss);
    rc = upload_data(ctx, data, n_slices, n_t);
    if (rc != ITJL_OK) return rc;
    ITJL_CUDA(ctx, ctx->scratch.reserve(win.size() * sizeof(float)));
    ITJL_CUDA(ctx, ctx->out3.reserve(tw.size() * sizeof(float2)));
    ITJL_CUDA(ctx, cudaMemcpyAsync(ctx->scratch.p, win.data(), win.size() * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    ITJL_CUDA(ctx, cudaMemcpyAsync(ctx->out3.p, tw.data(), tw.size() * sizeof(float2), cudaMemcpyHostToDevice, ctx->stream));
    const int64_t nb = n2 / 2 - 1;
    const size_t obytes = (size_t)n_slices * nb * sizeof(double);
    ITJL_CUDA(ctx, ctx->out.reserve(obytes));
    cudaError_t e = psd_data_launch((const double*)ctx->data.p, n_slices, (int)n_t, n2, log2_n2, (const float*)ctx->scratch.p,
                                    (const float2*)ctx->out3.p, scale, (double*)ctx->out.p, smem, ctx->stream);
    if (e != cudaSuccess) return fail(ctx, ITJL_ERR_CUDA, "launch k_psd_data: %s", cudaGetErrorString(e));
    ctx->launches++;
    ITJL_CUDA(ctx, cudaMemcpyAsync(out, ctx->out.p, obytes, cudaMemcpyDeviceToHost, ctx->stream));
    ITJL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));      // win / tw stay alive until here
    return ITJL_OK;
}

int64_t itjl_nextfastfft(int64_t n) { return n > 0 && n <= (1 << 30) ? (int64_t)next_fast_len((int)n) : -1; }

int itjl_psd_periodogram(itjl_ctx* ctx, const double* data, int64_t n_slices, int64_t n_t, double fs, double* out) {
    int rc = check_data_args(ctx, data, n_slices, n_t, "itjl_psd_periodogram");
    if (rc != ITJL_OK) return rc;
    if (!out || !(fs > 0.0)) return fail(ctx, ITJL_ERR_ARG, "itjl_psd_periodogram: null output or fs <= 0");
    ITJL_LOCK(ctx);
    ctx = primary(ctx);
    ITJL_LOCK(ctx);
    NvtxRange nvtx("itjl_psd_periodogram");
    ITJL_CUDA(ctx, cudaSetDevice(ctx->device));
    FftPlan plan;
    const int nfft = n_t <= 65536 ? next_fast_len((int)n_t) : 0;
    if (!nfft || !fft_plan_make(nfft, &plan)) return fail(ctx, ITJL_ERR_ARG, "itjl_psd_periodogram: n_t > 65536 is not supported");
    std::vector<double> win((size_t)n_t);
    std::vector<double2> roots((size_t)nfft);
    double win_ss = 0.0;
    pgram_make_tables((int)n_t, nfft, nullptr, win.data(), nullptr, roots.data(), &win_ss);
    rc = upload_data(ctx, data, n_slices, n_t);
    if (rc != ITJL_OK) return rc;
    const int grid = (int)std::min<int64_t>((n_slices + 1) / 2, (int64_t)ctx->sm_count * 4);
    ITJL_CUDA(ctx, ctx->scratch.reserve(win.size() * sizeof(double)));
    ITJL_CUDA(ctx, ctx->out3.reserve(roots.size() * sizeof(double2)));
    ITJL_CUDA(ctx, ctx->out2.reserve((size_t)grid * 2 * nfft * sizeof(double2)));
    ITJL_CUDA(ctx, cudaMemcpyAsync(ctx->scratch.p, win.data(), win.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    ITJL_CUDA(ctx, cudaMemcpyAsync(ctx->out3.p, roots.data(), roots.size() * sizeof(double2), cudaMemcpyHostToDevice, ctx->stream));
    const size_t obytes = (size_t)n_slices * (nfft / 2) * sizeof(double);
    ITJL_CUDA(ctx, ctx->out.reserve(obytes));
    cudaError_t e = pgram_data_launch((const double*)ctx->data.p, n_slices, (int)n_t, plan, (const double*)ctx->scratch.p,
                                      (const double2*)ctx->out3.p, 1.0 / (fs * win_ss), (double2*)ctx->out2.p, grid,
                                      (double*)ctx->out.p, ctx->stream);
    if (e != cudaSuccess) return fail(ctx, ITJL_ERR_CUDA, "launch k_pgram_data: %s", cudaGetErrorString(e));
    ctx->launches++;
    rc = download_large(ctx, out, ctx->out.p, obytes);
    if (rc != ITJL_OK) return rc;
    ITJL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));      // win / roots stay alive until here
    return ITJL_OK;
}

}  // extern "C"

// Lomb-Scargle spectra of the matrix in ctx->data into out_dev: chirp-z transforms where a 7-smooth length
// >= n_t + n_freq <= 65536 exists (tuning.psd_v2 bit 1 forces the direct O(n_t n_freq) sums), the direct kernel otherwise
static int lomb_scargle_dev(itjl_ctx* ctx, int64_t n_slices, int n_t, double dt, double f0, double df, int n_freq, double* out_dev) {
    const int M = (ctx->tuning.psd_v2 & 2) ? 0 : lomb_czt_length(n_t, n_freq);
    if (M > 0) {
        const int grid = (int)std::min<int64_t>(n_slices, (int64_t)ctx->sm_count * 2);
        ITJL_CUDA(ctx, ctx->out3.reserve(lomb_czt_table_elems(n_t, n_freq, M) * sizeof(double2)));
        ITJL_CUDA(ctx, ctx->out2.reserve(lomb_czt_scratch_elems(n_freq, M, grid) * sizeof(double2)));
        std::vector<double2> roots((size_t)M);
        for (int k = 0; k < M; ++k) {
            const double ang = -2.0 * M_PI * (double)k / (double)M;
            roots[(size_t)k] = make_double2(cos(ang), sin(ang));
        }
        ITJL_CUDA(ctx, cudaMemcpyAsync(ctx->out3.p, roots.data(), roots.size() * sizeof(double2), cudaMemcpyHostToDevice, ctx->stream));
        cudaError_t e = lomb_czt_launch((const double*)ctx->data.p, n_slices, n_t, dt, f0, df, n_freq, M, (double2*)ctx->out3.p,
                                        (double2*)ctx->out2.p, grid, out_dev, ctx->stream);
        if (e != cudaSuccess) return fail(ctx, ITJL_ERR_CUDA, "launch k_lomb_czt: %s", cudaGetErrorString(e));
        ctx->launches += 2;
        ITJL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));          // roots stays alive until here
        return ITJL_OK;
    }
    cudaError_t e = lomb_scargle_launch((const double*)ctx->data.p, n_slices, n_t, dt, f0, df, n_freq, out_dev, ctx->stream);
    if (e != cudaSuccess) return fail(ctx, ITJL_ERR_CUDA, "launch k_lomb_scargle: %s", cudaGetErrorString(e));
    ctx->launches++;
    return ITJL_OK;
}

extern "C" {

int itjl_psd_lombscargle(itjl_ctx* ctx, const double* data, int64_t n_slices, int64_t n_t, double dt,
                         double f0, double df, int64_t n_freq, double* out) {
    int rc = check_data_args(ctx, data, n_slices, n_t, "itjl_psd_lombscargle");
    if (rc != ITJL_OK) return rc;
    if (!out || !(dt > 0.0) || !(df > 0.0) || !(f0 >= 0.0) || n_freq < 1 || n_freq > (1 << 24) || n_slices > 65535)
        return fail(ctx, ITJL_ERR_ARG, "itjl_psd_lombscargle: invalid argument (need dt, df > 0, f0 >= 0, 1 <= n_freq <= 2^24, n_slices <= 65535)");
    ITJL_LOCK(ctx);
    ctx = primary(ctx);
    ITJL_LOCK(ctx);
    NvtxRange nvtx("itjl_psd_lombscargle");
    ITJL_CUDA(ctx, cudaSetDevice(ctx->device));
    rc = upload_data(ctx, data, n_slices, n_t);
    if (rc != ITJL_OK) return rc;
    const size_t obytes = (size_t)n_slices * n_freq * sizeof(double);
    ITJL_CUDA(ctx, ctx->out.reserve(obytes));
    rc = lomb_scargle_dev(ctx, n_slices, (int)n_t, dt, f0, df, (int)n_freq, (double*)ctx->out.p);
    if (rc != ITJL_OK) return rc;
    rc = download_large(ctx, out, ctx->out.p, obytes);
    if (rc != ITJL_OK) return rc;
    ITJL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ITJL_OK;
}

// knee fit of the rows in `rows_dev` (row-major, ld doubles apart) on columns [i0, i0 + n); results in ctx->stats as
// three consecutive arrays of n_rows doubles: knee, amp, peak
static int knee_fit_dev(itjl_ctx* ctx, const double* rows_dev, int64_t n_rows, int64_t ld, int i0, int n, const double* freqs_dev,
                        double min_freq, double max_freq, int mode, int max_peaks, double* out3 = nullptr) {
    const int grid = knee_fit_grid(ctx->sm_count, n_rows);
    ITJL_CUDA(ctx, ctx->part.reserve((size_t)grid * 3 * n * sizeof(double)));
    if (!out3) ITJL_CUDA(ctx, ctx->stats.reserve((size_t)n_rows * 3 * sizeof(double)));
    double* k = out3 ? out3 : (double*)ctx->stats.p;
    cudaError_t e = knee_fit_launch(rows_dev, n_rows, ld, i0, n, freqs_dev, min_freq, max_freq, mode, max_peaks, k, k + n_rows,
                                    k + 2 * n_rows, (double*)ctx->part.p, grid, ctx->stream);
    if (e != cudaSuccess) return fail(ctx, ITJL_ERR_CUDA, "launch k_knee_fit: %s", cudaGetErrorString(e));
    ctx->launches++;
    return ITJL_OK;
}

// fooof_fit's frequency mask as a column window (freqs ascending)
static bool freq_window(const double* freqs, int64_t n_freq, double lo, double hi, int* i0, int* n) {
    int64_t a = 0, b = n_freq;
    while (a < n_freq && !(freqs[a] >= lo)) ++a;
    while (b > a && !(freqs[b - 1] <= hi)) --b;
    *i0 = (int)a; *n = (int)(b - a);
    return b - a >= 3;
}

int itjl_fit_knee(itjl_ctx* ctx, const double* psd, int64_t n_rows, int64_t n_freq, int32_t col_major,
                  const double* freqs, double min_freq, double max_freq, int32_t fooof, int32_t max_peaks,
                  double* knee_out, double* amp_out, double* peak_out) {
    if (!ctx) return fail(nullptr, ITJL_ERR_ARG, "itjl_fit_knee: null context");
    if (!psd || !freqs || !knee_out || n_rows < 1 || n_freq < 3 || n_freq > (1 << 24) || max_peaks < 0)
        return fail(ctx, ITJL_ERR_ARG, "itjl_fit_knee: invalid argument (need psd, freqs, knee_out, n_rows >= 1, 3 <= n_freq <= 2^24)");
    for (int64_t i = 1; i < n_freq; ++i)
        if (!(freqs[i] > freqs[i - 1])) return fail(ctx, ITJL_ERR_ARG, "itjl_fit_knee: freqs must be strictly ascending");
    ITJL_LOCK(ctx);
    ctx = primary(ctx);
    ITJL_LOCK(ctx);
    NvtxRange nvtx("itjl_fit_knee");
    ITJL_CUDA(ctx, cudaSetDevice(ctx->device));
    int i0 = 0, n = (int)n_freq;
    if (fooof < 0 || fooof > 3) return fail(ctx, ITJL_ERR_ARG, "itjl_fit_knee: fooof must be 0 .. 3");
    if (fooof == 1 && !freq_window(freqs, n_freq, min_freq, max_freq, &i0, &n))
        return fail(ctx, ITJL_ERR_ARG, "itjl_fit_knee: fewer than three frequencies inside [min_freq, max_freq]");
    const size_t bytes = (size_t)n_rows * n_freq * sizeof(double);
    ITJL_CUDA(ctx, ctx->out2.reserve(bytes));
    ITJL_CUDA(ctx, ctx->out3.reserve((size_t)n_freq * sizeof(double)));
    ITJL_CUDA(ctx, cudaMemcpyAsync(ctx->out3.p, freqs, (size_t)n_freq * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    if (col_major) {
        ITJL_CUDA(ctx, ctx->out.reserve(bytes));
        int rc = upload_bytes(ctx, ctx->out.p, psd, bytes);
        if (rc != ITJL_OK) return rc;
        cudaError_t e = transpose_launch((const double*)ctx->out.p, n_rows, n_freq, (double*)ctx->out2.p, ctx->stream);
        if (e != cudaSuccess) return fail(ctx, ITJL_ERR_CUDA, "launch k_transpose: %s", cudaGetErrorString(e));
        ctx->launches++;
    } else {
        int rc = upload_bytes(ctx, ctx->out2.p, psd, bytes);
        if (rc != ITJL_OK) return rc;
    }
    int rc = knee_fit_dev(ctx, (const double*)ctx->out2.p, n_rows, n_freq, i0, n, (const double*)ctx->out3.p, min_freq, max_freq,
                          fooof, max_peaks);
    if (rc != ITJL_OK) return rc;
    const double* k = (const double*)ctx->stats.p;
    ITJL_CUDA(ctx, cudaMemcpyAsync(knee_out, k, (size_t)n_rows * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (amp_out) ITJL_CUDA(ctx, cudaMemcpyAsync(amp_out, k + n_rows, (size_t)n_rows * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (peak_out) ITJL_CUDA(ctx, cudaMemcpyAsync(peak_out, k + 2 * n_rows, (size_t)n_rows * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    ITJL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ITJL_OK;
}

// LombScargle.autofrequency(dt:dt:n_t*dt; maximum_frequency = 1 / (2 dt)): df = 1 / (5 (n_t - 1) dt), f0 = df / 2
static void ls_grid(int64_t n_t, double dt, double* f0, double* df, int64_t* n_freq) {
    *df = 1.0 / (5.0 * (double)(n_t - 1) * dt);
    *f0 = 0.5 * *df;
    *n_freq = (int64_t)floor((0.5 / dt - *f0) / *df + 1e-9) + 1;
}

int64_t itjl_acw_knee_nfreq(int64_t n_t, int32_t has_nan) {
    if (n_t < 2 || n_t > 65536) return -1;
    if (!has_nan) return next_fast_len((int)n_t) / 2;
    double f0, df; int64_t n;
    ls_grid(n_t, 1.0, &f0, &df, &n);
    return n;
}

int itjl_acw_knee(itjl_ctx* ctx, const double* data, int64_t n_slices, int64_t n_t, double fs, double freq_lo, double freq_hi,
                  uint32_t flags, int32_t max_peaks, double* knee_out, double* psd_out, int64_t psd_capacity,
                  double* freqs_out, int64_t* n_freq_out) {
    int rc = check_data_args(ctx, data, n_slices, n_t, "itjl_acw_knee");
    if (rc != ITJL_OK) return rc;
    if (!knee_out || !(fs > 0.0) || max_peaks < 0 || n_t > 65536) return fail(ctx, ITJL_ERR_ARG, "itjl_acw_knee: invalid argument (fs > 0, n_t <= 65536)");
    ITJL_LOCK(ctx);
    ctx = primary(ctx);
    ITJL_LOCK(ctx);
    NvtxRange nvtx("itjl_acw_knee");
    ITJL_CUDA(ctx, cudaSetDevice(ctx->device));
    bool has_nan = false;
    for (int64_t i = 0, m = n_slices * n_t; i < m; ++i) if (data[i] != data[i]) { has_nan = true; break; }
    const double dt = 1.0 / fs;
    rc = upload_data(ctx, data, n_slices, n_t);
    if (rc != ITJL_OK) return rc;
    int64_t nb = 0;
    std::vector<double> freqs;
    if (!has_nan) {
        FftPlan plan;
        const int nfft = next_fast_len((int)n_t);
        if (!fft_plan_make(nfft, &plan)) return fail(ctx, ITJL_ERR_ARG, "itjl_acw_knee: n_t too large");
        nb = nfft / 2;
        std::vector<double> win((size_t)n_t);
        std::vector<double2> roots((size_t)nfft);
        double win_ss = 0.0;
        pgram_make_tables((int)n_t, nfft, nullptr, win.data(), nullptr, roots.data(), &win_ss);
        const int grid = (int)std::min<int64_t>((n_slices + 1) / 2, (int64_t)ctx->sm_count * 4);
        ITJL_CUDA(ctx, ctx->scratch.reserve(win.size() * sizeof(double)));
        ITJL_CUDA(ctx, ctx->raw.reserve(roots.size() * sizeof(double2)));
        ITJL_CUDA(ctx, ctx->out2.reserve((size_t)grid * 2 * nfft * sizeof(double2)));
        ITJL_CUDA(ctx, cudaMemcpyAsync(ctx->scratch.p, win.data(), win.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        ITJL_CUDA(ctx, cudaMemcpyAsync(ctx->raw.p, roots.data(), roots.size() * sizeof(double2), cudaMemcpyHostToDevice, ctx->stream));
        ITJL_CUDA(ctx, ctx->out.reserve((size_t)n_slices * nb * sizeof(double)));
        cudaError_t e = pgram_data_launch((const double*)ctx->data.p, n_slices, (int)n_t, plan, (const double*)ctx->scratch.p,
                                          (const double2*)ctx->raw.p, 1.0 / (fs * win_ss), (double2*)ctx->out2.p, grid,
                                          (double*)ctx->out.p, ctx->stream);
        if (e != cudaSuccess) return fail(ctx, ITJL_ERR_CUDA, "launch k_pgram_data: %s", cudaGetErrorString(e));
        ctx->launches++;
        ITJL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));          // win / roots are block-local
        freqs.resize((size_t)nb);
        for (int64_t k = 0; k < nb; ++k) freqs[(size_t)k] = (double)(k + 1) * fs / (double)nfft;
    } else {
        if (n_slices > 65535) return fail(ctx, ITJL_ERR_ARG, "itjl_acw_knee: more than 65535 slices with missing samples");
        double f0, df;
        ls_grid(n_t, dt, &f0, &df, &nb);
        ITJL_CUDA(ctx, ctx->out.reserve((size_t)n_slices * nb * sizeof(double)));
        { const int rcl = lomb_scargle_dev(ctx, n_slices, (int)n_t, dt, f0, df, (int)nb, (double*)ctx->out.p); if (rcl != ITJL_OK) return rcl; }
        freqs.resize((size_t)nb);
        for (int64_t k = 0; k < nb; ++k) freqs[(size_t)k] = f0 + df * (double)k;
    }
    if (n_freq_out) *n_freq_out = nb;
    // average_over_trials: one mean spectrum
    const bool average = (flags & 2u) != 0;
    int64_t n_out = n_slices;
    const double* spec = (const double*)ctx->out.p;                  // (n_out x nb) column-major
    if (average) {
        ITJL_CUDA(ctx, ctx->gather.reserve((size_t)nb * sizeof(double)));
        cudaError_t e = col_mean_launch(spec, n_slices, nb, (double*)ctx->gather.p, ctx->stream);
        if (e != cudaSuccess) return fail(ctx, ITJL_ERR_CUDA, "launch k_col_mean: %s", cudaGetErrorString(e));
        ctx->launches++;
        spec = (const double*)ctx->gather.p; n_out = 1;
    }
    if (psd_out) {
        if (psd_capacity < n_out * nb) return fail(ctx, ITJL_ERR_ARG, "itjl_acw_knee: psd_out too small");
        rc = download_large(ctx, psd_out, spec, (size_t)n_out * nb * sizeof(double));
        if (rc != ITJL_OK) return rc;
    }
    if (freqs_out) memcpy(freqs_out, freqs.data(), (size_t)nb * sizeof(double));
    // freqlims default (freqs[1], freqs[end]) (acw.jl:255-257)
    const double lo = freq_lo == freq_lo ? freq_lo : freqs.front(), hi = freq_hi == freq_hi ? freq_hi : freqs.back();
    int i0 = 0, n = 0;
    if (!freq_window(freqs.data(), nb, lo, hi, &i0, &n)) return fail(ctx, ITJL_ERR_ARG, "itjl_acw_knee: fewer than three frequencies inside freqlims");
    // rows contiguous for the fit
    ITJL_CUDA(ctx, ctx->out2.reserve((size_t)n_out * nb * sizeof(double)));
    if (n_out > 1) {
        cudaError_t e = transpose_launch(spec, n_out, nb, (double*)ctx->out2.p, ctx->stream);
        if (e != cudaSuccess) return fail(ctx, ITJL_ERR_CUDA, "launch k_transpose: %s", cudaGetErrorString(e));
        ctx->launches++;
    } else {
        ITJL_CUDA(ctx, cudaMemcpyAsync(ctx->out2.p, spec, (size_t)nb * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    }
    ITJL_CUDA(ctx, ctx->out3.reserve((size_t)nb * sizeof(double)));
    ITJL_CUDA(ctx, cudaMemcpyAsync(ctx->out3.p, freqs.data(), (size_t)nb * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    // fooof_fit always masks to [lo, hi]; oscillation_peak = false stops after the first Lorentzian (utils.jl:622-624)
    rc = knee_fit_dev(ctx, (const double*)ctx->out2.p, n_out, nb, i0, n, (const double*)ctx->out3.p, lo, hi, (flags & 1u) ? 1 : 0,
                      (flags & 1u) ? max_peaks : 0);
    if (rc != ITJL_OK) return rc;
    ITJL_CUDA(ctx, cudaMemcpyAsync(knee_out, ctx->stats.p, (size_t)n_out * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    ITJL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));          // freqs stays alive until here
    return ITJL_OK;
}

int itjl_abc_distances_combined(itjl_model* m, const double* theta, int64_t n_particles, int32_t n_theta,
                                uint64_t seed, uint64_t generation, int64_t particle_offset, const double* lags_freqs,
                                const double* weights, int32_t n_weights, double data_tau, double data_osc, double* dist_out) {
    if (!m) return fail(nullptr, ITJL_ERR_ARG, "null model");
    if (!theta || !dist_out || !weights || n_particles < 1 || n_theta < 1) return fail(m->ctx, ITJL_ERR_ARG, "itjl_abc_distances_combined: null / empty argument");
    ITJL_LOCK(m->ctx);
    m = primary(m);
    itjl_ctx* ctx = m->ctx;
    ITJL_LOCK(ctx);
    NvtxRange nvtx("itjl_abc_distances_combined");
    const itjl_model_desc& d = m->d;
    const bool psd = d.summary == ITJL_SUMMARY_PSD || d.summary == ITJL_SUMMARY_PGRAM;
    const bool osc_psd = psd && d.kind == ITJL_KIND_OU_OSC;
    if (n_weights != (osc_psd ? 3 : 2)) return fail(ctx, ITJL_ERR_ARG, "itjl_abc_distances_combined: two weights (three for the PSD summary of the oscillation model)");
    if (psd && !lags_freqs) return fail(ctx, ITJL_ERR_ARG, "itjl_abc_distances_combined: lags_freqs (the model's frequencies) required for a PSD summary");
    if (psd && d.n_stats < 3) return fail(ctx, ITJL_ERR_ARG, "itjl_abc_distances_combined: fewer than three frequencies");
    ITJL_CUDA(ctx, cudaSetDevice(ctx->device));
    const int64_t ns = d.n_stats;
    if (psd) {
        for (int64_t i = 1; i < ns; ++i)
            if (!(lags_freqs[i] > lags_freqs[i - 1])) return fail(ctx, ITJL_ERR_ARG, "itjl_abc_distances_combined: lags_freqs must be strictly ascending");
        ITJL_CUDA(ctx, ctx->out3.reserve((size_t)ns * sizeof(double)));
        ITJL_CUDA(ctx, cudaMemcpyAsync(ctx->out3.p, lags_freqs, (size_t)ns * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    }
    int rc = upload(ctx, ctx->theta, theta, (size_t)n_particles * n_theta * sizeof(double));
    if (rc != ITJL_OK) return rc;
    ITJL_CUDA(ctx, ctx->dist.reserve((size_t)n_particles * sizeof(double)));
    // the statistics stay on the device: batches of at most 16384 particles (n_stats doubles each)
    const int64_t batch = std::min<int64_t>(n_particles, 16384);
    ITJL_CUDA(ctx, ctx->stats.reserve((size_t)batch * ns * sizeof(double)));
    ITJL_CUDA(ctx, ctx->gather.reserve((size_t)batch * 3 * sizeof(double)));
    for (int64_t lo = 0; lo < n_particles; lo += batch) {
        const int64_t nb = std::min<int64_t>(batch, n_particles - lo);
        rc = launch_distances(m, (const double*)ctx->theta.p + lo, nb, n_theta, seed, generation, particle_offset + lo,
                              nullptr, nullptr, nullptr, (double*)ctx->dist.p + lo, (double*)ctx->stats.p, n_particles);
        if (rc != ITJL_OK) return rc;
        double* fit = (double*)ctx->gather.p;                 // [knee or tau | amp | peak] x nb
        if (psd) {
            // find_knee_frequency defaults (min_freq = freqs[1], max_freq = freqs[end]); the peak search window of the
            // oscillation model is its freqlims, an open interval around every selected frequency: the whole axis
            rc = knee_fit_dev(ctx, (const double*)ctx->stats.p, nb, ns, 0, (int)ns, (const double*)ctx->out3.p, lags_freqs[0],
                              lags_freqs[ns - 1], 0, 0, fit);
            if (rc != ITJL_OK) return rc;
        } else {
            cudaError_t e = acw_tau_launch((const double*)ctx->stats.p, nb, ns, (int)ns, d.dt, fit, ctx->stream);
            if (e != cudaSuccess) return fail(ctx, ITJL_ERR_CUDA, "launch k_acw_tau: %s", cudaGetErrorString(e));
            ctx->launches++;
        }
        cudaError_t e = combine_launch((double*)ctx->dist.p + lo, nb, fit, psd ? 1 : 0, osc_psd ? fit + 2 * nb : nullptr, weights[0],
                                       weights[1], osc_psd ? weights[2] : 0.0, data_tau, data_osc, ctx->stream);
        if (e != cudaSuccess) return fail(ctx, ITJL_ERR_CUDA, "launch k_combine: %s", cudaGetErrorString(e));
        ctx->launches++;
    }
    ITJL_CUDA(ctx, cudaMemcpyAsync(dist_out, ctx->dist.p, (size_t)n_particles * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    ITJL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ITJL_OK;
}

int itjl_fit_expdecay(itjl_ctx* ctx, const double* acf, int64_t n_rows, int64_t n_lags, double dt, double* tau_out) {
    if (!ctx) return fail(nullptr, ITJL_ERR_ARG, "itjl_fit_expdecay: null context");
    if (!acf || !tau_out || n_rows < 1 || n_lags < 1 || !(dt > 0.0)) return fail(ctx, ITJL_ERR_ARG, "itjl_fit_expdecay: invalid argument");
    ITJL_LOCK(ctx);
    ctx = primary(ctx);
    ITJL_LOCK(ctx);
    ITJL_CUDA(ctx, cudaSetDevice(ctx->device));
    // rows are read with stride 1 along lags: transpose the column-major input on the host
    std::vector<double> rows((size_t)n_rows * n_lags);
    for (int64_t r = 0; r < n_rows; ++r)
        for (int64_t k = 0; k < n_lags; ++k) rows[(size_t)r * n_lags + k] = acf[r + k * n_rows];
    ITJL_CUDA(ctx, ctx->out.reserve(rows.size() * sizeof(double)));
    ITJL_CUDA(ctx, ctx->out3.reserve((size_t)n_rows * sizeof(double)));
    ITJL_CUDA(ctx, cudaMemcpyAsync(ctx->out.p, rows.data(), rows.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    cudaError_t e = acw_tau_launch((const double*)ctx->out.p, n_rows, n_lags, (int)n_lags, dt, (double*)ctx->out3.p, ctx->stream);
    if (e != cudaSuccess) return fail(ctx, ITJL_ERR_CUDA, "launch k_acw_tau: %s", cudaGetErrorString(e));
    ctx->launches++;
    ITJL_CUDA(ctx, cudaMemcpyAsync(tau_out, ctx->out3.p, (size_t)n_rows * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    ITJL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ITJL_OK;
}

int itjl_fit_expdecay_3_parameters(itjl_ctx* ctx, const double* acf, int64_t n_rows, int64_t n_lags, double dt, double* tau_out) {
    if (!ctx) return fail(nullptr, ITJL_ERR_ARG, "itjl_fit_expdecay_3_parameters: null context");
    if (!acf || !tau_out || n_rows < 1 || n_lags < 1 || n_lags > INT32_MAX || !(dt > 0.0))
        return fail(ctx, ITJL_ERR_ARG, "itjl_fit_expdecay_3_parameters: invalid argument");
    ITJL_LOCK(ctx);
    ctx = primary(ctx);
    ITJL_LOCK(ctx);
    ITJL_CUDA(ctx, cudaSetDevice(ctx->device));
    // device layout: [n_rows][n_lags] rows, then the lag axis
    std::vector<double> rows((size_t)(n_rows + 1) * n_lags);
    for (int64_t r = 0; r < n_rows; ++r)
        for (int64_t k = 0; k < n_lags; ++k) rows[(size_t)r * n_lags + k] = acf[r + k * n_rows];
    for (int64_t k = 0; k < n_lags; ++k) rows[(size_t)n_rows * n_lags + k] = (double)k * dt;
    ITJL_CUDA(ctx, ctx->out.reserve(rows.size() * sizeof(double)));
    ITJL_CUDA(ctx, ctx->out3.reserve((size_t)n_rows * sizeof(double)));
    ITJL_CUDA(ctx, cudaMemcpyAsync(ctx->out.p, rows.data(), rows.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    const double* d_rows = (const double*)ctx->out.p;
    cudaError_t e = expdecay3_launch(d_rows, n_rows, n_lags, (int)n_lags, d_rows + (size_t)n_rows * n_lags, (double*)ctx->out3.p,
                                     (int)std::min<int64_t>(n_rows, (int64_t)ctx->sm_count * 8), ctx->stream);
    if (e != cudaSuccess) return fail(ctx, ITJL_ERR_CUDA, "launch k_expdecay3_fit: %s", cudaGetErrorString(e));
    ctx->launches++;
    ITJL_CUDA(ctx, cudaMemcpyAsync(tau_out, ctx->out3.p, (size_t)n_rows * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    ITJL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ITJL_OK;
}

}  // extern "C"

// acw() on the matrix in ctx->data (n_slices x n_t, slice index fastest); `data` != null: upload it first
static int acw_core(itjl_ctx* ctx, const double* data, int64_t n_slices, int64_t n_t, int64_t ss, int64_t ts, double fs,
                    uint32_t typemask, int32_t average_over_trials, int64_t* n_lags_io, int32_t* k0, int32_t* k50,
                    int32_t* ke, double* auc, double* tau, double* acf_out, int64_t acf_capacity) {
    int rc = ITJL_OK;
    ctx->pending.clear(); ctx->pin_small_used = 0;        // leftovers of a call that failed half way
    if (!n_lags_io) return fail(ctx, ITJL_ERR_ARG, "itjl_acw: n_lags_io is required");
    if (!(fs > 0.0)) return fail(ctx, ITJL_ERR_ARG, "itjl_acw: fs must be positive");
    if ((typemask & ~63u) || typemask == 0) return fail(ctx, ITJL_ERR_ARG, "itjl_acw: No ACW types specified / unknown type bit");
    if ((typemask & ITJL_TAU) && (typemask & ITJL_TAU3)) return fail(ctx, ITJL_ERR_ARG, "itjl_acw: ITJL_TAU and ITJL_TAU3 share the tau output: request one");
    if (((typemask & ITJL_ACW0) && !k0) || ((typemask & ITJL_ACW50) && !k50) || ((typemask & ITJL_ACWEULER) && !ke) ||
        ((typemask & ITJL_AUC) && !auc) || ((typemask & (ITJL_TAU | ITJL_TAU3)) && !tau))
        return fail(ctx, ITJL_ERR_ARG, "itjl_acw: output pointer missing for a requested acwtype");
    const int64_t user_lags = *n_lags_io;
    if (user_lags < 0 || user_lags > n_t) return fail(ctx, ITJL_ERR_ARG, "itjl_acw: n_lags out of range");
    ITJL_CUDA(ctx, cudaSetDevice(ctx->device));
    const double dt = 1.0 / fs;
    const int nt = (int)n_t;
    const int reach = acf_max_reach(nt, ctx->max_smem_optin);
    if (reach < 40) return fail(ctx, ITJL_ERR_ARG, "itjl_acw: n_t too large for the shared-memory ACF kernel");

    if (data) {
        rc = upload_data(ctx, data, n_slices, n_t, ss, ts);
        if (rc != ITJL_OK) return rc;
    }
    ITJL_CUDA(ctx, ctx->flags.reserve(((size_t)n_slices * 6 + 4) * sizeof(int32_t)));
    int32_t* d_done = (int32_t*)ctx->flags.p;
    int32_t* d_nan = d_done + 6 * n_slices;
    int32_t* d_redo = d_done + 4 * n_slices;
    int32_t* d_list = d_done + 5 * n_slices;
    int32_t* d_k0 = d_done + n_slices;
    int32_t* d_k50 = d_k0 + n_slices;
    int32_t* d_ke = d_k50 + n_slices;

    const int64_t n_out = average_over_trials ? 1 : n_slices;
    std::vector<int32_t> h_k0((size_t)n_out), h_k50((size_t)n_out), h_ke((size_t)n_out);
    const double* acf_rows = nullptr;      // device rows used by the epilogues
    int64_t acf_ld = 0;
    int64_t lags_avail = 0;
    int64_t cap = 0;
    bool all_streamed = false;       // every slice holds the streaming pass's 32 lags
    int64_t n_redo = 0;              // slices redone by the fp64 kernel (their indices stay in d_list)
    int32_t has_nan = 0;             // some slice has a NaN (the reference then takes comp_ac_time_missing)
    int stream_w = 0;                // lags kept by the streaming pass of this call (0: not used)
    cudaError_t e;

    if (!average_over_trials) {
        // NaN-aware centring is the identity on complete data, so one code path serves both
        // comp_ac_fft (acw.jl:146) and comp_ac_time_missing (acw.jl:148-153).
        cap = reach;
        ITJL_CUDA(ctx, ctx->out.reserve((size_t)n_slices * cap * sizeof(double)));
        ITJL_CUDA(ctx, cudaMemsetAsync(d_nan, 0, sizeof(int32_t), ctx->stream));
        if (ctx->timing) cudaEventRecord(ctx->ev0, ctx->stream);
        const bool use_stream = ctx->tuning.acw_stream != 0 && nt >= 256 && acw_stream_lags() <= reach;
        std::vector<int32_t> h_redo;
        if (use_stream) {
            // one streaming read of the data: raw lagged products of the first stream_w lags
            stream_w = ctx->tuning.acw_stream == 8 ? 8 : (ctx->tuning.acw_stream == 16 ? 16 : 32);
            int n_seg = 1, seg_len = nt;
            acw_stream_segments(nt, n_slices, ctx->sm_count, &n_seg, &seg_len, ctx->tuning.acw_waves, stream_w);
            ITJL_CUDA(ctx, ctx->part.reserve(acw_stream_part_doubles(n_slices, n_seg) * sizeof(double)));
            e = acw_stream_launch((const double*)ctx->data.p, n_slices, nt, (double*)ctx->part.p, n_seg, seg_len,
                                  (double*)ctx->out.p, cap, d_done, d_k0, d_k50, d_ke, d_redo, stream_w, ctx->stream);
            if (e != cudaSuccess) return fail(ctx, ITJL_ERR_CUDA, "launch k_acw_stream: %s", cudaGetErrorString(e));
            ctx->launches += 2;
            if (ctx->timing) cudaEventRecord(ctx->ev1, ctx->stream);     // the streaming pass (k_acw_stream + k_acw_finish) alone
            h_redo.resize((size_t)n_slices);
            { const int rq = d2h_queue(ctx, h_redo.data(), d_redo, (size_t)n_slices * 4); if (rq != ITJL_OK) return rq; }
            { const int rq = d2h_flush(ctx); if (rq != ITJL_OK) return rq; }
            std::vector<int32_t> list;
            for (int64_t i = 0; i < n_slices; ++i) if (h_redo[i]) list.push_back((int32_t)i);
            all_streamed = list.empty();
            n_redo = (int64_t)list.size();
            if (!list.empty()) {
                // slices that need more than 32 lags (or exact centring): shared-memory kernel
                ITJL_CUDA(ctx, cudaMemcpyAsync(d_list, list.data(), list.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
                e = acw_scan_launch((const double*)ctx->data.p, n_slices, nt, 1, reach, (double*)ctx->out.p, cap,
                                    d_done, d_k0, d_k50, d_ke, d_nan, d_list, (int64_t)list.size(),
                                    ctx->max_smem_optin, ctx->stream);
                if (e != cudaSuccess) return fail(ctx, ITJL_ERR_CUDA, "launch k_acf_slices(scan list): %s", cudaGetErrorString(e));
                ctx->launches++;
                ITJL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));    // list stays alive until here
            }
        } else {
            e = acw_scan_launch((const double*)ctx->data.p, n_slices, nt, 1, reach, (double*)ctx->out.p, cap,
                                d_done, d_k0, d_k50, d_ke, d_nan, nullptr, 0, ctx->max_smem_optin, ctx->stream);
            if (e != cudaSuccess) return fail(ctx, ITJL_ERR_CUDA, "launch k_acf_slices(scan): %s", cudaGetErrorString(e));
            ctx->launches++;
            if (ctx->timing) cudaEventRecord(ctx->ev1, ctx->stream);
        }
        if (ctx->timing) {
            cudaEventSynchronize(ctx->ev1);
            float ms = 0.f;
            cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
            ctx->timed_ms += ms; ctx->timed_launches++;
        }
        {
            int rq = d2h_queue(ctx, h_k0.data(), d_k0, (size_t)n_slices * 4);
            if (rq == ITJL_OK) rq = d2h_queue(ctx, h_k50.data(), d_k50, (size_t)n_slices * 4);
            if (rq == ITJL_OK) rq = d2h_queue(ctx, h_ke.data(), d_ke, (size_t)n_slices * 4);
            if (rq == ITJL_OK) rq = d2h_queue(ctx, &has_nan, d_nan, sizeof(int32_t));
            if (rq == ITJL_OK) rq = d2h_flush(ctx);
            if (rq != ITJL_OK) return rq;
        }
        if (has_nan && user_lags > 0) {
            // NaN data + user n_lags: the reference only computes n_lags lags (acw.jl:151-152)
            for (int64_t i = 0; i < n_slices; ++i) {
                if (h_k0[i] >= user_lags) h_k0[i] = -1;
                if (h_k50[i] >= user_lags) h_k50[i] = -1;
                if (h_ke[i] >= user_lags) h_ke[i] = -1;
            }
        }
        acf_rows = (const double*)ctx->out.p; acf_ld = cap;
    } else {
        // acf = mean(acf, dims = trial_dims) over ALL lags (acw.jl:156-158): grow the lag
        // range until the mean ACF crosses zero.
        int L = (int)(user_lags > 0 ? user_lags : 64);
        if (L > reach) L = reach;
        bool first = true;
        ITJL_CUDA(ctx, cudaMemsetAsync(d_nan, 0, sizeof(int32_t), ctx->stream));
        for (;;) {
            cap = reach;
            rc = acf_all(ctx, n_slices, nt, L, 1, cap, first, first ? d_nan : nullptr);
            if (rc != ITJL_OK) return rc;
            if (first) ITJL_CUDA(ctx, cudaMemcpyAsync(&has_nan, d_nan, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
            first = false;
            ITJL_CUDA(ctx, ctx->out2.reserve((size_t)reach * sizeof(double)));
            e = acf_mean_launch((const double*)ctx->out.p, n_slices, cap, L, (double*)ctx->out2.p, ctx->stream);
            if (e != cudaSuccess) return fail(ctx, ITJL_ERR_CUDA, "launch k_acf_mean: %s", cudaGetErrorString(e));
            ctx->launches++;
            e = acf_cross_launch((const double*)ctx->out2.p, 1, reach, L, d_k0, d_k50, d_ke, ctx->stream);
            if (e != cudaSuccess) return fail(ctx, ITJL_ERR_CUDA, "launch k_acf_cross: %s", cudaGetErrorString(e));
            ctx->launches++;
            ITJL_CUDA(ctx, cudaMemcpyAsync(h_k0.data(), d_k0, 4, cudaMemcpyDeviceToHost, ctx->stream));
            ITJL_CUDA(ctx, cudaMemcpyAsync(h_k50.data(), d_k50, 4, cudaMemcpyDeviceToHost, ctx->stream));
            ITJL_CUDA(ctx, cudaMemcpyAsync(h_ke.data(), d_ke, 4, cudaMemcpyDeviceToHost, ctx->stream));
            ITJL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            if (h_k0[0] >= 0 || L >= reach) break;
            // NaN data + user n_lags: the reference only computes n_lags lags (acw.jl:151-152), a later
            // crossing of the mean ACF does not exist for it
            if (has_nan && user_lags > 0) break;
            L = (L * 4 > reach) ? reach : L * 4;
        }
        lags_avail = L;
        acf_rows = (const double*)ctx->out2.p; acf_ld = reach;
    }

    // n_lags (acw.jl:203-210)
    int64_t n_lags = user_lags;
    if (n_lags == 0) {
        int32_t mx = -1;
        for (int64_t i = 0; i < n_out; ++i) if (h_k0[i] > mx) mx = h_k0[i];
        n_lags = (mx < 0) ? n_t : (int64_t)ceil(1.1 * (double)mx);
        if (n_lags > n_t) n_lags = n_t;
    }
    if (n_lags > reach) return fail(ctx, ITJL_ERR_ARG, "itjl_acw: n_lags exceeds the shared-memory reach for this n_t");
    *n_lags_io = n_lags;

    // make sure every row holds n_lags lags - and, for the AUC of complete data, the first slice's whole
    // window: the reference integrates acf[0 : acw0) of the full-length comp_ac_fft ACF before it cuts the
    // rows to n_lags (acw.jl:194-213); with NaNs it only ever has n_lags lags
    int64_t fill_lags = n_lags;
    if ((typemask & ITJL_AUC) && !has_nan && h_k0[0] > fill_lags) fill_lags = h_k0[0];
    if (!average_over_trials) {
        if (!(all_streamed && fill_lags <= stream_w)) {
            // streamed slices already hold stream_w lags: when that is enough only the redone ones need filling
            const bool only_list = n_redo > 0 && fill_lags <= stream_w;
            e = acf_fill_launch((const double*)ctx->data.p, n_slices, nt, 1, (int)fill_lags, (double*)ctx->out.p, cap,
                                d_done, only_list ? d_list : nullptr, only_list ? n_redo : 0,
                                ctx->max_smem_optin, ctx->stream);
            if (e != cudaSuccess) return fail(ctx, ITJL_ERR_CUDA, "launch k_acf_slices(fill): %s", cudaGetErrorString(e));
            ctx->launches++;
        }
    } else if (fill_lags > lags_avail) {
        rc = acf_all(ctx, n_slices, nt, (int)fill_lags, 1, cap, false);
        if (rc != ITJL_OK) return rc;
        e = acf_mean_launch((const double*)ctx->out.p, n_slices, cap, (int)fill_lags, (double*)ctx->out2.p, ctx->stream);
        if (e != cudaSuccess) return fail(ctx, ITJL_ERR_CUDA, "launch k_acf_mean: %s", cudaGetErrorString(e));
        ctx->launches++;
    }

    if (typemask & ITJL_ACW0) memcpy(k0, h_k0.data(), (size_t)n_out * 4);
    if (typemask & ITJL_ACW50) memcpy(k50, h_k50.data(), (size_t)n_out * 4);
    if (typemask & ITJL_ACWEULER) memcpy(ke, h_ke.data(), (size_t)n_out * 4);

    ITJL_CUDA(ctx, ctx->out3.reserve((size_t)n_out * 2 * sizeof(double) + 64 * sizeof(int)));
    double* d_auc = (double*)ctx->out3.p;
    double* d_tau = d_auc + n_out;
    int* d_strides = (int*)(d_tau + n_out);
    std::vector<int> strides;                      // must outlive the async copy
    if (typemask & ITJL_AUC) {
        // window = floor(Int, acw0_sample[1]) of the FIRST slice (acw.jl:194)
        const int w = h_k0[0];
        if (w < 2 || w > fill_lags) {
            for (int64_t i = 0; i < n_out; ++i) auc[i] = NAN;    // reference throws here
        } else {
            prime_strides(w, strides);
            ITJL_CUDA(ctx, cudaMemcpyAsync(d_strides, strides.data(), strides.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
            e = acw_auc_launch(acf_rows, n_out, acf_ld, w, dt, d_strides, (int)strides.size(), d_auc, ctx->stream);
            if (e != cudaSuccess) return fail(ctx, ITJL_ERR_CUDA, "launch k_acw_auc: %s", cudaGetErrorString(e));
            ctx->launches++;
            { const int rq = d2h_queue(ctx, auc, d_auc, (size_t)n_out * sizeof(double)); if (rq != ITJL_OK) return rq; }
        }
    }
    if (typemask & ITJL_TAU) {
        e = acw_tau_launch(acf_rows, n_out, acf_ld, (int)n_lags, dt, d_tau, ctx->stream);
        if (e != cudaSuccess) return fail(ctx, ITJL_ERR_CUDA, "launch k_acw_tau: %s", cudaGetErrorString(e));
        ctx->launches++;
        { const int rq = d2h_queue(ctx, tau, d_tau, (size_t)n_out * sizeof(double)); if (rq != ITJL_OK) return rq; }
    }
    std::vector<double> lag_axis;                  // must outlive the async copy
    if (typemask & ITJL_TAU3) {
        // skip_zero_lag = true (acw.jl:212-216): fit_expdecay_3_parameters over lags[2:end]
        lag_axis.resize((size_t)n_lags);
        for (int64_t k = 0; k < n_lags; ++k) lag_axis[(size_t)k] = (double)k * dt;
        ITJL_CUDA(ctx, ctx->gather.reserve((size_t)n_lags * sizeof(double)));
        ITJL_CUDA(ctx, cudaMemcpyAsync(ctx->gather.p, lag_axis.data(), (size_t)n_lags * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        e = expdecay3_launch(acf_rows, n_out, acf_ld, (int)n_lags, (const double*)ctx->gather.p, d_tau,
                             (int)std::min<int64_t>(n_out, (int64_t)ctx->sm_count * 8), ctx->stream);
        if (e != cudaSuccess) return fail(ctx, ITJL_ERR_CUDA, "launch k_expdecay3_fit: %s", cudaGetErrorString(e));
        ctx->launches++;
        { const int rq = d2h_queue(ctx, tau, d_tau, (size_t)n_out * sizeof(double)); if (rq != ITJL_OK) return rq; }
    }
    // the ACF matrix (n_out x n_lags, column-major) stays on the device for itjl_acw_acf(); acf_out fetches it right away
    {
        if (acf_out && acf_capacity < n_out * n_lags) return fail(ctx, ITJL_ERR_ARG, "itjl_acw: acf_out capacity too small");
        const size_t obytes = (size_t)n_out * n_lags * sizeof(double);
        ctx->keep_rows = ctx->keep_lags = 0;
        ITJL_CUDA(ctx, ctx->acf_keep.reserve(obytes));
        e = acf_export_launch(acf_rows, n_out, acf_ld, (int)n_lags, (double*)ctx->acf_keep.p, ctx->stream);
        if (e != cudaSuccess) return fail(ctx, ITJL_ERR_CUDA, "launch k_acf_export: %s", cudaGetErrorString(e));
        ctx->launches++;
        ctx->keep_rows = n_out; ctx->keep_lags = n_lags;
        if (acf_out) { const int rcd = download_large(ctx, acf_out, ctx->acf_keep.p, obytes); if (rcd != ITJL_OK) return rcd; }
    }
    return d2h_flush(ctx);
}

extern "C" {

int itjl_acw_acf(itjl_ctx* ctx, double* acf_out, int64_t acf_capacity, int64_t* n_rows_out, int64_t* n_lags_out) {
    if (!ctx) return fail(nullptr, ITJL_ERR_ARG, "itjl_acw_acf: null context");
    ITJL_LOCK(ctx);
    ctx = primary(ctx);
    ITJL_LOCK(ctx);
    if (ctx->keep_rows < 1 || ctx->keep_lags < 1) return fail(ctx, ITJL_ERR_ARG, "itjl_acw_acf: no acw result on this context");
    if (n_rows_out) *n_rows_out = ctx->keep_rows;
    if (n_lags_out) *n_lags_out = ctx->keep_lags;
    if (!acf_out) return ITJL_OK;                          // shape query
    if (acf_capacity < ctx->keep_rows * ctx->keep_lags) return fail(ctx, ITJL_ERR_ARG, "itjl_acw_acf: acf_out capacity too small");
    ITJL_CUDA(ctx, cudaSetDevice(ctx->device));
    const int rc = download_large(ctx, acf_out, ctx->acf_keep.p, (size_t)ctx->keep_rows * ctx->keep_lags * sizeof(double));
    if (rc != ITJL_OK) return rc;
    ITJL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ITJL_OK;
}

int itjl_acw(itjl_ctx* ctx, const double* data, int64_t n_slices, int64_t n_t, double fs, uint32_t typemask,
             int32_t average_over_trials, int64_t* n_lags_io, int32_t* k0, int32_t* k50, int32_t* ke,
             double* auc, double* tau, double* acf_out, int64_t acf_capacity) {
    return itjl_acw_strided(ctx, data, n_slices, n_t, 1, n_slices, fs, typemask, average_over_trials, n_lags_io,
                            k0, k50, ke, auc, tau, acf_out, acf_capacity);
}

int itjl_acw_strided(itjl_ctx* ctx, const double* data, int64_t n_slices, int64_t n_t, int64_t slice_stride,
                     int64_t time_stride, double fs, uint32_t typemask, int32_t average_over_trials,
                     int64_t* n_lags_io, int32_t* k0, int32_t* k50, int32_t* ke, double* auc, double* tau,
                     double* acf_out, int64_t acf_capacity) {
    int rc = check_data_args(ctx, data, n_slices, n_t, "itjl_acw");
    if (rc != ITJL_OK) return rc;
    if (slice_stride < 1 || time_stride < 1) return fail(ctx, ITJL_ERR_ARG, "itjl_acw: strides must be positive");
    ITJL_LOCK(ctx);
    ctx = primary(ctx);
    ITJL_LOCK(ctx);
    NvtxRange nvtx("itjl_acw");
    return acw_core(ctx, data, n_slices, n_t, slice_stride, time_stride, fs, typemask, average_over_trials, n_lags_io,
                    k0, k50, ke, auc, tau, acf_out, acf_capacity);
}

int itjl_data_upload(itjl_ctx* ctx, const double* data, int64_t n_slices, int64_t n_t, int64_t slice_stride,
                     int64_t time_stride) {
    int rc = check_data_args(ctx, data, n_slices, n_t, "itjl_data_upload");
    if (rc != ITJL_OK) return rc;
    if (slice_stride < 1 || time_stride < 1) return fail(ctx, ITJL_ERR_ARG, "itjl_data_upload: strides must be positive");
    ITJL_LOCK(ctx);
    ctx = primary(ctx);
    ITJL_LOCK(ctx);
    NvtxRange nvtx("itjl_data_upload");
    ITJL_CUDA(ctx, cudaSetDevice(ctx->device));
    rc = upload_data(ctx, data, n_slices, n_t, slice_stride, time_stride);
    if (rc != ITJL_OK) return rc;
    ITJL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->res_slices = n_slices; ctx->res_nt = n_t;
    return ITJL_OK;
}

int itjl_acw_resident(itjl_ctx* ctx, double fs, uint32_t typemask, int32_t average_over_trials,
                      int64_t* n_lags_io, int32_t* k0, int32_t* k50, int32_t* ke, double* auc, double* tau,
                      double* acf_out, int64_t acf_capacity) {
    if (!ctx) return fail(nullptr, ITJL_ERR_ARG, "itjl_acw_resident: null context");
    ITJL_LOCK(ctx);
    ctx = primary(ctx);
    ITJL_LOCK(ctx);
    if (ctx->res_slices < 1) return fail(ctx, ITJL_ERR_ARG, "itjl_acw_resident: no resident data (call itjl_data_upload; other data calls on the context replace it)");
    NvtxRange nvtx("itjl_acw_resident");
    const int64_t ns = ctx->res_slices, nt = ctx->res_nt;
    const int rc = acw_core(ctx, nullptr, ns, nt, 1, ns, fs, typemask, average_over_trials, n_lags_io,
                            k0, k50, ke, auc, tau, acf_out, acf_capacity);
    return rc;
}

}  // extern "C"
