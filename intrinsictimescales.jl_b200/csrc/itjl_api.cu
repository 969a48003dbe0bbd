// C-ABI layer of libitjl_b200.so (see include/itjl_b200.h for the contract and the
// reference functions each entry point replaces).  Host logic only: argument checks,
// H2D/D2H staging on the context stream, launch geometry, status codes.
#include <math.h>
#include <stdlib.h>
#include <new>
#include <algorithm>
#include <vector>

#include "api_internal.h"
#include "philox.cuh"
#include "tc05.cuh"

namespace itjl {
char g_err[512] = {0};
}

using namespace itjl;

extern "C" {

const char* itjl_version(void) { return "itjl_b200 0.1.0 (sm_100a)"; }

int itjl_device_count(int* n_out) {
    if (!n_out) return fail(nullptr, ITJL_ERR_ARG, "itjl_device_count: null output");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { *n_out = 0; return fail(nullptr, ITJL_ERR_NO_DEVICE, "cudaGetDeviceCount: %s", cudaGetErrorString(e)); }
    *n_out = n;
    return ITJL_OK;
}

int itjl_tuning_default(itjl_tuning* t) {
    if (!t) return fail(nullptr, ITJL_ERR_ARG, "itjl_tuning_default: null output");
    memset(t, 0, sizeof(*t));
    t->abc_tc = -1; t->tc_tail_max = -1; t->tc_bias = 1; t->acw_stream = 1; t->pinned_upload = 1; t->pmc_fp32 = -1; t->psd_v2 = 1;
    t->tc_error_bound = 5e-6;
    return ITJL_OK;
}

int itjl_ctx_set_tuning(itjl_ctx* ctx, const itjl_tuning* t) {
    if (!ctx || !t) return fail(ctx, ITJL_ERR_ARG, "itjl_ctx_set_tuning: null argument");
    ITJL_LOCK(ctx);
    if (t->abc_tc < -1 || t->abc_tc > 1 || t->tc_terms < 0 || t->tc_terms > 3 || t->tc_tail_max > 60 || t->tc_groups < 0 ||
        t->tc_groups > 4 || t->tc_seg_trials < 0 || (t->abc_threads != 0 && t->abc_threads != 128 && t->abc_threads != 256) ||
        t->abc_bufs < 0 || t->abc_bufs > 2 || t->copy_threads < 0 || t->pmc_fp32 < -1 || t->pmc_fp32 > 2 || !(t->tc_error_bound > 0.0) ||
        (t->acw_stream != 0 && t->acw_stream != 1 && t->acw_stream != 8 && t->acw_stream != 16 && t->acw_stream != 32))
        return fail(ctx, ITJL_ERR_ARG, "itjl_ctx_set_tuning: value out of range");
    ctx->tuning = *t;
    for (itjl_ctx* sub : ctx->subs) sub->tuning = *t;
    return ITJL_OK;
}

int itjl_ctx_get_tuning(const itjl_ctx* ctx, itjl_tuning* out) {
    if (!ctx || !out) return fail(nullptr, ITJL_ERR_ARG, "itjl_ctx_get_tuning: null argument");
    *out = ctx->tuning;
    return ITJL_OK;
}

int itjl_ctx_create(int device, itjl_ctx** ctx_out) {
    if (!ctx_out) return fail(nullptr, ITJL_ERR_ARG, "itjl_ctx_create: null output");
    *ctx_out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(nullptr, ITJL_ERR_NO_DEVICE, "no CUDA device available: %s",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    if (device < 0 || device >= n) return fail(nullptr, ITJL_ERR_ARG, "itjl_ctx_create: device index out of range");
    itjl_ctx* ctx = new (std::nothrow) itjl_ctx();
    if (!ctx) return fail(nullptr, ITJL_ERR_ALLOC, "out of host memory");
    ctx->device = device;
    itjl_tuning_default(&ctx->tuning);
    if ((e = cudaSetDevice(device)) != cudaSuccess ||
        (e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaEventCreate(&ctx->ev0)) != cudaSuccess || (e = cudaEventCreate(&ctx->ev1)) != cudaSuccess ||
        (e = cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device)) != cudaSuccess ||
        (e = cudaDeviceGetAttribute(&ctx->max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device)) != cudaSuccess) {
        fail(nullptr, ITJL_ERR_CUDA, "itjl_ctx_create: %s", cudaGetErrorString(e));
        if (ctx->ev0) cudaEventDestroy(ctx->ev0);
        if (ctx->ev1) cudaEventDestroy(ctx->ev1);
        if (ctx->stream) cudaStreamDestroy(ctx->stream);
        delete ctx;
        return ITJL_ERR_CUDA;
    }
    *ctx_out = ctx;
    return ITJL_OK;
}

int itjl_ctx_destroy(itjl_ctx* ctx) {
    if (!ctx) return ITJL_OK;
    comm_destroy(ctx);                                   // NCCL communicator(s), if any (itjl_api_multi.cu)
    if (!ctx->subs.empty()) {                            // multi-GPU parent: no device resources of its own
        for (itjl_ctx* sub : ctx->subs) itjl_ctx_destroy(sub);
        delete ctx;
        return ITJL_OK;
    }
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    DevBuf* bufs[] = {&ctx->theta, &ctx->dist, &ctx->stats, &ctx->u0, &ctx->noise, &ctx->phases,
                      &ctx->flags, &ctx->scratch, &ctx->data, &ctx->out, &ctx->out2, &ctx->out3, &ctx->part, &ctx->gather, &ctx->raw, &ctx->acf_keep};
    for (DevBuf* b : bufs) b->release();
    cudaEventDestroy(ctx->ev0); cudaEventDestroy(ctx->ev1);
    for (int i = 0; i < 3; ++i) {
        if (ctx->pin[i]) cudaFreeHost(ctx->pin[i]);
        if (ctx->pin_ev[i]) cudaEventDestroy(ctx->pin_ev[i]);
    }
    if (ctx->pin_small) cudaFreeHost(ctx->pin_small);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
    return ITJL_OK;
}

const char* itjl_last_error(const itjl_ctx* ctx) { return ctx ? ctx->err : g_err; }

int itjl_ctx_synchronize(itjl_ctx* ctx) {
    if (!ctx) return fail(nullptr, ITJL_ERR_ARG, "null context");
    ITJL_LOCK(ctx);
    if (!ctx->subs.empty()) {
        for (itjl_ctx* sub : ctx->subs) { const int rc = itjl_ctx_synchronize(sub); if (rc != ITJL_OK) return rc; }
        return ITJL_OK;
    }
    ITJL_CUDA(ctx, cudaSetDevice(ctx->device));
    ITJL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ITJL_OK;
}

uint64_t itjl_ctx_stream(const itjl_ctx* ctx) { ctx = primary(const_cast<itjl_ctx*>(ctx)); return ctx ? (uint64_t)(uintptr_t)ctx->stream : 0; }
int64_t itjl_ctx_launch_count(const itjl_ctx* ctx) {
    if (!ctx) return 0;
    int64_t n = ctx->launches;
    for (const itjl_ctx* sub : ctx->subs) n += sub->launches;
    return n;
}
int itjl_ctx_sm_count(const itjl_ctx* ctx) { ctx = primary(const_cast<itjl_ctx*>(ctx)); return ctx ? ctx->sm_count : 0; }

int itjl_ctx_timing_reset(itjl_ctx* ctx) {
    if (!ctx) return fail(nullptr, ITJL_ERR_ARG, "null context");
    ITJL_LOCK(ctx);
    itjl_ctx* c = primary(ctx);
    c->timing = true; c->timed_launches = 0; c->timed_ms = 0.0;
    return ITJL_OK;
}
int itjl_ctx_timing_get(itjl_ctx* ctx, int64_t* n_launches, double* total_ms) {
    if (!ctx) return fail(nullptr, ITJL_ERR_ARG, "null context");
    ITJL_LOCK(ctx);
    itjl_ctx* c = primary(ctx);
    if (n_launches) *n_launches = c->timed_launches;
    if (total_ms) *total_ms = c->timed_ms;
    return ITJL_OK;
}
int itjl_ctx_timing_stop(itjl_ctx* ctx) {
    if (!ctx) return fail(nullptr, ITJL_ERR_ARG, "null context");
    ITJL_LOCK(ctx);
    primary(ctx)->timing = false;
    return ITJL_OK;
}

}  // extern "C"

// ---- helpers ---------------------------------------------------------------------------
int itjl::upload(itjl_ctx* ctx, DevBuf& buf, const void* host, size_t bytes) {
    ITJL_CUDA(ctx, buf.reserve(bytes ? bytes : 1));
    if (bytes) ITJL_CUDA(ctx, cudaMemcpyAsync(buf.p, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return ITJL_OK;
}

static int nextpow2_int(int64_t n) { int64_t p = 1; while (p < n) p <<= 1; return (int)p; }

extern "C" {

int itjl_model_create(itjl_ctx* ctx, const itjl_model_desc* desc, itjl_model** model_out) {
    if (!ctx || !desc || !model_out) return fail(ctx, ITJL_ERR_ARG, "itjl_model_create: null argument");
    *model_out = nullptr;
    const itjl_model_desc& d = *desc;
    if (d.kind != ITJL_KIND_OU && d.kind != ITJL_KIND_OU_OSC) return fail(ctx, ITJL_ERR_ARG, "model kind must be OU or OU_OSC");
    if (d.summary < ITJL_SUMMARY_ACF || d.summary > ITJL_SUMMARY_PGRAM) return fail(ctx, ITJL_ERR_ARG, "invalid summary method");
    if (d.distance != ITJL_DIST_LINEAR && d.distance != ITJL_DIST_LOG) return fail(ctx, ITJL_ERR_ARG, "Distance method must be :linear or :logarithmic");
    if (d.update != ITJL_UPDATE_EXACT && d.update != ITJL_UPDATE_EM) return fail(ctx, ITJL_ERR_ARG, "invalid update rule");
    if (d.n_trials < 1 || d.n_t < 4 || d.n_t > (1 << 24)) return fail(ctx, ITJL_ERR_ARG, "n_trials >= 1 and 4 <= n_t <= 2^24 required");
    if (d.n_stats < 1 || !d.target_stats) return fail(ctx, ITJL_ERR_ARG, "target_stats / n_stats missing");
    if (!(d.dt > 0.0)) return fail(ctx, ITJL_ERR_ARG, "dt must be positive");
    if (d.summary == ITJL_SUMMARY_ACF_MISSING && !d.missing_mask) return fail(ctx, ITJL_ERR_ARG, "missing_mask required for the missing-data ACF");
    if ((d.summary == ITJL_SUMMARY_PSD || d.summary == ITJL_SUMMARY_PGRAM) && !d.freq_bins) return fail(ctx, ITJL_ERR_ARG, "freq_bins required for the PSD summary");
    ITJL_LOCK(ctx);
    NvtxRange nvtx("itjl_model_create");
    if (!ctx->subs.empty()) return multi_model_create(ctx, desc, model_out);   // one model per device
    ITJL_CUDA(ctx, cudaSetDevice(ctx->device));

    itjl_model* m = new (std::nothrow) itjl_model();
    if (!m) return fail(ctx, ITJL_ERR_ALLOC, "out of host memory");
    m->ctx = ctx; m->d = d;
    m->d.target_stats = nullptr; m->d.missing_mask = nullptr; m->d.freq_bins = nullptr;
    int rc = ITJL_OK;
    std::vector<float> tgt((size_t)d.n_stats);
    for (int64_t i = 0; i < d.n_stats; ++i) tgt[i] = (float)d.target_stats[i];
    rc = upload(ctx, m->target, tgt.data(), tgt.size() * sizeof(float));

    if (rc == ITJL_OK && d.missing_mask) {
        int64_t masked = 0;
        for (int64_t i = 0; i < d.n_trials * d.n_t; ++i) masked += d.missing_mask[i] != 0;
        m->mask_frac = (double)masked / (double)(d.n_trials * d.n_t);
    }
    if (rc == ITJL_OK && d.summary != ITJL_SUMMARY_PSD && d.summary != ITJL_SUMMARY_PGRAM) {
        if (d.n_stats > d.n_t) { itjl_model_destroy(m); return fail(ctx, ITJL_ERR_ARG, "n_lags must not exceed n_t"); }
        int g = abc_acf_geometry((int)d.n_t, (int)d.n_stats, ctx->max_smem_optin, &m->geo, &ctx->tuning);
        if (g != 0) { itjl_model_destroy(m); return fail(ctx, ITJL_ERR_ARG, g == -2 ? "n_lags > 5120 is not supported by the fused ACF kernel" : "trial does not fit in shared memory (n_t too large)"); }
        // Tensor-core lag sums when the lag window fits the tensor-memory plans AND the error model
        // (abc_tc_predicted_error) puts the plan within tuning.tc_error_bound; otherwise - and with
        // tuning.abc_tc = 0 - the fp32 FFMA kernel.  abc_tc = 1 forces the tensor-core kernel.
        const int want_tc = ctx->tuning.abc_tc;
        const bool missing = d.summary == ITJL_SUMMARY_ACF_MISSING;
        const double mean_ratio = (missing && d.kind == ITJL_KIND_OU_OSC && d.data_sd != 0.0) ? d.data_mean / d.data_sd : 0.0;
        const int tcg = abc_tc_geometry((int)d.n_t, (int)d.n_stats, (int)d.n_trials, ctx->max_smem_optin, ctx->sm_count, &m->tc,
                                        &ctx->tuning, m->mask_frac, mean_ratio, missing);
        if (want_tc == 1 && tcg != 0) { itjl_model_destroy(m); return fail(ctx, ITJL_ERR_ARG, "tuning.abc_tc = 1 but the tensor-core kernel does not support this shape"); }
        m->use_tc = (tcg == 0) && (want_tc != 0);
        m->tc_pred_err = m->use_tc ? m->tc.pred_err : 0.0;
    }
    if (rc == ITJL_OK && d.missing_mask) {
        m->mask_words = (int)((d.n_t + 31) / 32);
        std::vector<uint32_t> bits((size_t)d.n_trials * m->mask_words, 0u);
        std::vector<int> nvalid((size_t)d.n_trials, 0);
        for (int64_t t = 0; t < d.n_t; ++t)
            for (int64_t tr = 0; tr < d.n_trials; ++tr) {
                if (d.missing_mask[tr + t * d.n_trials]) bits[(size_t)tr * m->mask_words + (t >> 5)] |= 1u << (t & 31);
                else nvalid[tr]++;
            }
        rc = upload(ctx, m->mask_bits, bits.data(), bits.size() * sizeof(uint32_t));
        if (rc == ITJL_OK) rc = upload(ctx, m->n_valid, nvalid.data(), nvalid.size() * sizeof(int));
        if (rc == ITJL_OK) rc = (cudaStreamSynchronize(ctx->stream) == cudaSuccess) ? ITJL_OK : ITJL_ERR_CUDA;
    }
    if (rc == ITJL_OK && d.summary == ITJL_SUMMARY_PSD) {
        m->n2 = 2 * nextpow2_int(d.n_t);
        m->log2_n2 = 0; while ((1 << m->log2_n2) < m->n2) m->log2_n2++;
        for (int64_t i = 0; i < d.n_stats; ++i)
            if (d.freq_bins[i] < 0 || d.freq_bins[i] >= m->n2 / 2 - 1) { itjl_model_destroy(m); return fail(ctx, ITJL_ERR_ARG, "freq_bins out of range"); }
        // k_abc_psd2 (warp-specialised, pruned first stage) when the shape is in its range and everything fits in
        // shared memory; tuning.psd_v2 = 0 keeps the first kernel
        if (ctx->tuning.psd_v2 != 0 && abc_psd2_supported((int)d.n_t, m->n2)) {
            const size_t sm2 = abc_psd2_smem_bytes((int)d.n_t, m->n2, (int)d.n_stats, &m->psd_chunk, &m->psd_xs_len);
            // (+ 2 KB: the kernel's static shared memory and the 1 KB the system reserves per CTA count against the same limit)
            if ((int64_t)sm2 + 2048 <= ctx->max_smem_optin && m->psd_chunk <= kQtab) { m->psd_v2 = true; m->psd_smem = sm2; }
        }
        if (!m->psd_v2) {
            m->psd_smem = abc_psd_smem_bytes((int)d.n_t, m->n2, (int)d.n_stats, &m->psd_chunk, &m->psd_xs_len);
            if ((int64_t)m->psd_smem + 2048 > ctx->max_smem_optin) {
                // many selected bins (e.g. fs = 250 Hz: 80 % of 16 384): keep the per-particle bin sums in global memory
                m->psd_smem = abc_psd_smem_bytes((int)d.n_t, m->n2, (int)d.n_stats, &m->psd_chunk, &m->psd_xs_len, false);
                m->psd_accb_global = true;
            }
            if ((int64_t)m->psd_smem + 2048 > ctx->max_smem_optin || m->psd_chunk > kQtab) { itjl_model_destroy(m); return fail(ctx, ITJL_ERR_ARG, "PSD FFT does not fit in shared memory (n_t too large)"); }
        }
        // window padded with zeros to the simulate buffer (k_abc_psd2 multiplies inside the simulator, 16 bytes at a time)
        std::vector<float> win((size_t)std::max<int64_t>(d.n_t + 1, m->psd_xs_len) + 4, 0.0f);
        std::vector<float2> tw((size_t)m->n2 / 2);
        double win_ss = 0.0;
        psd_make_tables((int)d.n_t, m->n2, win.data(), tw.data(), &win_ss);
        m->psd_scale = 1.0 / ((1.0 / d.dt) * win_ss);
        // bin table sorted by storage position: consecutive threads then read neighbouring shared-memory words
        // (in frequency order consecutive bins sit N/4 points apart - the same bank); binidx maps back
        std::vector<int4> binpos((size_t)d.n_stats);
        std::vector<int32_t> order((size_t)d.n_stats);
        for (int64_t i = 0; i < d.n_stats; ++i) order[(size_t)i] = (int32_t)i;
        const int N = m->n2 / 2;
        auto bin_addr = [&](int k, int* round) {
            int addr = 0; *round = 0;
            if (m->psd_v2) abc_psd2_bin(k, N, &addr, round); else addr = psd_fft_pos(k, N);
            return addr;
        };
        std::sort(order.begin(), order.end(), [&](int32_t a, int32_t b) {
            int ra, rb;
            const int pa = bin_addr(d.freq_bins[a] + 1, &ra), pb = bin_addr(d.freq_bins[b] + 1, &rb);
            return ra != rb ? ra < rb : pa < pb;
        });
        m->psd_round0_bins = 0;
        for (int64_t i = 0; i < d.n_stats; ++i) {
            const int k = d.freq_bins[order[(size_t)i]] + 1;
            int wx, wy, r0, r1;                             // the bin's twiddle exp(-2 pi i k / n2), as float bits
            memcpy(&wx, &tw[(size_t)k].x, 4); memcpy(&wy, &tw[(size_t)k].y, 4);
            const int pk = bin_addr(k, &r0), pn = bin_addr(N - k, &r1);
            binpos[(size_t)i] = make_int4(pk, pn, wx, wy);
            if (r0 == 0) m->psd_round0_bins = (int)i + 1;
        }
        rc = upload(ctx, m->binidx, order.data(), order.size() * sizeof(int32_t));
        if (rc == ITJL_OK)
            rc = upload(ctx, m->bins, binpos.data(), binpos.size() * sizeof(int4));
        if (rc == ITJL_OK) rc = upload(ctx, m->window, win.data(), win.size() * sizeof(float));
        if (rc == ITJL_OK) rc = upload(ctx, m->twiddle, tw.data(), tw.size() * sizeof(float2));
        // the host vectors above are block-local: finish the copies before they are destroyed
        if (rc == ITJL_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = fail(ctx, ITJL_ERR_CUDA, "model upload failed");
    }
    if (rc == ITJL_OK && d.summary == ITJL_SUMMARY_PGRAM) {
        // dsp.periodogram: nfft = nextfastfft(n_t); bins index power[2:end], i.e. rfft bin k = bin + 1 <= nfft / 2
        const int nfft = next_fast_len((int)d.n_t);
        if (d.n_t > 65536 || !fft_plan_make(nfft, &m->pg_plan)) { itjl_model_destroy(m); return fail(ctx, ITJL_ERR_ARG, "periodogram: n_t too large"); }
        for (int64_t i = 0; i < d.n_stats; ++i)
            if (d.freq_bins[i] < 0 || d.freq_bins[i] >= nfft / 2) { itjl_model_destroy(m); return fail(ctx, ITJL_ERR_ARG, "freq_bins out of range"); }
        m->pg_smem = abc_pgram_smem_bytes((int)d.n_t, nfft, (int)d.n_stats, &m->pg_chunk, &m->pg_bufb);
        if ((int64_t)m->pg_smem + 2048 > ctx->max_smem_optin || m->pg_chunk > kQtab) { itjl_model_destroy(m); return fail(ctx, ITJL_ERR_ARG, "periodogram FFT does not fit in shared memory (n_t too large)"); }
        std::vector<float> win((size_t)d.n_t);
        std::vector<float2> roots((size_t)roots_two_level_entries(nfft));
        std::vector<int32_t> bins((size_t)d.n_stats);
        for (int64_t i = 0; i < d.n_stats; ++i) bins[(size_t)i] = d.freq_bins[i] + 1;
        double win_ss = 0.0;
        pgram_make_tables((int)d.n_t, nfft, win.data(), nullptr, roots.data(), nullptr, &win_ss);
        m->psd_scale = 1.0 / ((1.0 / d.dt) * win_ss);
        rc = upload(ctx, m->binidx, bins.data(), bins.size() * sizeof(int32_t));
        if (rc == ITJL_OK) rc = upload(ctx, m->window, win.data(), win.size() * sizeof(float));
        if (rc == ITJL_OK) rc = upload(ctx, m->twiddle, roots.data(), roots.size() * sizeof(float2));
        if (rc == ITJL_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = fail(ctx, ITJL_ERR_CUDA, "model upload failed");
    }
    if (rc == ITJL_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = fail(ctx, ITJL_ERR_CUDA, "model upload failed");
    if (rc != ITJL_OK) { itjl_model_destroy(m); return rc; }
    *model_out = m;
    return ITJL_OK;
}

int itjl_model_destroy(itjl_model* m) {
    if (!m) return ITJL_OK;
    if (!m->subs.empty()) {
        for (itjl_model* sub : m->subs) itjl_model_destroy(sub);
        delete m;
        return ITJL_OK;
    }
    if (m->ctx) cudaSetDevice(m->ctx->device);
    m->target.release(); m->mask_bits.release(); m->n_valid.release();
    m->bins.release(); m->binidx.release(); m->window.release(); m->twiddle.release();
    delete m;
    return ITJL_OK;
}

int itjl_model_work(const itjl_model* m, double* flops_per_sample, int64_t* samples_per_particle) {
    if (!m) return fail(nullptr, ITJL_ERR_ARG, "null model");
    m = primary(const_cast<itjl_model*>(m));
    if (flops_per_sample) {
        if (m->d.summary == ITJL_SUMMARY_PSD) {
            // one n2-point real FFT per trial (2.5 n2 log2 n2) + |.|^2 on the selected bins + window etc.
            *flops_per_sample = (2.5 * m->n2 * m->log2_n2 + 3.0 * (double)m->d.n_stats) / (double)m->d.n_t + 12.0;
        } else if (m->d.summary == ITJL_SUMMARY_PGRAM) {
            // one nfft-point real FFT per trial (two trials per complex transform)
            const double nfft = (double)m->pg_plan.n;
            *flops_per_sample = (2.5 * nfft * log2(nfft) + 3.0 * (double)m->d.n_stats) / (double)m->d.n_t + 12.0;
        } else {
            *flops_per_sample = 2.0 * (double)m->d.n_stats + 8.0;
        }
    }
    if (samples_per_particle) *samples_per_particle = m->d.n_trials * m->d.n_t;
    return ITJL_OK;
}

int itjl_tc_plan(int64_t n_t, int64_t n_lags, int64_t n_trials, int32_t max_smem_optin, int32_t sm_count, int32_t* out) {
    if (!out || n_t > INT32_MAX || n_lags > INT32_MAX || n_trials > INT32_MAX) return fail(nullptr, ITJL_ERR_ARG, "itjl_tc_plan: bad argument");
    AbcTcGeometry g{};
    const int rc = abc_tc_geometry((int)n_t, (int)n_lags, (int)n_trials, max_smem_optin, sm_count, &g);
    if (rc != 0 && rc != -4) return fail(nullptr, ITJL_ERR_ARG, "itjl_tc_plan: shape outside the tensor-core kernel's plans");
    const int32_t v[16] = {g.groups, g.chunk, g.ksteps, g.rows_total, g.used_cols, g.mixed, g.seg_trials, g.n_terms,
                           g.tc_lags, g.tail_start, g.tail_lt, g.n_ops, (int32_t)g.smem_bytes, g.grid, g.n_windows,
                           rc == -4 ? 1 : 0 /* 1: the error model sends this shape to the CUDA-core kernel */};
    memcpy(out, v, sizeof(v));
    return ITJL_OK;
}

double itjl_tc_predicted_error(int64_t n_t, int64_t n_trials, int32_t n_terms, int32_t seg_rows, double mask_frac, double mean_ratio) {
    return abc_tc_predicted_error(n_t, n_trials, n_terms, seg_rows, mask_frac, mean_ratio);
}

int itjl_model_kernel_info(const itjl_model* m, char* name_out, double* tensor_flops_per_sample) {
    if (!m || !name_out || !tensor_flops_per_sample) return fail(m ? m->ctx : nullptr, ITJL_ERR_ARG, "itjl_model_kernel_info: null argument");
    m = primary(const_cast<itjl_model*>(m));
    *tensor_flops_per_sample = 0.0;
    if (m->d.summary == ITJL_SUMMARY_PSD) { snprintf(name_out, 32, m->psd_v2 ? "k_abc_psd2" : "k_abc_psd"); return ITJL_OK; }
    if (m->d.summary == ITJL_SUMMARY_PGRAM) { snprintf(name_out, 32, "k_abc_pgram"); return ITJL_OK; }
    if (!m->use_tc) { snprintf(name_out, 32, "k_abc_acf"); return ITJL_OK; }
    snprintf(name_out, 32, "k_abc_acf_tc");
    double f = 0.0;
    for (int o = 0; o < m->tc.n_ops; ++o) {
        const uint32_t id = m->tc.op_idesc[o];
        const double M = (double)(((id >> 24) & 0x1Fu) << 4), N = (double)(((id >> 17) & 0x3Fu) << 3);
        f += 2.0 * M * N * 16.0;
    }
    for (int w = 1; w < m->tc.n_windows; ++w) f += 2.0 * 128.0 * (w == m->tc.n_windows - 1 ? (double)m->tc.win1_used_cols : 512.0) * 16.0;
    *tensor_flops_per_sample = f * (double)m->tc.n_terms * m->tc.ksteps / (double)m->d.n_t;
    return ITJL_OK;
}

}  // extern "C"

// Launch the fused kernel for one batch; every pointer is a device pointer (or null).
int itjl::launch_distances(itjl_model* m, const double* theta_dev, int64_t n_particles, int32_t n_theta,
                            uint64_t seed, uint64_t generation, int64_t particle_offset,
                            const double* u0_dev, const double* noise_dev, const double* phases_dev,
                            double* dist_dev, double* stats_dev, int64_t theta_ld) {
    itjl_ctx* ctx = m->ctx;
    const itjl_model_desc& d = m->d;
    const int need_theta = (d.kind == ITJL_KIND_OU_OSC) ? 3 : 1;
    if (n_theta < need_theta) return fail(ctx, ITJL_ERR_ARG, "theta has too few columns for this model");
    if (n_particles < 1 || n_particles > 0x7fffffffll) return fail(ctx, ITJL_ERR_ARG, "n_particles out of range");
    cudaError_t e;
    NvtxRange nvtx(d.summary == ITJL_SUMMARY_PSD ? "k_abc_psd" : d.summary == ITJL_SUMMARY_PGRAM ? "k_abc_pgram" : (m->use_tc ? "k_abc_acf_tc" : "k_abc_acf"));
    int n_launched = 1;
    if (ctx->timing) cudaEventRecord(ctx->ev0, ctx->stream);
    if (d.summary == ITJL_SUMMARY_PSD) {
        AbcPsdParams P{};
        P.theta = theta_dev; P.n_particles = n_particles; P.theta_ld = theta_ld > 0 ? theta_ld : n_particles; P.n_theta = n_theta;
        P.n_t = (int)d.n_t; P.n_trials = (int)d.n_trials; P.n_stats = (int)d.n_stats;
        P.chunk = m->psd_chunk; P.xs_len = m->psd_xs_len; P.n2 = m->n2; P.log2_n2 = m->log2_n2;
        P.dt = d.dt; P.data_mean = d.data_mean; P.data_sd = d.data_sd;
        P.update = d.update; P.kind = d.kind; P.distance = d.distance;
        P.keys = philox_keys(seed);
        P.c3_noise = philox_c3(kStreamNoise, generation); P.c3_init = philox_c3(kStreamInit, generation);
        P.particle_offset = particle_offset;
        P.target = (const float*)m->target.p; P.binpos = (const int4*)m->bins.p; P.binidx = (const int32_t*)m->binidx.p;
        P.window = (const float*)m->window.p; P.twiddle = (const float2*)m->twiddle.p;
        P.psd_scale = m->psd_scale;
        P.u0 = u0_dev; P.noise = noise_dev; P.phases = phases_dev;
        P.dist_out = dist_dev; P.stats_out = stats_dev;
        P.round0_bins = m->psd_round0_bins;
        if (m->psd_v2) {
#ifdef ITJL_TUNING
            { const char* dbg = getenv("ITJL_PSD_DEBUG"); P.grid = dbg ? atoi(dbg) : 0; }   // tuning builds only
#endif
            e = abc_psd2_launch(P, d.kind == ITJL_KIND_OU_OSC, m->psd_smem, ctx->stream);
        } else if (!m->psd_accb_global) {
            e = abc_psd_launch(P, d.kind == ITJL_KIND_OU_OSC, m->psd_smem, ctx->stream);
        } else {
            // sub-batches of at most 2048 particles, each with its own row of bin sums in a scratch buffer
            const int64_t sub = 2048, pad = ((int64_t)d.n_stats + 3) & ~(int64_t)3;
            const cudaError_t er = ctx->scratch.reserve((size_t)std::min<int64_t>(sub, n_particles) * pad * sizeof(float));
            if (er != cudaSuccess) return fail(ctx, ITJL_ERR_CUDA, "PSD bin-sum scratch: %s", cudaGetErrorString(er));
            e = cudaSuccess;
            n_launched = 0;
            for (int64_t lo = 0; lo < n_particles && e == cudaSuccess; lo += sub, ++n_launched) {
                AbcPsdParams Q = P;
                Q.grid = std::min<int64_t>(sub, n_particles - lo);
                Q.accb_global = (float*)ctx->scratch.p;
                Q.theta = theta_dev + lo;                         // column-major: the leading dimension stays
                Q.particle_offset = particle_offset + lo;
                Q.dist_out = dist_dev + lo;
                Q.stats_out = stats_dev ? stats_dev + lo * d.n_stats : nullptr;
                Q.u0 = u0_dev ? u0_dev + lo * d.n_trials : nullptr;
                Q.noise = noise_dev ? noise_dev + lo * d.n_trials * d.n_t : nullptr;
                Q.phases = phases_dev ? phases_dev + lo * d.n_trials : nullptr;
                e = abc_psd_launch(Q, d.kind == ITJL_KIND_OU_OSC, m->psd_smem, ctx->stream);
            }
        }
    } else if (d.summary == ITJL_SUMMARY_PGRAM) {
        AbcPgramParams P{};
        P.theta = theta_dev; P.n_particles = n_particles; P.theta_ld = theta_ld > 0 ? theta_ld : n_particles; P.n_theta = n_theta;
        P.n_t = (int)d.n_t; P.n_trials = (int)d.n_trials; P.n_stats = (int)d.n_stats;
        P.chunk = m->pg_chunk; P.bufb_len = m->pg_bufb; P.plan = m->pg_plan;
        P.dt = d.dt; P.data_mean = d.data_mean; P.data_sd = d.data_sd;
        P.update = d.update; P.kind = d.kind; P.distance = d.distance;
        P.keys = philox_keys(seed);
        P.c3_noise = philox_c3(kStreamNoise, generation); P.c3_init = philox_c3(kStreamInit, generation);
        P.particle_offset = particle_offset;
        P.target = (const float*)m->target.p; P.bins = (const int32_t*)m->binidx.p;
        P.window = (const float*)m->window.p; P.roots = (const float2*)m->twiddle.p;
        P.psd_scale = m->psd_scale;
        P.u0 = u0_dev; P.noise = noise_dev; P.phases = phases_dev;
        P.dist_out = dist_dev; P.stats_out = stats_dev;
        e = abc_pgram_launch(P, d.kind == ITJL_KIND_OU_OSC, m->pg_smem, ctx->stream);
    } else if (m->use_tc) {
        AbcTcParams P{};
        const AbcTcGeometry& t = m->tc;
        P.theta = theta_dev; P.n_particles = n_particles; P.theta_ld = theta_ld > 0 ? theta_ld : n_particles; P.n_theta = n_theta;
        P.n_t = (int)d.n_t; P.n_trials = (int)d.n_trials; P.n_lags = (int)d.n_stats;
        P.dt = d.dt; P.update = d.update; P.kind = d.kind; P.distance = d.distance;
        P.keys = philox_keys(seed);
        P.c3_noise = philox_c3(kStreamNoise, generation); P.c3_init = philox_c3(kStreamInit, generation);
        P.particle_offset = particle_offset;
        P.target = (const float*)m->target.p;
        P.mask_bits = (const uint32_t*)m->mask_bits.p; P.n_valid = (const int*)m->n_valid.p;
        P.mask_words = m->mask_words;
        P.u0 = u0_dev; P.noise = noise_dev; P.phases = phases_dev;
        P.dist_out = dist_dev; P.stats_out = stats_dev;
        P.groups = t.groups; P.chunk = t.chunk; P.epi_offset = t.epi_offset;
        P.ksteps = t.ksteps; P.sbo_bytes = t.sbo_bytes;
        P.used_cols = t.used_cols; P.seg_trials = t.seg_trials; P.n_terms = t.n_terms;
        P.tc_lags = t.tc_lags; P.tail_start = t.tail_start; P.tail_lt = t.tail_lt; P.tail_steps = t.tail_steps;
        P.tail_steps_per_stream = t.tail_steps_per_stream;
        P.n_ops = t.n_ops;
        for (int o = 0; o < t.n_ops; ++o) {
            P.op_a_off[o] = (uint32_t)(t.op_a_blk[o] * (t.sbo_bytes / 16));
            P.op_b_off[o] = (uint32_t)(t.op_b_row[o] + t.op_b_blk[o] * (t.sbo_bytes / 16));
            P.op_idesc[o] = t.op_idesc[o]; P.op_taddr[o] = t.op_taddr[o];
        }
        P.tc_scale = t.tc_scale; P.tc_unscale = 1.0f / (t.tc_scale * t.tc_scale);
        P.tc_fix = t.tc_fix; P.tc_unfix = 1.0f / t.tc_fix;
        P.bias_a = t.bias_a; P.bias_b = t.bias_b;
        P.smem_bytes = t.smem_bytes;
        P.miss_mean_ratio = (d.kind == ITJL_KIND_OU_OSC && d.data_sd != 0.0) ? (float)(d.data_mean / d.data_sd) : 0.f;
#ifdef ITJL_TUNING
        { const char* dbg = getenv("ITJL_ABC_DEBUG"); P.debug = dbg ? atoi(dbg) : 0; }   // phase-isolation builds only (make EXTRA=-DITJL_TUNING)
#endif
        P.k_lo = 0; P.lag_base = 0; P.accumulate = 0;
        e = abc_tc_launch(P, d.summary == ITJL_SUMMARY_ACF_MISSING, d.kind == ITJL_KIND_OU_OSC, t.mixed != 0,
                          t.smem_bytes, t.grid, ctx->stream);
        for (int w = 1; w < t.n_windows && e == cudaSuccess; ++w) {
            // further lag windows: same trajectories (same Philox counters), B operand 3 w rows further
            // down, plan-A tiles; window w scores the lags from where window w - 1 stopped (449, then
            // 384 w + 1) and adds its share of the distance
            const int used = w == t.n_windows - 1 ? t.win1_used_cols : 512;
            P.lag_base = 128 * t.win1_s0 * w;
            P.k_lo = (w == 1 ? t.tc_lags : P.lag_base + 1) - P.lag_base;
            P.tc_lags = std::min((int)d.n_stats - P.lag_base, 385);
            P.used_cols = used;
            P.accumulate = 1;
            P.tail_lt = 0;
            const int n_tiles = (used + 127) / 128;
            P.n_ops = n_tiles;
            for (int s2 = 0; s2 < n_tiles; ++s2) {
                const int n = s2 == n_tiles - 1 ? used - 128 * s2 : 128;
                P.op_a_off[s2] = 0u;
                P.op_b_off[s2] = (uint32_t)(t.win1_s0 * w + s2);
                P.op_idesc[s2] = tc::make_idesc_f16_mn(128, n);
                P.op_taddr[s2] = (uint32_t)(128 * s2);
            }
            e = abc_tc_launch(P, d.summary == ITJL_SUMMARY_ACF_MISSING, d.kind == ITJL_KIND_OU_OSC, false,
                              t.smem_bytes, t.grid, ctx->stream);
            ++n_launched;
        }
    } else {
        AbcAcfParams P{};
        P.theta = theta_dev; P.n_particles = n_particles; P.theta_ld = theta_ld > 0 ? theta_ld : n_particles; P.n_theta = n_theta;
        P.n_t = (int)d.n_t; P.n_trials = (int)d.n_trials; P.n_lags = (int)d.n_stats;
        P.chunk = m->geo.chunk; P.xs_stride = m->geo.xs_stride; P.n_bufs = m->geo.n_bufs; P.lt = m->geo.lt;
        P.n_streams = m->geo.n_streams; P.tiles_per_stream = m->geo.tiles_per_stream; P.n_tiles = m->geo.n_tiles;
        P.dt = d.dt; P.data_mean = d.data_mean; P.data_sd = d.data_sd;
        P.update = d.update; P.kind = d.kind; P.distance = d.distance;
        P.keys = philox_keys(seed);
        P.c3_noise = philox_c3(kStreamNoise, generation); P.c3_init = philox_c3(kStreamInit, generation);
        P.particle_offset = particle_offset;
        P.target = (const float*)m->target.p;
        P.mask_bits = (const uint32_t*)m->mask_bits.p; P.n_valid = (const int*)m->n_valid.p;
        P.mask_words = m->mask_words;
        P.u0 = u0_dev; P.noise = noise_dev; P.phases = phases_dev;
        P.dist_out = dist_dev; P.stats_out = stats_dev;
#ifdef ITJL_TUNING
        { const char* dbg = getenv("ITJL_ABC_DEBUG"); P.debug = dbg ? atoi(dbg) : 0; }
#endif
        e = abc_acf_launch(P, m->geo.threads, d.summary == ITJL_SUMMARY_ACF_MISSING, d.kind == ITJL_KIND_OU_OSC,
                           m->geo.smem_bytes, ctx->stream);
    }
    if (e != cudaSuccess) return fail(ctx, ITJL_ERR_CUDA, "launch fused abc kernel: %s", cudaGetErrorString(e));
    ctx->launches += n_launched;                      // every lag window / PSD sub-batch is a launch
    if (ctx->timing) {
        cudaEventRecord(ctx->ev1, ctx->stream);
        cudaEventSynchronize(ctx->ev1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
        ctx->timed_ms += ms; ctx->timed_launches++;
    }
    return ITJL_OK;
}

static int stage_inputs(itjl_model* m, const double* theta, int64_t n_particles, int32_t n_theta,
                        const double* u0, const double* noise, const double* phases) {
    itjl_ctx* ctx = m->ctx;
    const itjl_model_desc& d = m->d;
    int rc = upload(ctx, ctx->theta, theta, (size_t)n_particles * n_theta * sizeof(double));
    const size_t pt = (size_t)n_particles * d.n_trials;
    if (rc == ITJL_OK && u0) rc = upload(ctx, ctx->u0, u0, pt * sizeof(double));
    if (rc == ITJL_OK && noise) rc = upload(ctx, ctx->noise, noise, pt * d.n_t * sizeof(double));
    if (rc == ITJL_OK && phases) rc = upload(ctx, ctx->phases, phases, pt * sizeof(double));
    return rc;
}

extern "C" {

int itjl_abc_distances(itjl_model* m, const double* theta, int64_t n_particles, int32_t n_theta,
                       uint64_t seed, uint64_t generation, int64_t particle_offset,
                       const double* u0, const double* noise, const double* phases,
                       double* dist_out, double* stats_out) {
    if (!m) return fail(nullptr, ITJL_ERR_ARG, "null model");
    if (!theta || !dist_out) return fail(m->ctx, ITJL_ERR_ARG, "itjl_abc_distances: null theta / dist_out");
    if (n_particles < 1 || n_theta < 1) return fail(m->ctx, ITJL_ERR_ARG, "itjl_abc_distances: empty batch");
    ITJL_LOCK(m->ctx);
    NvtxRange nvtx("itjl_abc_distances");
    if (is_sharded(m->ctx) && !u0 && !noise && !phases && !stats_out) {
        // particles sharded over the communicator, distances all-gathered device to device
        double* dist_dev = nullptr;
        int rcs = sharded_distances(m, theta, n_particles, n_theta, seed, generation, particle_offset, &dist_dev);
        if (rcs != ITJL_OK) return rcs;
        itjl_ctx* c0 = primary(m->ctx);
        ITJL_CUDA(m->ctx, cudaMemcpyAsync(dist_out, dist_dev, (size_t)n_particles * sizeof(double), cudaMemcpyDeviceToHost, c0->stream));
        ITJL_CUDA(m->ctx, cudaStreamSynchronize(c0->stream));
        return ITJL_OK;
    }
    m = primary(m);                                    // injected inputs / statistics: first device only
    itjl_ctx* ctx = m->ctx;
    ITJL_LOCK(ctx);
    ITJL_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc = stage_inputs(m, theta, n_particles, n_theta, u0, noise, phases);
    if (rc != ITJL_OK) return rc;
    ITJL_CUDA(ctx, ctx->dist.reserve((size_t)n_particles * sizeof(double)));
    if (stats_out) ITJL_CUDA(ctx, ctx->stats.reserve((size_t)n_particles * m->d.n_stats * sizeof(double)));
    rc = launch_distances(m, (const double*)ctx->theta.p, n_particles, n_theta, seed, generation, particle_offset,
                          u0 ? (const double*)ctx->u0.p : nullptr, noise ? (const double*)ctx->noise.p : nullptr,
                          phases ? (const double*)ctx->phases.p : nullptr, (double*)ctx->dist.p,
                          stats_out ? (double*)ctx->stats.p : nullptr);
    if (rc != ITJL_OK) return rc;
    ITJL_CUDA(ctx, cudaMemcpyAsync(dist_out, ctx->dist.p, (size_t)n_particles * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (stats_out)
        ITJL_CUDA(ctx, cudaMemcpyAsync(stats_out, ctx->stats.p, (size_t)n_particles * m->d.n_stats * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    ITJL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ITJL_OK;
}

int itjl_abc_distances_dev(itjl_model* m, const double* theta_dev, int64_t n_particles, int32_t n_theta,
                           uint64_t seed, uint64_t generation, int64_t particle_offset, double* dist_dev) {
    if (!m) return fail(nullptr, ITJL_ERR_ARG, "null model");
    if (!theta_dev || !dist_dev) return fail(m->ctx, ITJL_ERR_ARG, "itjl_abc_distances_dev: null device pointer");
    ITJL_LOCK(m->ctx);
    m = primary(m);
    ITJL_LOCK(m->ctx);
    ITJL_CUDA(m->ctx, cudaSetDevice(m->ctx->device));
    return launch_distances(m, theta_dev, n_particles, n_theta, seed, generation, particle_offset,
                            nullptr, nullptr, nullptr, dist_dev, nullptr);
}

int itjl_abc_generation(itjl_model* m, const double* theta, int64_t n_particles, int32_t n_theta,
                        uint64_t seed, uint64_t generation, int64_t particle_offset,
                        double epsilon, int64_t min_accepted, double* dist_out, uint8_t* isaccepted,
                        int64_t* n_total, int64_t* n_accepted) {
    if (!m) return fail(nullptr, ITJL_ERR_ARG, "null model");
    if (!theta || !dist_out || !isaccepted || !n_total || !n_accepted) return fail(m->ctx, ITJL_ERR_ARG, "itjl_abc_generation: null argument");
    if (n_particles < 1 || n_theta < 1) return fail(m->ctx, ITJL_ERR_ARG, "itjl_abc_generation: empty batch");
    ITJL_LOCK(m->ctx);
    NvtxRange nvtx("itjl_abc_generation");
    const bool sharded = is_sharded(m->ctx);
    itjl_ctx* ctx = primary(m->ctx);
    ITJL_LOCK(ctx);
    int rc = ITJL_OK;
    double* dist_dev = nullptr;
    if (sharded) {
        rc = sharded_distances(m, theta, n_particles, n_theta, seed, generation, particle_offset, &dist_dev);
        if (rc != ITJL_OK) return rc;
        ITJL_CUDA(ctx, cudaSetDevice(ctx->device));
    } else {
        ITJL_CUDA(ctx, cudaSetDevice(ctx->device));
        rc = stage_inputs(m, theta, n_particles, n_theta, nullptr, nullptr, nullptr);
        if (rc != ITJL_OK) return rc;
        ITJL_CUDA(ctx, ctx->dist.reserve((size_t)n_particles * sizeof(double)));
        rc = launch_distances(m, (const double*)ctx->theta.p, n_particles, n_theta, seed, generation, particle_offset,
                              nullptr, nullptr, nullptr, (double*)ctx->dist.p, nullptr);
        if (rc != ITJL_OK) return rc;
        dist_dev = (double*)ctx->dist.p;
    }
    ITJL_CUDA(ctx, ctx->flags.reserve((size_t)n_particles + 2 * sizeof(int64_t) + 16));
    ITJL_CUDA(ctx, ctx->scratch.reserve(2 * sizeof(int64_t)));
    cudaError_t e = abc_accept_launch(dist_dev, n_particles, epsilon, min_accepted,
                                      (uint8_t*)ctx->flags.p, (int64_t*)ctx->scratch.p, ctx->stream);
    if (e != cudaSuccess) return fail(ctx, ITJL_ERR_CUDA, "launch k_abc_accept: %s", cudaGetErrorString(e));
    ctx->launches++;
    int64_t out2[2] = {0, 0};
    ITJL_CUDA(ctx, cudaMemcpyAsync(dist_out, dist_dev, (size_t)n_particles * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    ITJL_CUDA(ctx, cudaMemcpyAsync(isaccepted, ctx->flags.p, (size_t)n_particles, cudaMemcpyDeviceToHost, ctx->stream));
    ITJL_CUDA(ctx, cudaMemcpyAsync(out2, ctx->scratch.p, sizeof(out2), cudaMemcpyDeviceToHost, ctx->stream));
    ITJL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *n_total = out2[0]; *n_accepted = out2[1];
    return ITJL_OK;
}

int itjl_abc_accept(itjl_ctx* ctx, const double* dist, int64_t n, double epsilon, int64_t min_accepted,
                    uint8_t* isaccepted, int64_t* n_total, int64_t* n_accepted) {
    if (!ctx) return fail(nullptr, ITJL_ERR_ARG, "null context");
    if (!dist || !isaccepted || !n_total || !n_accepted || n < 1) return fail(ctx, ITJL_ERR_ARG, "itjl_abc_accept: null / empty argument");
    ITJL_LOCK(ctx);
    ctx = primary(ctx);
    ITJL_LOCK(ctx);
    ITJL_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc = upload(ctx, ctx->dist, dist, (size_t)n * sizeof(double));
    if (rc != ITJL_OK) return rc;
    ITJL_CUDA(ctx, ctx->flags.reserve((size_t)n));
    ITJL_CUDA(ctx, ctx->scratch.reserve(2 * sizeof(int64_t)));
    cudaError_t e = abc_accept_launch((const double*)ctx->dist.p, n, epsilon, min_accepted,
                                      (uint8_t*)ctx->flags.p, (int64_t*)ctx->scratch.p, ctx->stream);
    if (e != cudaSuccess) return fail(ctx, ITJL_ERR_CUDA, "launch k_abc_accept: %s", cudaGetErrorString(e));
    ctx->launches++;
    int64_t out2[2] = {0, 0};
    ITJL_CUDA(ctx, cudaMemcpyAsync(isaccepted, ctx->flags.p, (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    ITJL_CUDA(ctx, cudaMemcpyAsync(out2, ctx->scratch.p, sizeof(out2), cudaMemcpyDeviceToHost, ctx->stream));
    ITJL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *n_total = out2[0]; *n_accepted = out2[1];
    return ITJL_OK;
}

// ---- OU generator ----------------------------------------------------------------------
static int generate_common(itjl_ctx* ctx, int kind, double tau, double freq, double coeff, double dt,
                           int64_t n_t, int64_t n_trials, int mode, double scale, double shift, int32_t update,
                           uint64_t seed, const double* u0, const double* noise, const double* phases, double* out,
                           uint32_t particle = 0) {
    if (!ctx) return fail(nullptr, ITJL_ERR_ARG, "null context");
    ITJL_LOCK(ctx);
    ctx = primary(ctx);
    ITJL_LOCK(ctx);
    NvtxRange nvtx("itjl_generate_ou");
    if (!out) return fail(ctx, ITJL_ERR_ARG, "generate_ou: null output");
    if (n_t < 4 || n_trials < 1 || !(dt > 0.0)) return fail(ctx, ITJL_ERR_ARG, "generate_ou: need n_t >= 4, n_trials >= 1, dt > 0");
    if (update != ITJL_UPDATE_EXACT && update != ITJL_UPDATE_EM) return fail(ctx, ITJL_ERR_ARG, "invalid update rule");
    ITJL_CUDA(ctx, cudaSetDevice(ctx->device));
    OuGenParams P{};
    P.n_t = (int)n_t; P.n_trials = (int)n_trials;
    int c = (int)((n_t + kThreads - 1) / kThreads);
    P.chunk = (c + 3) & ~3;
    const size_t smem = (size_t)P.chunk * kThreads * sizeof(float) + 2048;
    if ((int64_t)smem > ctx->max_smem_optin || P.chunk > kQtab) return fail(ctx, ITJL_ERR_ARG, "generate_ou: n_t too large for one CTA");
    P.tau = tau; P.freq = freq; P.coeff = coeff; P.dt = dt;
    P.update = update; P.kind = kind; P.mode = mode; P.scale = (float)scale; P.shift = (float)shift;
    P.keys = philox_keys(seed);
    P.c3_noise = philox_c3(kStreamNoise, 0); P.c3_init = philox_c3(kStreamInit, 0); P.particle = particle;
    int rc = ITJL_OK;
    if (u0) { rc = upload(ctx, ctx->u0, u0, (size_t)n_trials * sizeof(double)); P.u0 = (const double*)ctx->u0.p; }
    if (rc == ITJL_OK && noise) { rc = upload(ctx, ctx->noise, noise, (size_t)n_trials * n_t * sizeof(double)); P.noise = (const double*)ctx->noise.p; }
    if (rc == ITJL_OK && phases) { rc = upload(ctx, ctx->phases, phases, (size_t)n_trials * sizeof(double)); P.phases = (const double*)ctx->phases.p; }
    if (rc != ITJL_OK) return rc;
    const size_t bytes = (size_t)n_trials * n_t * sizeof(double);
    ITJL_CUDA(ctx, ctx->out.reserve(bytes));
    P.out = (double*)ctx->out.p;
    cudaError_t e = ou_generate_launch(P, ctx->stream);
    if (e != cudaSuccess) return fail(ctx, ITJL_ERR_CUDA, "launch k_ou_generate: %s", cudaGetErrorString(e));
    ctx->launches++;
    ITJL_CUDA(ctx, cudaMemcpyAsync(out, ctx->out.p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    ITJL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    // solver-failure contract (ou_process.jl:56-61): any non-finite value -> all-NaN matrix
    bool bad = false;
    for (size_t i = 0; i < (size_t)n_trials * n_t; ++i) if (!isfinite(out[i])) { bad = true; break; }
    if (bad) for (size_t i = 0; i < (size_t)n_trials * n_t; ++i) out[i] = NAN;
    return ITJL_OK;
}

int itjl_generate_ou(itjl_ctx* ctx, double tau, double true_D, double dt, int64_t n_t, int64_t n_trials,
                     int32_t standardize, int32_t update, uint64_t seed, const double* u0, const double* noise,
                     double* out) {
    return generate_common(ctx, ITJL_KIND_OU, tau, 0.0, 1.0, dt, n_t, n_trials, standardize ? 1 : 2,
                           true_D, 0.0, update, seed, u0, noise, nullptr, out);
}

int itjl_generate_ou_osc(itjl_ctx* ctx, double tau, double freq, double coeff, double dt, int64_t n_t,
                         int64_t n_trials, double data_mean, double data_sd, int32_t update, uint64_t seed,
                         const double* u0, const double* noise, const double* phases, double* out) {
    return generate_common(ctx, ITJL_KIND_OU_OSC, tau, freq, coeff, dt, n_t, n_trials, 1, data_sd, data_mean,
                           update, seed, u0, noise, phases, out);
}

int itjl_generate_ou_two(itjl_ctx* ctx, double tau1, double tau2, double coeff, double dt, int64_t n_t, int64_t n_trials,
                         double data_var, int32_t update, uint64_t seed, double* out) {
    // Models.generate_data(::TwoTimescaleModel, theta) (src/models/two_timescale.jl:30-51): two standardised OU
    // processes with D_i = 1 / tau_i, mixed sqrt(coeff) : sqrt(1 - coeff), scaled by sqrt(data_var).  The two
    // processes use the Philox streams of particles 0 and 1; the mix itself is one pass over the result.
    if (!ctx || !out) return fail(ctx, ITJL_ERR_ARG, "itjl_generate_ou_two: null argument");
    if (!(coeff >= 0.0 && coeff <= 1.0) || !(data_var >= 0.0) || n_t < 4 || n_trials < 1)
        return fail(ctx, ITJL_ERR_ARG, "itjl_generate_ou_two: need 0 <= coeff <= 1, data_var >= 0, n_t >= 4, n_trials >= 1");
    ITJL_LOCK(ctx);
    std::vector<double> second((size_t)n_trials * n_t);
    int rc = generate_common(ctx, ITJL_KIND_OU, tau1, 0.0, 1.0, dt, n_t, n_trials, 1, 1.0 / tau1, 0.0, update, seed, nullptr, nullptr, nullptr, out, 0);
    if (rc == ITJL_OK)
        rc = generate_common(ctx, ITJL_KIND_OU, tau2, 0.0, 1.0, dt, n_t, n_trials, 1, 1.0 / tau2, 0.0, update, seed, nullptr, nullptr, nullptr, second.data(), 1);
    if (rc != ITJL_OK) return rc;
    const double a = sqrt(coeff) * sqrt(data_var), b = sqrt(1.0 - coeff) * sqrt(data_var);
    for (size_t i = 0; i < second.size(); ++i) out[i] = a * out[i] + b * second[i];
    return ITJL_OK;
}

// ---- measurement -----------------------------------------------------------------------
int itjl_fp32_peak(itjl_ctx* ctx, int32_t iters, double* tflops_out) {
    if (!ctx || !tflops_out) return fail(ctx, ITJL_ERR_ARG, "itjl_fp32_peak: null argument");
    ITJL_LOCK(ctx);
    ctx = primary(ctx);
    ITJL_LOCK(ctx);
    if (iters < 1) iters = 1;
    ITJL_CUDA(ctx, cudaSetDevice(ctx->device));
    const int blocks = ctx->sm_count * 8;
    const int inner = 4096;
    ITJL_CUDA(ctx, ctx->scratch.reserve((size_t)blocks * kThreads * sizeof(float)));
    cudaError_t e = fp32_peak_launch((float*)ctx->scratch.p, blocks, inner, ctx->stream);   // warm-up
    if (e != cudaSuccess) return fail(ctx, ITJL_ERR_CUDA, "launch k_fp32_peak: %s", cudaGetErrorString(e));
    ctx->launches++;
    ITJL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    double best = 0.0;
    for (int i = 0; i < iters; ++i) {
        ITJL_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
        e = fp32_peak_launch((float*)ctx->scratch.p, blocks, inner, ctx->stream);
        if (e != cudaSuccess) return fail(ctx, ITJL_ERR_CUDA, "launch k_fp32_peak: %s", cudaGetErrorString(e));
        ctx->launches++;
        ITJL_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
        ITJL_CUDA(ctx, cudaEventSynchronize(ctx->ev1));
        float ms = 0.f;
        ITJL_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
        const double tf = fp32_peak_flops_per_launch(blocks, inner) / (ms * 1e-3) / 1e12;
        if (tf > best) best = tf;
    }
    *tflops_out = best;
    return ITJL_OK;
}

int itjl_tc_probe(itjl_ctx* ctx, const void* image, int64_t n_bytes, const int32_t* ops, int32_t n_ops,
                  uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t idesc, float* out) {
    if (!ctx || !image || !ops || !out) return fail(ctx, ITJL_ERR_ARG, "itjl_tc_probe: null argument");
    ITJL_LOCK(ctx);
    ctx = primary(ctx);
    ITJL_LOCK(ctx);
    if (n_bytes < 16 || (n_bytes & 15) || n_bytes > 200 * 1024 || n_ops < 1) return fail(ctx, ITJL_ERR_ARG, "itjl_tc_probe: bad image size / op count");
    ITJL_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc = upload(ctx, ctx->data, image, (size_t)n_bytes);
    if (rc == ITJL_OK) rc = upload(ctx, ctx->scratch, ops, (size_t)n_ops * 4 * sizeof(int32_t));
    if (rc != ITJL_OK) return rc;
    const size_t out_bytes = (size_t)128 * 512 * sizeof(float);
    ITJL_CUDA(ctx, ctx->out.reserve(out_bytes));
    ITJL_CUDA(ctx, cudaMemsetAsync(ctx->out.p, 0, out_bytes, ctx->stream));
    cudaError_t e = tc_probe_launch(ctx->data.p, (int)n_bytes, (const int4*)ctx->scratch.p, n_ops, lbo_bytes, sbo_bytes,
                                    idesc, (float*)ctx->out.p, ctx->stream);
    if (e != cudaSuccess) return fail(ctx, ITJL_ERR_CUDA, "launch k_tc_probe: %s", cudaGetErrorString(e));
    ctx->launches++;
    ITJL_CUDA(ctx, cudaMemcpyAsync(out, ctx->out.p, out_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    ITJL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ITJL_OK;
}

}  // extern "C"
