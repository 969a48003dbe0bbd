// C-ABI layer, part 3: particles of one ABC generation sharded over the GPUs of a box, inside the library.
//
// The reference has no distributed layer; this is new.  Two ways to get a communicator:
//   itjl_ctx_create_multi  one process drives n_dev GPUs (a Julia host): one sub-context per device,
//                          ncclCommInitAll
//   itjl_ctx_attach_comm   one process per GPU (torchrun): every rank attaches its context with
//                          ncclCommInitRank and the sharded calls become collective
// In both, GPU g scores the contiguous range [g P / G, (g + 1) P / G) of the global particle index into its
// slot of an all-gather buffer and ONE ncclAllGather (fp64, 8 B per particle, device to device over NVLink)
// gives every GPU the whole distance vector: accept / stop, epsilon and the weights then run on every rank
// without another exchange.  There is no other data-path collective.
//
// libnccl.so.2 is opened with dlopen on first use: single-GPU users do not need NCCL at all, and a host process
// that already carries an NCCL (PyTorch) shares that copy instead of loading a second one.
#include <dlfcn.h>
#include <nccl.h>

#include <new>

#include "api_internal.h"

using namespace itjl;

namespace {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};

NcclApi* nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        api.handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
        if (!api.handle) api.handle = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
        if (!api.handle) return;
        auto sym = [&](const char* n) { return dlsym(api.handle, n); };
        api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
        api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
        api.CommInitAll = (decltype(api.CommInitAll))sym("ncclCommInitAll");
        api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
        api.AllGather = (decltype(api.AllGather))sym("ncclAllGather");
        api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
        api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
        api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
        api.ok = api.GetUniqueId && api.CommInitRank && api.CommInitAll && api.CommDestroy && api.AllGather &&
                 api.GroupStart && api.GroupEnd && api.GetErrorString;
    });
    return &api;
}

int nccl_fail(itjl_ctx* ctx, const char* what, ncclResult_t r) {
    return fail(ctx, ITJL_ERR_CUDA, "%s: NCCL %s", what, nccl_api()->GetErrorString ? nccl_api()->GetErrorString(r) : "error");
}

}  // namespace

namespace itjl {

void comm_destroy(itjl_ctx* ctx) {
    if (!ctx) return;
    NcclApi* n = nccl_api();
    if (ctx->comm && n->ok) { cudaSetDevice(ctx->device); n->CommDestroy((ncclComm_t)ctx->comm); }
    ctx->comm = nullptr;
}

void shard_members(itjl_ctx* ctx, std::vector<ShardMember>* out, int* world) {
    out->clear();
    if (!ctx->subs.empty()) {
        for (size_t i = 0; i < ctx->subs.size(); ++i) out->push_back({ctx->subs[i], (int)i});
        *world = (int)ctx->subs.size();
    } else {
        out->push_back({ctx, ctx->rank});
        *world = ctx->world;
    }
}

// After every member has written its rows [lo_r, hi_r) at gather + rank * width: all-gather in place and, on
// the primary context, compact the padded [world][width] layout into ctx->dist (n doubles).
// in-place all-gather of `width` doubles per rank: member i holds its slice at bases[i] + rank * width and ends up
// with all world * width values at bases[i] (one ncclAllGather per member, stream-ordered; no-op on one GPU)
int shard_allgather_inplace(itjl_ctx* ctx, const std::vector<double*>& bases, int64_t width) {
    std::vector<ShardMember> mem;
    int world = 1;
    shard_members(ctx, &mem, &world);
    if (world <= 1) return ITJL_OK;
    NcclApi* nc = nccl_api();
    if (!nc->ok) return fail(ctx, ITJL_ERR_ARG, "libnccl.so.2 could not be loaded");
    ncclResult_t r = nc->GroupStart();
    if (r != ncclSuccess) return nccl_fail(ctx, "ncclGroupStart", r);
    for (size_t i = 0; i < mem.size(); ++i) {
        cudaSetDevice(mem[i].c->device);
        r = nc->AllGather(bases[i] + (int64_t)mem[i].rank * width, bases[i], (size_t)width, ncclDouble, (ncclComm_t)mem[i].c->comm, mem[i].c->stream);
        if (r != ncclSuccess) { nc->GroupEnd(); return nccl_fail(ctx, "ncclAllGather", r); }
        mem[i].c->launches++;
    }
    r = nc->GroupEnd();
    if (r != ncclSuccess) return nccl_fail(ctx, "ncclGroupEnd", r);
    return ITJL_OK;
}

int shard_allgather_compact(itjl_ctx* ctx, int64_t n, double** out_dev) {
    std::vector<ShardMember> mem;
    int world = 1;
    shard_members(ctx, &mem, &world);
    const int64_t width = (n + world - 1) / world;
    NcclApi* nc = nccl_api();
    if (world > 1) {
        if (!nc->ok) return fail(ctx, ITJL_ERR_ARG, "libnccl.so.2 could not be loaded");
        ncclResult_t r = nc->GroupStart();
        if (r != ncclSuccess) return nccl_fail(ctx, "ncclGroupStart", r);
        for (const ShardMember& sm : mem) {
            cudaSetDevice(sm.c->device);
            double* base = (double*)sm.c->gather.p;
            r = nc->AllGather(base + (int64_t)sm.rank * width, base, (size_t)width, ncclDouble, (ncclComm_t)sm.c->comm, sm.c->stream);
            if (r != ncclSuccess) { nc->GroupEnd(); return nccl_fail(ctx, "ncclAllGather", r); }
            sm.c->launches++;
        }
        r = nc->GroupEnd();
        if (r != ncclSuccess) return nccl_fail(ctx, "ncclGroupEnd", r);
    }
    itjl_ctx* c0 = primary(ctx);
    ITJL_CUDA(ctx, cudaSetDevice(c0->device));
    ITJL_CUDA(ctx, c0->dist.reserve((size_t)n * sizeof(double)));
    for (int r = 0; r < world; ++r) {
        int64_t lo, hi;
        shard_bounds(n, world, r, &lo, &hi);
        if (hi > lo)
            ITJL_CUDA(ctx, cudaMemcpyAsync((double*)c0->dist.p + lo, (const double*)c0->gather.p + (int64_t)r * width,
                                           (size_t)(hi - lo) * sizeof(double), cudaMemcpyDeviceToDevice, c0->stream));
    }
    *out_dev = (double*)c0->dist.p;
    return ITJL_OK;
}

int sharded_distances(itjl_model* m, const double* theta, int64_t n, int32_t n_theta, uint64_t seed,
                      uint64_t generation, int64_t particle_offset, double** dist_dev_out) {
    itjl_ctx* ctx = m->ctx;
    std::vector<ShardMember> mem;
    int world = 1;
    shard_members(ctx, &mem, &world);
    const int64_t width = (n + world - 1) / world;
    NvtxRange nvtx("sharded_distances");
    for (size_t i = 0; i < mem.size(); ++i) {
        itjl_ctx* c = mem[i].c;
        itjl_model* sm = m->subs.empty() ? m : m->subs[i];
        int64_t lo, hi;
        shard_bounds(n, world, mem[i].rank, &lo, &hi);
        ITJL_CUDA(ctx, cudaSetDevice(c->device));
        ITJL_CUDA(ctx, c->gather.reserve((size_t)world * width * sizeof(double)));
        if (hi <= lo) continue;
        const int64_t nl = hi - lo;
        ITJL_CUDA(ctx, c->theta.reserve((size_t)nl * n_theta * sizeof(double)));
        for (int32_t col = 0; col < n_theta; ++col)           // column-major shard with leading dimension nl
            ITJL_CUDA(ctx, cudaMemcpyAsync((double*)c->theta.p + (int64_t)col * nl, theta + (int64_t)col * n + lo,
                                           (size_t)nl * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        const int rc = launch_distances(sm, (const double*)c->theta.p, nl, n_theta, seed, generation, particle_offset + lo,
                                        nullptr, nullptr, nullptr, (double*)c->gather.p + (int64_t)mem[i].rank * width, nullptr);
        if (rc != ITJL_OK) { if (c != ctx) snprintf(ctx->err, sizeof(ctx->err), "%s", c->err); return rc; }
    }
    return shard_allgather_compact(ctx, n, dist_dev_out);
}

int multi_model_create(itjl_ctx* ctx, const itjl_model_desc* desc, itjl_model** model_out) {
    itjl_model* m = new (std::nothrow) itjl_model();
    if (!m) return fail(ctx, ITJL_ERR_ALLOC, "out of host memory");
    m->ctx = ctx; m->d = *desc;
    m->d.target_stats = nullptr; m->d.missing_mask = nullptr; m->d.freq_bins = nullptr;
    for (itjl_ctx* sub : ctx->subs) {
        itjl_model* sm = nullptr;
        const int rc = itjl_model_create(sub, desc, &sm);
        if (rc != ITJL_OK) {
            snprintf(ctx->err, sizeof(ctx->err), "device %d: %.480s", sub->device, sub->err);
            itjl_model_destroy(m);
            return rc;
        }
        m->subs.push_back(sm);
    }
    *model_out = m;
    return ITJL_OK;
}

}  // namespace itjl

extern "C" {

int itjl_comm_unique_id(void* id_out) {
    if (!id_out) return fail(nullptr, ITJL_ERR_ARG, "itjl_comm_unique_id: null output");
    NcclApi* nc = nccl_api();
    if (!nc->ok) return fail(nullptr, ITJL_ERR_ARG, "libnccl.so.2 could not be loaded");
    static_assert(sizeof(ncclUniqueId) == NCCL_UNIQUE_ID_BYTES && NCCL_UNIQUE_ID_BYTES == 128, "unique id is 128 bytes in the ABI");
    ncclUniqueId id;
    const ncclResult_t r = nc->GetUniqueId(&id);
    if (r != ncclSuccess) return nccl_fail(nullptr, "ncclGetUniqueId", r);
    memcpy(id_out, &id, sizeof(id));
    return ITJL_OK;
}

int itjl_ctx_attach_comm(itjl_ctx* ctx, int rank, int world, const void* id_bytes) {
    if (!ctx || !id_bytes) return fail(ctx, ITJL_ERR_ARG, "itjl_ctx_attach_comm: null argument");
    ITJL_LOCK(ctx);
    if (!ctx->subs.empty() || ctx->comm) return fail(ctx, ITJL_ERR_ARG, "itjl_ctx_attach_comm: the context already has a communicator");
    if (world < 1 || rank < 0 || rank >= world) return fail(ctx, ITJL_ERR_ARG, "itjl_ctx_attach_comm: rank / world out of range");
    if (world == 1) { ctx->rank = 0; ctx->world = 1; return ITJL_OK; }
    NcclApi* nc = nccl_api();
    if (!nc->ok) return fail(ctx, ITJL_ERR_ARG, "libnccl.so.2 could not be loaded");
    ITJL_CUDA(ctx, cudaSetDevice(ctx->device));
    ncclUniqueId id;
    memcpy(&id, id_bytes, sizeof(id));
    ncclComm_t comm = nullptr;
    const ncclResult_t r = nc->CommInitRank(&comm, world, id, rank);
    if (r != ncclSuccess) return nccl_fail(ctx, "ncclCommInitRank", r);
    ctx->comm = comm; ctx->rank = rank; ctx->world = world;
    return ITJL_OK;
}

int itjl_ctx_create_multi(int n_dev, const int* dev_ids, itjl_ctx** ctx_out) {
    if (!ctx_out) return fail(nullptr, ITJL_ERR_ARG, "itjl_ctx_create_multi: null output");
    *ctx_out = nullptr;
    int avail = 0;
    if (cudaGetDeviceCount(&avail) != cudaSuccess || avail == 0) return fail(nullptr, ITJL_ERR_NO_DEVICE, "no CUDA device available");
    if (n_dev < 1 || n_dev > avail || n_dev > 64) return fail(nullptr, ITJL_ERR_ARG, "itjl_ctx_create_multi: n_dev out of range");
    std::vector<int> devs((size_t)n_dev);
    for (int i = 0; i < n_dev; ++i) {
        devs[(size_t)i] = dev_ids ? dev_ids[i] : i;
        for (int k = 0; k < i; ++k) if (devs[(size_t)k] == devs[(size_t)i]) return fail(nullptr, ITJL_ERR_ARG, "itjl_ctx_create_multi: duplicate device");
    }
    itjl_ctx* parent = new (std::nothrow) itjl_ctx();
    if (!parent) return fail(nullptr, ITJL_ERR_ALLOC, "out of host memory");
    itjl_tuning_default(&parent->tuning);
    parent->device = devs[0];
    parent->world = n_dev;
    for (int i = 0; i < n_dev; ++i) {
        itjl_ctx* sub = nullptr;
        const int rc = itjl_ctx_create(devs[(size_t)i], &sub);
        if (rc != ITJL_OK) { itjl_ctx_destroy(parent); return rc; }
        sub->parent = parent; sub->rank = i; sub->world = n_dev;
        parent->subs.push_back(sub);
    }
    if (n_dev > 1) {
        NcclApi* nc = nccl_api();
        if (!nc->ok) { itjl_ctx_destroy(parent); return fail(nullptr, ITJL_ERR_ARG, "libnccl.so.2 could not be loaded"); }
        std::vector<ncclComm_t> comms((size_t)n_dev, nullptr);
        const ncclResult_t r = nc->CommInitAll(comms.data(), n_dev, devs.data());
        if (r != ncclSuccess) { itjl_ctx_destroy(parent); return nccl_fail(nullptr, "ncclCommInitAll", r); }
        for (int i = 0; i < n_dev; ++i) parent->subs[(size_t)i]->comm = comms[(size_t)i];
    }
    parent->sm_count = parent->subs[0]->sm_count;
    parent->max_smem_optin = parent->subs[0]->max_smem_optin;
    *ctx_out = parent;
    return ITJL_OK;
}

int itjl_abc_distances_sharded_dev(itjl_model* m, const double* theta_dev, int64_t n_particles, int32_t n_theta,
                                   uint64_t seed, uint64_t generation, int64_t particle_offset, double* dist_dev) {
    if (!m) return fail(nullptr, ITJL_ERR_ARG, "null model");
    itjl_ctx* ctx = m->ctx;
    if (!theta_dev || !dist_dev || n_particles < 1 || n_theta < 1) return fail(ctx, ITJL_ERR_ARG, "itjl_abc_distances_sharded_dev: invalid argument");
    if (!ctx->subs.empty()) return fail(ctx, ITJL_ERR_ARG, "itjl_abc_distances_sharded_dev: device pointers belong to one device; use a single-device context (itjl_ctx_attach_comm)");
    ITJL_LOCK(ctx);
    NvtxRange nvtx("itjl_abc_distances_sharded_dev");
    ITJL_CUDA(ctx, cudaSetDevice(ctx->device));
    const int world = ctx->world;
    const int64_t width = (n_particles + world - 1) / world;
    int64_t lo, hi;
    shard_bounds(n_particles, world, ctx->rank, &lo, &hi);
    ITJL_CUDA(ctx, ctx->gather.reserve((size_t)world * width * sizeof(double)));
    if (hi > lo) {
        const int rc = launch_distances(m, theta_dev + lo, hi - lo, n_theta, seed, generation, particle_offset + lo,
                                        nullptr, nullptr, nullptr, (double*)ctx->gather.p + (int64_t)ctx->rank * width, nullptr,
                                        n_particles);
        if (rc != ITJL_OK) return rc;
    }
    double* all = nullptr;
    const int rc = shard_allgather_compact(ctx, n_particles, &all);
    if (rc != ITJL_OK) return rc;
    ITJL_CUDA(ctx, cudaMemcpyAsync(dist_dev, all, (size_t)n_particles * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    return ITJL_OK;
}

int itjl_ctx_world(const itjl_ctx* ctx, int* rank, int* world, int* n_local) {
    if (!ctx) return fail(nullptr, ITJL_ERR_ARG, "null context");
    if (rank) *rank = ctx->subs.empty() ? ctx->rank : 0;
    if (world) *world = ctx->subs.empty() ? ctx->world : (int)ctx->subs.size();
    if (n_local) *n_local = ctx->subs.empty() ? 1 : (int)ctx->subs.size();
    return ITJL_OK;
}

}  // extern "C"
