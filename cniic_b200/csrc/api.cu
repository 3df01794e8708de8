// api.cu -- context management, error reporting, NCCL (dlopen) plumbing of the cniic_b200 C ABI.
#include <dlfcn.h>

#include <algorithm>
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <vector>

#include "common.cuh"
#include "nccl_dyn.h"

int cniic_set_error(cniic_ctx *ctx, int code, const char *fmt, ...) {
    if (ctx) {
        char buf[512];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(buf, sizeof(buf), fmt, ap);
        va_end(ap);
        ctx->err = buf;
    }
    return code;
}

void *cniic_cache_alloc(cniic_ctx *ctx, size_t bytes) {
    bytes = (std::max<size_t>(bytes, 16) + 255) & ~size_t(255);
    int best = -1;
    for (size_t i = 0; i < ctx->cache.size(); i++) {
        const cniic_ctx::Block &b = ctx->cache[i];
        if (!b.used && b.bytes >= bytes && b.bytes <= 2 * bytes + (size_t(1) << 20) && (best < 0 || b.bytes < ctx->cache[best].bytes)) best = (int)i;
    }
    if (best >= 0) {
        ctx->cache[best].used = true;
        return ctx->cache[best].p;
    }
    void *p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) {
        // give cached-but-unused blocks back to the driver and retry once
        cudaGetLastError();
        cudaStreamSynchronize(ctx->stream);
        for (size_t i = 0; i < ctx->cache.size();) {
            if (!ctx->cache[i].used) { cudaFree(ctx->cache[i].p); ctx->cache.erase(ctx->cache.begin() + i); }
            else i++;
        }
        e = cudaMalloc(&p, bytes);
    }
    if (e != cudaSuccess) {
        cniic_set_error(ctx, CNIIC_ERR_CUDA, "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
        return nullptr;
    }
    ctx->cache.push_back({p, bytes, true});
    return p;
}

void cniic_cache_free(cniic_ctx *ctx, void *p) {
    if (!p) return;
    for (cniic_ctx::Block &b : ctx->cache)
        if (b.p == p) { b.used = false; return; }
}

void *cniic_pinned_get(cniic_ctx *ctx) {
    if (!ctx->pinned_free.empty()) {
        void *p = ctx->pinned_free.back();
        ctx->pinned_free.pop_back();
        return p;
    }
    void *p = nullptr;
    if (cudaMallocHost(&p, 256) != cudaSuccess) return nullptr;
    return p;
}

void cniic_pinned_put(cniic_ctx *ctx, void *p) {
    if (p) ctx->pinned_free.push_back(p);
}

void *cniic_pinned_big_get(cniic_ctx *ctx) {
    if (!ctx->pinned_big_free.empty()) {
        void *p = ctx->pinned_big_free.back();
        ctx->pinned_big_free.pop_back();
        return p;
    }
    void *p = nullptr;
    if (cudaMallocHost(&p, size_t(CNIIC_MAX_K) * 5 * 4) != cudaSuccess) return nullptr;
    return p;
}

void cniic_pinned_big_put(cniic_ctx *ctx, void *p) {
    if (p) ctx->pinned_big_free.push_back(p);
}

int cniic_launch_bump(cniic_ctx *ctx, uint32_t n) {
    ctx->launches += n;
    return CNIIC_OK;
}

// ---- NCCL through dlopen -------------------------------------------------------------------------------------
typedef struct { char internal[128]; } nccl_uid_t;
struct NcclApi {
    void *handle = nullptr;
    int (*GetUniqueId)(nccl_uid_t *) = nullptr;
    int (*CommInitRank)(void **, int, nccl_uid_t, int) = nullptr;
    int (*CommDestroy)(void *) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, void *, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
};
static const int NCCL_UINT8 = 1, NCCL_UINT64 = 5, NCCL_SUM = 0;  // ncclDataType_t / ncclRedOp_t values (nccl.h)

static NcclApi *nccl_load(std::string *why) {
    static NcclApi api;
    static bool tried = false, ok = false;
    static std::string err;
    if (!tried) {
        tried = true;
        const char *names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char *nm : names) {
            api.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
            if (api.handle) break;
        }
        if (!api.handle) {
            err = std::string("dlopen(libnccl.so.2) failed: ") + dlerror();
        } else {
            api.GetUniqueId = (int (*)(nccl_uid_t *))dlsym(api.handle, "ncclGetUniqueId");
            api.CommInitRank = (int (*)(void **, int, nccl_uid_t, int))dlsym(api.handle, "ncclCommInitRank");
            api.CommDestroy = (int (*)(void *))dlsym(api.handle, "ncclCommDestroy");
            api.AllReduce = (int (*)(const void *, void *, size_t, int, int, void *, cudaStream_t))dlsym(api.handle, "ncclAllReduce");
            api.AllGather = (int (*)(const void *, void *, size_t, int, void *, cudaStream_t))dlsym(api.handle, "ncclAllGather");
            api.GetErrorString = (const char *(*)(int))dlsym(api.handle, "ncclGetErrorString");
            ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllReduce && api.AllGather;
            if (!ok) err = "libnccl.so.2 lacks a required symbol";
        }
    }
    if (!ok && why) *why = err;
    return ok ? &api : nullptr;
}

extern "C" int cniic_nccl_unique_id(uint8_t out_id[128]) {
    std::string why;
    NcclApi *api = nccl_load(&why);
    if (!api || !out_id) return CNIIC_ERR_NCCL;
    nccl_uid_t id;
    if (api->GetUniqueId(&id) != 0) return CNIIC_ERR_NCCL;
    memcpy(out_id, id.internal, 128);
    return CNIIC_OK;
}

int cniic_nccl_init(cniic_ctx *ctx, int rank, int world, const uint8_t unique_id[128]) {
    std::string why;
    NcclApi *api = nccl_load(&why);
    if (!api) return cniic_set_error(ctx, CNIIC_ERR_NCCL, "%s", why.c_str());
    nccl_uid_t id;
    memcpy(id.internal, unique_id, 128);
    const int rc = api->CommInitRank(&ctx->comm, world, id, rank);
    if (rc != 0) return cniic_set_error(ctx, CNIIC_ERR_NCCL, "ncclCommInitRank: %s", api->GetErrorString ? api->GetErrorString(rc) : "error");
    ctx->nccl = api;
    ctx->rank = rank;
    ctx->world = world;
    return CNIIC_OK;
}

void cniic_nccl_destroy(cniic_ctx *ctx) {
    if (ctx->comm && ctx->nccl) ctx->nccl->CommDestroy(ctx->comm);
    ctx->comm = nullptr;
}

int cniic_nccl_allreduce_u64(cniic_ctx *ctx, unsigned long long *d_buf, size_t count) {
    if (!ctx->comm) return cniic_set_error(ctx, CNIIC_ERR_NCCL, "context has no communicator");
    const int rc = ctx->nccl->AllReduce(d_buf, d_buf, count, NCCL_UINT64, NCCL_SUM, ctx->comm, ctx->stream);
    if (rc != 0) return cniic_set_error(ctx, CNIIC_ERR_NCCL, "ncclAllReduce: %s", ctx->nccl->GetErrorString ? ctx->nccl->GetErrorString(rc) : "error");
    return CNIIC_OK;
}

int cniic_nccl_allgather_bytes(cniic_ctx *ctx, const void *d_send, void *d_recv, size_t bytes_per_rank) {
    if (!ctx->comm) return cniic_set_error(ctx, CNIIC_ERR_NCCL, "context has no communicator");
    const int rc = ctx->nccl->AllGather(d_send, d_recv, bytes_per_rank, NCCL_UINT8, ctx->comm, ctx->stream);
    if (rc != 0) return cniic_set_error(ctx, CNIIC_ERR_NCCL, "ncclAllGather: %s", ctx->nccl->GetErrorString ? ctx->nccl->GetErrorString(rc) : "error");
    return CNIIC_OK;
}

// ---- peer-memory exchange region (CUDA IPC between the per-GPU processes) ---------------------------------------------------
// layout (kmeans.cu: p2p_xcount_off): recv[2 parities][world source ranks][P2P_SUMS_MAX cells of 16 bytes] | u32 exchange counter.
// 6.3 MB for 8 ranks.
static const size_t P2P_SUMS_MAX_HOST = CNIIC_MAX_K * 6 + 8;
static size_t p2p_region_bytes(int world) { return 4 * size_t(world) * P2P_SUMS_MAX_HOST * 8 + 256; }

extern "C" int cniic_ctx_p2p_export(cniic_ctx *ctx, uint8_t out_handle[64]) {
    if (!ctx || !out_handle) return CNIIC_ERR_BAD_ARG;
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    if (!ctx->p2p_local) {
        void *p = nullptr;
        CU_TRY(ctx, cudaMalloc(&p, p2p_region_bytes(ctx->world)));
        CU_TRY(ctx, cudaMemset(p, 0, p2p_region_bytes(ctx->world)));
        ctx->p2p_local = static_cast<unsigned long long *>(p);
    }
    cudaIpcMemHandle_t h;
    CU_TRY(ctx, cudaIpcGetMemHandle(&h, ctx->p2p_local));
    static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
    memcpy(out_handle, &h, 64);
    return CNIIC_OK;
}

extern "C" int cniic_ctx_p2p_connect(cniic_ctx *ctx, const uint8_t *handles /* world x 64 */) {
    if (!ctx || !handles || !ctx->p2p_local) return CNIIC_ERR_BAD_ARG;
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    std::vector<unsigned long long *> bases(ctx->world);
    for (int r = 0; r < ctx->world; r++) {
        if (r == ctx->rank) { bases[r] = ctx->p2p_local; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, handles + 64 * r, 64);
        void *p = nullptr;
        CU_TRY(ctx, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        ctx->p2p_opened.push_back(p);
        bases[r] = static_cast<unsigned long long *>(p);
    }
    void *tab = nullptr;
    CU_TRY(ctx, cudaMalloc(&tab, sizeof(void *) * ctx->world));
    CU_TRY(ctx, cudaMemcpy(tab, bases.data(), sizeof(void *) * ctx->world, cudaMemcpyHostToDevice));
    ctx->p2p_peer_table = static_cast<unsigned long long **>(tab);
    ctx->p2p_ready = true;
    return CNIIC_OK;
}

// ---- context ---------------------------------------------------------------------------------------------------
extern "C" int cniic_version(void) { return 100; }

// ---- device-wide histogram bins (declared in common.cuh) ---------------------------------------------------------------
constexpr size_t BINS_PAGE = 4096;  // bins per page flag (stages.cu: PAGE)
struct DeviceBins {
    std::mutex lease[2];                      // held by the context that is counting into / compacting key space `kind`
    uint32_t *p[2] = {nullptr, nullptr};      // bins, then one flag byte per page
    cudaEvent_t clean_ev[2] = {nullptr, nullptr};  // recorded by the last holder behind the work that left the bins zero
    int contexts = 0;
};
static std::mutex g_bins_mu;
static std::map<int, DeviceBins> g_bins;  // by device ordinal (map nodes do not move)

static size_t bins_count(int kind) { return kind == 0 ? (size_t(1) << 24) : (size_t)511 * 511 * 511; }
static size_t bins_bytes(int kind) { return bins_count(kind) * 4 + (bins_count(kind) + BINS_PAGE - 1) / BINS_PAGE + 16; }

int cniic_bins_acquire(cniic_ctx *ctx, int kind, uint32_t **bins, uint8_t **flags, size_t *nbins) {
    DeviceBins *db = ctx->dev_bins;
    if (!ctx->bins_held[kind]) {
        db->lease[kind].lock();
        cudaError_t e = cudaSuccess;
        if (!db->p[kind]) {  // first use on this device: allocate, clear on my stream (later holders wait for my release event)
            void *p = nullptr;
            e = cudaMalloc(&p, bins_bytes(kind));
            if (e == cudaSuccess) e = cudaMemsetAsync(p, 0, bins_bytes(kind), ctx->stream);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&db->clean_ev[kind], cudaEventDisableTiming);
            if (e == cudaSuccess) e = cudaEventRecord(db->clean_ev[kind], ctx->stream);
            if (e != cudaSuccess) {
                if (p) cudaFree(p);
                if (db->clean_ev[kind]) { cudaEventDestroy(db->clean_ev[kind]); db->clean_ev[kind] = nullptr; }
                db->lease[kind].unlock();
                return cniic_set_error(ctx, CNIIC_ERR_CUDA, "histogram bins (%zu MB): %s", bins_bytes(kind) >> 20, cudaGetErrorString(e));
            }
            db->p[kind] = static_cast<uint32_t *>(p);
        } else if ((e = cudaStreamWaitEvent(ctx->stream, db->clean_ev[kind], 0)) != cudaSuccess) {
            db->lease[kind].unlock();
            return cniic_set_error(ctx, CNIIC_ERR_CUDA, "cudaStreamWaitEvent failed: %s", cudaGetErrorString(e));
        }
        ctx->bins_held[kind] = true;
    }
    *bins = db->p[kind];
    *nbins = bins_count(kind);
    *flags = reinterpret_cast<uint8_t *>(*bins + *nbins);
    return CNIIC_OK;
}

void cniic_bins_release(cniic_ctx *ctx, int kind, bool clean) {
    if (!ctx->bins_held[kind]) return;
    DeviceBins *db = ctx->dev_bins;
    if (!clean) cudaMemsetAsync(db->p[kind], 0, bins_bytes(kind), ctx->stream);
    cudaEventRecord(db->clean_ev[kind], ctx->stream);
    ctx->bins_held[kind] = false;
    db->lease[kind].unlock();
}

const uint32_t *cniic_bins_peek(const cniic_ctx *ctx, int kind) { return ctx->dev_bins ? ctx->dev_bins->p[kind] : nullptr; }

static int ctx_create_common(int device, cniic_ctx **out) {
    if (!out) return CNIIC_ERR_BAD_ARG;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) return CNIIC_ERR_CUDA;  // no CPU fallback: fail loudly
    if (device < 0) {
        if (cudaGetDevice(&device) != cudaSuccess) return CNIIC_ERR_CUDA;
    }
    if (device >= count) return CNIIC_ERR_BAD_ARG;
    if (cudaSetDevice(device) != cudaSuccess) return CNIIC_ERR_CUDA;
    cniic_ctx *ctx = new cniic_ctx();
    ctx->device = device;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
        delete ctx;
        return CNIIC_ERR_CUDA;
    }
    ctx->sm_count = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete ctx;
        return CNIIC_ERR_CUDA;
    }
    if (getenv("CNIIC_TLOG")) {
        if (cudaMalloc(&ctx->tlog, 64 * 8 * 8) == cudaSuccess) cudaMemset(ctx->tlog, 0, 64 * 8 * 8);
        else ctx->tlog = nullptr;
    }
    {
        std::lock_guard<std::mutex> g(g_bins_mu);
        ctx->dev_bins = &g_bins[device];
        ctx->dev_bins->contexts++;
    }
    *out = ctx;
    return CNIIC_OK;
}

extern "C" int cniic_ctx_create(int device, cniic_ctx **out) { return ctx_create_common(device, out); }

extern "C" int cniic_ctx_create_dist(int device, int rank, int world, const uint8_t nccl_unique_id[128], cniic_ctx **out) {
    if (world < 1 || rank < 0 || rank >= world || !nccl_unique_id) return CNIIC_ERR_BAD_ARG;
    ST_TRY(ctx_create_common(device, out));
    if (world == 1) return CNIIC_OK;
    const int rc = cniic_nccl_init(*out, rank, world, nccl_unique_id);
    if (rc != CNIIC_OK) {
        fprintf(stderr, "cniic_b200: %s\n", (*out)->err.c_str());
        cniic_ctx_destroy(*out);
        *out = nullptr;
    }
    return rc;
}

extern "C" void cniic_ctx_destroy(cniic_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    for (int kind = 0; kind < 2; kind++) cniic_bins_release(ctx, kind, false);  // (only held here if a call was abandoned half way)
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    cniic_nccl_destroy(ctx);
    for (cniic_ctx::Block &b : ctx->cache) cudaFree(b.p);
    if (ctx->dev_bins) {  // the last context on the device takes the bins with it
        std::lock_guard<std::mutex> g(g_bins_mu);
        if (--ctx->dev_bins->contexts == 0)
            for (int kind = 0; kind < 2; kind++) {
                if (ctx->dev_bins->p[kind]) cudaFree(ctx->dev_bins->p[kind]);
                if (ctx->dev_bins->clean_ev[kind]) cudaEventDestroy(ctx->dev_bins->clean_ev[kind]);
                ctx->dev_bins->p[kind] = nullptr;
                ctx->dev_bins->clean_ev[kind] = nullptr;
            }
    }
    for (void *p : ctx->p2p_opened) cudaIpcCloseMemHandle(p);
    if (ctx->p2p_peer_table) cudaFree(ctx->p2p_peer_table);
    if (ctx->p2p_local) cudaFree(ctx->p2p_local);
    if (ctx->tlog) cudaFree(ctx->tlog);
    for (void *p : ctx->pinned_free) cudaFreeHost(p);
    for (void *p : ctx->pinned_big_free) cudaFreeHost(p);
    for (cudaEvent_t e : ctx->event_pool) cudaEventDestroy(e);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

extern "C" const char *cniic_last_error(const cniic_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }

extern "C" int cniic_ctx_sync(cniic_ctx *ctx) {
    if (!ctx) return CNIIC_ERR_BAD_ARG;
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return CNIIC_OK;
}

extern "C" void *cniic_ctx_stream(cniic_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }
extern "C" int cniic_ctx_rank(const cniic_ctx *ctx) { return ctx ? ctx->rank : -1; }
extern "C" int cniic_ctx_world(const cniic_ctx *ctx) { return ctx ? ctx->world : 0; }
extern "C" uint32_t cniic_ctx_launches(const cniic_ctx *ctx) { return ctx ? ctx->launches : 0; }

extern "C" int cniic_ctx_set_max_iters(cniic_ctx *ctx, uint32_t max_iters) {
    if (!ctx) return CNIIC_ERR_BAD_ARG;
    ctx->codec_max_iters = max_iters;
    return CNIIC_OK;
}

extern "C" void *cniic_device_alloc(cniic_ctx *ctx, size_t bytes) {
    if (!ctx) return nullptr;
    void *p = nullptr;
    if (cudaSetDevice(ctx->device) != cudaSuccess) return nullptr;
    if (cudaMalloc(&p, bytes ? bytes : 16) != cudaSuccess) {
        cniic_set_error(ctx, CNIIC_ERR_CUDA, "cudaMalloc(%zu) failed", bytes);
        return nullptr;
    }
    return p;
}

extern "C" void cniic_device_free(cniic_ctx *ctx, void *p) {
    if (!ctx || !p) return;
    cudaSetDevice(ctx->device);
    cudaFree(p);
}

extern "C" int cniic_memcpy_h2d(cniic_ctx *ctx, void *d, const void *h, size_t bytes) {
    if (!ctx) return CNIIC_ERR_BAD_ARG;
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    CU_TRY(ctx, cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return CNIIC_OK;
}

extern "C" int cniic_memcpy_d2h(cniic_ctx *ctx, void *h, const void *d, size_t bytes) {
    if (!ctx) return CNIIC_ERR_BAD_ARG;
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    CU_TRY(ctx, cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return CNIIC_OK;
}
