// common.cuh -- shared declarations of the cniic_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <string>
#include <vector>

#include "../../include/cniic_b200.h"

#ifndef __CUDA_ARCH__
#define CNIIC_HOST 1
#endif

struct NcclApi;

struct cniic_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    std::string err;
    int rank = 0, world = 1;
    void *comm = nullptr;  // ncclComm_t
    NcclApi *nccl = nullptr;
    uint32_t codec_max_iters = 0;
    uint32_t launches = 0;  // kernels launched on this ctx (bench.py reports it as gpu_launches)
    // device scratch cache: blocks are reused across calls (all work is ordered on `stream`, so reuse is safe);
    // nothing is returned to the driver before cniic_ctx_destroy -- cudaMalloc/cudaFree cost milliseconds per call
    struct Block { void *p; size_t bytes; bool used; };
    std::vector<Block> cache;
    std::vector<void *> pinned_free;  // 256-byte pinned host slots
    std::vector<void *> pinned_big_free;  // CNIIC_MAX_K * 5 * 4-byte pinned host staging buffers (initial centroids of sharded sessions)
    std::vector<cudaEvent_t> event_pool;
    // peer-memory exchange (multi-GPU): my IPC region (receive areas, flags, exchange counter), the peer-mapped bases of all ranks (device table)
    unsigned long long *p2p_local = nullptr;
    unsigned long long **p2p_peer_table = nullptr;  // device array [world]
    std::vector<void *> p2p_opened;
    bool p2p_ready = false;
    size_t xy_cull_smem = 0;  // dynamic shared memory the culled D = 5 kernels were last configured for on this context, and their
    int xy_cull_per_sm = 0;   // resident CTAs per SM at that size (cached: a session would otherwise repeat three driver calls)
    unsigned long long *tlog = nullptr;  // CNIIC_TLOG=1: device timeline buffer of the Lloyd loop (64 iterations x 8 timestamps)
    std::vector<uint8_t> pending_stream;  // cniic_codec_encode result that did not fit the caller's buffer (cniic_codec_encode_fetch)
    bool has_pending_stream = false;
    // dense histogram bins (+ page flags): ONE set per device, shared by every context on it and borrowed for the span of a
    // counting pass + its compaction (cniic_bins_acquire / cniic_bins_release in api.cu); all zero whenever nobody holds them
    struct DeviceBins *dev_bins = nullptr;
    bool bins_held[2] = {false, false};
};

// The two key spaces (kind 0: 2^24 colours, kind 1: 511^3 delta symbols) cost 64 MB and 534 MB.  A process that follows the
// reference's threading (bench.rs:27: one codec call per rayon worker, one context per worker) would hold them once per worker if
// they were a context's own; they belong to the device instead.  acquire: blocks until no other context on the device is using
// the key space, makes this context's stream wait for the work that left the bins zero, and returns them (idempotent while
// held).  release: `clean` says the work queued on ctx->stream so far zeroes them again (the compaction kernels do); otherwise --
// an error between counting and compaction -- they are cleared here.  Both are called on the thread that runs the C-ABI call.
int cniic_bins_acquire(cniic_ctx *ctx, int kind, uint32_t **bins, uint8_t **flags, size_t *nbins);
void cniic_bins_release(cniic_ctx *ctx, int kind, bool clean);
const uint32_t *cniic_bins_peek(const cniic_ctx *ctx, int kind);  // the device's bins of that kind if they exist (no lease), else nullptr

void *cniic_cache_alloc(cniic_ctx *ctx, size_t bytes);  // nullptr + error set on failure
void cniic_cache_free(cniic_ctx *ctx, void *p);
void *cniic_pinned_get(cniic_ctx *ctx);
void cniic_pinned_put(cniic_ctx *ctx, void *p);
void *cniic_pinned_big_get(cniic_ctx *ctx);  // CNIIC_MAX_K * 5 * 4 bytes
void cniic_pinned_big_put(cniic_ctx *ctx, void *p);

struct DevBuf {  // RAII scratch from the ctx cache
    cniic_ctx *ctx;
    void *p = nullptr;
    explicit DevBuf(cniic_ctx *c) : ctx(c) {}
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    ~DevBuf() { if (p) cniic_cache_free(ctx, p); }
    cudaError_t alloc(size_t bytes) { p = cniic_cache_alloc(ctx, bytes); return p ? cudaSuccess : cudaErrorMemoryAllocation; }
    template <class T> T *as() { return static_cast<T *>(p); }
};

int cniic_set_error(cniic_ctx *ctx, int code, const char *fmt, ...);

#define CU_TRY(ctx, expr)                                                                                       \
    do {                                                                                                        \
        cudaError_t e__ = (expr);                                                                               \
        if (e__ != cudaSuccess)                                                                                 \
            return cniic_set_error((ctx), CNIIC_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), \
                                   __FILE__, __LINE__);                                                         \
    } while (0)

#define ST_TRY(expr)             \
    do {                         \
        int s__ = (expr);        \
        if (s__ != CNIIC_OK) return s__; \
    } while (0)

// ---- integer dot-product instructions (IDP.4A / IDP.2A on sm_100a) ----
// dp4a: c + sum_i a.u8[i] * b.u8[i]   (two's-complement wrap makes the signed accumulator exact)
__device__ __forceinline__ int dp4a_uu(uint32_t a, uint32_t b, int c) {
    int d;
    asm("dp4a.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
// dp2a.lo: c + a.s16[0] * b.u8[0] + a.s16[1] * b.u8[1]
__device__ __forceinline__ int dp2a_lo_su(uint32_t a, uint32_t b, int c) {
    int d;
    asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
// ---- programmatic dependent launch (sm_90+): a kernel launched with launch_pdl() may become resident while its predecessor in
// the stream still runs; pdl_wait() (griddepcontrol.wait) returns once that predecessor has completed and its writes are visible,
// so it must precede the first access to anything an earlier kernel wrote; pdl_trigger() lets the NEXT kernel of the stream
// start the same way.  Removes the launch gap between the three short dependent kernels of a Lloyd iteration. ----
__device__ __forceinline__ void pdl_wait() {
#if defined(__CUDA_ARCH__)
    asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
}
__device__ __forceinline__ void pdl_trigger() {
#if defined(__CUDA_ARCH__)
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}
// start fetching a line into L1 without holding a destination register (the later load then hits L1)
__device__ __forceinline__ void prefetch_l1(const void *p) {
#if defined(__CUDA_ARCH__)
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
#endif
}
template <class... KArgs, class... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args &&...args) {
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cudaLaunchConfig_t cfg = {grid, block, smem, stream, at, 1u};  // {gridDim, blockDim, dynamicSmemBytes, stream, attrs, numAttrs}
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

__device__ __forceinline__ int max3i(int a, int b, int c) { return max(a, max(b, c)); }  // VIMNMX3

static inline uint32_t round_up_u32(uint32_t v, uint32_t m) { return (v + m - 1) / m * m; }

// internal entry points shared between translation units
int cniic_launch_bump(cniic_ctx *ctx, uint32_t n);
