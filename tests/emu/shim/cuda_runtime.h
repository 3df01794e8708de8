// cuda_runtime.h (tests/emu shim) -- TEST INFRASTRUCTURE, NOT PART OF THE PRODUCT.
//
// A host-side stand-in for the CUDA runtime and the device intrinsics the cniic_b200 kernels use, so that the
// UNMODIFIED kernel sources (cniic_b200/csrc/*.cu, passed through tests/emu/build_emu.py) can be compiled with g++ and
// their logic -- barriers, warp collectives, shared-memory protocols, index arithmetic -- exercised on a machine
// without a GPU.  Every thread of a CTA is a cooperatively scheduled fiber; CTAs of a grid run one after another;
// warp collectives are rendezvous points between the fibers of a warp (tests/emu/emu_runtime.cpp).
//
// Nothing under cniic_b200/ loads the emulated library: it is built into tests/emu/_build/ and loaded only by
// tests/test_emu_kernels.py.  It says nothing about performance and is not a CPU fallback of the product.
#pragma once
#include <limits.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <functional>

#define CNIIC_EMU 1

// ---- qualifiers -----------------------------------------------------------------------------------------------
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __launch_bounds__(...)
#define __shared__ static
#define __constant__ static const
#define __align__(n) __attribute__((aligned(n)))

// ---- vector types -----------------------------------------------------------------------------------------------
struct alignas(8) uint2 { unsigned x, y; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
struct alignas(8) int2 { int x, y; };
struct alignas(16) int4 { int x, y, z, w; };
struct alignas(16) ulonglong2 { unsigned long long x, y; };
struct uchar4 { unsigned char x, y, z, w; };
struct ushort2 { unsigned short x, y; };
static inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
static inline int2 make_int2(int x, int y) { return int2{x, y}; }
static inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }
static inline ulonglong2 make_ulonglong2(unsigned long long x, unsigned long long y) { return ulonglong2{x, y}; }

struct dim3 {
    unsigned x = 1, y = 1, z = 1;
    dim3() {}
    dim3(unsigned long long x_, unsigned y_ = 1, unsigned z_ = 1) : x((unsigned)x_), y(y_), z(z_) {}
};

// ---- runtime API ------------------------------------------------------------------------------------------------
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorInvalidValue = 1, cudaErrorNotSupported = 801 };
typedef struct emu_stream *cudaStream_t;
typedef struct emu_event *cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaStreamNonBlocking = 1, cudaIpcMemLazyEnablePeerAccess = 1 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8, cudaFuncAttributePreferredSharedMemoryCarveout = 9 };
enum { cudaSharedmemCarveoutMaxShared = 100 };
struct cudaIpcMemHandle_t { char reserved[64]; };
struct cudaDeviceProp {
    char name[256];
    int multiProcessorCount, major, minor;
    size_t totalGlobalMem, sharedMemPerBlockOptin;
};

cudaError_t cudaMalloc(void **p, size_t bytes);
template <class T> static inline cudaError_t cudaMalloc(T **p, size_t bytes) { return cudaMalloc(reinterpret_cast<void **>(p), bytes); }
cudaError_t cudaFree(void *p);
cudaError_t cudaMallocHost(void **p, size_t bytes);
template <class T> static inline cudaError_t cudaMallocHost(T **p, size_t bytes) { return cudaMallocHost(reinterpret_cast<void **>(p), bytes); }
cudaError_t cudaFreeHost(void *p);
cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind k);
cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind k, cudaStream_t st = nullptr);
cudaError_t cudaMemset(void *d, int v, size_t n);
cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t st = nullptr);
cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned flags);
cudaError_t cudaStreamDestroy(cudaStream_t s);
cudaError_t cudaStreamSynchronize(cudaStream_t s);
cudaError_t cudaDeviceSynchronize();
cudaError_t cudaEventCreate(cudaEvent_t *e);
enum { cudaEventDisableTiming = 2 };
cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned flags);
cudaError_t cudaStreamWaitEvent(cudaStream_t s, cudaEvent_t e, unsigned flags = 0);
cudaError_t cudaEventDestroy(cudaEvent_t e);
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t s = nullptr);
cudaError_t cudaEventSynchronize(cudaEvent_t e);
cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b);
cudaError_t cudaGetLastError();
const char *cudaGetErrorString(cudaError_t e);
cudaError_t cudaGetDeviceCount(int *n);
cudaError_t cudaGetDevice(int *d);
cudaError_t cudaSetDevice(int d);
cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int dev);
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t *h, void *p);
cudaError_t cudaIpcOpenMemHandle(void **p, cudaIpcMemHandle_t h, unsigned flags);
cudaError_t cudaIpcCloseMemHandle(void *p);
int emu_max_dyn_smem();
int emu_blocks_per_sm();
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, int, int v) { return v <= emu_max_dyn_smem() ? cudaSuccess : cudaErrorInvalidValue; }
template <class F> static inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int *n, F, int, size_t smem) {
    *n = smem <= (size_t)emu_max_dyn_smem() ? emu_blocks_per_sm() : 0;
    return cudaSuccess;
}

// ---- the fiber machine ------------------------------------------------------------------------------------------
namespace emu {
struct Idx { unsigned x, y, z; };
struct ThreadCtx {
    Idx tid, bid, bdim, gdim;
    unsigned lane, warp, linear;
};
extern ThreadCtx *g_cur;  // the fiber that is running
void *dyn_smem();
void cta_barrier();
int cta_barrier_reduce(int pred, int op);  // op 0: or, 1: and, 2: count
// warp rendezvous: every lane named in `mask` (and still alive) deposits `v`; returns the deposited values and who deposited
struct WarpVals { unsigned long long v[32]; unsigned present; };
void warp_exchange(unsigned mask, unsigned long long v, int op, WarpVals *out);
struct LaunchCfg {
    dim3 grid, block;
    size_t smem;
    LaunchCfg(dim3 g, dim3 b, size_t s = 0, cudaStream_t = nullptr) : grid(g), block(b), smem(s) {}
};
void launch(const LaunchCfg &cfg, const std::function<void()> &body, const char *name);
void spin_while_equal(const volatile unsigned *p, unsigned val, const char *what);
}  // namespace emu

// cudaLaunchKernelEx with launch attributes (programmatic dependent launch): the emulation runs launches one after another
enum cudaLaunchAttributeID { cudaLaunchAttributeProgrammaticStreamSerialization = 4 };
struct cudaLaunchAttributeValue { int programmaticStreamSerializationAllowed; };
struct cudaLaunchAttribute { cudaLaunchAttributeID id; cudaLaunchAttributeValue val; };
struct cudaLaunchConfig_t { dim3 gridDim, blockDim; size_t dynamicSmemBytes; cudaStream_t stream; cudaLaunchAttribute *attrs; unsigned numAttrs; };
template <class... KA, class... A> static inline cudaError_t cudaLaunchKernelEx(const cudaLaunchConfig_t *c, void (*k)(KA...), A &&...a) {
    emu::launch(emu::LaunchCfg(c->gridDim, c->blockDim, c->dynamicSmemBytes, c->stream), [&]() { k(KA(a)...); }, "cudaLaunchKernelEx");
    return cudaSuccess;
}

// ---- stand-ins of cniic_b200/csrc/tma.cuh (tensor-map tile loads completing on an mbarrier): the copy is performed at once by
// the issuing thread; the barrier word holds the parity of the phase in progress (bit 0) ----
#define __grid_constant__
struct CUtensorMap { const unsigned char *base; unsigned long long inner_bytes, rows, row_stride; unsigned box_inner, box_rows; unsigned long long pad[11]; };
static inline bool tma_encode_2d_u8(CUtensorMap *m, const void *base, unsigned long long inner_bytes, unsigned long long rows,
                                    unsigned long long row_stride_bytes, unsigned box_inner, unsigned box_rows) {
    if (row_stride_bytes % 16 || box_inner % 16 || box_inner > 256 || box_rows > 256 || (reinterpret_cast<uintptr_t>(base) & 15)) return false;
    m->base = static_cast<const unsigned char *>(base); m->inner_bytes = inner_bytes; m->rows = rows; m->row_stride = row_stride_bytes;
    m->box_inner = box_inner; m->box_rows = box_rows;
    return true;
}
// bit 0: parity of the phase in progress; bits 32..46: arrivals so far; bit 47: bytes announced (expect_tx) and not yet delivered;
// bits 48..63: arrivals a phase needs.  A phase completes when all arrivals are in and no bytes are outstanding.
static inline void mbar_init(unsigned long long *bar, unsigned arrivals) { *bar = (unsigned long long)arrivals << 48; }
static inline void emu_mbar_update(unsigned long long *bar, unsigned long long got, bool tx_pending) {
    const unsigned long long need = *bar >> 48, phase = *bar & 1ull;
    if (got > need) { fprintf(stderr, "emu: more arrivals on an mbarrier than it was initialised for\n"); abort(); }
    if (got == need && !tx_pending) *bar = (need << 48) | (phase ^ 1ull);
    else *bar = (need << 48) | (tx_pending ? 1ull << 47 : 0ull) | (got << 32) | phase;
}
static inline void mbar_arrive(unsigned long long *bar) { emu_mbar_update(bar, ((*bar >> 32) & 0x7fffull) + 1, (*bar >> 47) & 1ull); }
static inline void mbar_fence_init() {}
static inline void mbar_arrive_expect_tx(unsigned long long *bar, unsigned) { emu_mbar_update(bar, ((*bar >> 32) & 0x7fffull) + 1, true); }
static inline void mbar_wait(unsigned long long *bar, unsigned parity) {
    emu::spin_while_equal(reinterpret_cast<const volatile unsigned *>(bar), parity, "mbar_wait (mbarrier phase not complete)");
}
static inline void tma_load_2d(void *dst, const CUtensorMap *m, int c_inner, int c_row, unsigned long long *bar) {
    if (reinterpret_cast<uintptr_t>(dst) & 127) { fprintf(stderr, "emu: TMA destination not 128-byte aligned\n"); abort(); }
    unsigned char *d = static_cast<unsigned char *>(dst);
    for (unsigned r = 0; r < m->box_rows; r++)
        for (unsigned b = 0; b < m->box_inner; b++) {
            const long long y = (long long)c_row + r, x = (long long)c_inner + b;
            d[(size_t)r * m->box_inner + b] = (y >= 0 && (unsigned long long)y < m->rows && x >= 0 && (unsigned long long)x < m->inner_bytes)
                                                  ? m->base[(size_t)y * m->row_stride + x] : 0;
        }
    emu_mbar_update(bar, (*bar >> 32) & 0x7fffull, false);  // the bytes are there: the phase completes once every arrival is in
}

#define threadIdx (emu::g_cur->tid)
#define blockIdx (emu::g_cur->bid)
#define blockDim (emu::g_cur->bdim)
#define gridDim (emu::g_cur->gdim)

static inline void __syncthreads() { emu::cta_barrier(); }
static inline int __syncthreads_or(int pred) { return emu::cta_barrier_reduce(pred, 0); }
static inline int __syncthreads_and(int pred) { return emu::cta_barrier_reduce(pred, 1); }
static inline int __syncthreads_count(int pred) { return emu::cta_barrier_reduce(pred, 2); }
static inline void __syncwarp(unsigned mask = 0xffffffffu) { emu::WarpVals w; emu::warp_exchange(mask, 0, 0, &w); }
static inline void __threadfence() {}
static inline void __threadfence_block() {}
static inline void __threadfence_system() {}

// ---- warp collectives ------------------------------------------------------------------------------------------------
namespace emu {
template <class T> static inline unsigned long long to_bits(T v) { unsigned long long b = 0; static_assert(sizeof(T) <= 8, ""); memcpy(&b, &v, sizeof(T)); return b; }
template <class T> static inline T from_bits(unsigned long long b) { T v; memcpy(&v, &b, sizeof(T)); return v; }
}
static inline unsigned __ballot_sync(unsigned mask, int pred) {
    emu::WarpVals w; emu::warp_exchange(mask, pred ? 1 : 0, 1, &w);
    unsigned r = 0;
    for (int i = 0; i < 32; i++) if ((w.present >> i & 1) && w.v[i]) r |= 1u << i;
    return r;
}
static inline int __all_sync(unsigned mask, int pred) {
    emu::WarpVals w; emu::warp_exchange(mask, pred ? 1 : 0, 2, &w);
    for (int i = 0; i < 32; i++) if ((w.present >> i & 1) && !w.v[i]) return 0;
    return 1;
}
static inline int __any_sync(unsigned mask, int pred) {
    emu::WarpVals w; emu::warp_exchange(mask, pred ? 1 : 0, 3, &w);
    for (int i = 0; i < 32; i++) if ((w.present >> i & 1) && w.v[i]) return 1;
    return 0;
}
template <class T> static inline T __shfl_sync(unsigned mask, T v, int src, int width = 32) {
    emu::WarpVals w; emu::warp_exchange(mask, emu::to_bits(v), 4, &w);
    const int lane = emu::g_cur->lane, base = lane & ~(width - 1);
    const int s = base + (src & (width - 1));
    return (w.present >> s & 1) ? emu::from_bits<T>(w.v[s]) : v;
}
template <class T> static inline T __shfl_xor_sync(unsigned mask, T v, int lanemask, int width = 32) {
    emu::WarpVals w; emu::warp_exchange(mask, emu::to_bits(v), 5, &w);
    const int lane = emu::g_cur->lane, s = lane ^ lanemask;
    if ((s & ~(width - 1)) != (lane & ~(width - 1)) || !(w.present >> s & 1)) return v;
    return emu::from_bits<T>(w.v[s]);
}
template <class T> static inline T __shfl_down_sync(unsigned mask, T v, unsigned delta, int width = 32) {
    emu::WarpVals w; emu::warp_exchange(mask, emu::to_bits(v), 6, &w);
    const int lane = emu::g_cur->lane, s = lane + (int)delta;
    if (s >= (lane & ~(width - 1)) + width || !(w.present >> s & 1)) return v;
    return emu::from_bits<T>(w.v[s]);
}
template <class T> static inline T __shfl_up_sync(unsigned mask, T v, unsigned delta, int width = 32) {
    emu::WarpVals w; emu::warp_exchange(mask, emu::to_bits(v), 7, &w);
    const int lane = emu::g_cur->lane, s = lane - (int)delta;
    if (s < (lane & ~(width - 1)) || !(w.present >> s & 1)) return v;
    return emu::from_bits<T>(w.v[s]);
}
template <class T> static inline T __reduce_add_sync(unsigned mask, T v) {
    emu::WarpVals w; emu::warp_exchange(mask, emu::to_bits(v), 8, &w);
    T r = 0;
    for (int i = 0; i < 32; i++) if (w.present >> i & 1) r += emu::from_bits<T>(w.v[i]);
    return r;
}
template <class T> static inline T __reduce_min_sync(unsigned mask, T v) {
    emu::WarpVals w; emu::warp_exchange(mask, emu::to_bits(v), 10, &w);
    T r = v;
    for (int i = 0; i < 32; i++) if (w.present >> i & 1) r = std::min(r, emu::from_bits<T>(w.v[i]));
    return r;
}
template <class T> static inline T __reduce_max_sync(unsigned mask, T v) {
    emu::WarpVals w; emu::warp_exchange(mask, emu::to_bits(v), 11, &w);
    T r = v;
    for (int i = 0; i < 32; i++) if (w.present >> i & 1) r = std::max(r, emu::from_bits<T>(w.v[i]));
    return r;
}
template <class T> static inline unsigned __match_any_sync(unsigned mask, T v) {
    emu::WarpVals w; emu::warp_exchange(mask, emu::to_bits(v), 9, &w);
    unsigned r = 0;
    const unsigned long long mine = emu::to_bits(v);
    for (int i = 0; i < 32; i++) if ((w.present >> i & 1) && w.v[i] == mine) r |= 1u << i;
    return r;
}

// address-space hints of the device compiler: no-ops here
#define __builtin_assume(x) ((void)0)
template <class T> static inline bool __isGlobal(const T *) { return true; }

// ---- scalar intrinsics ------------------------------------------------------------------------------------------------
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
static inline unsigned __brev(unsigned v) { unsigned r = 0; for (int i = 0; i < 32; i++) r |= ((v >> i) & 1u) << (31 - i); return r; }
static inline unsigned __byte_perm(unsigned a, unsigned b, unsigned s) {
    const unsigned long long src = (unsigned long long)b << 32 | a;
    unsigned r = 0;
    for (int i = 0; i < 4; i++) {
        const unsigned sel = (s >> (4 * i)) & 0xf;
        unsigned byte = (unsigned)(src >> (8 * (sel & 7))) & 0xff;
        if (sel & 8) byte = (byte & 0x80) ? 0xff : 0x00;
        r |= byte << (8 * i);
    }
    return r;
}
static inline unsigned __vminu4(unsigned a, unsigned b) {
    unsigned r = 0;
    for (int i = 0; i < 4; i++) r |= std::min((a >> 8 * i) & 0xff, (b >> 8 * i) & 0xff) << 8 * i;
    return r;
}
static inline unsigned __vmaxu4(unsigned a, unsigned b) {
    unsigned r = 0;
    for (int i = 0; i < 4; i++) r |= std::max((a >> 8 * i) & 0xff, (b >> 8 * i) & 0xff) << 8 * i;
    return r;
}
static inline unsigned __vabsdiffu4(unsigned a, unsigned b) {
    unsigned r = 0;
    for (int i = 0; i < 4; i++) {
        const int x = (a >> 8 * i) & 0xff, y = (b >> 8 * i) & 0xff;
        r |= unsigned(x > y ? x - y : y - x) << 8 * i;
    }
    return r;
}
static inline unsigned __vsub2(unsigned a, unsigned b) { return ((a - b) & 0xffffu) | (((a >> 16) - (b >> 16)) << 16); }
static inline unsigned __vcmpltu2(unsigned a, unsigned b) { return ((a & 0xffffu) < (b & 0xffffu) ? 0xffffu : 0u) | ((a >> 16) < (b >> 16) ? 0xffff0000u : 0u); }
static inline unsigned __vadd2(unsigned a, unsigned b) { return ((a + b) & 0xffffu) | (((a >> 16) + (b >> 16)) << 16); }
static inline unsigned __dp4a(unsigned a, unsigned b, unsigned c) {
    for (int i = 0; i < 4; i++) c += ((a >> 8 * i) & 0xff) * ((b >> 8 * i) & 0xff);
    return c;
}
static inline int __dp4a(int a, int b, int c) {
    for (int i = 0; i < 4; i++) c += int(int8_t((unsigned)a >> 8 * i)) * int(int8_t((unsigned)b >> 8 * i));
    return c;
}
// PTX forms named by the inline asm of common.cuh (build_emu.py rewrites `asm("dp4a.u32.u32 ...")` to these calls)
static inline int emu_ptx_dp4a_u32_u32(unsigned a, unsigned b, int c) { return (int)__dp4a(a, b, (unsigned)c); }
static inline int emu_ptx_dp2a_lo_s32_u32(unsigned a, unsigned b, int c) {  // c + a.s16[0]*b.u8[0] + a.s16[1]*b.u8[1]
    return c + int(int16_t(a & 0xffff)) * int(b & 0xff) + int(int16_t(a >> 16)) * int((b >> 8) & 0xff);
}
static inline long long clock64() { static long long t = 0; return ++t; }
template <class T> static inline T __ldg(const T *p) { return *p; }
template <class T> static inline T __ldcv(const T *p) { return *p; }
template <class T> static inline T __ldcg(const T *p) { return *p; }
template <class T> static inline T __ldcs(const T *p) { return *p; }
template <class T> static inline void __stcs(T *p, T v) { *p = v; }

// atomics: one fiber runs at a time, so plain read-modify-write is atomic
template <class T, class U> static inline T atomicAdd(T *p, U v) { const T o = *p; *p = T(o + T(v)); return o; }
template <class T, class U> static inline T atomicSub(T *p, U v) { const T o = *p; *p = T(o - T(v)); return o; }
template <class T, class U> static inline T atomicMin(T *p, U v) { const T o = *p; if (T(v) < o) *p = T(v); return o; }
template <class T, class U> static inline T atomicMax(T *p, U v) { const T o = *p; if (T(v) > o) *p = T(v); return o; }
template <class T, class U> static inline T atomicOr(T *p, U v) { const T o = *p; *p = T(o | T(v)); return o; }
template <class T, class U> static inline T atomicAnd(T *p, U v) { const T o = *p; *p = T(o & T(v)); return o; }
template <class T, class U> static inline T atomicExch(T *p, U v) { const T o = *p; *p = T(v); return o; }
template <class T, class U> static inline T atomicExch_system(T *p, U v) { return atomicExch(p, v); }
template <class T, class U> static inline T atomicMax_system(T *p, U v) { return atomicMax(p, v); }
template <class T, class U> static inline T atomicCAS(T *p, U cmp, U v) { const T o = *p; if (o == T(cmp)) *p = T(v); return o; }

// CUDA's integer min / max overloads (mixed signedness promotes to unsigned, like the device headers)
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
static inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
static inline unsigned min(int a, unsigned b) { return min((unsigned)a, b); }
static inline unsigned min(unsigned a, int b) { return min(a, (unsigned)b); }
static inline unsigned max(int a, unsigned b) { return max((unsigned)a, b); }
static inline unsigned max(unsigned a, int b) { return max(a, (unsigned)b); }
static inline long long min(long long a, long long b) { return a < b ? a : b; }
static inline long long max(long long a, long long b) { return a > b ? a : b; }
static inline unsigned long long min(unsigned long long a, unsigned long long b) { return a < b ? a : b; }
static inline unsigned long long max(unsigned long long a, unsigned long long b) { return a > b ? a : b; }
static inline unsigned long min(unsigned long a, unsigned long b) { return a < b ? a : b; }
static inline unsigned long max(unsigned long a, unsigned long b) { return a > b ? a : b; }
static inline long min(long a, long b) { return a < b ? a : b; }
static inline long max(long a, long b) { return a > b ? a : b; }
static inline float min(float a, float b) { return a < b ? a : b; }
static inline float max(float a, float b) { return a > b ? a : b; }
static inline double min(double a, double b) { return a < b ? a : b; }
static inline double max(double a, double b) { return a > b ? a : b; }
