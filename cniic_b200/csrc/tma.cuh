// tma.cuh -- the few sm_90+/sm_100a primitives the tile kernels use to move 2-D tiles with the Tensor Memory Accelerator:
// a tensor map of a row-major byte image (host, cuTensorMapEncodeTiled through the runtime's driver entry point -- no libcuda
// link), cp.async.bulk.tensor.2d global -> shared completing on an mbarrier, and the mbarrier init / arm / wait wrappers.
// SASS: UTMALDG (tile load), SYNCS.ARRIVE.TRANS64 (expect_tx), SYNCS.PHASECHK.TRANS64.TRYWAIT (wait).
#pragma once
#ifdef __CUDACC__
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

// 2-D tensor map of `rows` rows of `inner_bytes` bytes each, `row_stride_bytes` apart (multiple of 16), box = box_inner bytes x
// box_rows rows (box_inner a multiple of 16, <= 256), no swizzle, out-of-bounds bytes read as zero.
static inline bool tma_encode_2d_u8(CUtensorMap *map, const void *base, uint64_t inner_bytes, uint64_t rows, uint64_t row_stride_bytes,
                                    uint32_t box_inner, uint32_t box_rows) {
    typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static encode_fn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) return false;
        fn = reinterpret_cast<encode_fn>(p);
    }
    const cuuint64_t dims[2] = {inner_bytes, rows};
    const cuuint64_t strides[1] = {row_stride_bytes};
    const cuuint32_t box[2] = {box_inner, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(unsigned long long *bar, uint32_t arrivals) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(arrivals) : "memory");
}
// makes the initialised barriers visible to the async proxy (the TMA unit) before the first copy names them
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// one arrival + the number of bytes the copies armed on this phase will deliver
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// plain arrival (release): the consumers of a tile stage hand it back to the thread that refills it
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!done);
}
// box of the tensor map whose first element is (c_inner, c_row) -> dst (128-byte aligned shared memory), completes on `bar`
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int c_inner, int c_row, unsigned long long *bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
                 "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c_inner), "r"(c_row)
                 : "memory");
}
#endif  // __CUDACC__
