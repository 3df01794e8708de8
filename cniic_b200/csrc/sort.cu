// sort.cu -- one-time colour sort of the points of a D = 3 K-means session (preprocessing of the culled path).
// Points are sorted by the 24-bit Morton code of (r, g, b), so that consecutive points are close in colour space at
// every scale: tiles get small bounding boxes and the 8 points of one thread usually fall into one cluster.
// The sort itself is a library call (cub::DeviceRadixSort, 3 x 8-bit passes); key generation and unpacking are ours.
#include <cub/device/device_radix_sort.cuh>

#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "sort.cuh"

namespace {

__device__ __forceinline__ uint32_t part1by2(uint32_t x) {
    x &= 0xff;
    x = (x ^ (x << 16)) & 0xff0000ffu;
    x = (x ^ (x << 8)) & 0x0300f00fu;
    x = (x ^ (x << 4)) & 0x030c30c3u;
    x = (x ^ (x << 2)) & 0x09249249u;
    return x;
}
__device__ __forceinline__ uint32_t compact1by2(uint32_t x) {
    x &= 0x09249249u;
    x = (x ^ (x >> 2)) & 0x030c30c3u;
    x = (x ^ (x >> 4)) & 0x0300f00fu;
    x = (x ^ (x >> 8)) & 0xff0000ffu;
    x = (x ^ (x >> 16)) & 0x3ffu;
    return x;
}

__global__ void __launch_bounds__(256) sort_keys_kernel(const uint8_t *__restrict__ rgb, size_t n, uint32_t *keys, uint32_t *vals) {
    const bool al = (reinterpret_cast<uintptr_t>(rgb) & 3) == 0;
    const size_t quads = n / 4;
    for (size_t q = blockIdx.x * (size_t)blockDim.x + threadIdx.x; q < quads; q += (size_t)gridDim.x * blockDim.x) {
        uint32_t pk[4];
        if (al) {  // 4 points = three aligned words
            const uint32_t *p = reinterpret_cast<const uint32_t *>(rgb + q * 12);
            const uint32_t w0 = __ldg(p), w1 = __ldg(p + 1), w2 = __ldg(p + 2);
            pk[0] = w0 & 0xffffff; pk[1] = (w0 >> 24) | ((w1 & 0xffff) << 8); pk[2] = (w1 >> 16) | ((w2 & 0xff) << 16); pk[3] = w2 >> 8;
        } else {
            for (int j = 0; j < 4; j++) { const uint8_t *p = rgb + (q * 4 + j) * 3; pk[j] = uint32_t(p[0]) | (uint32_t(p[1]) << 8) | (uint32_t(p[2]) << 16); }
        }
        uint4 kv, iv;
        kv.x = (part1by2(pk[0]) << 2) | (part1by2(pk[0] >> 8) << 1) | part1by2(pk[0] >> 16);
        kv.y = (part1by2(pk[1]) << 2) | (part1by2(pk[1] >> 8) << 1) | part1by2(pk[1] >> 16);
        kv.z = (part1by2(pk[2]) << 2) | (part1by2(pk[2] >> 8) << 1) | part1by2(pk[2] >> 16);
        kv.w = (part1by2(pk[3]) << 2) | (part1by2(pk[3] >> 8) << 1) | part1by2(pk[3] >> 16);
        iv = make_uint4(uint32_t(q * 4), uint32_t(q * 4 + 1), uint32_t(q * 4 + 2), uint32_t(q * 4 + 3));
        reinterpret_cast<uint4 *>(keys)[q] = kv;
        reinterpret_cast<uint4 *>(vals)[q] = iv;
    }
    if (blockIdx.x == 0 && threadIdx.x < n - quads * 4) {
        const size_t i = quads * 4 + threadIdx.x;
        const uint8_t *p = rgb + i * 3;
        keys[i] = (part1by2(p[0]) << 2) | (part1by2(p[1]) << 1) | part1by2(p[2]);
        vals[i] = (uint32_t)i;
    }
}

__global__ void __launch_bounds__(256) sort_finish_kernel(const uint32_t *__restrict__ keys, const uint32_t *__restrict__ perm,
                                                          const uint32_t *__restrict__ wts, size_t n, uint32_t *pts_sorted, uint32_t *wts_sorted) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t key = keys[i];
        pts_sorted[i] = compact1by2(key >> 2) | (compact1by2(key >> 1) << 8) | (compact1by2(key) << 16);
        if (wts) wts_sorted[i] = wts[perm[i]];
    }
}

}  // namespace

int cniic_dev_sort_colours(cniic_ctx *ctx, const uint8_t *d_rgb, const uint32_t *d_wts, size_t n, uint32_t *d_sorted, uint32_t *d_perm,
                           uint32_t *d_wsorted, uint32_t *launches) {
    // Only the tiles' bounding boxes depend on the order, never the results: sorting on the top bits of the Morton code alone
    // (CNIIC_SORT_BITS, 8..24) saves radix passes at the price of slightly larger boxes inside a Morton cell.
    int sort_bits = 24;
    if (const char *e = getenv("CNIIC_SORT_BITS")) sort_bits = std::min(24, std::max(8, atoi(e)));
    const int lo_bit = 24 - sort_bits;
    DevBuf keys_in(ctx), keys_out(ctx), vals_in(ctx), tmp(ctx);
    CU_TRY(ctx, keys_in.alloc((n + 4) * 4));
    CU_TRY(ctx, keys_out.alloc((n + 4) * 4));
    CU_TRY(ctx, vals_in.alloc((n + 4) * 4));
    size_t tmp_bytes = 0;
    CU_TRY(ctx, cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys_in.as<uint32_t>(), keys_out.as<uint32_t>(), vals_in.as<uint32_t>(), d_perm,
                                                (int)n, lo_bit, 24, ctx->stream));
    CU_TRY(ctx, tmp.alloc(tmp_bytes));
    const int grid = (int)std::max<size_t>(1, std::min<size_t>((n / 4 + 255) / 256, (size_t)ctx->sm_count * 16));
    sort_keys_kernel<<<grid, 256, 0, ctx->stream>>>(d_rgb, n, keys_in.as<uint32_t>(), vals_in.as<uint32_t>());
    CU_TRY(ctx, cudaGetLastError());
    CU_TRY(ctx, cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, keys_in.as<uint32_t>(), keys_out.as<uint32_t>(), vals_in.as<uint32_t>(), d_perm,
                                                (int)n, lo_bit, 24, ctx->stream));
    sort_finish_kernel<<<grid, 256, 0, ctx->stream>>>(keys_out.as<uint32_t>(), d_perm, d_wts, n, d_sorted, d_wsorted);
    CU_TRY(ctx, cudaGetLastError());
    if (launches) *launches += 2 + 4;  // our two kernels + the library's histogram / onesweep passes
    return CNIIC_OK;
}
