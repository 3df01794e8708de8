// synth.cu -- synthetic "photo-like" RGB images, integer-only so host and device agree bit for bit.
// Two octaves of hashed-lattice value noise (bilinear, integer arithmetic) plus +-8 per-channel hashed noise.
#include "common.cuh"

namespace {

__host__ __device__ inline uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

__host__ __device__ inline uint32_t lattice(uint64_t seed, uint32_t gx, uint32_t gy) {
    return (uint32_t)splitmix64(seed ^ (((uint64_t)gy << 32) | gx));
}

__host__ __device__ inline void bilerp(uint64_t seed, uint32_t x, uint32_t y, uint32_t cell, uint32_t out[3]) {
    const uint32_t gx = x / cell, gy = y / cell, fx = x % cell, fy = y % cell;
    const uint32_t c00 = lattice(seed, gx, gy), c10 = lattice(seed, gx + 1, gy), c01 = lattice(seed, gx, gy + 1),
                   c11 = lattice(seed, gx + 1, gy + 1);
    const uint64_t w00 = (uint64_t)(cell - fx) * (cell - fy), w10 = (uint64_t)fx * (cell - fy),
                   w01 = (uint64_t)(cell - fx) * fy, w11 = (uint64_t)fx * fy;
    for (int ch = 0; ch < 3; ch++) {
        const uint64_t v = w00 * ((c00 >> (8 * ch)) & 0xff) + w10 * ((c10 >> (8 * ch)) & 0xff) +
                           w01 * ((c01 >> (8 * ch)) & 0xff) + w11 * ((c11 >> (8 * ch)) & 0xff);
        out[ch] = (uint32_t)(v / ((uint64_t)cell * cell));
    }
}

__host__ __device__ inline uint32_t isqrt_u64(uint64_t v) {
    uint64_t r = 0, bit = 1ull << 62;
    while (bit > v) bit >>= 2;
    while (bit) {
        if (v >= r + bit) { v -= r + bit; r = (r >> 1) + bit; }
        else r >>= 1;
        bit >>= 2;
    }
    return (uint32_t)r;
}

__host__ __device__ inline uint32_t cell_of(uint32_t w, uint32_t h_total, uint32_t n_blobs) {
    if (n_blobs == 0) n_blobs = 1;
    uint32_t c = isqrt_u64((uint64_t)w * h_total / n_blobs);
    return c < 8 ? 8 : c;
}

__host__ __device__ inline void synth_pixel(uint64_t seed, uint32_t x, uint32_t y, uint32_t w, uint32_t cell, uint8_t out[3]) {
    uint32_t a[3], b[3];
    const uint32_t fine = cell / 4 < 2 ? 2 : cell / 4;
    bilerp(seed, x, y, cell, a);
    bilerp(seed ^ 0xA5A5A5A5DEADBEEFull, x, y, fine, b);
    const uint64_t nz = splitmix64(seed ^ (0x51ED270B4C3Dull + (uint64_t)y * w + x));
    for (int ch = 0; ch < 3; ch++) {
        int v = (int)((3 * a[ch] + b[ch]) / 4) + (int)((nz >> (8 * ch)) % 17) - 8;
        out[ch] = (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
    }
}

__global__ void synth_kernel(uint8_t *rgb, uint32_t w, uint32_t h, uint32_t y0, uint32_t cell, uint64_t seed) {
    const uint64_t n = (uint64_t)w * h;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint8_t px[3];
        synth_pixel(seed, (uint32_t)(i % w), y0 + (uint32_t)(i / w), w, cell, px);
        rgb[3 * i] = px[0]; rgb[3 * i + 1] = px[1]; rgb[3 * i + 2] = px[2];
    }
}

}  // namespace

extern "C" int cniic_synth_image_device(cniic_ctx *ctx, uint8_t *d_rgb, uint32_t w, uint32_t h, uint32_t y0, uint32_t h_total,
                                        uint64_t seed, uint32_t n_blobs) {
    if (!ctx || !d_rgb || w == 0 || h == 0) return CNIIC_ERR_BAD_ARG;
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    synth_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(d_rgb, w, h, y0, cell_of(w, h_total ? h_total : h, n_blobs), seed);
    ctx->launches++;
    CU_TRY(ctx, cudaGetLastError());
    return CNIIC_OK;
}

extern "C" int cniic_synth_image_host(uint8_t *rgb, uint32_t w, uint32_t h, uint32_t y0, uint32_t h_total, uint64_t seed,
                                      uint32_t n_blobs) {
    if (!rgb || w == 0 || h == 0) return CNIIC_ERR_BAD_ARG;
    const uint32_t cell = cell_of(w, h_total ? h_total : h, n_blobs);
    for (uint32_t y = 0; y < h; y++)
        for (uint32_t x = 0; x < w; x++) synth_pixel(seed, x, y0 + y, w, cell, rgb + 3 * ((size_t)y * w + x));
    return CNIIC_OK;
}
