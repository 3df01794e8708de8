// stages.cu -- the per-pixel stages around K-means: voronoi fill, colour / delta histograms, recolour,
// Hilbert index map, delta stream, SSE.  All HBM-bound integer/byte kernels except the fill (ALU).
#include <algorithm>
#include <vector>

#include "common.cuh"
#include "stages.cuh"
#include "tma.cuh"

namespace {

// ============================================================================================================
// voronoi decode fill  (reference src/codec/clusterc.rs:179-186)
//   per pixel: FIRST centroid minimising (cx-x)^2 + (cy-y)^2 ; exact u32 integers.
//   Tile = 64x64 pixels per CTA.  Exact culling: U = min_c maxdist^2(c, tile) bounds every pixel's minimum, so only
//   centroids with mindist^2(c, tile) <= U can win; they are compacted IN INDEX ORDER (first minimum = lowest index).
// ============================================================================================================
constexpr int FT = 64;

__device__ __forceinline__ uint32_t block_rank256(bool flag, uint32_t *s_warp, uint32_t *total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t bal = __ballot_sync(0xffffffffu, flag);
    __syncthreads();
    if (lane == 0) s_warp[warp] = __popc(bal);
    __syncthreads();
    uint32_t before = 0, tot = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const uint32_t v = s_warp[i];
        if (i < warp) before += v;
        tot += v;
    }
    *total = tot;
    return before + __popc(bal & ((1u << lane) - 1));
}

constexpr int FSW = 512, FSH = 256;  // supertile (8 x 4 tiles): level 1 of the culling

// Three levels of the same exact argument (U = min_c maxdist^2(c, box) bounds every pixel's minimum over the box, so a centroid with
// mindist^2(c, box) > U can neither win nor tie), each compacting IN ID ORDER so that strict "<" keeps the FIRST minimum = lowest id
// (Iterator::min_by_key, clusterc.rs:182-184):
//   level 1  512x256 supertile, over all k centroids, once per supertile a CTA enters (tiles are enumerated supertile by supertile
//            and a CTA owns a contiguous range: one or two supertiles) -> coordinates + colours of ~20 centroids in shared memory;
//   level 2  64x64 tile, over the supertile's list (shared memory to shared memory: no dependent global loads per tile -- the first
//            version chased list -> coordinates -> colours through global memory for every tile and was latency bound at 70 us);
//   level 3  64x8 strip of a warp, over the tile's list, survivors broadcast by shuffle and scored: ~5 centroids per pixel.
__global__ void __launch_bounds__(256, 4) fill_kernel(const uint32_t *__restrict__ cxy, const uint8_t *__restrict__ crgb, uint32_t k,
                                                   uint32_t w, uint32_t y0, uint32_t h_local, uint8_t *__restrict__ out) {
    extern __shared__ uint4 fsm[];
    int2 *s_sc = reinterpret_cast<int2 *>(fsm);                      // level-1 survivors: coordinates
    uint32_t *s_scol = reinterpret_cast<uint32_t *>(s_sc + k);       // ... colours, packed r | g<<8 | b<<16
    uint16_t *s_ti = reinterpret_cast<uint16_t *>(s_scol + k);       // level-2 survivors: positions in the level-1 list
    __shared__ uint32_t s_warp[8];
    __shared__ uint32_t s_U[2];
    __shared__ uint32_t s_wcol[8][64];  // per warp: colours of its level-3 survivors by slot
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t tiles_x = (w + FT - 1) / FT, tiles_y = (h_local + FT - 1) / FT;
    const uint32_t super_x = (w + FSW - 1) / FSW;
    constexpr uint32_t STX = FSW / FT, STY = FSH / FT;
    const uint32_t tiles = tiles_x * tiles_y;
    const uint32_t per_cta = (tiles + gridDim.x - 1) / gridDim.x;
    const uint32_t t_begin = blockIdx.x * per_cta, t_end = min(tiles, t_begin + per_cta);
    uint32_t cur_sup = 0xffffffffu, ms = 0;
    for (uint32_t tile_seq = t_begin; tile_seq < t_end; tile_seq++) {
        uint32_t tx, ty, sup;
        {   // supertile-major enumeration of the tiles
            static_assert(STX == 8 && STY == 4, "the decode below shifts by 3 and 5");  // (runtime divisions only in edge supertiles)
            const uint32_t row_tiles = tiles_x * STY;
            const uint32_t sy = tile_seq / row_tiles, rem = tile_seq - sy * row_tiles;
            const uint32_t rows_in = min(STY, tiles_y - sy * STY);
            const uint32_t sx = rows_in == STY ? rem >> 5 : rem / (STX * rows_in), rem2 = rem - sx * STX * rows_in;
            const uint32_t cols_in = min(STX, tiles_x - sx * STX);
            const uint32_t ly = cols_in == STX ? rem2 >> 3 : rem2 / cols_in;
            tx = sx * STX + (rem2 - ly * cols_in);
            ty = sy * STY + ly;
            sup = sy * super_x + sx;
        }
        if (sup != cur_sup) {  // ---- level 1 ----
            __syncthreads();   // the previous tile is done with the lists
            const int sx0 = (sup % super_x) * FSW, syl0 = (sup / super_x) * FSH;
            const int sx1 = min(sx0 + FSW, (int)w) - 1, sy0 = y0 + syl0, sy1 = y0 + min(syl0 + FSH, (int)h_local) - 1;
            if (tid == 0) s_U[0] = 0xffffffffu;
            __syncthreads();
            uint32_t um = 0xffffffffu;
            for (uint32_t c = tid; c < k; c += 256) {
                const int cx = (int)__ldg(cxy + 2 * c), cy = (int)__ldg(cxy + 2 * c + 1);
                const uint32_t dx = max(abs(cx - sx0), abs(cx - sx1)), dy = max(abs(cy - sy0), abs(cy - sy1));
                um = min(um, dx * dx + dy * dy);
            }
            for (int o = 16; o > 0; o >>= 1) um = min(um, __shfl_xor_sync(0xffffffffu, um, o));
            if (lane == 0) atomicMin(&s_U[0], um);
            __syncthreads();
            const uint32_t US = s_U[0];
            uint32_t placed = 0;
            for (uint32_t cb = 0; cb < k; cb += 256) {
                const uint32_t c = cb + tid;
                bool keep = false;
                int cx = 0, cy = 0;
                if (c < k) {
                    cx = (int)__ldg(cxy + 2 * c); cy = (int)__ldg(cxy + 2 * c + 1);
                    const uint32_t dx = max(0, max(sx0 - cx, cx - sx1)), dy = max(0, max(sy0 - cy, cy - sy1));
                    keep = dx * dx + dy * dy <= US;
                }
                uint32_t tot;
                const uint32_t r = block_rank256(keep, s_warp, &tot);
                if (keep) {
                    s_sc[placed + r] = make_int2(cx, cy);
                    s_scol[placed + r] = uint32_t(crgb[3 * c]) | (uint32_t(crgb[3 * c + 1]) << 8) | (uint32_t(crgb[3 * c + 2]) << 16);
                }
                placed += tot;
            }
            ms = placed;
            cur_sup = sup;
        }
        const int x0 = tx * FT, yl0 = ty * FT;
        const int x1 = min(x0 + FT, (int)w) - 1, yl1 = min(yl0 + FT, (int)h_local) - 1;
        const int gy0 = y0 + yl0, gy1 = y0 + yl1;
        // ---- level 2 ----
        __syncthreads();  // level-1 lists complete / the previous tile is done with s_ti
        if (tid == 0) s_U[1] = 0xffffffffu;
        __syncthreads();
        uint32_t umin = 0xffffffffu;
        for (uint32_t j = tid; j < ms; j += 256) {
            const int2 c = s_sc[j];
            const uint32_t dx = max(abs(c.x - x0), abs(c.x - x1)), dy = max(abs(c.y - gy0), abs(c.y - gy1));
            umin = min(umin, dx * dx + dy * dy);
        }
        for (int o = 16; o > 0; o >>= 1) umin = min(umin, __shfl_xor_sync(0xffffffffu, umin, o));
        if (lane == 0) atomicMin(&s_U[1], umin);
        __syncthreads();
        const uint32_t U = s_U[1];
        uint32_t ncand = 0;
        for (uint32_t jb = 0; jb < ms; jb += 256) {
            const uint32_t j = jb + tid;
            bool keep = false;
            if (j < ms) {
                const int2 c = s_sc[j];
                const uint32_t dx = max(0, max(x0 - c.x, c.x - x1)), dy = max(0, max(gy0 - c.y, c.y - gy1));
                keep = dx * dx + dy * dy <= U;
            }
            uint32_t tot;
            const uint32_t r = block_rank256(keep, s_warp, &tot);
            if (keep) s_ti[ncand + r] = (uint16_t)j;
            ncand += tot;
        }
        __syncthreads();
        // ---- level 3: warp -> a 64 x 8 strip of the tile; lane -> one row of it, 16 consecutive pixels ----
        const int wy0 = gy0 + warp * 8, wy1 = min(wy0 + 7, gy1);
        if (wy0 > gy1) continue;  // (warp-uniform; the strip lies below the image / shard)
        uint32_t uw = 0xffffffffu;
        for (uint32_t j = lane; j < ncand; j += 32) {
            const int2 c = s_sc[s_ti[j]];
            const uint32_t dx = max(abs(c.x - x0), abs(c.x - x1)), dy = max(abs(c.y - wy0), abs(c.y - wy1));
            uw = min(uw, dx * dx + dy * dy);
        }
        uw = __reduce_min_sync(0xffffffffu, uw);
        const int row = lane >> 2, seg = (lane & 3) * 16;
        const int gy = wy0 + row, xs = x0 + seg;
        uint32_t bc[16];
        // Fast path (any realistic image): the strip's bound uw < 2^25 means every survivor lies within 5 793 + 65 pixels of every
        // pixel of the strip, so d^2 < 2^26 and `d^2 * 64 + slot` fits 32 bits: the running minimum of that key IS the first
        // minimum (slots are handed out in ascending id order), three instructions per pixel and centroid (IADD, IMAD, VIMNMX)
        // instead of five and half the registers.  More than 64 survivors or a huge bound take the compare/select loop below.
        uint32_t *w_col = s_wcol[warp];
        bool packed_ok = uw < (1u << 25);
        uint32_t key[16];
#pragma unroll
        for (int p = 0; p < 16; p++) key[p] = 0xffffffffu;
        uint32_t nslot = 0;
        for (uint32_t jb = 0; jb < ncand && packed_ok; jb += 32) {
            const uint32_t j = jb + lane;
            int2 c = make_int2(0, 0);
            uint32_t col = 0;
            bool keep = false;
            if (j < ncand) {
                const uint32_t e = s_ti[j];
                c = s_sc[e];
                col = s_scol[e];
                const uint32_t dx = max(0, max(x0 - c.x, c.x - x1)), dy = max(0, max(wy0 - c.y, c.y - wy1));
                keep = dx * dx + dy * dy <= uw;
            }
            uint32_t mask = __ballot_sync(0xffffffffu, keep);
            if (nslot + __popc(mask) > 64u) { packed_ok = false; break; }
            if (keep) w_col[nslot + __popc(mask & ((1u << lane) - 1))] = col;
            while (mask) {  // warp-uniform, ascending ids
                const int src = __ffs(mask) - 1;
                mask &= mask - 1;
                const int ccx = __shfl_sync(0xffffffffu, c.x, src), ccy = __shfl_sync(0xffffffffu, c.y, src);
                const int dy = ccy - gy;
                const uint32_t base = uint32_t(dy * dy) * 64u + nslot;
                const int c8 = (ccx - xs) * 8;
#pragma unroll
                for (int p = 0; p < 16; p++) {
                    const int d8 = c8 - 8 * p;
                    key[p] = min(key[p], uint32_t(d8 * d8) + base);
                }
                nslot++;
            }
        }
        if (packed_ok) {
            __syncwarp();
#pragma unroll
            for (int p = 0; p < 16; p++) bc[p] = w_col[key[p] & 63u];
        } else {
            uint32_t bd[16];
#pragma unroll
            for (int p = 0; p < 16; p++) { bd[p] = 0xffffffffu; bc[p] = 0; }
            for (uint32_t jb = 0; jb < ncand; jb += 32) {
                const uint32_t j = jb + lane;
                int2 c = make_int2(0, 0);
                uint32_t col = 0;
                bool keep = false;
                if (j < ncand) {
                    const uint32_t e = s_ti[j];
                    c = s_sc[e];
                    col = s_scol[e];
                    const uint32_t dx = max(0, max(x0 - c.x, c.x - x1)), dy = max(0, max(wy0 - c.y, c.y - wy1));
                    keep = dx * dx + dy * dy <= uw;
                }
                uint32_t mask = __ballot_sync(0xffffffffu, keep);
                while (mask) {  // warp-uniform, ascending ids
                    const int src = __ffs(mask) - 1;
                    mask &= mask - 1;
                    const int ccx = __shfl_sync(0xffffffffu, c.x, src), ccy = __shfl_sync(0xffffffffu, c.y, src);
                    const uint32_t ccol = __shfl_sync(0xffffffffu, col, src);
                    const int dy = ccy - gy;
                    const uint32_t dy2 = dy * dy;
#pragma unroll
                    for (int p = 0; p < 16; p++) {
                        const int dx = ccx - (xs + p);
                        const uint32_t dd = dx * dx + dy2;
                        if (dd < bd[p]) { bd[p] = dd; bc[p] = ccol; }
                    }
                }
            }
        }
        __syncwarp();  // (w_col is rewritten by the next tile)
        if (gy > gy1) continue;
        uint8_t *o = out + ((size_t)(gy - (int)y0) * w + xs) * 3;
        if (xs + 15 <= x1 && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {  // 16 pixels = three 128-bit stores
            uint32_t wd[12];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const uint32_t a = bc[4 * q], b = bc[4 * q + 1], c = bc[4 * q + 2], e = bc[4 * q + 3];
                wd[3 * q] = a | (b << 24); wd[3 * q + 1] = (b >> 8) | (c << 16); wd[3 * q + 2] = (c >> 16) | (e << 8);
            }
            uint4 *o4 = reinterpret_cast<uint4 *>(o);
            o4[0] = make_uint4(wd[0], wd[1], wd[2], wd[3]); o4[1] = make_uint4(wd[4], wd[5], wd[6], wd[7]); o4[2] = make_uint4(wd[8], wd[9], wd[10], wd[11]);
        } else {
#pragma unroll
            for (int p = 0; p < 16; p++)
                if (xs + p <= x1) { o[3 * p] = bc[p]; o[3 * p + 1] = bc[p] >> 8; o[3 * p + 2] = bc[p] >> 16; }
        }
    }
}

// ============================================================================================================
// dense histogram + ordered compaction (utils::count_freqs, reference src/utils.rs:4-16)
// ============================================================================================================
__device__ __forceinline__ uint32_t rgb_key(const uint8_t *p) { return (uint32_t(p[0]) << 16) | (uint32_t(p[1]) << 8) | p[2]; }

// warp-aggregated atomic increment: lanes holding the same key elect a leader that adds the group's population
// `flags` marks the 4096-bin pages that received a count, so compaction and re-zeroing touch only those pages.
constexpr int PAGE_SHIFT = 12;
__device__ __forceinline__ void warp_hist_add(uint32_t *bins, uint8_t *flags, uint32_t key, bool valid) {
    const uint32_t act = __ballot_sync(0xffffffffu, valid);
    if (!valid) return;
    const uint32_t peers = __match_any_sync(act, key);
    if ((__ffs(peers) - 1) == (threadIdx.x & 31)) {
        // first touch of a bin marks its page (one flag store per distinct key instead of one per pixel)
        if (atomicAdd(&bins[key], (uint32_t)__popc(peers)) == 0) flags[key >> PAGE_SHIFT] = 1;
    }
}

__global__ void hist_rgb_kernel(const uint8_t *__restrict__ rgb, size_t n, uint32_t *bins, uint8_t *flags) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t n_round = (n + 31) / 32 * 32;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n_round; i += stride) {
        const bool valid = i < n;
        warp_hist_add(bins, flags, valid ? rgb_key(rgb + 3 * i) : 0, valid);
    }
}

// single-block exclusive scan of block_counts (u32 -> u64 offsets); total to offsets[nblocks]
__global__ void __launch_bounds__(1024) scan_blocks_kernel(const uint32_t *__restrict__ counts, size_t nblocks, unsigned long long *offsets) {
    __shared__ unsigned long long s_warp[32];
    __shared__ unsigned long long s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (size_t base = 0; base < nblocks; base += 1024) {
        const size_t i = base + threadIdx.x;
        unsigned long long v = i < nblocks ? counts[i] : 0ull, x = v;
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) s_warp[warp] = x;
        __syncthreads();
        unsigned long long before = s_carry;
        for (int j = 0; j < warp; j++) before += s_warp[j];
        if (i < nblocks) offsets[i] = before + x - v;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = before + x;
        __syncthreads();
    }
    if (threadIdx.x == 0) offsets[nblocks] = s_carry;
}

// ---- paged variant: the bins of the two key spaces live in the context for its whole life and are ALL ZERO between
// calls; a histogram pass marks the pages it touched, compaction visits only those pages and zeroes them again.
constexpr int PAGE = 1 << PAGE_SHIFT;

__global__ void __launch_bounds__(1024) list_pages_kernel(const uint8_t *__restrict__ flags, uint32_t npages, uint32_t *list, uint32_t *count) {
    __shared__ uint32_t s_w[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t placed = 0;
    for (uint32_t base = 0; base < npages; base += 1024) {
        const uint32_t pg = base + threadIdx.x;
        const bool f = pg < npages && flags[pg];
        const uint32_t bal = __ballot_sync(0xffffffffu, f);
        __syncthreads();
        if (lane == 0) s_w[warp] = __popc(bal);
        __syncthreads();
        uint32_t before = 0, tot = 0;
        for (int i = 0; i < 32; i++) { const uint32_t v = s_w[i]; if (i < warp) before += v; tot += v; }
        if (f) list[placed + before + __popc(bal & ((1u << lane) - 1))] = pg;
        placed += tot;
    }
    if (threadIdx.x == 0) *count = placed;
}

__global__ void __launch_bounds__(256) page_count_kernel(const uint32_t *__restrict__ bins, size_t nbins, const uint32_t *__restrict__ list,
                                                         const uint32_t *__restrict__ count, uint32_t *block_counts) {
    if (blockIdx.x >= *count) { block_counts[blockIdx.x] = 0; return; }
    const size_t base = (size_t)list[blockIdx.x] * PAGE;
    uint32_t c = 0;
    for (int j = 0; j < PAGE / 256; j++) {
        const size_t i = base + (size_t)j * 256 + threadIdx.x;
        if (i < nbins && bins[i]) c++;
    }
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    __shared__ uint32_t s[8];
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int i = 0; i < 8; i++) t += s[i];
        block_counts[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(256) page_compact_kernel(uint32_t *bins, size_t nbins, const uint32_t *__restrict__ list, const uint32_t *__restrict__ count,
                                                           const unsigned long long *__restrict__ offsets, uint8_t *flags, uint32_t *out_keys,
                                                           unsigned long long *out_counts, size_t cap) {
    __shared__ uint32_t s_warp[8];
    if (blockIdx.x >= *count) return;
    const uint32_t pg = list[blockIdx.x];
    const size_t base = (size_t)pg * PAGE;
    unsigned long long pos = offsets[blockIdx.x];
    for (int j = 0; j < PAGE / 256; j++) {
        const size_t i = base + (size_t)j * 256 + threadIdx.x;
        const uint32_t v = i < nbins ? bins[i] : 0;
        uint32_t tot;
        const uint32_t r = block_rank256(v != 0, s_warp, &tot);
        if (v) {
            if (pos + r < cap) { out_keys[pos + r] = (uint32_t)i; out_counts[pos + r] = v; }
            bins[i] = 0;  // restore the all-zero invariant
        }
        pos += tot;
    }
    if (threadIdx.x == 0) flags[pg] = 0;
}

// ============================================================================================================
// cluster-colors helpers (reference src/codec/clusterc.rs:19-47)
// ============================================================================================================

// ---- unique colours in the order the culled D = 3 K-means wants, without a sort (clusterc.rs:19-28, utils.rs:4-16) ------------
// count_freqs turns the pixels into (colour, count) points.  The dense bins are indexed by the 24-bit MORTON code of (r, g, b), so
// the ordered compaction of the bins IS the deduplicated, weighted, Morton-sorted point list the culled kernel scans:
// no radix sort and, on photo-like images, a fraction of the pixels as points.  The reference's HashMap order is random (SURVEY
// F5); the canonical order of the unique colours of cluster-colors is DEFINED as this one -- ascending Morton code, r on the most
// significant bit of every triple (header, oracle_cluster_colors) -- so the list needs no permutation at all: the chunked init
// (kmeans.rs:61-108) and the empty-cluster rule index it directly.
__device__ __forceinline__ uint32_t spread3(uint32_t x) {  // 8 bits -> every third bit
    x &= 0xff;
    x = (x ^ (x << 16)) & 0xff0000ffu;
    x = (x ^ (x << 8)) & 0x0300f00fu;
    x = (x ^ (x << 4)) & 0x030c30c3u;
    x = (x ^ (x << 2)) & 0x09249249u;
    return x;
}
__device__ __forceinline__ uint32_t gather3(uint32_t x) {
    x &= 0x09249249u;
    x = (x ^ (x >> 2)) & 0x030c30c3u;
    x = (x ^ (x >> 4)) & 0x0300f00fu;
    x = (x ^ (x >> 8)) & 0xff0000ffu;
    x = (x ^ (x >> 16)) & 0x3ffu;
    return x;
}
__device__ __forceinline__ uint32_t morton_rgb(uint32_t r, uint32_t g, uint32_t b) { return (spread3(r) << 2) | (spread3(g) << 1) | spread3(b); }

// one fire-and-forget reduction (RED.ADD) per pixel, or per run of equal neighbours; nothing is read back
__global__ void __launch_bounds__(256) dedup_hist_kernel(const uint8_t *__restrict__ rgb, size_t n, uint32_t *bins) {
    const bool al = (reinterpret_cast<uintptr_t>(rgb) & 3) == 0;
    const size_t quads = n / 4;
    for (size_t q = blockIdx.x * (size_t)blockDim.x + threadIdx.x; q < quads; q += (size_t)gridDim.x * blockDim.x) {
        uint32_t pk[4];  // r | g << 8 | b << 16
        if (al) {        // 4 pixels = three aligned words
            const uint32_t *p = reinterpret_cast<const uint32_t *>(rgb + q * 12);
            const uint32_t w0 = __ldg(p), w1 = __ldg(p + 1), w2 = __ldg(p + 2);
            pk[0] = w0 & 0xffffff; pk[1] = (w0 >> 24) | ((w1 & 0xffff) << 8); pk[2] = (w1 >> 16) | ((w2 & 0xff) << 16); pk[3] = w2 >> 8;
        } else {
            for (int j = 0; j < 4; j++) { const uint8_t *p = rgb + (q * 4 + j) * 3; pk[j] = uint32_t(p[0]) | (uint32_t(p[1]) << 8) | (uint32_t(p[2]) << 16); }
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {
            if (j > 0 && pk[j] == pk[j - 1]) continue;  // a run of equal neighbours (flat areas) costs one reduction
            uint32_t cnt = 1;
            for (int t = j + 1; t < 4 && pk[t] == pk[j]; t++) cnt++;
            atomicAdd(&bins[morton_rgb(pk[j] & 0xff, (pk[j] >> 8) & 0xff, pk[j] >> 16)], cnt);
        }
    }
    if (blockIdx.x == 0 && threadIdx.x < n - quads * 4) {
        const uint8_t *p = rgb + (quads * 4 + threadIdx.x) * 3;
        atomicAdd(&bins[morton_rgb(p[0], p[1], p[2])], 1u);
    }
}

// non-empty bins per 4096-bin page (every page is visited: a photo-like image touches nearly all of them)
__global__ void __launch_bounds__(256) dedup_count_kernel(const uint32_t *__restrict__ bins, uint32_t *page_counts) {
    const uint4 *p = reinterpret_cast<const uint4 *>(bins + (size_t)blockIdx.x * PAGE) + threadIdx.x * 4;
    uint32_t c = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const uint4 v = p[j];
        c += (v.x != 0) + (v.y != 0) + (v.z != 0) + (v.w != 0);
    }
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    __shared__ uint32_t s[8];
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int i = 0; i < 8; i++) t += s[i];
        page_counts[blockIdx.x] = t;
    }
}

// ordered compaction: thread t owns 16 consecutive bins of the page; packed colour r | g<<8 | b<<16 (the layout the K-means
// kernels read) and count; the bins go back to all-zero
__global__ void __launch_bounds__(256) dedup_compact_kernel(uint32_t *bins, const unsigned long long *__restrict__ offsets, uint32_t *out_pts, uint32_t *out_wts) {
    __shared__ uint32_t s_warp[8];
    uint4 *p = reinterpret_cast<uint4 *>(bins + (size_t)blockIdx.x * PAGE) + threadIdx.x * 4;
    uint32_t v[16];
#pragma unroll
    for (int j = 0; j < 4; j++) { const uint4 q = p[j]; v[4 * j] = q.x; v[4 * j + 1] = q.y; v[4 * j + 2] = q.z; v[4 * j + 3] = q.w; }
    uint32_t c = 0;
#pragma unroll
    for (int j = 0; j < 16; j++) c += v[j] != 0;
    // exclusive prefix of the per-thread counts over the block (thread order = bin order)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t x = c;
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) s_warp[warp] = x;
    __syncthreads();
    uint32_t before = x - c;
    for (int j = 0; j < warp; j++) before += s_warp[j];
    if (c) {
        unsigned long long pos = offsets[blockIdx.x] + before;
        const uint32_t m0 = blockIdx.x * PAGE + threadIdx.x * 16;
#pragma unroll
        for (int j = 0; j < 16; j++)
            if (v[j]) {
                const uint32_t m = m0 + j;
                out_pts[pos] = gather3(m >> 2) | (gather3(m >> 1) << 8) | (gather3(m) << 16);
                out_wts[pos] = v[j];  // clusterc.rs:23 "count as u32"
                pos++;
            }
#pragma unroll
        for (int j = 0; j < 4; j++) p[j] = make_uint4(0u, 0u, 0u, 0u);  // restore the all-zero invariant
    }
}

// lut[key] = centroid colour (packed r | g<<8 | b<<16) straight from the sorted point list and its assignment
__global__ void build_lut_sorted_kernel(const uint32_t *__restrict__ pts_sorted, const uint16_t *__restrict__ assign_sorted, size_t n,
                                        const int32_t *__restrict__ cen, uint32_t *lut) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t v = pts_sorted[i];
        const int32_t *c = cen + 3 * assign_sorted[i];
        lut[((v & 0xff) << 16) | (v & 0xff00) | (v >> 16)] = uint32_t(c[0]) | (uint32_t(c[1]) << 8) | (uint32_t(c[2]) << 16);
    }
}

__global__ void keys_to_points_kernel(const uint32_t *__restrict__ keys, const unsigned long long *__restrict__ counts, size_t n,
                                      uint8_t *rgb, uint32_t *wts) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t key = keys[i];
        rgb[3 * i] = key >> 16; rgb[3 * i + 1] = key >> 8; rgb[3 * i + 2] = key;
        wts[i] = (uint32_t)counts[i];  // clusterc.rs:23 "count as u32"
    }
}

// lut[key] = centroid colour packed r | g<<8 | b<<16
__global__ void build_lut_kernel(const uint32_t *__restrict__ keys, const uint16_t *__restrict__ assign, size_t n,
                                 const int32_t *__restrict__ cen, uint32_t *lut) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int32_t *c = cen + 3 * assign[i];
        lut[keys[i]] = uint32_t(c[0]) | (uint32_t(c[1]) << 8) | (uint32_t(c[2]) << 16);
    }
}

__global__ void build_lut_u8_kernel(const uint32_t *__restrict__ keys, const uint16_t *__restrict__ assign, size_t n,
                                    const uint8_t *__restrict__ cen, uint32_t *lut) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const uint8_t *c = cen + 3 * assign[i];
        lut[keys[i]] = uint32_t(c[0]) | (uint32_t(c[1]) << 8) | (uint32_t(c[2]) << 16);
    }
}

__global__ void recolor_kernel(const uint8_t *__restrict__ rgb, size_t n, const uint32_t *__restrict__ lut, uint8_t *out) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t v = lut[rgb_key(rgb + 3 * i)];
        out[3 * i] = v; out[3 * i + 1] = v >> 8; out[3 * i + 2] = v >> 16;
    }
}

// ============================================================================================================
// Hilbert scan (reference src/hilbert.rs:40-43 -> zhang_hilbert, PARITY UNPINNED; same curve as the oracle)
// ============================================================================================================
struct HRect { int x, y, ax, ay, bx, by; };

__device__ __forceinline__ int isgn(int v) { return (v > 0) - (v < 0); }
__device__ __forceinline__ int half_floor(int v) { return v >> 1; }  // arithmetic shift == floor division by 2

// generic rectangle: descend the recursive halving until the index lands in a 1-wide strip
__device__ void hilbert_d2xy_generic(uint32_t w, uint32_t h, unsigned long long d, uint32_t *ox, uint32_t *oy) {
    int x = 0, y = 0, ax, ay, bx, by;
    if (w >= h) { ax = w; ay = 0; bx = 0; by = h; }
    else { ax = 0; ay = h; bx = w; by = 0; }
    for (;;) {
        const int ww = abs(ax + ay), hh = abs(bx + by);
        const int dax = isgn(ax), day = isgn(ay), dbx = isgn(bx), dby = isgn(by);
        if (hh == 1) { *ox = x + dax * (int)d; *oy = y + day * (int)d; return; }
        if (ww == 1) { *ox = x + dbx * (int)d; *oy = y + dby * (int)d; return; }
        int ax2 = half_floor(ax), ay2 = half_floor(ay), bx2 = half_floor(bx), by2 = half_floor(by);
        const int w2 = abs(ax2 + ay2), h2 = abs(bx2 + by2);
        if (2 * ww > 3 * hh) {
            if ((w2 & 1) && ww > 2) { ax2 += dax; ay2 += day; }
            const unsigned long long n1 = (unsigned long long)abs(ax2 + ay2) * hh;
            if (d < n1) { ax = ax2; ay = ay2; }
            else { d -= n1; x += ax2; y += ay2; ax -= ax2; ay -= ay2; }
        } else {
            if ((h2 & 1) && hh > 2) { bx2 += dbx; by2 += dby; }
            const unsigned long long n1 = (unsigned long long)abs(bx2 + by2) * abs(ax2 + ay2);
            const unsigned long long n2 = (unsigned long long)ww * abs((bx - bx2) + (by - by2));
            if (d < n1) {
                const int tax = bx2, tay = by2;
                bx = ax2; by = ay2; ax = tax; ay = tay;
            } else if (d < n1 + n2) {
                d -= n1; x += bx2; y += by2; bx -= bx2; by -= by2;
            } else {
                d -= n1 + n2;
                x += (ax - dax) + (bx2 - dbx); y += (ay - day) + (by2 - dby);
                const int nax = -bx2, nay = -by2, nbx = -(ax - ax2), nby = -(ay - ay2);
                ax = nax; ay = nay; bx = nbx; by = nby;
            }
        }
    }
}

// 2^n x 2^n squares: the scan above IS the classic Hilbert curve (first step +y for odd n, +x for even n)
__device__ __forceinline__ void hilbert_d2xy_pow2(uint32_t n, unsigned long long d, uint32_t *ox, uint32_t *oy) {
    uint32_t x = 0, y = 0;
    unsigned long long t = d;
    for (uint32_t s = 1; s < n; s <<= 1) {
        const uint32_t rx = 1u & (uint32_t)(t >> 1), ry = 1u & ((uint32_t)t ^ rx);
        if (ry == 0) {
            if (rx == 1) { x = s - 1 - x; y = s - 1 - y; }
            const uint32_t tmp = x; x = y; y = tmp;
        }
        x += s * rx; y += s * ry;
        t >>= 2;
    }
    *ox = x; *oy = y;
}

__device__ __forceinline__ void hilbert_d2xy(uint32_t w, uint32_t h, bool pow2, unsigned long long d, uint32_t *ox, uint32_t *oy) {
    if (pow2) hilbert_d2xy_pow2(w, d, ox, oy);
    else hilbert_d2xy_generic(w, h, d, ox, oy);
}

__global__ void hilbert_xy_kernel(uint32_t w, uint32_t h, bool pow2, uint32_t *out) {
    const unsigned long long n = (unsigned long long)w * h;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * blockDim.x) {
        uint32_t x, y;
        hilbert_d2xy(w, h, pow2, i, &x, &y);
        out[2 * i] = x; out[2 * i + 1] = y;
    }
}

// mode 0: gather rgb along the curve; mode 1: delta stream (i16 x 3); mode 2: delta histogram only (fused)
template <int MODE>
__global__ void hilbert_stream_kernel(const uint8_t *__restrict__ rgb, uint32_t w, uint32_t h, bool pow2, uint8_t *out_rgb,
                                      int16_t *out_delta, uint32_t *bins, uint8_t *flags, unsigned long long i_begin, unsigned long long i_end) {
    // curve indices [i_begin, i_end) (a rank's share of the curve; the whole curve for a single GPU); outputs are
    // written relative to i_begin.  The predecessor of a warp's first index is recomputed, so a range needs no halo.
    const unsigned long long n = i_end;
    const unsigned long long n_round = i_begin + (i_end - i_begin + 31) / 32 * 32;
    for (unsigned long long i = i_begin + blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n_round;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const bool valid = i < n;
        uint32_t x = 0, y = 0;
        int c0 = 0, c1 = 0, c2 = 0;
        if (valid) {
            hilbert_d2xy(w, h, pow2, i, &x, &y);
            const uint8_t *p = rgb + ((size_t)y * w + x) * 3;
            c0 = p[0]; c1 = p[1]; c2 = p[2];
        }
        if (MODE == 0) {
            if (valid) { out_rgb[3 * (i - i_begin)] = c0; out_rgb[3 * (i - i_begin) + 1] = c1; out_rgb[3 * (i - i_begin) + 2] = c2; }
        } else {
            // predecessor along the curve: neighbouring lane, or one extra index map for lane 0
            int p0 = __shfl_up_sync(0xffffffffu, c0, 1), p1 = __shfl_up_sync(0xffffffffu, c1, 1), p2 = __shfl_up_sync(0xffffffffu, c2, 1);
            if ((threadIdx.x & 31) == 0) {
                p0 = p1 = p2 = 0;  // hilbertc.rs:445 START = [0;3]
                if (valid && i > 0) {
                    uint32_t px, py;
                    hilbert_d2xy(w, h, pow2, i - 1, &px, &py);
                    const uint8_t *q = rgb + ((size_t)py * w + px) * 3;
                    p0 = q[0]; p1 = q[1]; p2 = q[2];
                }
            }
            const int d0 = c0 - p0, d1 = c1 - p1, d2 = c2 - p2;
            if (MODE == 1) {
                if (valid) { out_delta[3 * (i - i_begin)] = d0; out_delta[3 * (i - i_begin) + 1] = d1; out_delta[3 * (i - i_begin) + 2] = d2; }
            } else {
                warp_hist_add(bins, flags, valid ? uint32_t(((d0 + 255) * 511 + (d1 + 255)) * 511 + (d2 + 255)) : 0, valid);
            }
        }
    }
}

// ---- fast path for 2^n x 2^n images (n >= 6): one CTA per 64x64 block = 4096 consecutive curve indices ----------
// The classic index->(x,y) map builds coordinates from the LOW base-4 digits upward, and every higher level applies
// the same swap / flip / translate to the partial result.  The 16 indices of one thread (a 4x4 cell) share all digits
// above the lowest two, so the thread folds levels 2..L-1 once into an affine map  x = ax + sx*u, y = ay + sy*v
// ((u,v) = the 4x4 pattern, possibly transposed) and applies it to the 16 pattern entries.  The block's pixels are
// staged in shared memory with coalesced 128-bit row loads; outputs are 128-bit stores of 16 consecutive symbols.
__constant__ uint8_t HIL4_X[16] = {0, 1, 1, 0, 0, 0, 1, 1, 2, 2, 3, 3, 3, 2, 2, 3};
__constant__ uint8_t HIL4_Y[16] = {0, 0, 1, 1, 2, 3, 3, 2, 2, 3, 3, 2, 1, 1, 0, 0};

constexpr int HT = 64;            // block side
constexpr int HT_STRIDE = 68;     // words per staged row (64 + 4 pad: rows shift by 4 banks, 128-bit aligned)

// near-zero delta symbols (|d| <= 15: 97 % of the symbols of the noisy benchmark image, more for photographs) are counted in
// shared memory: 31^3 counters packed two per word = 58 KB (+ 17 KB of staged pixels: 2 CTAs per SM).  A field holds 15 bits of
// count plus a guard bit: the increment that sets the guard moves 2^15 counts to the global bin and clears the guard again, so
// a CTA flushes ONCE, at the end of its persistent loop, whatever the image.  (Between that increment and its subtraction at
// most 4095 other increments of the tile can land -- the thread reaches the tile's next barrier first -- far from the 2^15 that
// would carry into the neighbouring field.)  The first version flushed every 12 tiles, ~20 M contended global atomics per
// 8192^2 image, and launched 3 CTAs per SM where only 2 fit (a 1.5-wave tail).
constexpr int CUBE_R = 15, CUBE_S = 2 * CUBE_R + 1, CUBE_N = CUBE_S * CUBE_S * CUBE_S;

template <int MODE>
__global__ void __launch_bounds__(256) hilbert_tile_kernel(const uint8_t *__restrict__ rgb, uint32_t n, uint8_t *out_rgb,
                                                           int16_t *out_delta, uint32_t *bins, uint8_t *flags, unsigned long long blk_begin,
                                                           unsigned long long blk_end) {
    extern __shared__ uint32_t s_cube[];  // MODE 2 only: CUBE_N 15-bit counters (+ guard bit), two per word
    __shared__ __align__(16) uint32_t s_px[HT * HT_STRIDE];  // one word per pixel (r | g<<8 | b<<16)
    __shared__ uint32_t s_last[8];
    __shared__ int s_top[5];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned long long nblocks = blk_end;  // 4096-index blocks [blk_begin, blk_end): a rank's share of the curve; outputs relative to it
    if (MODE == 2) {
        for (int i = tid; i < (CUBE_N + 1) / 2; i += 256) s_cube[i] = 0;
    }
    for (unsigned long long blk = blk_begin + blockIdx.x; blk < nblocks; blk += gridDim.x) {
        const unsigned long long B = blk * 4096;
        const unsigned long long i0 = B + (unsigned long long)tid * 16;
        // The block's 4096 indices share every base-4 digit above the lowest six: lane 0 of warp 0 folds those levels
        // (6..L-1) into one affine map of the 64x64 block, every thread folds only its own levels 2..5.
        __syncthreads();  // previous block is done with s_px / s_last / s_top
        if (tid == 0) {
            int bx = 0, by = 0, tx = 1, ty = 1, sw = 0;
            unsigned long long t = B >> 12;
            for (uint32_t sft = HT; sft < n; sft <<= 1) {
                const int sl = (int)sft;
                const uint32_t rx = 1u & (uint32_t)(t >> 1), ry = 1u & ((uint32_t)t ^ rx);
                if (ry == 0) {
                    if (rx == 1) {
                        const int nbx = sl - 1 - by, ntx = -ty, nby = sl - 1 - bx, nty = -tx;
                        bx = nbx; tx = ntx; by = nby; ty = nty;
                    } else {
                        const int q = bx, qs = tx;
                        bx = by; tx = ty; by = q; ty = qs;
                    }
                    sw ^= 1;
                }
                bx += sl * (int)rx;
                by += sl * (int)ry;
                t >>= 2;
            }
            s_top[0] = bx; s_top[1] = tx; s_top[2] = by; s_top[3] = ty; s_top[4] = sw;
        }
        int ax = 0, ay = 0, sx = 1, sy = 1;
        bool swapped = false;
        {
            uint32_t t = (uint32_t)(i0 >> 4) & 0xff;  // digits 2..5
#pragma unroll
            for (int sl = 4; sl < HT; sl <<= 1) {
                const uint32_t rx = 1u & (t >> 1), ry = 1u & (t ^ rx);
                if (ry == 0) {
                    if (rx == 1) {
                        const int nax = sl - 1 - ay, nsx = -sy, nay = sl - 1 - ax, nsy = -sx;
                        ax = nax; sx = nsx; ay = nay; sy = nsy;
                    } else {
                        const int q = ax, qs = sx;
                        ax = ay; sx = sy; ay = q; sy = qs;
                    }
                    swapped = !swapped;
                }
                ax += sl * (int)rx;
                ay += sl * (int)ry;
                t >>= 2;
            }
        }
        __syncthreads();
        {   // compose: (x, y) = top(local(u, v)); x = bx + tx * (sw ? yl : xl), y = by + ty * (sw ? xl : yl)
            const int bx = s_top[0], tx = s_top[1], by = s_top[2], ty = s_top[3], sw = s_top[4];
            const int nax = bx + tx * (sw ? ay : ax), nsx = tx * (sw ? sy : sx);
            const int nay = by + ty * (sw ? ax : ay), nsy = ty * (sw ? sx : sy);
            ax = nax; sx = nsx; ay = nay; sy = nsy;
            swapped = swapped != (sw != 0);
        }
        const int u0 = swapped ? HIL4_Y[0] : HIL4_X[0], v0 = swapped ? HIL4_X[0] : HIL4_Y[0];
        const int X0 = (ax + sx * u0) & ~(HT - 1), Y0 = (ay + sy * v0) & ~(HT - 1);
        // stage the 64x64 block as one 32-bit word per pixel: thread t expands 16 pixels (three 128-bit loads) of row t/4
        {
            const int r = tid >> 2, c16 = (tid & 3) * 16;
            const uint4 *src = reinterpret_cast<const uint4 *>(rgb + ((size_t)(Y0 + r) * n + X0 + c16) * 3);
            const uint4 a = __ldg(src), b = __ldg(src + 1), c = __ldg(src + 2);
            const uint32_t wd[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w};
            uint32_t *dst = s_px + r * HT_STRIDE + c16;
#pragma unroll
            for (int q = 0; q < 4; q++) {  // 3 words -> 4 pixels
                const uint32_t w0 = wd[3 * q], w1 = wd[3 * q + 1], w2 = wd[3 * q + 2];
                *reinterpret_cast<uint4 *>(dst + 4 * q) =
                    make_uint4(w0 & 0xffffff, __byte_perm(w0, w1, 0x4543) & 0xffffff, __byte_perm(w1, w2, 0x4432) & 0xffffff, w2 >> 8);
            }
        }
        __syncthreads();
        uint32_t pix[16];
#pragma unroll
        for (int j = 0; j < 16; j++) {
            const int u = swapped ? HIL4_Y[j] : HIL4_X[j], v = swapped ? HIL4_X[j] : HIL4_Y[j];
            const int lx = (ax + sx * u) - X0, ly = (ay + sy * v) - Y0;
            pix[j] = s_px[ly * HT_STRIDE + lx];
        }
        if (MODE == 0) {
            uint32_t wd[12];
#pragma unroll
            for (int q = 0; q < 4; q++) {  // 4 pixels -> 3 words
                const uint32_t a = pix[4 * q], b = pix[4 * q + 1], c = pix[4 * q + 2], e = pix[4 * q + 3];
                wd[3 * q] = a | (b << 24);
                wd[3 * q + 1] = (b >> 8) | (c << 16);
                wd[3 * q + 2] = (c >> 16) | (e << 8);
            }
            uint4 *o = reinterpret_cast<uint4 *>(out_rgb + (i0 - blk_begin * 4096) * 3);
            o[0] = make_uint4(wd[0], wd[1], wd[2], wd[3]);
            o[1] = make_uint4(wd[4], wd[5], wd[6], wd[7]);
            o[2] = make_uint4(wd[8], wd[9], wd[10], wd[11]);
            continue;
        }
        // predecessor of this thread's first symbol
        uint32_t prev = __shfl_up_sync(0xffffffffu, pix[15], 1);
        if (lane == 31) s_last[warp] = pix[15];
        __syncthreads();
        if (lane == 0) {
            if (warp > 0) prev = s_last[warp - 1];
            else if (B == 0) prev = 0;  // hilbertc.rs:445 START = [0;3]
            else {
                uint32_t px, py;
                hilbert_d2xy_pow2(n, B - 1, &px, &py);
                const uint8_t *q = rgb + ((size_t)py * n + px) * 3;
                prev = uint32_t(q[0]) | (uint32_t(q[1]) << 8) | (uint32_t(q[2]) << 16);
            }
        }
        if (MODE == 1) {
            // 16-bit SIMD lanes: A = (r, b), G = (g, 0); per-lane wrap-around subtraction gives the i16 differences
            uint32_t wd[24];  // 48 i16 packed two per word
            uint32_t pa = prev & 0x00ff00ffu, pg = (prev >> 8) & 0xffu;
#pragma unroll
            for (int j = 0; j < 16; j += 2) {
                const uint32_t a0 = pix[j] & 0x00ff00ffu, g0 = (pix[j] >> 8) & 0xffu;
                const uint32_t a1 = pix[j + 1] & 0x00ff00ffu, g1 = (pix[j + 1] >> 8) & 0xffu;
                const uint32_t da0 = __vsub2(a0, pa), dg0 = __vsub2(g0, pg), da1 = __vsub2(a1, a0), dg1 = __vsub2(g1, g0);
                wd[3 * (j / 2)] = __byte_perm(da0, dg0, 0x5410);      // dr0, dg0
                wd[3 * (j / 2) + 1] = __byte_perm(da0, da1, 0x5432);  // db0, dr1
                wd[3 * (j / 2) + 2] = __byte_perm(dg1, da1, 0x7610);  // dg1, db1
                pa = a1; pg = g1;
            }
            uint4 *o = reinterpret_cast<uint4 *>(out_delta + (i0 - blk_begin * 4096) * 3);
#pragma unroll
            for (int q = 0; q < 6; q++) o[q] = make_uint4(wd[4 * q], wd[4 * q + 1], wd[4 * q + 2], wd[4 * q + 3]);
        } else {
            // near-zero symbols are counted in shared memory (ATOMS), the rest goes to the global bins directly
#pragma unroll
            for (int j = 0; j < 16; j++) {
                const uint32_t c = pix[j], p = j ? pix[j - 1] : prev;
                const int d0 = int(c & 0xff) - int(p & 0xff), d1 = int((c >> 8) & 0xff) - int((p >> 8) & 0xff),
                          d2 = int((c >> 16) & 0xff) - int((p >> 16) & 0xff);
                if ((unsigned)(d0 + CUBE_R) < (unsigned)CUBE_S && (unsigned)(d1 + CUBE_R) < (unsigned)CUBE_S && (unsigned)(d2 + CUBE_R) < (unsigned)CUBE_S) {
                    const int ci = ((d0 + CUBE_R) * CUBE_S + (d1 + CUBE_R)) * CUBE_S + (d2 + CUBE_R);
                    const int sh = 16 * (ci & 1);
                    const uint32_t old = atomicAdd(&s_cube[ci >> 1], 1u << sh);
                    if (((old >> sh) & 0x7fffu) == 0x7fffu) {  // my increment set the guard bit: 2^15 counts leave the field
                        atomicSub(&s_cube[ci >> 1], 0x8000u << sh);
                        const uint32_t key = uint32_t(((d0 + 255) * 511 + (d1 + 255)) * 511 + (d2 + 255));
                        atomicAdd(&bins[key], 32768u);
                        if (!flags[key >> PAGE_SHIFT]) flags[key >> PAGE_SHIFT] = 1;
                    }
                } else {
                    const uint32_t key = uint32_t(((d0 + 255) * 511 + (d1 + 255)) * 511 + (d2 + 255));
                    atomicAdd(&bins[key], 1u);
                    if (!flags[key >> PAGE_SHIFT]) flags[key >> PAGE_SHIFT] = 1;
                }
            }
        }
    }
    if (MODE == 2) {  // flush the CTA's near-zero counters into the global bins
        __syncthreads();
        for (int ci = tid; ci < CUBE_N; ci += 256) {
            const uint32_t cnt = (s_cube[ci >> 1] >> (16 * (ci & 1))) & 0xffffu;
            if (cnt) {
                const int d2 = ci % CUBE_S - CUBE_R, d1 = (ci / CUBE_S) % CUBE_S - CUBE_R, d0 = ci / (CUBE_S * CUBE_S) - CUBE_R;
                const uint32_t key = uint32_t(((d0 + 255) * 511 + (d1 + 255)) * 511 + (d2 + 255));
                atomicAdd(&bins[key], cnt);
                if (!flags[key >> PAGE_SHIFT]) flags[key >> PAGE_SHIFT] = 1;
            }
        }
    }
}

// ---- the same three stages with the tile brought in by the Tensor Memory Accelerator (default) -------------------------------------
// One elected thread issues cp.async.bulk.tensor.2d for the 64 x 64 pixel block (box = 192 bytes x 64 rows of a 2-D tensor map over
// the packed-RGB image) into one of two shared-memory stages and arms its mbarrier with the 12 288 bytes to come; the CTA is
// persistent, so the tile of block i+1 is in flight while block i is expanded, gathered along the curve, differenced and stored:
// the 64 row segments at a 3*w-byte stride -- what made the plain-load version latency bound (profiles/r02_ncu_full_c5.txt: 43 % of
// HBM, top stalls barrier + long scoreboard) -- are one asynchronous request that no warp waits for.
constexpr uint32_t HT_TILE_BYTES = HT * HT * 3;  // 12 288

template <int MODE>
__global__ void __launch_bounds__(256) hilbert_tile_tma_kernel(const __grid_constant__ CUtensorMap tmap, const uint8_t *__restrict__ rgb, uint32_t n,
                                                               uint8_t *out_rgb, int16_t *out_delta, uint32_t *bins, uint8_t *flags,
                                                               unsigned long long blk_begin, unsigned long long blk_end) {
    // two raw tile stages (packed RGB rows of 192 bytes; TMA destinations: 128-byte aligned), then (MODE 2) the counter cube.  The
    // alignment is asked of the declaration -- aligning the pointer by hand made it a generic address, and every tile load and
    // counter update a generic LD / ATOM instead of LDS / ATOMS
    extern __shared__ __align__(128) uint8_t s_raw[];
    uint32_t *s_cube = reinterpret_cast<uint32_t *>(s_raw + 2 * HT_TILE_BYTES);
    __shared__ __align__(16) uint32_t s_px[HT * HT_STRIDE];  // one word per pixel (r | g<<8 | b<<16)
    __shared__ uint32_t s_last[8];
    __shared__ int s_top[2][5];
    __shared__ __align__(8) unsigned long long s_bar[2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned long long nblocks = blk_end;
    if (MODE == 2) {
        for (int i = tid; i < (CUBE_N + 1) / 2; i += 256) s_cube[i] = 0;
    }
    if (tid == 0) {
        mbar_init(&s_bar[0], 1);
        mbar_init(&s_bar[1], 1);
        mbar_fence_init();
    }
    // this thread's 16 indices (a 4x4 cell) share the base-4 digits 2..5 = its thread id: fold those levels once, for every block
    int lax = 0, lay = 0, lsx = 1, lsy = 1;
    bool lswapped = false;
    {
        uint32_t t = (uint32_t)tid;
#pragma unroll
        for (int sl = 4; sl < HT; sl <<= 1) {
            const uint32_t rx = 1u & (t >> 1), ry = 1u & (t ^ rx);
            if (ry == 0) {
                if (rx == 1) {
                    const int nax = sl - 1 - lay, nsx = -lsy, nay = sl - 1 - lax, nsy = -lsx;
                    lax = nax; lsx = nsx; lay = nay; lsy = nsy;
                } else {
                    const int q = lax, qs = lsx;
                    lax = lay; lsx = lsy; lay = q; lsy = qs;
                }
                lswapped = !lswapped;
            }
            lax += sl * (int)rx;
            lay += sl * (int)ry;
            t >>= 2;
        }
    }
    __syncthreads();
    // thread 0: fold the levels above the block (6..L-1) into an affine map of the 64x64 block, request the block's pixels
    auto issue = [&](unsigned long long blk, int stage) {
        int bx = 0, by = 0, tx = 1, ty = 1, sw = 0;
        unsigned long long t = blk;
        for (uint32_t sft = HT; sft < n; sft <<= 1) {
            const int sl = (int)sft;
            const uint32_t rx = 1u & (uint32_t)(t >> 1), ry = 1u & ((uint32_t)t ^ rx);
            if (ry == 0) {
                if (rx == 1) {
                    const int nbx = sl - 1 - by, ntx = -ty, nby = sl - 1 - bx, nty = -tx;
                    bx = nbx; tx = ntx; by = nby; ty = nty;
                } else {
                    const int q = bx, qs = tx;
                    bx = by; tx = ty; by = q; ty = qs;
                }
                sw ^= 1;
            }
            bx += sl * (int)rx;
            by += sl * (int)ry;
            t >>= 2;
        }
        s_top[stage][0] = bx; s_top[stage][1] = tx; s_top[stage][2] = by; s_top[stage][3] = ty; s_top[stage][4] = sw;
        mbar_arrive_expect_tx(&s_bar[stage], HT_TILE_BYTES);
        tma_load_2d(s_raw + (size_t)stage * HT_TILE_BYTES, &tmap, (bx & ~(HT - 1)) * 3, by & ~(HT - 1), &s_bar[stage]);
    };
    const unsigned long long first = blk_begin + blockIdx.x;
    if (tid == 0 && first < nblocks) issue(first, 0);
    uint32_t it = 0;
    for (unsigned long long blk = first; blk < nblocks; blk += gridDim.x, it++) {
        const int stage = it & 1;
        const unsigned long long B = blk * 4096;
        const unsigned long long i0 = B + (unsigned long long)tid * 16;
        // the other stage was consumed before the barrier that ended the previous iteration's expansion: refill it now
        if (tid == 0 && blk + gridDim.x < nblocks) issue(blk + gridDim.x, stage ^ 1);
        mbar_wait(&s_bar[stage], (it >> 1) & 1);
        int ax, ay, sx, sy;
        bool swapped;
        {   // compose: (x, y) = top(local(u, v)); x = bx + tx * (sw ? yl : xl), y = by + ty * (sw ? xl : yl)
            const int bx = s_top[stage][0], tx = s_top[stage][1], by = s_top[stage][2], ty = s_top[stage][3], sw = s_top[stage][4];
            ax = bx + tx * (sw ? lay : lax); sx = tx * (sw ? lsy : lsx);
            ay = by + ty * (sw ? lax : lay); sy = ty * (sw ? lsx : lsy);
            swapped = lswapped != (sw != 0);
        }
        const int X0 = ax & ~(HT - 1), Y0 = ay & ~(HT - 1);  // (every coordinate of the block shares the bits above the low six)
        {   // expand the landed tile to one 32-bit word per pixel: thread t takes 16 pixels (three 128-bit words) of row t/4
            const int r = tid >> 2, c16 = (tid & 3) * 16;
            const uint4 *src = reinterpret_cast<const uint4 *>(s_raw + (size_t)stage * HT_TILE_BYTES + r * (HT * 3) + c16 * 3);
            const uint4 a = src[0], b = src[1], c = src[2];
            const uint32_t wd[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w};
            uint32_t *dst = s_px + r * HT_STRIDE + c16;
#pragma unroll
            for (int q = 0; q < 4; q++) {  // 3 words -> 4 pixels
                const uint32_t w0 = wd[3 * q], w1 = wd[3 * q + 1], w2 = wd[3 * q + 2];
                *reinterpret_cast<uint4 *>(dst + 4 * q) =
                    make_uint4(w0 & 0xffffff, __byte_perm(w0, w1, 0x4543) & 0xffffff, __byte_perm(w1, w2, 0x4432) & 0xffffff, w2 >> 8);
            }
        }
        __syncthreads();  // s_px complete; the raw stage is free again
        uint32_t pix[16];
#pragma unroll
        for (int j = 0; j < 16; j++) {
            const int u = swapped ? HIL4_Y[j] : HIL4_X[j], v = swapped ? HIL4_X[j] : HIL4_Y[j];
            const int lx = (ax + sx * u) - X0, ly = (ay + sy * v) - Y0;
            pix[j] = s_px[ly * HT_STRIDE + lx];
        }
        if (MODE == 0) {
            uint32_t wd[12];
#pragma unroll
            for (int q = 0; q < 4; q++) {  // 4 pixels -> 3 words
                const uint32_t a = pix[4 * q], b = pix[4 * q + 1], c = pix[4 * q + 2], e = pix[4 * q + 3];
                wd[3 * q] = a | (b << 24);
                wd[3 * q + 1] = (b >> 8) | (c << 16);
                wd[3 * q + 2] = (c >> 16) | (e << 8);
            }
            uint4 *o = reinterpret_cast<uint4 *>(out_rgb + (i0 - blk_begin * 4096) * 3);
            o[0] = make_uint4(wd[0], wd[1], wd[2], wd[3]);
            o[1] = make_uint4(wd[4], wd[5], wd[6], wd[7]);
            o[2] = make_uint4(wd[8], wd[9], wd[10], wd[11]);
            __syncthreads();  // before the next block's expansion overwrites s_px
            continue;
        }
        // predecessor of this thread's first symbol
        uint32_t prev = __shfl_up_sync(0xffffffffu, pix[15], 1);
        if (lane == 31) s_last[warp] = pix[15];
        __syncthreads();  // s_last complete; every thread has gathered its pixels, so s_px may be overwritten after this point
        if (lane == 0) {
            if (warp > 0) prev = s_last[warp - 1];
            else if (B == 0) prev = 0;  // hilbertc.rs:445 START = [0;3]
            else {
                uint32_t px, py;
                hilbert_d2xy_pow2(n, B - 1, &px, &py);
                const uint8_t *q = rgb + ((size_t)py * n + px) * 3;
                prev = uint32_t(q[0]) | (uint32_t(q[1]) << 8) | (uint32_t(q[2]) << 16);
            }
        }
        if (MODE == 1) {
            // 16-bit SIMD lanes: A = (r, b), G = (g, 0); per-lane wrap-around subtraction gives the i16 differences
            uint32_t wd[24];  // 48 i16 packed two per word
            uint32_t pa = prev & 0x00ff00ffu, pg = (prev >> 8) & 0xffu;
#pragma unroll
            for (int j = 0; j < 16; j += 2) {
                const uint32_t a0 = pix[j] & 0x00ff00ffu, g0 = (pix[j] >> 8) & 0xffu;
                const uint32_t a1 = pix[j + 1] & 0x00ff00ffu, g1 = (pix[j + 1] >> 8) & 0xffu;
                const uint32_t da0 = __vsub2(a0, pa), dg0 = __vsub2(g0, pg), da1 = __vsub2(a1, a0), dg1 = __vsub2(g1, g0);
                wd[3 * (j / 2)] = __byte_perm(da0, dg0, 0x5410);      // dr0, dg0
                wd[3 * (j / 2) + 1] = __byte_perm(da0, da1, 0x5432);  // db0, dr1
                wd[3 * (j / 2) + 2] = __byte_perm(dg1, da1, 0x7610);  // dg1, db1
                pa = a1; pg = g1;
            }
            uint4 *o = reinterpret_cast<uint4 *>(out_delta + (i0 - blk_begin * 4096) * 3);
#pragma unroll
            for (int q = 0; q < 6; q++) o[q] = make_uint4(wd[4 * q], wd[4 * q + 1], wd[4 * q + 2], wd[4 * q + 3]);
        } else {
            // near-zero symbols are counted in shared memory (ATOMS), the rest goes to the global bins directly.  The 16 shared
            // atomics of a thread are ISSUED back to back and their return values (the guard-bit check) examined afterwards: checking
            // each one right behind its atomic made every warp wait out the atomic's latency 16 times per block, with 16 warps per SM
            uint32_t old_v[16], slot[16];  // slot = cube index (bit 31 set: outside the cube, already counted globally)
#pragma unroll
            for (int j = 0; j < 16; j++) {
                const uint32_t c = pix[j], p = j ? pix[j - 1] : prev;
                const int d0 = int(c & 0xff) - int(p & 0xff), d1 = int((c >> 8) & 0xff) - int((p >> 8) & 0xff),
                          d2 = int((c >> 16) & 0xff) - int((p >> 16) & 0xff);
                old_v[j] = 0;
                if ((unsigned)(d0 + CUBE_R) < (unsigned)CUBE_S && (unsigned)(d1 + CUBE_R) < (unsigned)CUBE_S && (unsigned)(d2 + CUBE_R) < (unsigned)CUBE_S) {
                    const uint32_t ci = uint32_t(((d0 + CUBE_R) * CUBE_S + (d1 + CUBE_R)) * CUBE_S + (d2 + CUBE_R));
                    slot[j] = ci;
                    old_v[j] = atomicAdd(&s_cube[ci >> 1], 1u << (16 * (ci & 1)));
                } else {
                    slot[j] = 0x80000000u;
                    const uint32_t key = uint32_t(((d0 + 255) * 511 + (d1 + 255)) * 511 + (d2 + 255));
                    atomicAdd(&bins[key], 1u);
                    if (!flags[key >> PAGE_SHIFT]) flags[key >> PAGE_SHIFT] = 1;
                }
            }
#pragma unroll
            for (int j = 0; j < 16; j++) {
                const uint32_t ci = slot[j];
                const int sh = 16 * (ci & 1);
                if (!(ci >> 31) && ((old_v[j] >> sh) & 0x7fffu) == 0x7fffu) {  // my increment set the guard bit: 2^15 counts leave the field
                    atomicSub(&s_cube[ci >> 1], 0x8000u << sh);
                    const int d2 = int(ci % CUBE_S) - CUBE_R, d1 = int((ci / CUBE_S) % CUBE_S) - CUBE_R, d0 = int(ci / (CUBE_S * CUBE_S)) - CUBE_R;
                    const uint32_t key = uint32_t(((d0 + 255) * 511 + (d1 + 255)) * 511 + (d2 + 255));
                    atomicAdd(&bins[key], 32768u);
                    if (!flags[key >> PAGE_SHIFT]) flags[key >> PAGE_SHIFT] = 1;
                }
            }
        }
    }
    if (MODE == 2) {  // flush the CTA's near-zero counters into the global bins
        __syncthreads();
        for (int ci = tid; ci < CUBE_N; ci += 256) {
            const uint32_t cnt = (s_cube[ci >> 1] >> (16 * (ci & 1))) & 0xffffu;
            if (cnt) {
                const int d2 = ci % CUBE_S - CUBE_R, d1 = (ci / CUBE_S) % CUBE_S - CUBE_R, d0 = ci / (CUBE_S * CUBE_S) - CUBE_R;
                const uint32_t key = uint32_t(((d0 + 255) * 511 + (d1 + 255)) * 511 + (d2 + 255));
                atomicAdd(&bins[key], cnt);
                if (!flags[key >> PAGE_SHIFT]) flags[key >> PAGE_SHIFT] = 1;
            }
        }
    }
}

// ---- second version of the TMA tile stages: no expansion pass, no CTA barrier ------------------------------------------------------
// A thread's 16 curve indices are one 4 x 4 pixel cell of the block, i.e. four 12-byte row segments of the landed tile: it reads them
// as 12 words straight from the raw stage, unpacks the 16 pixels in registers, and brings them into curve order with two
// conditional register permutations -- on the Hilbert curve a cell is the base motif either as it is, rotated by 180 degrees
// (negative step), transposed (swapped axes), or both; sx == sy always (the fold below only ever transposes or anti-transposes).
// What the first version spent on the word-per-pixel copy of the tile (3 LDS.128 + 4 STS.128 + 16 scattered LDS per thread, 16.6 KB
// of shared memory, two CTA barriers per block -- its top stall, profiles/r02_ncu_full_c5_tma.txt) is gone: the predecessor of a
// warp's first symbol is read from the raw tile too (its position in the block is a per-thread constant), and the stages are handed
// back through a second pair of mbarriers (256 arrivals) that only the refilling thread waits on, so warps drift freely.
// A 15-bit counter of the shared-memory cube reached 2^15 (its guard bit is set by the increment that got there): move 2^15 counts
// of that symbol to the global bins.  Out of line on purpose -- it runs once per 32 768 equal symbols, and sixteen inlined copies of
// its divisions made up a third of the kernel's code.
template <int CR>
__device__ __noinline__ void cube_spill(uint32_t *s_cube, uint32_t ci, uint32_t *bins, uint8_t *flags) {
    constexpr int CS = 2 * CR + 1;
    atomicSub(&s_cube[ci >> 1], 0x8000u << (16 * (ci & 1)));
    const int d2 = int(ci % CS) - CR, d1 = int((ci / CS) % CS) - CR, d0 = int(ci / (CS * CS)) - CR;
    const uint32_t key = uint32_t(((d0 + 255) * 511 + (d1 + 255)) * 511 + (d2 + 255));
    atomicAdd(&bins[key], 32768u);
    if (!flags[key >> PAGE_SHIFT]) flags[key >> PAGE_SHIFT] = 1;
}

template <int MODE, int NB, int CR, int NS>  // CR: radius of the shared-memory counter cube (MODE 2); NS: tile stages
__global__ void __launch_bounds__(256, NB) hilbert_tile_tma2_kernel(const __grid_constant__ CUtensorMap tmap, const uint8_t *__restrict__ rgb,
                                                                                   uint32_t n, uint8_t *out_rgb, int16_t *out_delta, uint32_t *bins,
                                                                                   uint8_t *flags, unsigned long long blk_begin, unsigned long long blk_end) {
    // two raw tile stages (packed RGB rows of 192 bytes; TMA destinations: 128-byte aligned), then (MODE 2) the counter cube.  The
    // alignment is asked of the declaration -- aligning the pointer by hand made it a generic address, and every tile load and
    // counter update a generic LD / ATOM instead of LDS / ATOMS
    extern __shared__ __align__(128) uint8_t s_raw[];
    uint32_t *s_cube = reinterpret_cast<uint32_t *>(s_raw + NS * HT_TILE_BYTES);
    __shared__ int s_top[NS][4];  // per stage: block origin (bx, by), step sign, swapped axes
    // MODE 2: which pages of the global bins this CTA touched, one bit per 8 pages (511^3 / 4096 / 8 < 4096 bits).  The page flags
    // tell the compaction where to look; marking them from the counting path cost either a dependent global load per far symbol
    // (test, then set) or -- as plain stores -- millions of writes to a few hot lines (0.75 -> 1.1 ms for the whole call).  Bits
    // are set here with ATOMS.OR (no return value, nothing to wait for) and written out once, when the CTA is done.
    __shared__ uint32_t s_dirty[128];
    __shared__ __align__(8) unsigned long long s_full[NS], s_empty[NS];
    __shared__ uint32_t s_prev[NS];  // per stage: the pixel in front of the block's first one (the last pixel of the previous block)
    constexpr int CS = 2 * CR + 1, CN = CS * CS * CS;
    const int tid = threadIdx.x, lane = tid & 31;
    const unsigned long long nblocks = blk_end;
    if (MODE == 2) {
        for (int i = tid; i < (CN + 1) / 2; i += 256) s_cube[i] = 0;
        if (tid < 128) s_dirty[tid] = 0;
    }
    if (tid == 0) {
        for (int i = 0; i < NS; i++) {
            mbar_init(&s_full[i], MODE == 0 ? 1 : 2);  // the tile's bytes (+ one arrival), and the arrival that publishes s_prev
            mbar_init(&s_empty[i], 256);
        }
        mbar_fence_init();
    }
    // base-4 digits 2..5 of a cell's first index = the cell number: fold those levels once, for every block
    auto fold_cell = [](uint32_t t, int &cx, int &cy, int &cs, bool &csw) {
        cx = 0; cy = 0; cs = 1; csw = false;
#pragma unroll
        for (int sl = 4; sl < HT; sl <<= 1) {
            const uint32_t rx = 1u & (t >> 1), ry = 1u & (t ^ rx);
            if (ry == 0) {
                if (rx == 1) {
                    const int nx = sl - 1 - cy, ny = sl - 1 - cx;
                    cx = nx; cy = ny; cs = -cs;
                } else {
                    const int q = cx;
                    cx = cy; cy = q;
                }
                csw = !csw;
            }
            cx += sl * (int)rx;
            cy += sl * (int)ry;
            t >>= 2;
        }
    };
    int lax, lay, ls, plx = 0, ply = 0;
    bool lswapped;
    fold_cell((uint32_t)tid, lax, lay, ls, lswapped);
    if (lane == 0 && tid > 0) {  // where the LAST pixel of the previous cell lies (digit 15 of the motif is (3, 0))
        int qx, qy, qs;
        bool qsw;
        fold_cell((uint32_t)tid - 1, qx, qy, qs, qsw);
        plx = qx + qs * (qsw ? 0 : 3);
        ply = qy + qs * (qsw ? 3 : 0);
    }
    __syncthreads();
    // One thread per block: fold the levels above the block (6..L-1) into an affine map of the 64x64 block, request the block's
    // pixels, fetch the pixel in front of its first one.  That is ~250 dependent instructions and a global load -- a fifth of what a
    // warp spends on a block -- and a block is only done when its slowest warp is, so the duty ROTATES over the eight warps
    // (with thread 0 doing all of it, the other warps spent 22 % of the kernel's issue slots polling the tile barrier:
    // profiles/r02_ncu_full_c5_tma2.txt).
    auto issue = [&](unsigned long long blk, int stage) {
        int bx = 0, by = 0, ts = 1, sw = 0;
        unsigned long long t = blk;
        for (uint32_t sft = HT; sft < n; sft <<= 1) {
            const int sl = (int)sft;
            const uint32_t rx = 1u & (uint32_t)(t >> 1), ry = 1u & ((uint32_t)t ^ rx);
            if (ry == 0) {
                if (rx == 1) {
                    const int nbx = sl - 1 - by, nby = sl - 1 - bx;
                    bx = nbx; by = nby; ts = -ts;
                } else {
                    const int q = bx;
                    bx = by; by = q;
                }
                sw ^= 1;
            }
            bx += sl * (int)rx;
            by += sl * (int)ry;
            t >>= 2;
        }
        s_top[stage][0] = bx; s_top[stage][1] = by; s_top[stage][2] = ts; s_top[stage][3] = sw;
        mbar_arrive_expect_tx(&s_full[stage], HT_TILE_BYTES);
        tma_load_2d(s_raw + (size_t)stage * HT_TILE_BYTES, &tmap, (bx & ~(HT - 1)) * 3, by & ~(HT - 1), &s_full[stage]);
        if (MODE != 0) {
            uint32_t pv = 0;  // hilbertc.rs:445 START = [0;3] in front of the very first pixel
            if (blk != 0) {
                uint32_t px, py;
                hilbert_d2xy_pow2(n, blk * 4096 - 1, &px, &py);
                const uint8_t *q = rgb + ((size_t)py * n + px) * 3;
                pv = uint32_t(q[0]) | (uint32_t(q[1]) << 8) | (uint32_t(q[2]) << 16);
            }
            s_prev[stage] = pv;
            mbar_arrive(&s_full[stage]);  // (release: s_prev and s_top are visible to whoever sees the phase complete)
        }
    };
    // NS - 1 blocks are requested ahead of the one being worked on, each as soon as every thread has left the stage it goes to
    const unsigned long long first = blk_begin + blockIdx.x;
    if (tid == 0)
        for (int i = 0; i < NS - 1; i++)
            if (first + (unsigned long long)i * gridDim.x < nblocks) issue(first + (unsigned long long)i * gridDim.x, i);
    uint32_t it = 0;
    int stage = 0, pstage = NS - 1;  // stage of block `it`; stage of block it - 1 = the one block it + NS - 1 goes to
    uint32_t phase = 0, pphase = 0;  // parities of those stages' current uses
    for (unsigned long long blk = first; blk < nblocks; blk += gridDim.x, it++) {
        if (tid == int(it & 7) * 32 && blk + (unsigned long long)(NS - 1) * gridDim.x < nblocks) {  // refill the stage block it - 1 was read from
            if (it > 0) mbar_wait(&s_empty[pstage], pphase);  // ... once every thread has left it
            issue(blk + (unsigned long long)(NS - 1) * gridDim.x, pstage);
        }
        const unsigned long long B = blk * 4096;
        const unsigned long long i0 = B + (unsigned long long)tid * 16;
        mbar_wait(&s_full[stage], phase);
        uint32_t prev = MODE != 0 && tid == 0 ? s_prev[stage] : 0u;
        const uint8_t *tile = s_raw + (size_t)stage * HT_TILE_BYTES;
        const int bx = s_top[stage][0], by = s_top[stage][1], ts = s_top[stage][2], sw = s_top[stage][3];
        // compose: (x, y) = top(local(u, v)); x = bx + ts * (sw ? yl : xl), y = by + ts * (sw ? xl : yl)
        const int ax = bx + ts * (sw ? lay : lax), ay = by + ts * (sw ? lax : lay);
        const bool neg = ts * ls < 0, swapped = lswapped != (sw != 0);
        const int mx = (neg ? ax - 3 : ax) & (HT - 1), my = (neg ? ay - 3 : ay) & (HT - 1);  // the cell's low corner inside the block
        uint32_t p[16];
        {
            const uint32_t *src = reinterpret_cast<const uint32_t *>(tile + my * (HT * 3) + mx * 3);
#pragma unroll
            for (int v = 0; v < 4; v++) {  // 3 words -> 4 pixels (r | g<<8 | b<<16)
                const uint32_t w0 = src[v * (HT * 3 / 4)], w1 = src[v * (HT * 3 / 4) + 1], w2 = src[v * (HT * 3 / 4) + 2];
                p[4 * v] = w0 & 0xffffff;
                p[4 * v + 1] = __byte_perm(w0, w1, 0x4543) & 0xffffff;
                p[4 * v + 2] = __byte_perm(w1, w2, 0x4432) & 0xffffff;
                p[4 * v + 3] = w2 >> 8;
            }
        }
        if (MODE != 0 && lane == 0 && tid > 0) {
            const int qx = (bx + ts * (sw ? ply : plx)) & (HT - 1), qy = (by + ts * (sw ? plx : ply)) & (HT - 1);
            const uint8_t *q = tile + qy * (HT * 3) + qx * 3;
            prev = uint32_t(q[0]) | (uint32_t(q[1]) << 8) | (uint32_t(q[2]) << 16);
        }
        mbar_arrive(&s_empty[stage]);  // this thread is done with the stage
        pstage = stage; pphase = phase;
        if (++stage == NS) { stage = 0; phase ^= 1; }
        // bring the cell into curve order: rotate by 180 degrees (negative step), then transpose (swapped axes), then the base motif
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const uint32_t a = p[i], b = p[15 - i];
            p[i] = neg ? b : a;
            p[15 - i] = neg ? a : b;
        }
#pragma unroll
        for (int a = 0; a < 4; a++)
#pragma unroll
            for (int b = a + 1; b < 4; b++) {
                const uint32_t u = p[4 * a + b], v = p[4 * b + a];
                p[4 * a + b] = swapped ? v : u;
                p[4 * b + a] = swapped ? u : v;
            }
        constexpr int ORD[16] = {0, 1, 5, 4, 8, 12, 13, 9, 10, 14, 15, 11, 7, 6, 2, 3};  // 4 * HIL4_Y[j] + HIL4_X[j]
        uint32_t pix[16];
#pragma unroll
        for (int j = 0; j < 16; j++) pix[j] = p[ORD[j]];
        if (MODE == 0) {
            uint32_t wd[12];
#pragma unroll
            for (int q = 0; q < 4; q++) {  // 4 pixels -> 3 words
                const uint32_t a = pix[4 * q], b = pix[4 * q + 1], c = pix[4 * q + 2], e = pix[4 * q + 3];
                wd[3 * q] = a | (b << 24);
                wd[3 * q + 1] = (b >> 8) | (c << 16);
                wd[3 * q + 2] = (c >> 16) | (e << 8);
            }
            uint4 *o = reinterpret_cast<uint4 *>(out_rgb + (i0 - blk_begin * 4096) * 3);
            o[0] = make_uint4(wd[0], wd[1], wd[2], wd[3]);
            o[1] = make_uint4(wd[4], wd[5], wd[6], wd[7]);
            o[2] = make_uint4(wd[8], wd[9], wd[10], wd[11]);
            continue;
        }
        {   // predecessor of this thread's first symbol: the neighbour lane's last pixel (lane 0 read it from the tile / from global)
            const uint32_t up = __shfl_up_sync(0xffffffffu, pix[15], 1);
            if (lane != 0) prev = up;
        }
        if (MODE == 1) {
            // 16-bit SIMD lanes: A = (r, b), G = (g, 0); per-lane wrap-around subtraction gives the i16 differences
            uint32_t wd[24];  // 48 i16 packed two per word
            uint32_t pa = prev & 0x00ff00ffu, pg = (prev >> 8) & 0xffu;
#pragma unroll
            for (int j = 0; j < 16; j += 2) {
                const uint32_t a0 = pix[j] & 0x00ff00ffu, g0 = (pix[j] >> 8) & 0xffu;
                const uint32_t a1 = pix[j + 1] & 0x00ff00ffu, g1 = (pix[j + 1] >> 8) & 0xffu;
                const uint32_t da0 = __vsub2(a0, pa), dg0 = __vsub2(g0, pg), da1 = __vsub2(a1, a0), dg1 = __vsub2(g1, g0);
                wd[3 * (j / 2)] = __byte_perm(da0, dg0, 0x5410);      // dr0, dg0
                wd[3 * (j / 2) + 1] = __byte_perm(da0, da1, 0x5432);  // db0, dr1
                wd[3 * (j / 2) + 2] = __byte_perm(dg1, da1, 0x7610);  // dg1, db1
                pa = a1; pg = g1;
            }
            uint4 *o = reinterpret_cast<uint4 *>(out_delta + (i0 - blk_begin * 4096) * 3);
#pragma unroll
            for (int q = 0; q < 6; q++) o[q] = make_uint4(wd[4 * q], wd[4 * q + 1], wd[4 * q + 2], wd[4 * q + 3]);
        } else {
            // near-zero symbols are counted in shared memory (ATOMS), the rest goes to the global bins directly; the shared atomics
            // are issued back to back, their return values (the guard-bit check) examined afterwards.  Per symbol: one VABSDIFF4 and
            // three logic operations decide whether all three |differences| are <= CR; the cube index is linear in the channels,
            // ((d0 + CR) * CS + (d1 + CR)) * CS + (d2 + CR) = (A(c) - A(p)) * CS + (b(c) - b(p)) + K with A(x) = CS * r + g (one
            // IDP.4A per pixel), so no channel is extracted on the common path.  The global key of a far symbol is linear in the same
            // way: ((d0 + 255) * 511 + (d1 + 255)) * 511 + (d2 + 255) = (W(c) - W(p)) * 511 + (b(c) - b(p)) + K' with W(x) = 511 r + g
            // = IDP.4A(x, (255, 1)) + (r << 8) -- three instructions per pixel instead of a dozen per far symbol, which matters
            // because on a noisy image nine warp-instructions in ten have at least one far lane and run that path.
            constexpr uint32_t SPLAT_LIM = (127u - CR) * 0x01010101u;
            constexpr int K_CUBE = (CR * CS + CR) * CS + CR;
            constexpr int K_FAR = (255 * 511 + 255) * 511 + 255;
            uint32_t old_v[16], slot[16];  // slot = cube index (bit 31 set: outside the cube, already counted globally)
            int pa = dp4a_uu(prev, uint32_t(CS) | (1u << 8), 0), pb = int(prev >> 16);
            int pw = dp4a_uu(prev, 255u | (1u << 8), 0) + int(__byte_perm(prev, 0u, 0x4404));
#pragma unroll
            for (int j = 0; j < 16; j++) {
                const uint32_t c = pix[j], q = j ? pix[j - 1] : prev;
                const int ca = dp4a_uu(c, uint32_t(CS) | (1u << 8), 0), cb = int(c >> 16);
                const int cw = dp4a_uu(c, 255u | (1u << 8), 0) + int(__byte_perm(c, 0u, 0x4404));
                const uint32_t ad = __vabsdiffu4(c, q);  // |difference| per channel
                old_v[j] = 0;
                if (((((ad & 0x7f7f7f7fu) + SPLAT_LIM) | ad) & 0x80808080u) == 0) {  // no channel differs by more than CR
                    const uint32_t ci = uint32_t((ca - pa) * CS + (cb - pb + K_CUBE));
                    slot[j] = ci;
                    old_v[j] = atomicAdd(&s_cube[ci >> 1], 1u << (16 * (ci & 1)));
                } else {
                    slot[j] = 0x80000000u;
                    const uint32_t key = uint32_t((cw - pw) * 511 + (cb - pb + K_FAR));
                    atomicAdd(&bins[key], 1u);
                    atomicOr(&s_dirty[key >> (PAGE_SHIFT + 8)], 1u << ((key >> (PAGE_SHIFT + 3)) & 31));
                }
                pa = ca; pb = cb; pw = cw;
            }
            // guard bits: a counter that read 0x7fff before my increment is the one I pushed to 2^15 (far symbols read 0).  One OR
            // over the sixteen answers decides whether anything has to be looked at
            uint32_t any_full = 0;
#pragma unroll
            for (int j = 0; j < 16; j++) any_full |= ((old_v[j] >> (16 * (slot[j] & 1))) & 0x7fffu) + 1u;
            if (any_full & 0x8000u) {
#pragma unroll
                for (int j = 0; j < 16; j++) {
                    const uint32_t ci = slot[j];
                    const int sh = 16 * (ci & 1);
                    if (!(ci >> 31) && ((old_v[j] >> sh) & 0x7fffu) == 0x7fffu)  // my increment set the guard bit: 2^15 counts leave the field
                        cube_spill<CR>(s_cube, ci, bins, flags);
                }
            }
        }
    }
    if (MODE == 2) {  // flush the CTA's near-zero counters into the global bins, then its dirty-page bits into the page flags
        __syncthreads();
        // (the three digits of ci in base CS are carried along instead of divided out: ci advances by 256 = Q * CS + Rm)
        constexpr int Q = 256 / CS, Rm = 256 % CS;
        static_assert(Q + 1 < CS, "one carry per digit");
        int e2 = tid % CS, e1 = (tid / CS) % CS, e0 = tid / (CS * CS);
        for (int ci = tid; ci < CN; ci += 256) {
            const uint32_t cnt = (s_cube[ci >> 1] >> (16 * (ci & 1))) & 0xffffu;
            if (cnt) {
                const uint32_t key = uint32_t(((e0 - CR + 255) * 511 + (e1 - CR + 255)) * 511 + (e2 - CR + 255));
                atomicAdd(&bins[key], cnt);
                atomicOr(&s_dirty[key >> (PAGE_SHIFT + 8)], 1u << ((key >> (PAGE_SHIFT + 3)) & 31));
            }
            e2 += Rm; e1 += Q;
            if (e2 >= CS) { e2 -= CS; e1++; }
            if (e1 >= CS) { e1 -= CS; e0++; }
        }
        __syncthreads();
        constexpr uint32_t NPAGES = (511u * 511u * 511u + PAGE - 1) / PAGE;
        for (uint32_t c = tid; c < 128 * 32; c += 256)
            if ((s_dirty[c >> 5] >> (c & 31)) & 1u)
                for (uint32_t pg = c * 8; pg < c * 8 + 8 && pg < NPAGES; pg++)
                    if (!flags[pg]) flags[pg] = 1;  // (a page of the group that holds no symbol costs the compaction one scan of zeros)
    }
}

// inverse: segmented prefix sum along the curve is sequential per channel; done as a 3-kernel scan over i16 diffs.
// Arithmetic is modulo 2^32 (unsigned): FromDiff (hilbertc.rs:482-509) panics at the FIRST reconstructed channel outside 0..255
// (`try_into().unwrap()`), every prefix before that one is in 0..255, so the wrapped value at the first offender equals the true
// value and is detected whatever a damaged stream does afterwards (err_flag -> CNIIC_ERR_DECODE).
__global__ void __launch_bounds__(256) undelta_partial_kernel(const int16_t *__restrict__ diff, unsigned long long n, uint32_t *block_sums) {
    const unsigned long long base = (unsigned long long)blockIdx.x * 4096;
    uint32_t s0 = 0, s1 = 0, s2 = 0;
    for (int j = 0; j < 16; j++) {
        const unsigned long long i = base + (unsigned long long)j * 256 + threadIdx.x;
        if (i < n) { s0 += uint32_t(int(diff[3 * i])); s1 += uint32_t(int(diff[3 * i + 1])); s2 += uint32_t(int(diff[3 * i + 2])); }
    }
    for (int o = 16; o > 0; o >>= 1) {
        s0 += __shfl_xor_sync(0xffffffffu, s0, o); s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    __shared__ uint32_t s[8][3];
    if ((threadIdx.x & 31) == 0) { s[threadIdx.x >> 5][0] = s0; s[threadIdx.x >> 5][1] = s1; s[threadIdx.x >> 5][2] = s2; }
    __syncthreads();
    if (threadIdx.x < 3) {
        uint32_t t = 0;
        for (int i = 0; i < 8; i++) t += s[i][threadIdx.x];
        block_sums[3 * blockIdx.x + threadIdx.x] = t;
    }
}

__global__ void undelta_scan_blocks_kernel(uint32_t *block_sums, size_t nblocks) {
    // one thread per channel: nblocks <= N/4096 (16K for 8192^2) -- sequential is fine
    if (threadIdx.x < 3) {
        uint32_t run = 0;
        for (size_t b = 0; b < nblocks; b++) {
            const uint32_t v = block_sums[3 * b + threadIdx.x];
            block_sums[3 * b + threadIdx.x] = run;
            run += v;
        }
    }
}

__global__ void __launch_bounds__(256) undelta_apply_kernel(const int16_t *__restrict__ diff, unsigned long long n, const uint32_t *__restrict__ block_sums,
                                                            uint32_t w, uint32_t h, bool pow2, uint8_t *out, uint32_t *err_flag) {
    // each thread owns 16 consecutive stream positions of the 4096-position block
    const unsigned long long base = (unsigned long long)blockIdx.x * 4096 + (unsigned long long)threadIdx.x * 16;
    uint32_t l0 = 0, l1 = 0, l2 = 0;
    for (int j = 0; j < 16; j++) {
        const unsigned long long i = base + j;
        if (i < n) { l0 += uint32_t(int(diff[3 * i])); l1 += uint32_t(int(diff[3 * i + 1])); l2 += uint32_t(int(diff[3 * i + 2])); }
    }
    // exclusive scan of the per-thread sums across the block
    __shared__ uint32_t s_w[8][3];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t x0 = l0, x1 = l1, x2 = l2;
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y0 = __shfl_up_sync(0xffffffffu, x0, o), y1 = __shfl_up_sync(0xffffffffu, x1, o), y2 = __shfl_up_sync(0xffffffffu, x2, o);
        if (lane >= o) { x0 += y0; x1 += y1; x2 += y2; }
    }
    if (lane == 31) { s_w[warp][0] = x0; s_w[warp][1] = x1; s_w[warp][2] = x2; }
    __syncthreads();
    uint32_t b0 = block_sums[3 * blockIdx.x], b1 = block_sums[3 * blockIdx.x + 1], b2 = block_sums[3 * blockIdx.x + 2];
    for (int j = 0; j < warp; j++) { b0 += s_w[j][0]; b1 += s_w[j][1]; b2 += s_w[j][2]; }
    uint32_t r0 = b0 + x0 - l0, r1 = b1 + x1 - l1, r2 = b2 + x2 - l2;
    bool bad = false;
    for (int j = 0; j < 16; j++) {
        const unsigned long long i = base + j;
        if (i < n) {
            r0 += uint32_t(int(diff[3 * i])); r1 += uint32_t(int(diff[3 * i + 1])); r2 += uint32_t(int(diff[3 * i + 2]));
            bad |= (r0 | r1 | r2) > 255u;  // hilbertc.rs:503-506: Rgb<u8>::try_from(SignedColor).unwrap()
            uint32_t x, y;
            hilbert_d2xy(w, h, pow2, i, &x, &y);
            uint8_t *p = out + ((size_t)y * w + x) * 3;
            p[0] = (uint8_t)r0; p[1] = (uint8_t)r1; p[2] = (uint8_t)r2;
        }
    }
    if (bad) *err_flag = 1u;
}

// ============================================================================================================
// exact integer SSE (reference src/bench.rs:95-104 sums sqrt(n)^2 in f64; the integer sum is the exact value)
// ============================================================================================================
__global__ void sse_kernel(const uint8_t *__restrict__ a, const uint8_t *__restrict__ b, size_t nbytes, unsigned long long *out) {
    unsigned long long s = 0;
    const size_t nw = nbytes / 4;
    const bool al = ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 3) == 0;
    const size_t stride = (size_t)gridDim.x * blockDim.x, t0 = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (al) {
        for (size_t i = t0; i < nw; i += stride) {
            const uint32_t x = reinterpret_cast<const uint32_t *>(a)[i], y = reinterpret_cast<const uint32_t *>(b)[i];
            const uint32_t ad = __vabsdiffu4(x, y);
            s += __dp4a(ad, ad, 0u);
        }
        for (size_t i = nw * 4 + t0; i < nbytes; i += stride) { const int d = int(a[i]) - int(b[i]); s += d * d; }
    } else {
        for (size_t i = t0; i < nbytes; i += stride) { const int d = int(a[i]) - int(b[i]); s += d * d; }
    }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0 && s) atomicAdd(out, s);
}

// ============================================================================================================
// Huffman payload packing on the device (reference src/huf.rs:36-41 + src/bit.rs:209-253: MSB-first, zero padded)
//   symbol stream -> (code, length) by binary search in the ascending key table -> exclusive scan of the lengths
//   -> every thread assembles the bits of its 16 symbols in registers and stores whole 32-bit words; only the first and
//   last word of a thread can be shared with a neighbour and are merged with atomicOr.
// ============================================================================================================
template <int SRC>  // 0: packed RGB8 pixels in raster order, 1: i16 x 3 delta symbols
__device__ __forceinline__ uint32_t pack_key(const void *src, size_t i) {
    if (SRC == 0) {
        const uint8_t *p = static_cast<const uint8_t *>(src) + 3 * i;
        return (uint32_t(p[0]) << 16) | (uint32_t(p[1]) << 8) | p[2];
    }
    const int16_t *d = static_cast<const int16_t *>(src) + 3 * i;
    return uint32_t(((d[0] + 255) * 511 + (d[1] + 255)) * 511 + (d[2] + 255));
}

__device__ __forceinline__ uint32_t find_symbol(const uint32_t *__restrict__ keys, uint32_t nsym, uint32_t key) {
    uint32_t lo = 0, hi = nsym;  // keys[lo] <= key < keys[hi]
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (__ldg(keys + mid) <= key) lo = mid;
        else hi = mid;
    }
    return lo;
}

template <int SRC>
__global__ void __launch_bounds__(256) pack_len_kernel(const void *__restrict__ src, size_t n, const uint32_t *__restrict__ keys, uint32_t nsym,
                                                       const uint8_t *__restrict__ lens, uint32_t *block_bits) {
    const size_t base = (size_t)blockIdx.x * 4096 + (size_t)threadIdx.x * 16;
    uint32_t bits = 0;
    for (int j = 0; j < 16; j++)
        if (base + j < n) bits += lens[find_symbol(keys, nsym, pack_key<SRC>(src, base + j))];
    for (int o = 16; o > 0; o >>= 1) bits += __shfl_xor_sync(0xffffffffu, bits, o);
    __shared__ uint32_t s[8];
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = bits;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int i = 0; i < 8; i++) t += s[i];
        block_bits[blockIdx.x] = t;
    }
}

struct BitEmitter {
    uint32_t *out;                 // big-endian bit stream viewed as 32-bit words (byte-swapped on store)
    unsigned long long word;       // index of the word being filled
    uint32_t acc;                  // bits collected for that word, left aligned
    int nb;                        // number of valid bits in acc
    bool shared_first;             // the word being filled may also be written by the previous thread
    __device__ __forceinline__ void flush_word() {
        const uint32_t v = __byte_perm(acc, 0, 0x0123);
        if (shared_first) { if (v) atomicOr(out + word, v); shared_first = false; }
        else out[word] = v;
        word++; acc = 0; nb = 0;
    }
    __device__ __forceinline__ void put(uint32_t bits, int len) {  // len <= 32, bits right aligned
        while (len > 0) {
            const int room = 32 - nb, take = len < room ? len : room;
            const uint32_t chunk = (take == 32) ? bits : ((bits >> (len - take)) & ((1u << take) - 1u));
            acc |= (take == 32) ? chunk : (chunk << (room - take));
            nb += take; len -= take;
            if (nb == 32) flush_word();
        }
    }
    __device__ __forceinline__ void finish() {  // last, partially filled word: may be shared with the next thread
        if (nb) { const uint32_t v = __byte_perm(acc, 0, 0x0123); if (v) atomicOr(out + word, v); }
    }
};

template <int SRC>
__global__ void __launch_bounds__(256) pack_write_kernel(const void *__restrict__ src, size_t n, const uint32_t *__restrict__ keys, uint32_t nsym,
                                                         const uint8_t *__restrict__ lens, const unsigned long long *__restrict__ codes,
                                                         const unsigned long long *__restrict__ block_off, uint32_t *out) {
    __shared__ uint32_t s_w[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t base = (size_t)blockIdx.x * 4096 + (size_t)threadIdx.x * 16;
    uint32_t sym[16];
    uint32_t mine = 0;
#pragma unroll
    for (int j = 0; j < 16; j++) {
        sym[j] = 0xffffffffu;
        if (base + j < n) { sym[j] = find_symbol(keys, nsym, pack_key<SRC>(src, base + j)); mine += lens[sym[j]]; }
    }
    uint32_t x = mine;
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) s_w[warp] = x;
    __syncthreads();
    unsigned long long p = block_off[blockIdx.x] + (x - mine);
    for (int j = 0; j < warp; j++) p += s_w[j];
    if (mine == 0) return;
    BitEmitter em{out, p >> 5, 0u, int(p & 31), true};
#pragma unroll
    for (int j = 0; j < 16; j++) {
        if (sym[j] != 0xffffffffu) {
            const int len = lens[sym[j]];
            const unsigned long long code = codes[sym[j]];
            if (len > 32) { em.put(uint32_t(code >> 32), len - 32); em.put(uint32_t(code), 32); }
            else if (len > 0) em.put(uint32_t(code), len);
        }
    }
    em.finish();
}

inline int grid_for(cniic_ctx *ctx, size_t n, int per_thread = 1) {
    const size_t blocks = (n + 256 * (size_t)per_thread - 1) / (256 * (size_t)per_thread);
    return (int)std::max<size_t>(1, std::min<size_t>(blocks, (size_t)ctx->sm_count * 16));
}


// ============================================================================================================
// exact run-length coding along the Hilbert stream (reference src/codec/hilbertc.rs:99-196, decoder :304-333)
//   record = u8 count (1..=255) + Rgb as a serialised slice (u64 length 3 + 3 bytes) = 12 bytes (ser.rs:164-172).
//   The reference's iterator is greedy: a maximal run of L equal pixels becomes floor(L/255) records of 255 and one of
//   L mod 255 (if non-zero), cut from the START of the run.  With r = position of a pixel inside its maximal run, pixel j
//   ends a record iff j is the last pixel, or pixel j+1 differs, or (r+1) mod 255 == 0 -- and that record's count is
//   r mod 255 + 1.  So records fall out of a segmented scan (run starts) and an ordinary scan (record index).
// ============================================================================================================
constexpr int RLE_PX = 16;                 // pixels per thread
constexpr int RLE_BLOCK = 256 * RLE_PX;    // pixels per CTA

__device__ __forceinline__ uint32_t rle_px(const uint8_t *__restrict__ lin, unsigned long long i) {
    const uint8_t *p = lin + 3 * i;
    return uint32_t(p[0]) | (uint32_t(p[1]) << 8) | (uint32_t(p[2]) << 16);
}

// inclusive scan over the 256 threads of a CTA; MAX: running maximum, else running sum.  Returns the inclusive value,
// *total = value over the whole CTA.  s_w: 8 words of shared scratch.
template <bool MAX>
__device__ __forceinline__ uint32_t block_scan256(uint32_t v, uint32_t *s_w, uint32_t *total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v = MAX ? max(v, y) : v + y;
    }
    __syncthreads();  // s_w may still be read from a previous call
    if (lane == 31) s_w[warp] = v;
    __syncthreads();
    uint32_t before = 0, tot = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const uint32_t x = s_w[i];
        if (i < warp) before = MAX ? max(before, x) : before + x;
        tot = MAX ? max(tot, x) : tot + x;
    }
    *total = tot;
    return MAX ? max(v, before) : v + before;
}

// pass 1: per CTA block, (index + 1) of the last run start inside the block (0 = none)
__global__ void __launch_bounds__(256) rle_heads_kernel(const uint8_t *__restrict__ lin, uint32_t n, uint32_t *blk_last) {
    __shared__ uint32_t s_w[8];
    const uint32_t i0 = blockIdx.x * RLE_BLOCK + threadIdx.x * RLE_PX;
    uint32_t last = 0;
    if (i0 < n) {
        uint32_t prev = i0 ? rle_px(lin, i0 - 1) : 0xffffffffu;
        for (int j = 0; j < RLE_PX && i0 + j < n; j++) {
            const uint32_t c = rle_px(lin, i0 + j);
            if (c != prev) last = i0 + j + 1;
            prev = c;
        }
    }
    uint32_t tot;
    block_scan256<true>(last, s_w, &tot);
    if (threadIdx.x == 0) blk_last[blockIdx.x] = tot;
}

// single-CTA exclusive scan over the per-block values (MAX: running maximum; else sum, with the grand total in out[nb])
template <bool MAX>
__global__ void __launch_bounds__(256) rle_scan_blocks_kernel(const uint32_t *__restrict__ in, uint32_t nb, uint32_t *out) {
    __shared__ uint32_t s_w[8];
    uint32_t carry = 0;
    for (uint32_t base = 0; base < nb; base += 256) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = i < nb ? in[i] : 0u;
        uint32_t tot;
        const uint32_t inc = block_scan256<MAX>(v, s_w, &tot);
        // exclusive value = scan of everything before me
        const uint32_t up = __shfl_up_sync(0xffffffffu, inc, 1);
        uint32_t excl;
        if ((threadIdx.x & 31) != 0) excl = up;
        else {
            excl = 0;
            for (int w = 0; w < int(threadIdx.x >> 5); w++) excl = MAX ? max(excl, s_w[w]) : excl + s_w[w];
        }
        if (i < nb) out[i] = MAX ? max(carry, excl) : carry + excl;
        carry = MAX ? max(carry, tot) : carry + tot;
    }
    if (!MAX && threadIdx.x == 0) out[nb] = carry;
}

// pass 2 (WRITE = false): records per block; pass 3 (WRITE = true): the records themselves
template <bool WRITE>
__global__ void __launch_bounds__(256) rle_emit_kernel(const uint8_t *__restrict__ lin, uint32_t n, const uint32_t *__restrict__ blk_carry,
                                                       uint32_t *blk_nrec, const uint32_t *__restrict__ blk_off, uint32_t *records) {
    __shared__ uint32_t s_w[8];
    const uint32_t i0 = blockIdx.x * RLE_BLOCK + threadIdx.x * RLE_PX;
    uint32_t px[RLE_PX + 2];  // px[0] = predecessor, px[RLE_PX + 1] = successor
    const int nv = i0 < n ? int(min(uint32_t(RLE_PX), n - i0)) : 0;
    px[0] = (nv && i0) ? rle_px(lin, i0 - 1) : 0xffffffffu;
#pragma unroll
    for (int j = 0; j < RLE_PX; j++) px[j + 1] = j < nv ? rle_px(lin, i0 + j) : 0xfffffffeu;
    px[RLE_PX + 1] = (nv == RLE_PX && i0 + RLE_PX < n) ? rle_px(lin, i0 + RLE_PX) : 0xfffffffdu;
    uint32_t last = 0;  // (index + 1) of the last run start among my pixels
#pragma unroll
    for (int j = 0; j < RLE_PX; j++)
        if (j < nv && px[j + 1] != px[j]) last = i0 + j + 1;
    uint32_t tot;
    const uint32_t inc = block_scan256<true>(last, s_w, &tot);
    // run start in force at my first pixel: the last start before me in this block, else the carry of the earlier blocks
    uint32_t up = __shfl_up_sync(0xffffffffu, inc, 1);
    if ((threadIdx.x & 31) == 0) {
        up = 0;
        for (int w = 0; w < int(threadIdx.x >> 5); w++) up = max(up, s_w[w]);
    }
    uint32_t start1 = max(up, blk_carry[blockIdx.x]);  // index + 1
    uint32_t ends = 0, endmask = 0;
    uint32_t cnt_of[RLE_PX];
#pragma unroll
    for (int j = 0; j < RLE_PX; j++) {
        cnt_of[j] = 0;
        if (j < nv) {
            const uint32_t i = i0 + j;
            if (px[j + 1] != px[j]) start1 = i + 1;
            const uint32_t r = i - (start1 - 1);
            const bool is_end = i + 1 == n || px[j + 2] != px[j + 1] || (r + 1) % 255 == 0;
            if (is_end) { ends++; endmask |= 1u << j; cnt_of[j] = r % 255 + 1; }
        }
    }
    const uint32_t inc_e = block_scan256<false>(ends, s_w, &tot);
    if (!WRITE) {
        if (threadIdx.x == 0) blk_nrec[blockIdx.x] = tot;
        return;
    }
    uint32_t idx = blk_off[blockIdx.x] + inc_e - ends;
#pragma unroll
    for (int j = 0; j < RLE_PX; j++)
        if (endmask >> j & 1) {
            uint32_t *rec = records + 3 * (size_t)idx++;   // count, u64 length = 3 (little endian), r, g, b
            rec[0] = cnt_of[j] | (3u << 8);
            rec[1] = 0u;
            rec[2] = px[j + 1] << 8;
        }
}

// ---- decoder: parse the 12-byte records, scan the counts, paint every curve index from its record ----
__global__ void rle_parse_kernel(const uint8_t *__restrict__ recs, uint32_t nrec, uint32_t *cnt, uint8_t *bad) {
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < nrec; r += gridDim.x * blockDim.x) {
        const uint32_t *w = reinterpret_cast<const uint32_t *>(recs + 12 * (size_t)r);
        const uint32_t w0 = w[0], w1 = w[1], w2 = w[2];
        cnt[r] = w0 & 0xff;
        bad[r] = ((w0 >> 8) != 3u || w1 != 0u || (w2 & 0xff) != 0u) ? 1 : 0;  // Rgb slice length must be 3 (ser.rs:210-214)
    }
}

// per-block sums of the counts (256 records per thread block of 256 threads x 1) and, after the block scan, the starts
__global__ void __launch_bounds__(256) rle_cnt_blocks_kernel(const uint32_t *__restrict__ cnt, uint32_t nrec, uint32_t *blk_sum) {
    __shared__ uint32_t s_w[8];
    const uint32_t r = blockIdx.x * 256 + threadIdx.x;
    uint32_t tot;
    block_scan256<false>(r < nrec ? cnt[r] : 0u, s_w, &tot);
    if (threadIdx.x == 0) blk_sum[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(256) rle_starts_kernel(const uint32_t *__restrict__ cnt, const uint8_t *__restrict__ bad, uint32_t nrec,
                                                         const unsigned long long *__restrict__ blk_off, unsigned long long n,
                                                         unsigned long long *start, uint32_t *err) {
    __shared__ uint32_t s_w[8];
    const uint32_t r = blockIdx.x * 256 + threadIdx.x;
    const uint32_t c = r < nrec ? cnt[r] : 0u;
    uint32_t tot;
    const uint32_t inc = block_scan256<false>(c, s_w, &tot);
    if (r < nrec) {
        const unsigned long long st = blk_off[blockIdx.x] + (inc - c);
        start[r] = st;
        // the sequential decoder reads record r only while pixels are missing; a record it reads must be well formed and must
        // not have count 0 (hilbertc.rs:325 assert!(self.count > 0))
        if (st < n && (bad[r] || c == 0)) atomicOr(err, 1u);
    }
}

// exclusive scan of u32 block sums into u64 offsets, single CTA (totals can pass 2^32 for hostile streams); out[nb] = total
__global__ void __launch_bounds__(256) rle_scan_u64_kernel(const uint32_t *__restrict__ in, uint32_t nb, unsigned long long *out) {
    __shared__ unsigned long long s_part[256];
    __shared__ unsigned long long s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < nb; base += 256) {
        const uint32_t i = base + threadIdx.x;
        s_part[threadIdx.x] = i < nb ? in[i] : 0ull;
        __syncthreads();
        if (threadIdx.x == 0) {  // 256 additions per batch: negligible next to the passes over the records
            unsigned long long run = s_carry;
            for (int j = 0; j < 256; j++) { const unsigned long long v = s_part[j]; s_part[j] = run; run += v; }
            s_carry = run;
        }
        __syncthreads();
        if (i < nb) out[i] = s_part[threadIdx.x];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[nb] = s_carry;
}

__global__ void zero_curve_tail_kernel(uint32_t w, uint32_t h, bool pow2, unsigned long long from, uint8_t *out) {
    const unsigned long long n = (unsigned long long)w * h;
    for (unsigned long long i = from + blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * blockDim.x) {
        uint32_t x, y;
        hilbert_d2xy(w, h, pow2, i, &x, &y);
        uint8_t *o = out + ((size_t)y * w + x) * 3;
        o[0] = 0; o[1] = 0; o[2] = 0;
    }
}

__global__ void rle_paint_kernel(const uint8_t *__restrict__ recs, const unsigned long long *__restrict__ start, uint32_t nrec, uint32_t w, uint32_t h,
                                 bool pow2, unsigned long long total, uint8_t *out) {
    const unsigned long long n = (unsigned long long)w * h;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * blockDim.x) {
        uint32_t x, y;
        hilbert_d2xy(w, h, pow2, i, &x, &y);
        uint8_t *o = out + ((size_t)y * w + x) * 3;
        if (i >= total) { o[0] = 0; o[1] = 0; o[2] = 0; continue; }  // the records ran out cleanly: ImageBuffer::new leaves zeros (hilbertc.rs:57)
        // last record whose start is <= i (records with count 0 share their start with the next one and are skipped)
        uint32_t lo = 0, hi = nrec;
        while (hi - lo > 1) {
            const uint32_t mid = lo + (hi - lo) / 2;
            if (__ldg(start + mid) <= i) lo = mid;
            else hi = mid;
        }
        const uint8_t *rec = recs + 12 * (size_t)lo + 9;
        o[0] = rec[0]; o[1] = rec[1]; o[2] = rec[2];
    }
}

inline bool is_pow2_square(uint32_t w, uint32_t h) { return w == h && (w & (w - 1)) == 0; }

}  // namespace

// ---- device-level building blocks (declared in stages.cuh) -----------------------------------------------------

// the device's bins of a key space (kind 0: 2^24 colours, kind 1: 511^3 delta symbols), borrowed until the compaction that
// follows the counting pass has zeroed them again (common.cuh: cniic_bins_acquire / cniic_bins_release)
static int hist_space(cniic_ctx *ctx, int kind, uint32_t **bins, uint8_t **flags, size_t *nbins) {
    static_assert(PAGE == 4096, "api.cu sizes the page flags with the same constant");
    return cniic_bins_acquire(ctx, kind, bins, flags, nbins);
}

static int dense_compact_body(cniic_ctx *ctx, int kind, uint32_t **d_keys, unsigned long long **d_counts, size_t *n_unique) {
    uint32_t *bins;
    uint8_t *flags;
    size_t nbins;
    ST_TRY(hist_space(ctx, kind, &bins, &flags, &nbins));
    const uint32_t npages = (uint32_t)((nbins + PAGE - 1) / PAGE);
    *d_keys = nullptr;
    *d_counts = nullptr;
    DevBuf list(ctx), bc(ctx), off(ctx);
    CU_TRY(ctx, list.alloc((size_t(npages) + 1) * 4));
    CU_TRY(ctx, bc.alloc(size_t(npages) * 4));
    CU_TRY(ctx, off.alloc((size_t(npages) + 1) * 8));
    uint32_t *d_count = list.as<uint32_t>() + npages;
    list_pages_kernel<<<1, 1024, 0, ctx->stream>>>(flags, npages, list.as<uint32_t>(), d_count);
    page_count_kernel<<<npages, 256, 0, ctx->stream>>>(bins, nbins, list.as<uint32_t>(), d_count, bc.as<uint32_t>());
    scan_blocks_kernel<<<1, 1024, 0, ctx->stream>>>(bc.as<uint32_t>(), npages, off.as<unsigned long long>());
    ctx->launches += 3;
    unsigned long long total = 0;
    CU_TRY(ctx, cudaMemcpyAsync(&total, off.as<unsigned long long>() + npages, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    *n_unique = (size_t)total;
    if (!(*d_keys = static_cast<uint32_t *>(cniic_cache_alloc(ctx, std::max<size_t>(16, total * 4))))) return CNIIC_ERR_CUDA;
    if (!(*d_counts = static_cast<unsigned long long *>(cniic_cache_alloc(ctx, std::max<size_t>(16, total * 8))))) return CNIIC_ERR_CUDA;
    page_compact_kernel<<<npages, 256, 0, ctx->stream>>>(bins, nbins, list.as<uint32_t>(), d_count, off.as<unsigned long long>(), flags, *d_keys,
                                                         *d_counts, total);
    ctx->launches++;
    CU_TRY(ctx, cudaGetLastError());
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return CNIIC_OK;
}

// (key, count) lists of the non-empty bins of the key space `d_bins_in` belongs to, ascending; leaves the bins zero and hands
// them back to the device (the lease was taken by the counting pass)
int cniic_dev_dense_compact(cniic_ctx *ctx, const uint32_t *d_bins_in, size_t nbins, uint32_t **d_keys, unsigned long long **d_counts, size_t *n_unique) {
    (void)nbins;
    const int kind = d_bins_in == cniic_bins_peek(ctx, 0) ? 0 : 1;
    const int rc = dense_compact_body(ctx, kind, d_keys, d_counts, n_unique);
    cniic_bins_release(ctx, kind, rc == CNIIC_OK);
    return rc;
}

void cniic_unique_colours_free(cniic_ctx *ctx, UniqueColours *uc) {
    cniic_cache_free(ctx, uc->d_pts);
    cniic_cache_free(ctx, uc->d_wts);
    *uc = UniqueColours();
}

// count_freqs over the pixels (utils.rs:4-16 at clusterc.rs:21) with the result already in the form the culled D = 3 K-means scans:
// see dedup_hist_kernel.  One host synchronisation (the number of unique colours sizes the outputs and the K-means grid).
static int unique_colours_body(cniic_ctx *ctx, const uint8_t *d_rgb, size_t n, UniqueColours *out) {
    *out = UniqueColours();
    uint32_t *bins;
    uint8_t *flags;
    size_t nbins;
    ST_TRY(hist_space(ctx, 0, &bins, &flags, &nbins));  // the 2^24 colour bins, indexed by Morton code for this pass
    const uint32_t npages = (uint32_t)(nbins / PAGE);
    DevBuf bc(ctx), off(ctx);
    if (bc.alloc(size_t(npages) * 4) != cudaSuccess || off.alloc((size_t(npages) + 1) * 8) != cudaSuccess)
        return cniic_set_error(ctx, CNIIC_ERR_CUDA, "cudaMalloc failed");
    if (n) {
        dedup_hist_kernel<<<grid_for(ctx, n / 4 + 1, 2), 256, 0, ctx->stream>>>(d_rgb, n, bins);
        ctx->launches++;
    }
    dedup_count_kernel<<<npages, 256, 0, ctx->stream>>>(bins, bc.as<uint32_t>());
    scan_blocks_kernel<<<1, 1024, 0, ctx->stream>>>(bc.as<uint32_t>(), npages, off.as<unsigned long long>());
    ctx->launches += 2;
    unsigned long long total = 0;
    CU_TRY(ctx, cudaMemcpyAsync(&total, off.as<unsigned long long>() + npages, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    UniqueColours uc;
    uc.u = (size_t)total;
    uc.d_pts = static_cast<uint32_t *>(cniic_cache_alloc(ctx, (uc.u + 8) * 4));
    uc.d_wts = static_cast<uint32_t *>(cniic_cache_alloc(ctx, (uc.u + 8) * 4));
    if (!uc.d_pts || !uc.d_wts) { cniic_unique_colours_free(ctx, &uc); return CNIIC_ERR_CUDA; }
    dedup_compact_kernel<<<npages, 256, 0, ctx->stream>>>(bins, off.as<unsigned long long>(), uc.d_pts, uc.d_wts);
    ctx->launches++;
    if (cudaGetLastError() != cudaSuccess) { cniic_unique_colours_free(ctx, &uc); return cniic_set_error(ctx, CNIIC_ERR_CUDA, "unique-colour compaction failed"); }
    *out = uc;
    return CNIIC_OK;
}

int cniic_dev_unique_colours(cniic_ctx *ctx, const uint8_t *d_rgb, size_t n, UniqueColours *out) {
    const int rc = unique_colours_body(ctx, d_rgb, n, out);
    cniic_bins_release(ctx, 0, rc == CNIIC_OK);  // the compaction kernel queued above zeroes the bins as it reads them
    return rc;
}

int cniic_dev_hist_rgb_bins(cniic_ctx *ctx, const uint8_t *d_rgb, size_t n, uint32_t **d_bins) {
    uint8_t *flags;
    size_t nbins;
    ST_TRY(hist_space(ctx, 0, d_bins, &flags, &nbins));
    if (n) {
        hist_rgb_kernel<<<grid_for(ctx, n, 4), 256, 0, ctx->stream>>>(d_rgb, n, *d_bins, flags);
        ctx->launches++;
    }
    if (cudaGetLastError() != cudaSuccess) {
        cniic_bins_release(ctx, 0, false);
        return cniic_set_error(ctx, CNIIC_ERR_CUDA, "colour histogram launch failed");
    }
    return CNIIC_OK;  // the lease ends in cniic_dev_dense_compact
}

int cniic_dev_recolor(cniic_ctx *ctx, const uint8_t *d_rgb, size_t n, const uint32_t *d_keys, const uint16_t *d_assign, size_t n_unique,
                      const int32_t *d_cen_i32, const uint8_t *d_cen_u8, uint32_t *d_lut, uint8_t *d_out) {
    if (n_unique) {
        if (d_cen_i32) build_lut_kernel<<<grid_for(ctx, n_unique), 256, 0, ctx->stream>>>(d_keys, d_assign, n_unique, d_cen_i32, d_lut);
        else build_lut_u8_kernel<<<grid_for(ctx, n_unique), 256, 0, ctx->stream>>>(d_keys, d_assign, n_unique, d_cen_u8, d_lut);
        ctx->launches++;
    }
    if (n) {
        recolor_kernel<<<grid_for(ctx, n, 4), 256, 0, ctx->stream>>>(d_rgb, n, d_lut, d_out);
        ctx->launches++;
    }
    CU_TRY(ctx, cudaGetLastError());
    return CNIIC_OK;
}

int cniic_dev_keys_to_points(cniic_ctx *ctx, const uint32_t *d_keys, const unsigned long long *d_counts, size_t n, uint8_t *d_rgb, uint32_t *d_wts) {
    if (n) {
        keys_to_points_kernel<<<grid_for(ctx, n), 256, 0, ctx->stream>>>(d_keys, d_counts, n, d_rgb, d_wts);
        ctx->launches++;
    }
    CU_TRY(ctx, cudaGetLastError());
    return CNIIC_OK;
}

// 2^n squares: the three tile stages, fed by TMA (default) or by plain loads (CNIIC_STAGES_NO_TMA=1, kept for A/B measurements)
template <int MODE>
static int launch_tile_stage(cniic_ctx *ctx, const uint8_t *d_rgb, uint32_t n, uint8_t *out_rgb, int16_t *out_delta, uint32_t *bins, uint8_t *flags,
                             unsigned long long blk_begin, unsigned long long blk_end) {
    static const bool no_tma = getenv("CNIIC_STAGES_NO_TMA") != nullptr;
    const size_t cube_v1 = MODE == 2 ? size_t((CUBE_N + 1) / 2) * 4 : 0;
    const unsigned long long nblk = blk_end - blk_begin;
    if (no_tma) {
        const size_t cube = cube_v1;
        if (MODE == 2) {
            CU_TRY(ctx, cudaFuncSetAttribute(hilbert_tile_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cube));
            int per_sm = 0;
            CU_TRY(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, hilbert_tile_kernel<MODE>, 256, cube));
            if (per_sm < 1) return cniic_set_error(ctx, CNIIC_ERR_UNSUPPORTED, "delta histogram kernel does not fit an SM");
            const unsigned grid = (unsigned)std::min<size_t>((size_t)nblk, (size_t)ctx->sm_count * per_sm);  // persistent: one wave
            hilbert_tile_kernel<MODE><<<grid, 256, cube, ctx->stream>>>(d_rgb, n, out_rgb, out_delta, bins, flags, blk_begin, blk_end);
        } else {
            hilbert_tile_kernel<MODE><<<(unsigned)nblk, 256, 0, ctx->stream>>>(d_rgb, n, out_rgb, out_delta, bins, flags, blk_begin, blk_end);
        }
    } else {
        CUtensorMap tmap;
        if (!tma_encode_2d_u8(&tmap, d_rgb, (uint64_t)n * 3, n, (uint64_t)n * 3, HT * 3, HT))
            return cniic_set_error(ctx, CNIIC_ERR_CUDA, "cuTensorMapEncodeTiled failed for a %u x %u image", n, n);
        static const bool tile_v1 = getenv("CNIIC_TILE_V1") != nullptr;  // first TMA version (word-per-pixel copy of the tile), for A/B runs
        // counter cube of the histogram stage: |d| <= 14 per channel (48.8 KB, 3 CTAs/SM) by default; CNIIC_HIST_CUBE_R=15 -> 59.6 KB, 2 CTAs/SM
        static const int cube_r = getenv("CNIIC_HIST_CUBE_R") ? atoi(getenv("CNIIC_HIST_CUBE_R")) : 14;
        void (*kern)(const CUtensorMap, const uint8_t *, uint32_t, uint8_t *, int16_t *, uint32_t *, uint8_t *, unsigned long long, unsigned long long);
        size_t cube = 0;
        int stages = 2;
        if (tile_v1) { kern = hilbert_tile_tma_kernel<MODE>; cube = cube_v1; }
        // 40 registers, two 12 KB stages: six CTAs (48 warps) per SM.  A third stage (two blocks requested ahead) measured SLOWER,
        // 0.169 against 0.157 ms at 8192^2: it costs the sixth CTA, and the resident warps are what hides this kernel's latencies
        else if constexpr (MODE != 2) { kern = hilbert_tile_tma2_kernel<MODE, 6, CUBE_R, 2>; }
        else if (cube_r == 15) { kern = hilbert_tile_tma2_kernel<2, 2, CUBE_R, 2>; cube = cube_v1; }
        else { kern = hilbert_tile_tma2_kernel<2, 3, 14, 2>; cube = size_t((29 * 29 * 29 + 1) / 2) * 4; }
        const size_t smem = stages * size_t(HT_TILE_BYTES) + cube;  // the tile stages + the counter cube
        CU_TRY(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CU_TRY(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        int per_sm = 0;
        CU_TRY(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, smem));
        if (per_sm < 1) return cniic_set_error(ctx, CNIIC_ERR_UNSUPPORTED, "tile stage kernel does not fit an SM");
        const unsigned grid = (unsigned)std::min<size_t>((size_t)nblk, (size_t)ctx->sm_count * per_sm);  // persistent: one wave
        kern<<<grid, 256, smem, ctx->stream>>>(tmap, d_rgb, n, out_rgb, out_delta, bins, flags, blk_begin, blk_end);
    }
    ctx->launches++;
    CU_TRY(ctx, cudaGetLastError());
    return CNIIC_OK;
}

static inline bool tile_path(const void *in, const void *out, uint32_t w, uint32_t h) {
    return w == h && (w & (w - 1)) == 0 && w >= 64 && ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
}

int cniic_dev_hilbert_gather(cniic_ctx *ctx, const uint8_t *d_rgb, uint32_t w, uint32_t h, uint8_t *d_out) {
    if (tile_path(d_rgb, d_out, w, h)) {
        return launch_tile_stage<0>(ctx, d_rgb, w, d_out, nullptr, nullptr, nullptr, 0ull, (unsigned long long)w * h / 4096);
    }
    hilbert_stream_kernel<0><<<grid_for(ctx, (size_t)w * h), 256, 0, ctx->stream>>>(d_rgb, w, h, is_pow2_square(w, h), d_out, nullptr, nullptr, nullptr, 0ull, (unsigned long long)w * h);
    ctx->launches++;
    CU_TRY(ctx, cudaGetLastError());
    return CNIIC_OK;
}

int cniic_dev_zero_curve_tail(cniic_ctx *ctx, uint32_t w, uint32_t h, unsigned long long from, uint8_t *d_img) {
    const unsigned long long n = (unsigned long long)w * h;
    if (from >= n) return CNIIC_OK;
    zero_curve_tail_kernel<<<grid_for(ctx, (size_t)(n - from)), 256, 0, ctx->stream>>>(w, h, is_pow2_square(w, h), from, d_img);
    ctx->launches++;
    CU_TRY(ctx, cudaGetLastError());
    return CNIIC_OK;
}

// hilbertc.rs:26-38 + 99-196: run-length records of the linearised image `d_lin` (n pixels, device) appended to *out
int cniic_dev_rle_encode(cniic_ctx *ctx, const uint8_t *d_lin, size_t n, std::vector<uint8_t> *out) {
    if (n == 0) return CNIIC_OK;
    const uint32_t nb = (uint32_t)((n + RLE_BLOCK - 1) / RLE_BLOCK);
    DevBuf d_last(ctx), d_carry(ctx), d_nrec(ctx), d_off(ctx), d_rec(ctx);
    CU_TRY(ctx, d_last.alloc(size_t(nb) * 4));
    CU_TRY(ctx, d_carry.alloc(size_t(nb) * 4));
    CU_TRY(ctx, d_nrec.alloc(size_t(nb) * 4));
    CU_TRY(ctx, d_off.alloc((size_t(nb) + 1) * 4));
    rle_heads_kernel<<<nb, 256, 0, ctx->stream>>>(d_lin, (uint32_t)n, d_last.as<uint32_t>());
    rle_scan_blocks_kernel<true><<<1, 256, 0, ctx->stream>>>(d_last.as<uint32_t>(), nb, d_carry.as<uint32_t>());
    rle_emit_kernel<false><<<nb, 256, 0, ctx->stream>>>(d_lin, (uint32_t)n, d_carry.as<uint32_t>(), d_nrec.as<uint32_t>(), nullptr, nullptr);
    rle_scan_blocks_kernel<false><<<1, 256, 0, ctx->stream>>>(d_nrec.as<uint32_t>(), nb, d_off.as<uint32_t>());
    ctx->launches += 4;
    CU_TRY(ctx, cudaGetLastError());
    uint32_t nrec = 0;
    CU_TRY(ctx, cudaMemcpyAsync(&nrec, d_off.as<uint32_t>() + nb, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    CU_TRY(ctx, d_rec.alloc(size_t(nrec) * 12));
    rle_emit_kernel<true><<<nb, 256, 0, ctx->stream>>>(d_lin, (uint32_t)n, d_carry.as<uint32_t>(), nullptr, d_off.as<uint32_t>(), d_rec.as<uint32_t>());
    ctx->launches += 1;
    CU_TRY(ctx, cudaGetLastError());
    const size_t at = out->size();
    out->resize(at + size_t(nrec) * 12);
    CU_TRY(ctx, cudaMemcpyAsync(out->data() + at, d_rec.p, size_t(nrec) * 12, cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return CNIIC_OK;
}

// hilbertc.rs:55-79 + 304-333: `recs` = HOST bytes behind the dimensions; paints the w x h image d_out (device).
// Reference semantics: the decoder iterator is zipped with the curve over a zero-initialised image.  If the bytes end exactly at
// a record boundary the iterator just ends and the remaining pixels STAY ZERO (decode still returns the image); a record that is
// read but has count 0, or whose colour is truncated / malformed, makes the reference panic -> CNIIC_ERR_DECODE here.
int cniic_dev_rle_decode(cniic_ctx *ctx, const uint8_t *recs, size_t len, uint32_t w, uint32_t h, uint8_t *d_out) {
    const unsigned long long n = (unsigned long long)w * h;
    if (n == 0) return CNIIC_OK;
    const size_t nrec_all = len / 12;
    const bool ragged = len % 12 != 0;  // bytes of an incomplete record behind the last complete one
    if (nrec_all == 0) {
        if (ragged) return cniic_set_error(ctx, CNIIC_ERR_DECODE, "truncated RLE record");
        CU_TRY(ctx, cudaMemsetAsync(d_out, 0, n * 3, ctx->stream));  // no record at all: the image stays zero
        return CNIIC_OK;
    }
    // records behind the first n can never be read (every count that matters is >= 0; n records of count >= 1 suffice only if
    // all are non-zero, so keep up to 2^31 - 1 and let the scan decide)
    const uint32_t nrec = (uint32_t)std::min<size_t>(nrec_all, (size_t(1) << 31) - 1);
    const uint32_t nb = (nrec + 255) / 256;
    DevBuf d_recs(ctx), d_cnt(ctx), d_bad(ctx), d_bsum(ctx), d_boff(ctx), d_start(ctx), d_err(ctx);
    CU_TRY(ctx, d_recs.alloc(size_t(nrec) * 12));
    CU_TRY(ctx, d_cnt.alloc(size_t(nrec) * 4));
    CU_TRY(ctx, d_bad.alloc(nrec));
    CU_TRY(ctx, d_bsum.alloc(size_t(nb) * 4));
    CU_TRY(ctx, d_boff.alloc((size_t(nb) + 1) * 8));
    CU_TRY(ctx, d_start.alloc(size_t(nrec) * 8));
    CU_TRY(ctx, d_err.alloc(256));
    CU_TRY(ctx, cudaMemcpyAsync(d_recs.p, recs, size_t(nrec) * 12, cudaMemcpyHostToDevice, ctx->stream));
    CU_TRY(ctx, cudaMemsetAsync(d_err.p, 0, 4, ctx->stream));
    rle_parse_kernel<<<grid_for(ctx, nrec), 256, 0, ctx->stream>>>(d_recs.as<uint8_t>(), nrec, d_cnt.as<uint32_t>(), d_bad.as<uint8_t>());
    rle_cnt_blocks_kernel<<<nb, 256, 0, ctx->stream>>>(d_cnt.as<uint32_t>(), nrec, d_bsum.as<uint32_t>());
    rle_scan_u64_kernel<<<1, 256, 0, ctx->stream>>>(d_bsum.as<uint32_t>(), nb, d_boff.as<unsigned long long>());
    rle_starts_kernel<<<nb, 256, 0, ctx->stream>>>(d_cnt.as<uint32_t>(), d_bad.as<uint8_t>(), nrec, d_boff.as<unsigned long long>(), n,
                                                   d_start.as<unsigned long long>(), d_err.as<uint32_t>());
    ctx->launches += 4;
    CU_TRY(ctx, cudaGetLastError());
    unsigned long long total = 0;
    uint32_t err = 0;
    CU_TRY(ctx, cudaMemcpyAsync(&total, d_boff.as<unsigned long long>() + nb, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(ctx, cudaMemcpyAsync(&err, d_err.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (err) return cniic_set_error(ctx, CNIIC_ERR_DECODE, "malformed or zero-count RLE record (the reference panics)");
    if (total < n && ragged) return cniic_set_error(ctx, CNIIC_ERR_DECODE, "truncated RLE record (the reference panics)");
    rle_paint_kernel<<<grid_for(ctx, n), 256, 0, ctx->stream>>>(d_recs.as<uint8_t>(), d_start.as<unsigned long long>(), nrec, w, h, is_pow2_square(w, h),
                                                                std::min<unsigned long long>(total, n), d_out);
    ctx->launches += 1;
    CU_TRY(ctx, cudaGetLastError());
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return CNIIC_OK;
}

int cniic_dev_hist_delta_bins(cniic_ctx *ctx, const uint8_t *d_rgb, uint32_t w, uint32_t h, uint32_t **d_bins, size_t *nbins) {
    return cniic_dev_hist_delta_bins_range(ctx, d_rgb, w, h, 0, (unsigned long long)w * h, d_bins, nbins);
}

// histogram of the delta symbols of curve indices [i0, i1) only (one rank's share; partial histograms add up)
int cniic_dev_hist_delta_bins_range(cniic_ctx *ctx, const uint8_t *d_rgb, uint32_t w, uint32_t h, unsigned long long i0, unsigned long long i1,
                                    uint32_t **d_bins, size_t *nbins) {
    uint8_t *flags;
    ST_TRY(hist_space(ctx, 1, d_bins, &flags, nbins));  // the lease ends in cniic_dev_dense_compact
    if (i0 >= i1) return CNIIC_OK;
    int rc = CNIIC_OK;
    if (tile_path(d_rgb, nullptr, w, h) && i0 % 4096 == 0 && i1 % 4096 == 0) {
        rc = launch_tile_stage<2>(ctx, d_rgb, w, nullptr, nullptr, *d_bins, flags, i0 / 4096, i1 / 4096);
    } else {
        hilbert_stream_kernel<2><<<grid_for(ctx, (size_t)(i1 - i0)), 256, 0, ctx->stream>>>(d_rgb, w, h, is_pow2_square(w, h), nullptr, nullptr, *d_bins, flags, i0, i1);
        ctx->launches++;
        if (cudaGetLastError() != cudaSuccess) rc = cniic_set_error(ctx, CNIIC_ERR_CUDA, "delta histogram launch failed");
    }
    if (rc != CNIIC_OK) cniic_bins_release(ctx, 1, false);
    return rc;
}

// ---- C ABI -----------------------------------------------------------------------------------------------------


extern "C" int cniic_voronoi_fill_device(cniic_ctx *ctx, const uint32_t *d_cxy, const uint8_t *d_crgb, uint32_t k, uint32_t w,
                                         uint32_t h, uint32_t y0, uint32_t h_local, uint8_t *d_out_rgb) {
    if (!ctx || !d_cxy || !d_crgb || !d_out_rgb) return CNIIC_ERR_BAD_ARG;
    if (k == 0 || k > CNIIC_MAX_K) return cniic_set_error(ctx, CNIIC_ERR_BAD_ARG, "k must be in 1..%d (clusterc.rs:182-184 unwraps on k = 0)", CNIIC_MAX_K);
    if (w == 0 || h_local == 0) return CNIIC_OK;
    if (w > CNIIC_MAX_DIM || h > CNIIC_MAX_DIM || (uint64_t)y0 + h_local > h) return cniic_set_error(ctx, CNIIC_ERR_BAD_ARG, "bad image dimensions");
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    const size_t smem = (size_t)k * 14 + 32;  // level-1 coordinates + colours, level-2 index list
    CU_TRY(ctx, cudaFuncSetAttribute(fill_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    CU_TRY(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fill_kernel, 256, smem));
    if (per_sm < 1) return cniic_set_error(ctx, CNIIC_ERR_UNSUPPORTED, "fill kernel does not fit an SM");
    const size_t tiles = (size_t)((w + FT - 1) / FT) * ((h_local + FT - 1) / FT);
    const int grid = (int)std::min<size_t>(tiles, (size_t)ctx->sm_count * per_sm);
    fill_kernel<<<grid, 256, smem, ctx->stream>>>(d_cxy, d_crgb, k, w, y0, h_local, d_out_rgb);
    ctx->launches++;
    CU_TRY(ctx, cudaGetLastError());
    return CNIIC_OK;
}

extern "C" int cniic_voronoi_fill(cniic_ctx *ctx, const uint32_t *cxy, const uint8_t *crgb, uint32_t k, uint32_t w, uint32_t h,
                                  uint8_t *out_rgb) {
    if (!ctx || !cxy || !crgb || (!out_rgb && (size_t)w * h)) return CNIIC_ERR_BAD_ARG;
    if (k == 0 || k > CNIIC_MAX_K) return cniic_set_error(ctx, CNIIC_ERR_BAD_ARG, "k must be in 1..%d", CNIIC_MAX_K);
    for (uint32_t c = 0; c < k; c++)
        if (cxy[2 * c] >= 32768 || cxy[2 * c + 1] >= 32768)
            return cniic_set_error(ctx, CNIIC_ERR_UNSUPPORTED, "centroid coordinate >= 32768 (u32 wrap-around of the reference is not reproduced)");
    if ((size_t)w * h == 0) return CNIIC_OK;
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    DevBuf dxy(ctx), drgb(ctx), dout(ctx);
    CU_TRY(ctx, dxy.alloc((size_t)k * 8));
    CU_TRY(ctx, drgb.alloc((size_t)k * 3));
    CU_TRY(ctx, dout.alloc((size_t)w * h * 3));
    CU_TRY(ctx, cudaMemcpyAsync(dxy.p, cxy, (size_t)k * 8, cudaMemcpyHostToDevice, ctx->stream));
    CU_TRY(ctx, cudaMemcpyAsync(drgb.p, crgb, (size_t)k * 3, cudaMemcpyHostToDevice, ctx->stream));
    ST_TRY(cniic_voronoi_fill_device(ctx, dxy.as<uint32_t>(), drgb.as<uint8_t>(), k, w, h, 0, h, dout.as<uint8_t>()));
    CU_TRY(ctx, cudaMemcpyAsync(out_rgb, dout.p, (size_t)w * h * 3, cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return CNIIC_OK;
}

static int hist_out(cniic_ctx *ctx, uint32_t *d_bins, size_t nbins, uint32_t *out_keys, uint64_t *out_counts, size_t cap, size_t *out_n) {
    uint32_t *d_keys = nullptr;
    unsigned long long *d_counts = nullptr;
    size_t u = 0;
    int rc = cniic_dev_dense_compact(ctx, d_bins, nbins, &d_keys, &d_counts, &u);
    if (rc == CNIIC_OK) {
        if (out_n) *out_n = u;
        if (u > cap) rc = cniic_set_error(ctx, CNIIC_ERR_BUFFER_TOO_SMALL, "%zu distinct symbols, capacity %zu", u, cap);
        else if (u) {
            cudaMemcpyAsync(out_keys, d_keys, u * 4, cudaMemcpyDeviceToHost, ctx->stream);
            cudaMemcpyAsync(out_counts, d_counts, u * 8, cudaMemcpyDeviceToHost, ctx->stream);
            if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = cniic_set_error(ctx, CNIIC_ERR_CUDA, "histogram copy failed");
        }
    }
    cniic_cache_free(ctx, d_keys);
    cniic_cache_free(ctx, d_counts);
    return rc;
}

extern "C" int cniic_hist_rgb(cniic_ctx *ctx, const uint8_t *rgb, size_t n, uint32_t *out_keys, uint64_t *out_counts, size_t cap,
                              size_t *out_n) {
    if (!ctx || (!rgb && n) || !out_n || (cap && (!out_keys || !out_counts))) return CNIIC_ERR_BAD_ARG;
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    DevBuf din(ctx);
    CU_TRY(ctx, din.alloc(n * 3));
    CU_TRY(ctx, cudaMemcpyAsync(din.p, rgb, n * 3, cudaMemcpyHostToDevice, ctx->stream));
    uint32_t *d_bins = nullptr;
    int rc = cniic_dev_hist_rgb_bins(ctx, din.as<uint8_t>(), n, &d_bins);
    if (rc == CNIIC_OK) rc = hist_out(ctx, d_bins, size_t(1) << 24, out_keys, out_counts, cap, out_n);
    cniic_cache_free(ctx, d_bins);
    return rc;
}

extern "C" int cniic_recolor_rgb(cniic_ctx *ctx, const uint8_t *rgb, size_t n, const uint32_t *keys, const uint16_t *assign,
                                 size_t n_unique, const uint8_t *centroids, uint32_t k, uint8_t *out_rgb) {
    if (!ctx || (!rgb && n) || !keys || !assign || !centroids || (!out_rgb && n) || k == 0) return CNIIC_ERR_BAD_ARG;
    for (size_t i = 0; i < n_unique; i++)
        if (assign[i] >= k || keys[i] >= (1u << 24)) return cniic_set_error(ctx, CNIIC_ERR_BAD_ARG, "assignment / key out of range");
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    DevBuf din(ctx), dkeys(ctx), dasg(ctx), dcen(ctx), dlut(ctx), dout(ctx);
    CU_TRY(ctx, din.alloc(n * 3));
    CU_TRY(ctx, dkeys.alloc(n_unique * 4));
    CU_TRY(ctx, dasg.alloc(n_unique * 2));
    CU_TRY(ctx, dcen.alloc((size_t)k * 3));
    CU_TRY(ctx, dlut.alloc((size_t(1) << 24) * 4));
    CU_TRY(ctx, dout.alloc(n * 3));
    CU_TRY(ctx, cudaMemcpyAsync(din.p, rgb, n * 3, cudaMemcpyHostToDevice, ctx->stream));
    CU_TRY(ctx, cudaMemcpyAsync(dkeys.p, keys, n_unique * 4, cudaMemcpyHostToDevice, ctx->stream));
    CU_TRY(ctx, cudaMemcpyAsync(dasg.p, assign, n_unique * 2, cudaMemcpyHostToDevice, ctx->stream));
    CU_TRY(ctx, cudaMemcpyAsync(dcen.p, centroids, (size_t)k * 3, cudaMemcpyHostToDevice, ctx->stream));
    CU_TRY(ctx, cudaMemsetAsync(dlut.p, 0, (size_t(1) << 24) * 4, ctx->stream));
    ST_TRY(cniic_dev_recolor(ctx, din.as<uint8_t>(), n, dkeys.as<uint32_t>(), dasg.as<uint16_t>(), n_unique, nullptr, dcen.as<uint8_t>(),
                             dlut.as<uint32_t>(), dout.as<uint8_t>()));
    CU_TRY(ctx, cudaMemcpyAsync(out_rgb, dout.p, n * 3, cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return CNIIC_OK;
}

// device-resident cluster-colors front half; d_out may alias nothing; returns centroids (k x 3 i32 on host)
int cniic_dev_cluster_colors(cniic_ctx *ctx, const uint8_t *d_rgb, size_t n, uint32_t k, uint32_t max_iters, int tie_rule, uint8_t *d_out,
                             std::vector<int32_t> *cen_host, cniic_kmeans_stats *stats, size_t *n_unique) {
    // clusterc.rs:19-47: count_freqs -> kmeans::cluster over (colour, count) points -> recolour.  The unique colours leave the
    // histogram already deduplicated and Morton-sorted (cniic_dev_unique_colours), which is what the culled K-means kernel scans.
    UniqueColours uc;
    int rc = cniic_dev_unique_colours(ctx, d_rgb, n, &uc);
    const size_t u = uc.u;
    if (n_unique) *n_unique = u;
    cniic_kmeans *km = nullptr;
    if (rc == CNIIC_OK && u / k == 0) rc = cniic_set_error(ctx, CNIIC_ERR_TOO_FEW_POINTS, "only %zu distinct colours for k = %u (kmeans.rs:67-68)", u, k);
    if (rc == CNIIC_OK) rc = cniic_kmeans_open_unique(ctx, &uc, k, tie_rule, &km);
    cniic_unique_colours_free(ctx, &uc);  // (whatever the session did not take over)
    if (rc == CNIIC_OK) rc = cniic_kmeans_reset(km, nullptr);
    if (rc == CNIIC_OK) rc = cniic_kmeans_run(km, max_iters, stats);
    std::vector<uint64_t> wts(k);
    std::vector<int32_t> cen(size_t(k) * 3);
    if (rc == CNIIC_OK) rc = cniic_kmeans_get(km, cen.data(), wts.data(), nullptr);
    if (rc == CNIIC_OK && d_out) {
        // colour -> centroid colour lookup table (every colour that occurs is written before it is read)
        DevBuf dcen(ctx), dlut(ctx);
        if (dcen.alloc(cen.size() * 4) != cudaSuccess || dlut.alloc((size_t(1) << 24) * 4) != cudaSuccess) rc = cniic_set_error(ctx, CNIIC_ERR_CUDA, "cudaMalloc failed");
        if (rc == CNIIC_OK) {
            cudaMemcpyAsync(dcen.p, cen.data(), cen.size() * 4, cudaMemcpyHostToDevice, ctx->stream);
            const uint32_t *pts_sorted;
            const uint16_t *assign_sorted;
            cniic_kmeans_sorted_view(km, &pts_sorted, &assign_sorted);
            if (u) {
                build_lut_sorted_kernel<<<grid_for(ctx, u), 256, 0, ctx->stream>>>(pts_sorted, assign_sorted, u, dcen.as<int32_t>(), dlut.as<uint32_t>());
                ctx->launches++;
            }
            if (n) {
                recolor_kernel<<<grid_for(ctx, n, 4), 256, 0, ctx->stream>>>(d_rgb, n, dlut.as<uint32_t>(), d_out);
                ctx->launches++;
            }
            if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = cniic_set_error(ctx, CNIIC_ERR_CUDA, "recolour failed");
        }
    }
    if (km) cniic_kmeans_close(km);
    if (rc != CNIIC_OK) return rc;
    if (cen_host) *cen_host = cen;
    // kmeans.rs:41-57
    uint64_t active = 0;
    for (uint32_t c = 0; c < k; c++) active += wts[c] > 0;
    uint64_t min_cc = (uint64_t)(0.99 * (double)k);
    if (u < min_cc) min_cc = u;
    if (active < min_cc) return cniic_set_error(ctx, CNIIC_ERR_TOO_FEW_ACTIVE, "Not enough active clusters: requested %u, got %llu", k, (unsigned long long)active);
    return CNIIC_OK;
}

// device-resident form (bench.py's C2 `value`): image in HBM, recoloured image (nullable) to HBM, centroids to the host
extern "C" int cniic_cluster_colors_device(cniic_ctx *ctx, const uint8_t *d_rgb, size_t n_pixels, uint32_t k, uint32_t max_iters, int tie_rule,
                                           uint8_t *d_out_rgb, int32_t *out_centroids, size_t *out_n_unique, cniic_kmeans_stats *stats) {
    if (!ctx || !d_rgb) return CNIIC_ERR_BAD_ARG;
    if (k == 0 || k > CNIIC_MAX_K) return cniic_set_error(ctx, CNIIC_ERR_BAD_ARG, "k must be in 1..%d", CNIIC_MAX_K);
    if (n_pixels >= (size_t(1) << 31)) return cniic_set_error(ctx, CNIIC_ERR_BAD_ARG, "image too large");
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    std::vector<int32_t> cen;
    const int rc = cniic_dev_cluster_colors(ctx, d_rgb, n_pixels, k, max_iters, tie_rule, d_out_rgb, &cen, stats, out_n_unique);
    if ((rc == CNIIC_OK || rc == CNIIC_ERR_TOO_FEW_ACTIVE) && out_centroids && !cen.empty()) memcpy(out_centroids, cen.data(), cen.size() * 4);
    return rc;
}

extern "C" int cniic_cluster_colors(cniic_ctx *ctx, const uint8_t *rgb, uint32_t w, uint32_t h, uint32_t k, uint32_t max_iters,
                                    int tie_rule, uint8_t *out_rgb, uint8_t *out_centroids, cniic_kmeans_stats *stats) {
    if (!ctx || !rgb) return CNIIC_ERR_BAD_ARG;
    if (k == 0 || k > CNIIC_MAX_K) return cniic_set_error(ctx, CNIIC_ERR_BAD_ARG, "k must be in 1..%d", CNIIC_MAX_K);
    const size_t n = (size_t)w * h;
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    DevBuf din(ctx), dout(ctx);
    CU_TRY(ctx, din.alloc(n * 3));
    if (out_rgb) CU_TRY(ctx, dout.alloc(n * 3));
    CU_TRY(ctx, cudaMemcpyAsync(din.p, rgb, n * 3, cudaMemcpyHostToDevice, ctx->stream));
    std::vector<int32_t> cen;
    const int rc = cniic_dev_cluster_colors(ctx, din.as<uint8_t>(), n, k, max_iters, tie_rule, out_rgb ? dout.as<uint8_t>() : nullptr, &cen, stats, nullptr);
    if (rc != CNIIC_OK && rc != CNIIC_ERR_TOO_FEW_ACTIVE) return rc;
    if (out_rgb) {
        CU_TRY(ctx, cudaMemcpyAsync(out_rgb, dout.p, n * 3, cudaMemcpyDeviceToHost, ctx->stream));
        CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    }
    if (out_centroids)
        for (size_t i = 0; i < cen.size(); i++) out_centroids[i] = (uint8_t)cen[i];
    return rc;
}

static int check_dims(cniic_ctx *ctx, uint32_t w, uint32_t h) {
    if ((uint64_t)w * h >= (1ull << 31) || w > 32768 || h > 32768) return cniic_set_error(ctx, CNIIC_ERR_BAD_ARG, "image too large");
    return CNIIC_OK;
}

extern "C" int cniic_hilbert_xy(cniic_ctx *ctx, uint32_t w, uint32_t h, uint32_t *out_xy) {
    if (!ctx) return CNIIC_ERR_BAD_ARG;
    ST_TRY(check_dims(ctx, w, h));
    const size_t n = (size_t)w * h;
    if (n == 0) return CNIIC_OK;
    if (!out_xy) return CNIIC_ERR_BAD_ARG;
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    DevBuf d(ctx);
    CU_TRY(ctx, d.alloc(n * 8));
    hilbert_xy_kernel<<<grid_for(ctx, n), 256, 0, ctx->stream>>>(w, h, is_pow2_square(w, h), d.as<uint32_t>());
    ctx->launches++;
    CU_TRY(ctx, cudaGetLastError());
    CU_TRY(ctx, cudaMemcpyAsync(out_xy, d.p, n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return CNIIC_OK;
}

extern "C" int cniic_hilbert_gather_rgb(cniic_ctx *ctx, const uint8_t *rgb, uint32_t w, uint32_t h, uint8_t *out_rgb) {
    if (!ctx) return CNIIC_ERR_BAD_ARG;
    ST_TRY(check_dims(ctx, w, h));
    const size_t n = (size_t)w * h;
    if (n == 0) return CNIIC_OK;
    if (!rgb || !out_rgb) return CNIIC_ERR_BAD_ARG;
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    DevBuf din(ctx), dout(ctx);
    CU_TRY(ctx, din.alloc(n * 3));
    CU_TRY(ctx, dout.alloc(n * 3));
    CU_TRY(ctx, cudaMemcpyAsync(din.p, rgb, n * 3, cudaMemcpyHostToDevice, ctx->stream));
    ST_TRY(cniic_dev_hilbert_gather(ctx, din.as<uint8_t>(), w, h, dout.as<uint8_t>()));
    CU_TRY(ctx, cudaMemcpyAsync(out_rgb, dout.p, n * 3, cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return CNIIC_OK;
}

extern "C" int cniic_delta_i16_device(cniic_ctx *ctx, const uint8_t *d_rgb, uint32_t w, uint32_t h, int16_t *d_out) {
    return cniic_delta_i16_range_device(ctx, d_rgb, w, h, 0, (uint64_t)w * h, d_out);
}

extern "C" int cniic_delta_i16_range_device(cniic_ctx *ctx, const uint8_t *d_rgb, uint32_t w, uint32_t h, uint64_t i_begin, uint64_t i_end,
                                            int16_t *d_out) {
    if (!ctx) return CNIIC_ERR_BAD_ARG;
    ST_TRY(check_dims(ctx, w, h));
    if (i_begin > i_end || i_end > (uint64_t)w * h) return cniic_set_error(ctx, CNIIC_ERR_BAD_ARG, "curve range outside the image");
    if (i_begin == i_end) return CNIIC_OK;
    if (!d_rgb || !d_out) return CNIIC_ERR_BAD_ARG;
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    if (tile_path(d_rgb, d_out, w, h) && i_begin % 4096 == 0 && i_end % 4096 == 0)
        return launch_tile_stage<1>(ctx, d_rgb, w, nullptr, d_out, nullptr, nullptr, i_begin / 4096, i_end / 4096);
    else
        hilbert_stream_kernel<1><<<grid_for(ctx, (size_t)(i_end - i_begin)), 256, 0, ctx->stream>>>(d_rgb, w, h, is_pow2_square(w, h), nullptr, d_out, nullptr, nullptr, i_begin, i_end);
    ctx->launches++;
    CU_TRY(ctx, cudaGetLastError());
    return CNIIC_OK;
}

extern "C" int cniic_delta_i16(cniic_ctx *ctx, const uint8_t *rgb, uint32_t w, uint32_t h, int16_t *out) {
    if (!ctx) return CNIIC_ERR_BAD_ARG;
    ST_TRY(check_dims(ctx, w, h));
    const size_t n = (size_t)w * h;
    if (n == 0) return CNIIC_OK;
    if (!rgb || !out) return CNIIC_ERR_BAD_ARG;
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    DevBuf din(ctx), dout(ctx);
    CU_TRY(ctx, din.alloc(n * 3));
    CU_TRY(ctx, dout.alloc(n * 6));
    CU_TRY(ctx, cudaMemcpyAsync(din.p, rgb, n * 3, cudaMemcpyHostToDevice, ctx->stream));
    ST_TRY(cniic_delta_i16_device(ctx, din.as<uint8_t>(), w, h, dout.as<int16_t>()));
    CU_TRY(ctx, cudaMemcpyAsync(out, dout.p, n * 6, cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return CNIIC_OK;
}

int cniic_dev_undelta(cniic_ctx *ctx, const int16_t *d_diff, uint32_t w, uint32_t h, uint8_t *d_out) {
    const unsigned long long n = (unsigned long long)w * h;
    const size_t nblocks = (n + 4095) / 4096;
    DevBuf bs(ctx);
    CU_TRY(ctx, bs.alloc(nblocks * 12 + 4));
    uint32_t *d_err = bs.as<uint32_t>() + nblocks * 3;
    CU_TRY(ctx, cudaMemsetAsync(d_err, 0, 4, ctx->stream));
    undelta_partial_kernel<<<(unsigned)nblocks, 256, 0, ctx->stream>>>(d_diff, n, bs.as<uint32_t>());
    undelta_scan_blocks_kernel<<<1, 32, 0, ctx->stream>>>(bs.as<uint32_t>(), nblocks);
    undelta_apply_kernel<<<(unsigned)nblocks, 256, 0, ctx->stream>>>(d_diff, n, bs.as<uint32_t>(), w, h, is_pow2_square(w, h), d_out, d_err);
    ctx->launches += 3;
    CU_TRY(ctx, cudaGetLastError());
    uint32_t bad = 0;
    CU_TRY(ctx, cudaMemcpyAsync(&bad, d_err, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    // FromDiff (hilbertc.rs:497-507) unwraps the conversion to Rgb<u8>: the reference panics on such a stream
    if (bad) return cniic_set_error(ctx, CNIIC_ERR_DECODE, "delta stream reconstructs a colour channel outside 0..255");
    return CNIIC_OK;
}

extern "C" int cniic_undelta_rgb(cniic_ctx *ctx, const int16_t *diff, uint32_t w, uint32_t h, uint8_t *out_rgb) {
    if (!ctx) return CNIIC_ERR_BAD_ARG;
    ST_TRY(check_dims(ctx, w, h));
    const size_t n = (size_t)w * h;
    if (n == 0) return CNIIC_OK;
    if (!diff || !out_rgb) return CNIIC_ERR_BAD_ARG;
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    DevBuf din(ctx), dout(ctx);
    CU_TRY(ctx, din.alloc(n * 6));
    CU_TRY(ctx, dout.alloc(n * 3));
    CU_TRY(ctx, cudaMemcpyAsync(din.p, diff, n * 6, cudaMemcpyHostToDevice, ctx->stream));
    ST_TRY(cniic_dev_undelta(ctx, din.as<int16_t>(), w, h, dout.as<uint8_t>()));
    CU_TRY(ctx, cudaMemcpyAsync(out_rgb, dout.p, n * 3, cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return CNIIC_OK;
}

// histogram of the delta symbols of curve indices [i_begin, i_end) of a DEVICE-resident image, (key, count) lists to the host:
// one rank's partial histogram of a curve-sharded run; the ranks' lists are merged by adding counts of equal keys
extern "C" int cniic_hist_delta_range_device(cniic_ctx *ctx, const uint8_t *d_rgb, uint32_t w, uint32_t h, uint64_t i_begin, uint64_t i_end,
                                             uint32_t *out_keys, uint64_t *out_counts, size_t cap, size_t *out_n) {
    if (!ctx || !d_rgb || !out_n || (cap && (!out_keys || !out_counts))) return CNIIC_ERR_BAD_ARG;
    ST_TRY(check_dims(ctx, w, h));
    if (i_begin > i_end || i_end > (uint64_t)w * h) return cniic_set_error(ctx, CNIIC_ERR_BAD_ARG, "curve range outside the image");
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    *out_n = 0;
    uint32_t *d_bins = nullptr;
    size_t nbins = 0;
    int rc = cniic_dev_hist_delta_bins_range(ctx, d_rgb, w, h, i_begin, i_end, &d_bins, &nbins);
    if (rc == CNIIC_OK) rc = hist_out(ctx, d_bins, nbins, out_keys, out_counts, cap, out_n);
    cniic_cache_free(ctx, d_bins);
    return rc;
}

extern "C" int cniic_hist_delta_device(cniic_ctx *ctx, const uint8_t *d_rgb, uint32_t w, uint32_t h, size_t *out_n) {
    if (!ctx || !d_rgb || !out_n) return CNIIC_ERR_BAD_ARG;
    ST_TRY(check_dims(ctx, w, h));
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    uint32_t *d_bins = nullptr, *d_keys = nullptr;
    unsigned long long *d_counts = nullptr;
    size_t nbins = 0;
    int rc = cniic_dev_hist_delta_bins(ctx, d_rgb, w, h, &d_bins, &nbins);
    if (rc == CNIIC_OK) rc = cniic_dev_dense_compact(ctx, d_bins, nbins, &d_keys, &d_counts, out_n);
    cniic_cache_free(ctx, d_bins);
    cniic_cache_free(ctx, d_keys);
    cniic_cache_free(ctx, d_counts);
    return rc;
}

extern "C" int cniic_hist_delta(cniic_ctx *ctx, const uint8_t *rgb, uint32_t w, uint32_t h, uint32_t *out_keys, uint64_t *out_counts,
                                size_t cap, size_t *out_n) {
    if (!ctx || !out_n || (cap && (!out_keys || !out_counts))) return CNIIC_ERR_BAD_ARG;
    ST_TRY(check_dims(ctx, w, h));
    const size_t n = (size_t)w * h;
    *out_n = 0;
    if (n == 0) return CNIIC_OK;
    if (!rgb) return CNIIC_ERR_BAD_ARG;
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    DevBuf din(ctx);
    CU_TRY(ctx, din.alloc(n * 3));
    CU_TRY(ctx, cudaMemcpyAsync(din.p, rgb, n * 3, cudaMemcpyHostToDevice, ctx->stream));
    uint32_t *d_bins = nullptr;
    size_t nbins = 0;
    int rc = cniic_dev_hist_delta_bins(ctx, din.as<uint8_t>(), w, h, &d_bins, &nbins);
    if (rc == CNIIC_OK) rc = hist_out(ctx, d_bins, nbins, out_keys, out_counts, cap, out_n);
    cniic_cache_free(ctx, d_bins);
    return rc;
}

int cniic_dev_sse(cniic_ctx *ctx, const uint8_t *d_a, const uint8_t *d_b, size_t nbytes, uint64_t *out) {
    DevBuf acc(ctx);
    CU_TRY(ctx, acc.alloc(8));
    CU_TRY(ctx, cudaMemsetAsync(acc.p, 0, 8, ctx->stream));
    if (nbytes) {
        sse_kernel<<<grid_for(ctx, nbytes / 4 + 1, 4), 256, 0, ctx->stream>>>(d_a, d_b, nbytes, acc.as<unsigned long long>());
        ctx->launches++;
    }
    CU_TRY(ctx, cudaGetLastError());
    CU_TRY(ctx, cudaMemcpyAsync(out, acc.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return CNIIC_OK;
}

extern "C" int cniic_sse_rgb(cniic_ctx *ctx, const uint8_t *a, const uint8_t *b, size_t n_pixels, uint64_t *out_sse) {
    if (!ctx || !out_sse || ((!a || !b) && n_pixels)) return CNIIC_ERR_BAD_ARG;
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    DevBuf da(ctx), db(ctx);
    CU_TRY(ctx, da.alloc(n_pixels * 3));
    CU_TRY(ctx, db.alloc(n_pixels * 3));
    CU_TRY(ctx, cudaMemcpyAsync(da.p, a, n_pixels * 3, cudaMemcpyHostToDevice, ctx->stream));
    CU_TRY(ctx, cudaMemcpyAsync(db.p, b, n_pixels * 3, cudaMemcpyHostToDevice, ctx->stream));
    return cniic_dev_sse(ctx, da.as<uint8_t>(), db.as<uint8_t>(), n_pixels * 3, out_sse);
}

// Huffman payload of a symbol stream, packed on the device.  d_keys: ascending symbol keys (device); codes/lens: per symbol id (host).
// *out receives the payload bytes (MSB-first, zero padded to a whole byte).
int cniic_dev_huffman_pack(cniic_ctx *ctx, int src_kind, const void *d_src, size_t n, const uint32_t *d_keys, size_t nsym,
                           const std::vector<uint64_t> &codes, const std::vector<uint8_t> &lens, std::vector<uint8_t> *out) {
    if (n == 0 || nsym == 0) return CNIIC_OK;  // the payload is APPENDED to *out (the stream under construction)
    const size_t nblocks = (n + 4095) / 4096;
    DevBuf dl(ctx), dc(ctx), bb(ctx), off(ctx), dout(ctx);
    CU_TRY(ctx, dl.alloc(nsym));
    CU_TRY(ctx, dc.alloc(nsym * 8));
    CU_TRY(ctx, bb.alloc(nblocks * 4));
    CU_TRY(ctx, off.alloc((nblocks + 1) * 8));
    CU_TRY(ctx, cudaMemcpyAsync(dl.p, lens.data(), nsym, cudaMemcpyHostToDevice, ctx->stream));
    CU_TRY(ctx, cudaMemcpyAsync(dc.p, codes.data(), nsym * 8, cudaMemcpyHostToDevice, ctx->stream));
    if (src_kind == 0) pack_len_kernel<0><<<(unsigned)nblocks, 256, 0, ctx->stream>>>(d_src, n, d_keys, (uint32_t)nsym, dl.as<uint8_t>(), bb.as<uint32_t>());
    else pack_len_kernel<1><<<(unsigned)nblocks, 256, 0, ctx->stream>>>(d_src, n, d_keys, (uint32_t)nsym, dl.as<uint8_t>(), bb.as<uint32_t>());
    scan_blocks_kernel<<<1, 1024, 0, ctx->stream>>>(bb.as<uint32_t>(), nblocks, off.as<unsigned long long>());
    ctx->launches += 2;
    unsigned long long total_bits = 0;
    CU_TRY(ctx, cudaMemcpyAsync(&total_bits, off.as<unsigned long long>() + nblocks, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    const size_t nbytes = (size_t)((total_bits + 7) / 8), nwords = (nbytes + 3) / 4;
    if (nbytes == 0) return CNIIC_OK;
    CU_TRY(ctx, dout.alloc(nwords * 4 + 16));
    CU_TRY(ctx, cudaMemsetAsync(dout.p, 0, nwords * 4 + 16, ctx->stream));
    if (src_kind == 0) pack_write_kernel<0><<<(unsigned)nblocks, 256, 0, ctx->stream>>>(d_src, n, d_keys, (uint32_t)nsym, dl.as<uint8_t>(), dc.as<unsigned long long>(), off.as<unsigned long long>(), dout.as<uint32_t>());
    else pack_write_kernel<1><<<(unsigned)nblocks, 256, 0, ctx->stream>>>(d_src, n, d_keys, (uint32_t)nsym, dl.as<uint8_t>(), dc.as<unsigned long long>(), off.as<unsigned long long>(), dout.as<uint32_t>());
    ctx->launches++;
    CU_TRY(ctx, cudaGetLastError());
    const size_t at = out->size();
    out->resize(at + nbytes);
    CU_TRY(ctx, cudaMemcpyAsync(out->data() + at, dout.p, nbytes, cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return CNIIC_OK;
}
