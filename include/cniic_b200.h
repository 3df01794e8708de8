/*
 * cniic_b200.h -- C ABI of the B200-native (sm_100a CUDA) implementation of cniic's data-parallel hot path.
 *
 * The reference (hkapp/cniic, Rust) has NO FFI/plugin ABI; this header is the boundary a thin
 * `cniic-cuda-sys` Rust crate binds (INTEGRATION.md shows the bindings).  Every entry point names the
 * reference interface it replaces (paths relative to the reference's src/).
 *
 * Conventions
 *   - plain pointers and sizes only; every function returns an int status (CNIIC_OK == 0) and never unwinds.
 *   - images are row-major packed RGB8: u8 rgb[3*w*h] == image::DynamicImage::to_rgb8().as_raw().
 *   - "host" entry points take HOST buffers, copy H->D, run the kernels and copy results back; the
 *     caller owns every buffer; the library keeps no pointer after return.
 *   - `cniic_ctx` owns one CUDA device + stream + scratch.  A ctx is NOT thread-safe; create one ctx per calling
 *     thread (bench.rs:27 calls codecs from rayon workers, one image per worker).  Different ctx objects may be
 *     used concurrently; the two dense histogram key spaces (64 MB of colours, 534 MB of delta symbols) belong to the
 *     DEVICE and are lent to one counting call at a time, so contexts share them instead of holding a copy each.
 *   - there is NO CPU fallback: without a usable CUDA device cniic_ctx_create fails with CNIIC_ERR_CUDA.
 *
 * Deterministic rules where the reference is random / unordered are stated at each function
 * and in DESIGN.md; they are identical to the rules of the CPU oracle used by the parity tests.
 */
#ifndef CNIIC_B200_H
#define CNIIC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- status codes ---- */
#define CNIIC_OK 0
#define CNIIC_ERR_BAD_ARG 1          /* null pointer, k == 0, unsupported size                                   */
#define CNIIC_ERR_TOO_FEW_POINTS 2   /* kmeans.rs:67-68   assert!(points_per_cluster > 0)  (N < k)               */
#define CNIIC_ERR_TOO_FEW_ACTIVE 3   /* kmeans.rs:41-57   check_enough_active_clusters panics                    */
#define CNIIC_ERR_CUDA 4             /* any CUDA runtime failure; see cniic_last_error                           */
#define CNIIC_ERR_NCCL 5             /* NCCL failure / NCCL not loadable                                         */
#define CNIIC_ERR_DECODE 6           /* malformed stream: Codec::decode returns None (codec.rs:16)               */
#define CNIIC_ERR_BUFFER_TOO_SMALL 7 /* out buffer too small; *out_len holds the required size                   */
#define CNIIC_ERR_UNSUPPORTED 8

/* ---- limits ---- */
#define CNIIC_MAX_K 4096   /* clusters per K-means run                                                          */
#define CNIIC_MAX_DIM 16384 /* image width / height for the (x,y,r,g,b) path (int32-exact scores)               */

/* tie rule of the assignment step (exact ties of integer squared distance) */
#define CNIIC_TIE_KEEP_CURRENT 0 /* kmeans.rs:350-378: move only to a STRICTLY closer centroid, else lowest index */
#define CNIIC_TIE_LOWEST_INDEX 1 /* plain lowest-index argmin                                                    */

typedef struct cniic_ctx cniic_ctx;
typedef struct cniic_kmeans cniic_kmeans;

typedef struct {
    uint32_t iterations;   /* kmeans.rs:24-33 "#iterations"                                                     */
    uint32_t empty_events; /* kmeans.rs:117-134 empty-cluster repairs over the run                               */
    uint64_t moved_last;   /* kmeans.rs:401 "Moved" of the last assignment pass                                  */
    uint64_t moved_total;
    uint32_t converged;    /* 1 if the last pass moved nothing (kmeans.rs:25 loop exit)                          */
    uint32_t gpu_launches; /* kernels launched by this run                                                      */
    float device_ms;       /* CUDA-event time of the iteration loop (kernels + collective), H2D/D2H excluded    */
    float assign_ms_avg;   /* mean CUDA-event duration of the fused assign+accumulate kernel (sampled launches)       */
    uint64_t pairs_scored; /* point-centroid pairs actually scored since the last reset (N*k per iteration without culling) */
} cniic_kmeans_stats;

/* ---- context ---- */
int cniic_version(void);
/* device: CUDA ordinal, or -1 = current device.  */
int cniic_ctx_create(int device, cniic_ctx **out);
/* One rank of a multi-GPU job (one process per GPU).  nccl_unique_id = the 128-byte ncclUniqueId created by rank 0
 * (cniic_nccl_unique_id) and distributed by the host (torch.distributed store / MPI / file).                   */
int cniic_ctx_create_dist(int device, int rank, int world, const uint8_t nccl_unique_id[128], cniic_ctx **out);
int cniic_nccl_unique_id(uint8_t out_id[128]);
/* Peer-memory exchange (replaces the per-iteration NCCL call of the row-sharded Lloyd loop; cniic_b200/dist.py enables it by
 * default, without it the library calls ncclAllReduce): every rank exports a CUDA IPC handle of its exchange region, the host
 * gathers the world x 64 bytes, every rank connects.  The update kernel of every rank then PUSHES its partial sums into all
 * ranks' receive areas over NVLink as self-validating 16-byte cells and adds what it received in rank order (DESIGN.md
 * section 6).  Call both on every rank, in the same order, before the first session.                                     */
int cniic_ctx_p2p_export(cniic_ctx *ctx, uint8_t out_handle[64]);
int cniic_ctx_p2p_connect(cniic_ctx *ctx, const uint8_t *handles /* world x 64 bytes, rank order */);
void cniic_ctx_destroy(cniic_ctx *ctx);
const char *cniic_last_error(const cniic_ctx *ctx);
int cniic_ctx_sync(cniic_ctx *ctx);
/* raw cudaStream_t of the ctx (for timing with events on the launching stream) */
void *cniic_ctx_stream(cniic_ctx *ctx);
int cniic_ctx_rank(const cniic_ctx *ctx);
int cniic_ctx_world(const cniic_ctx *ctx);
/* kernels launched so far on this ctx (bench.py reports the per-step delta as gpu_launches) */
uint32_t cniic_ctx_launches(const cniic_ctx *ctx);

/* ---- K-means (replaces kmeans::cluster<T: Point>(Vec<T>, k) -> Clusters<T>, kmeans.rs:21-39) -------------
 *
 * One-shot host-buffer forms.  max_iters == 0 => run until an assignment pass moves nothing (kmeans.rs:25-32).
 * Init = the reference's chunked init (kmeans.rs:61-108).  Exact integer nearest-centroid search (the reference's
 * truncated-neighbour heuristic, kmeans.rs:150-260, is NOT reproduced: results equal the oracle's exact mode).
 * Empty cluster (kmeans.rs:117-134, random in the reference): the j-th empty cluster in id order copies the member
 * with the (j mod m)-th lowest point index of the heaviest cluster (m = its member count, lowest id on ties).
 * Returns CNIIC_ERR_TOO_FEW_ACTIVE (outputs still written) when kmeans.rs:41-57 would panic.                     */

/* ColorCount points (clusterc.rs:68-114). rgb = n x 3 bytes; counts = n weights or NULL (all 1 = per-pixel).
 * out_centroids = k x 3 bytes; out_weight = k x u64 (sum of counts per cluster, nullable);
 * out_assign = n x u16 cluster id (nullable).                                                                     */
int cniic_kmeans_rgb(cniic_ctx *ctx, const uint8_t *rgb, const uint32_t *counts, size_t n, uint32_t k, uint32_t max_iters,
                     int tie_rule, uint8_t *out_centroids, uint64_t *out_weight, uint16_t *out_assign,
                     cniic_kmeans_stats *stats);

/* ColorPos points (clusterc.rs:148-153, 200-248): one point per pixel of a w x h image, raster order.
 * out_xy = k x {u32 x, u32 y}; out_rgb = k x 3 bytes.                                                           */
int cniic_kmeans_xyrgb(cniic_ctx *ctx, const uint8_t *rgb, uint32_t w, uint32_t h, uint32_t k, uint32_t max_iters,
                       int tie_rule, uint32_t *out_xy, uint8_t *out_rgb, uint64_t *out_weight, uint16_t *out_assign,
                       cniic_kmeans_stats *stats);

/* Session form: points stay resident in HBM across calls (what bench.py's device-resident `value` times, and what
 * the row-sharded multi-GPU path uses).                                                                          */
#define CNIIC_POINTS_RGB 0   /* D = 3, points = packed RGB8 bytes (3 B / point)                                  */
#define CNIIC_KMEANS_FORCE_CULL 2 /* RGB: use the colour-sorted culled kernel even for small problems (default: brute force below 2^27 pairs) */
#define CNIIC_KMEANS_NO_CULL 1 /* score all k centroids for every point (brute-force kernels) instead of exact culling; same results */
#define CNIIC_POINTS_XYRGB 1 /* D = 5, points = pixels of a raster image (x, y synthesised from the index)       */

typedef struct {
    int kind;                /* CNIIC_POINTS_*                                                                  */
    uint32_t k;
    int tie_rule;
    uint64_t n_local;        /* points held by this rank (RGB), or w*h_local (XYRGB)                             */
    uint64_t n_total;        /* points over all ranks (== n_local for a single GPU)                              */
    uint64_t first_index;    /* global index of this rank's first point (row-sharding: y0*w)                     */
    uint32_t w, h_local;     /* XYRGB: image width, rows held by this rank                                       */
    uint32_t y0;             /* XYRGB: first global row of this shard                                            */
    const uint8_t *rgb;      /* HOST (points_on_device == 0) or DEVICE pointer to 3*n_local bytes                */
    const uint32_t *weights; /* RGB only, nullable; same memory space as rgb                                     */
    int points_on_device;    /* 1: rgb/weights are device pointers that outlive the session (no copy is made)    */
    int flags;               /* CNIIC_KMEANS_* bits                                                              */
} cniic_kmeans_desc;

/* kmeans::cluster (kmeans.rs:21-39) on the points of `desc` in one call = open + reset(host_init_centroids) + run(max_iters) +
 * get (outputs nullable: k x D int32 centroids, k x u64 weights, n_local x u16 assignment) + close.  Single GPU or one rank of a
 * row-sharded image (then collective: every rank calls it with the same initial centroids).                              */
int cniic_kmeans_cluster(cniic_ctx *ctx, const cniic_kmeans_desc *desc, const int32_t *host_init_centroids, uint32_t max_iters,
                         int32_t *out_centroids, uint64_t *out_weight, uint16_t *out_assign, cniic_kmeans_stats *stats);
int cniic_kmeans_open(cniic_ctx *ctx, const cniic_kmeans_desc *desc, cniic_kmeans **out);
/* (Re)start from the reference's chunked init.  Single GPU: gathered on the device.  Multi-GPU: every rank must
 * pass the same k x D int32 initial centroids (host) computed from the global point list (kmeans.rs:101-108).    */
int cniic_kmeans_reset(cniic_kmeans *km, const int32_t *host_init_centroids /* nullable on a single GPU */);
/* Run Lloyd iterations: fused assign+accumulate kernel, [multi-GPU: all-reduce of the k x (D+1) u64 partial sums, over peer memory
 * inside the finalize kernel or with NCCL],
 * finalize kernel.  Stops after max_iters (0 = unbounded) or at convergence.                                    */
int cniic_kmeans_run(cniic_kmeans *km, uint32_t max_iters, cniic_kmeans_stats *stats);
/* centroids: k x D int32 (D = 3: r,g,b ; D = 5: x,y,r,g,b), weights k x u64, assign n_local x u16 (nullable each) */
int cniic_kmeans_get(cniic_kmeans *km, int32_t *out_centroids, uint64_t *out_weight, uint16_t *out_assign);
/* device pointer to the n_local x u16 assignment (valid until close) */
const uint16_t *cniic_kmeans_device_assign(cniic_kmeans *km);
void cniic_kmeans_close(cniic_kmeans *km);

/* ---- batches of independent images (bench.rs:15-34: `measure_all` hands one image to each rayon worker) -------------
 * `count` K-means problems advance in lock step: every stage is ONE launch for the whole batch (blockIdx.y = problem), so
 * small images fill the GPU and cost one launch per iteration instead of one per image.  Results are those of `count`
 * separate cniic_kmeans_run calls, bit for bit.  All sessions must belong to one single-GPU ctx and agree in kind, k, tie
 * rule, weighted-ness and kernel variant (CNIIC_KMEANS_NO_CULL / FORCE_CULL make the variant explicit).
 * stats (nullable) receives `count` entries; device_ms, assign_ms_avg and gpu_launches describe the whole batch.        */
int cniic_kmeans_reset_batch(cniic_kmeans *const *sessions, uint32_t count);
int cniic_kmeans_run_batch(cniic_kmeans *const *sessions, uint32_t count, uint32_t max_iters, cniic_kmeans_stats *stats);
/* Host-buffer form for RGB images (per-pixel points): rgb[i] = n[i] x 3 bytes.  out_centroids = count x k x 3 bytes,
 * out_weight = count x k (nullable), out_assign = count pointers to n[i] x u16 (nullable).  Returns
 * CNIIC_ERR_TOO_FEW_ACTIVE (outputs written) if kmeans.rs:41-57 would panic for any image.                              */
int cniic_kmeans_rgb_batch(cniic_ctx *ctx, const uint8_t *const *rgb, const size_t *n, uint32_t count, uint32_t k, uint32_t max_iters,
                           int tie_rule, uint8_t *out_centroids, uint64_t *out_weight, uint16_t *const *out_assign,
                           cniic_kmeans_stats *stats);
/* ColorPos points: rgb[i] = image of w[i] x h[i] pixels; out_xy = count x k x {u32 x, u32 y}, out_rgb = count x k x 3 bytes. */
int cniic_kmeans_xyrgb_batch(cniic_ctx *ctx, const uint8_t *const *rgb, const uint32_t *w, const uint32_t *h, uint32_t count, uint32_t k,
                             uint32_t max_iters, int tie_rule, uint32_t *out_xy, uint8_t *out_rgb, uint64_t *out_weight,
                             uint16_t *const *out_assign, cniic_kmeans_stats *stats);

/* ---- cluster-colors stages (clusterc.rs:18-52) ---------------------------------------------------------- */
/* utils::count_freqs over pixels (utils.rs:4-16 as called at clusterc.rs:21 and huf.rs:30 via hufc.rs:15-16).
 * HashMap order is unspecified in the reference; canonical order here = ascending key r<<16|g<<8|b.
 * out_keys/out_counts hold cap entries; *out_n = number of distinct colours (may exceed cap ->
 * CNIIC_ERR_BUFFER_TOO_SMALL).                                                                                  */
int cniic_hist_rgb(cniic_ctx *ctx, const uint8_t *rgb, size_t n, uint32_t *out_keys, uint64_t *out_counts, size_t cap,
                   size_t *out_n);
/* clusterc.rs:31-47: every pixel -> colour of the centroid its colour was assigned to.
 * keys/assign describe the clustered unique colours (ascending keys), centroids = k x 3.                          */
int cniic_recolor_rgb(cniic_ctx *ctx, const uint8_t *rgb, size_t n, const uint32_t *keys, const uint16_t *assign,
                      size_t n_unique, const uint8_t *centroids, uint32_t k, uint8_t *out_rgb);
/* Whole front half of ClusterColors::encode: unique colours -> weighted K-means -> recolour (clusterc.rs:19-47).  The reference
 * feeds K-means the unique colours in HashMap order (random, clusterc.rs:21-27); the deterministic stand-in here and in the oracle is
 * ASCENDING MORTON CODE of (r, g, b), r on the most significant bit of every bit triple -- the order the Morton-indexed histogram
 * bins compact into, so the deduplicated point list needs neither a sort nor a permutation.  The chunked init (kmeans.rs:61-108)
 * and the empty-cluster rule index that list.  out_rgb == NULL skips the recolour pass and its copy.                           */
int cniic_cluster_colors(cniic_ctx *ctx, const uint8_t *rgb, uint32_t w, uint32_t h, uint32_t k, uint32_t max_iters,
                         int tie_rule, uint8_t *out_rgb /* nullable */, uint8_t *out_centroids /* 3k, nullable */,
                         cniic_kmeans_stats *stats);

/* ---- voronoi decode fill (clusterc.rs:179-186) ------------------------------------------------------------
 * per pixel the FIRST centroid minimising (cx-x)^2 + (cy-y)^2, paint its colour.  k >= 1.                        */
int cniic_voronoi_fill(cniic_ctx *ctx, const uint32_t *cxy, const uint8_t *crgb, uint32_t k, uint32_t w, uint32_t h,
                       uint8_t *out_rgb);

/* ---- Hilbert / delta / histograms (hilbert.rs:34-43, hilbertc.rs:402-477, utils.rs:4-16) ------------------
 * The curve of the reference comes from the un-vendored crate zhang_hilbert 0.1.1 (PARITY UNPINNED);
 * this library and the oracle implement the same generalized Hilbert scan (classic Hilbert curve on 2^n squares), which is
 * NOT known to equal Zhang's block scan.  Everything built on the curve -- cniic_hilbert_xy, the delta stream, the "delta" and
 * "hilbert(rle)" codecs -- is therefore self-consistent (encode -> decode round-trips) but not guaranteed to interchange
 * with the reference's CPU codecs until tools/pin_hilbert.sh has compared the two curves on a box with cargo.                */
int cniic_hilbert_xy(cniic_ctx *ctx, uint32_t w, uint32_t h, uint32_t *out_xy /* 2*w*h */);
int cniic_hilbert_gather_rgb(cniic_ctx *ctx, const uint8_t *rgb, uint32_t w, uint32_t h, uint8_t *out_rgb /* 3*w*h */);
/* DiffStream: out[i] = rgb(H(i)) - rgb(H(i-1)) per channel as i16, rgb(H(-1)) = 0.                               */
int cniic_delta_i16(cniic_ctx *ctx, const uint8_t *rgb, uint32_t w, uint32_t h, int16_t *out /* 3*w*h */);
/* inverse (hilbertc.rs:417-431, 482-509) */
int cniic_undelta_rgb(cniic_ctx *ctx, const int16_t *diff, uint32_t w, uint32_t h, uint8_t *out_rgb);
/* histogram of the joint SignedColor symbols of the delta stream, fused (no delta stream materialised).
 * key = ((dr+255)*511 + (dg+255))*511 + (db+255), ascending.                                                     */
int cniic_hist_delta(cniic_ctx *ctx, const uint8_t *rgb, uint32_t w, uint32_t h, uint32_t *out_keys, uint64_t *out_counts,
                     size_t cap, size_t *out_n);
/* exact integer sum of squared channel errors; MSE = sse / (w*h)  (bench.rs:95-104)                              */
int cniic_sse_rgb(cniic_ctx *ctx, const uint8_t *a, const uint8_t *b, size_t n_pixels, uint64_t *out_sse);

/* ---- whole codecs (codec.rs:14-19 Codec::encode / Codec::decode) -------------------------------------------
 * codec = the reference's own expression strings: "cluster-colors(N)", "voronoi(N)", "delta", "hufman",
 * "hilbert(rle)".  Containers, tries and payloads follow ser.rs / huf.rs / bit.rs byte for byte (wire formats), with
 * the deterministic Huffman tie rule: leaves enter in ascending symbol order, heap ordered by (freq, creation seq).  The ORDER
 * of the symbols of "delta" and "hilbert(rle)" is the Hilbert order of this library (see the caveat above).
 * encode: returns CNIIC_ERR_BUFFER_TOO_SMALL with *out_len = required size when cap is too small.               */
int cniic_codec_encode(cniic_ctx *ctx, const char *codec, const uint8_t *rgb, uint32_t w, uint32_t h, uint8_t *out,
                       size_t cap, size_t *out_len);
/* After cniic_codec_encode returned CNIIC_ERR_BUFFER_TOO_SMALL the finished stream stays in the ctx: this call copies it out
 * (no second H2D / K-means / packing pass) and releases it.  CNIIC_ERR_BAD_ARG if no stream is pending on this ctx.    */
int cniic_codec_encode_fetch(cniic_ctx *ctx, uint8_t *out, size_t cap, size_t *out_len);
/* decode: *w,*h receive the dimensions; out_rgb must hold cap_pixels pixels (call with out_rgb NULL to query dims).
 * Short or damaged streams follow the reference decoder of each codec: hufman / cluster-colors return CNIIC_ERR_DECODE
 * when code words are missing (hufc.rs:24-36 -> None); delta and hilbert(rle) zip their symbol iterator with the curve
 * over a zero image (hilbertc.rs:55-79, 417-431), so a payload that simply ENDS (for RLE: at a record boundary) gives the
 * pixels it reached, zeros elsewhere, and CNIIC_OK; wherever the reference would panic (bad trie: .unwrap(); RLE record
 * with count 0 or a truncated colour: assert! / .unwrap()) the call returns CNIIC_ERR_DECODE -- nothing unwinds across
 * the FFI.  Bytes behind the last needed symbol are ignored, as in the reference.                                       */
int cniic_codec_decode(cniic_ctx *ctx, const char *codec, const uint8_t *data, size_t len, uint32_t *w, uint32_t *h,
                       uint8_t *out_rgb, size_t cap_pixels);
/* Codec::name() (clusterc.rs:59-61,191-193; hilbertc.rs:433-435): writes a NUL-terminated name                  */
int cniic_codec_name(const char *codec, char *out, size_t cap);
/* max_iters knob for the K-means codecs (0 = until converged, the reference behaviour) */
int cniic_ctx_set_max_iters(cniic_ctx *ctx, uint32_t max_iters);

/* ---- device-resident helpers used by bench.py (inputs already in HBM) ------------------------------------- */
/* synthetic "photo-like" image generator: counter-based, identical on host and device.              */
int cniic_synth_image_device(cniic_ctx *ctx, uint8_t *d_rgb, uint32_t w, uint32_t h, uint32_t y0, uint32_t h_total,
                             uint64_t seed, uint32_t n_blobs);
int cniic_synth_image_host(uint8_t *rgb, uint32_t w, uint32_t h, uint32_t y0, uint32_t h_total, uint64_t seed,
                           uint32_t n_blobs);
void *cniic_device_alloc(cniic_ctx *ctx, size_t bytes);
void cniic_device_free(cniic_ctx *ctx, void *p);
int cniic_memcpy_h2d(cniic_ctx *ctx, void *d, const void *h, size_t bytes);
int cniic_memcpy_d2h(cniic_ctx *ctx, void *h, const void *d, size_t bytes);
/* device-pointer forms of the per-pixel stages (same semantics as the host forms above) */
/* cniic_cluster_colors on an image resident in HBM: d_out_rgb (nullable) receives the recoloured image in HBM, out_centroids
 * (host, nullable) k x 3 int32, *out_n_unique (nullable) the number of distinct colours K-means iterated over.            */
int cniic_cluster_colors_device(cniic_ctx *ctx, const uint8_t *d_rgb, size_t n_pixels, uint32_t k, uint32_t max_iters, int tie_rule,
                                uint8_t *d_out_rgb, int32_t *out_centroids, size_t *out_n_unique, cniic_kmeans_stats *stats);
int cniic_voronoi_fill_device(cniic_ctx *ctx, const uint32_t *d_cxy, const uint8_t *d_crgb, uint32_t k, uint32_t w,
                              uint32_t h, uint32_t y0, uint32_t h_local, uint8_t *d_out_rgb);
int cniic_delta_i16_device(cniic_ctx *ctx, const uint8_t *d_rgb, uint32_t w, uint32_t h, int16_t *d_out);
int cniic_hist_delta_device(cniic_ctx *ctx, const uint8_t *d_rgb, uint32_t w, uint32_t h, size_t *out_n);
/* Curve-sharded forms (GPU g of G owns curve indices [g*N/G, (g+1)*N/G); every rank holds the image).  The
 * predecessor of a range's first symbol is recomputed from the image, so no halo is exchanged; d_out receives
 * 3*(i_end-i_begin) values; partial histograms are merged by adding the counts of equal keys.  Ranges aligned to 4096
 * indices take the tile kernels on 2^n squares.  (hilbert.rs:34-43, hilbertc.rs:445-477, utils.rs:4-16)                 */
int cniic_delta_i16_range_device(cniic_ctx *ctx, const uint8_t *d_rgb, uint32_t w, uint32_t h, uint64_t i_begin, uint64_t i_end,
                                 int16_t *d_out);
int cniic_hist_delta_range_device(cniic_ctx *ctx, const uint8_t *d_rgb, uint32_t w, uint32_t h, uint64_t i_begin, uint64_t i_end,
                                  uint32_t *out_keys, uint64_t *out_counts, size_t cap, size_t *out_n);

#ifdef __cplusplus
}
#endif
#endif /* CNIIC_B200_H */
