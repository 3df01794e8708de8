/*
 * cniic_oracle.h -- CPU restatement of cniic's K-means / voronoi / pre-Huffman hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load this library.  The product
 * (cniic_b200/, include/cniic_b200.h) never links, imports or calls it.
 *
 * Every function cites the reference file:line (paths relative to /root/reference/src) it follows.
 * The reference is Rust and cannot be built in this image (no cargo/rustc), so this restatement is
 * pinned against the reference's own unit-test vectors (tests/test_oracle_kat.py):
 *   kmeans.rs:491-580 (8 tests), clusterc.rs:304-337 (5 tests), huf.rs:417-539, bit.rs:299-349.
 * Parity status per area:
 *   K-means / ColorCount / ColorPos arithmetic ....... pinned by the reference KATs above
 *   Huffman code lengths, bit packing, trie format ... pinned by huf.rs / bit.rs KATs
 *   Hilbert order (zhang_hilbert 0.1.1, un-vendored) . PARITY UNPINNED (see oracle_hilbert_xy)
 *   HashMap / BinaryHeap / thread_rng dependent order . unpinnable by construction; deterministic
 *                                                        stand-ins documented at each function
 */
#ifndef CNIIC_ORACLE_H
#define CNIIC_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* K-means modes */
#define ORACLE_MODE_EXACT 0    /* full scan of all centroids, integer d^2, keep-current on ties then lowest index */
#define ORACLE_MODE_VERBATIM 1 /* kmeans.rs as written: f64 sqrt distances + truncated sorted neighbour lists      */

/* tie rules (exact mode only) */
#define ORACLE_TIE_KEEP_CURRENT 0 /* kmeans.rs:350-378: a point moves only to a STRICTLY closer centroid           */
#define ORACLE_TIE_LOWEST_INDEX 1 /* pure lowest-index argmin, current cluster ignored                             */

/* status codes */
#define ORACLE_OK 0
#define ORACLE_ERR_BAD_ARG 1
#define ORACLE_ERR_TOO_FEW_POINTS 2   /* kmeans.rs:67-68  assert!(points_per_cluster > 0)                           */
#define ORACLE_ERR_TOO_FEW_ACTIVE 3   /* kmeans.rs:41-57  check_enough_active_clusters                              */

typedef struct {
    uint32_t iterations;       /* kmeans.rs:24-33 "#iterations"                                                    */
    uint32_t empty_events;     /* kmeans.rs:117-134 number of empty-cluster repairs over the whole run              */
    uint64_t moved_last;       /* points that changed cluster in the last assignment pass                          */
    uint64_t dist_evals;       /* point-centroid distance evaluations (work counter for the CPU baseline)          */
    uint64_t moved_total;      /* sum of moved over all passes                                                     */
} oracle_kmeans_stats;

/* Test-only 2-D point type of kmeans.rs:451-477 ((i32,i32), truncating i64 mean). pts = n x {x,y}. */
int oracle_kmeans_i32x2(const int32_t *pts, size_t n, size_t k, int mode, int tie_rule, uint32_t max_iters,
                        int32_t *out_centroids /*2k*/, uint32_t *out_assign /*n, nullable*/,
                        double *out_radii /*k certainty radii, nullable*/, oracle_kmeans_stats *stats);

/* ColorCount points (clusterc.rs:68-114): rgb = n x 3 bytes, counts = n weights (NULL => all 1, i.e. per-pixel
 * K-means).  Centroids out = k x 3 bytes.  out_cluster_weight = sum of counts per cluster (nullable). */
int oracle_kmeans_rgb(const uint8_t *rgb, const uint32_t *counts, size_t n, size_t k, int mode, int tie_rule,
                      uint32_t max_iters, uint8_t *out_centroids /*3k*/, uint64_t *out_cluster_weight /*k, nullable*/,
                      uint32_t *out_assign /*n, nullable*/, oracle_kmeans_stats *stats);

/* ColorPos points (clusterc.rs:148-153, 200-248): one point per pixel of a w x h raster RGB image (x fastest).
 * out_xy = k x {x,y}; out_rgb = k x 3. */
int oracle_kmeans_xyrgb(const uint8_t *rgb, uint32_t w, uint32_t h, size_t k, int mode, int tie_rule,
                        uint32_t max_iters, uint32_t *out_xy /*2k*/, uint8_t *out_rgb /*3k*/,
                        uint64_t *out_cluster_weight /*k, nullable*/, uint32_t *out_assign /*w*h, nullable*/,
                        oracle_kmeans_stats *stats);

/* Distances exactly as the reference computes them (f64). geom.rs:8-23 ; clusterc.rs:206-213 */
double oracle_dist_rgb(const uint8_t a[3], const uint8_t b[3]);
double oracle_dist_colorpos(uint32_t ax, uint32_t ay, const uint8_t a[3], uint32_t bx, uint32_t by, const uint8_t b[3]);
/* ColorCount::mean clusterc.rs:81-114; returns 0 if n==0 (None). out_count = resulting count field. */
int oracle_mean_colorcount(const uint8_t *rgb, const uint32_t *counts, size_t n, uint8_t out[3], uint32_t *out_count);
/* ColorPos::mean clusterc.rs:215-248 */
int oracle_mean_colorpos(const uint32_t *xy, const uint8_t *rgb, size_t n, uint32_t out_xy[2], uint8_t out_rgb[3]);

/* utils.rs:4-16 count_freqs over RGB pixels. Canonical order stand-in for HashMap iteration order:
 * ascending packed key r<<16|g<<8|b.  out_keys/out_counts must hold min(n, 2^24) entries. Returns #unique. */
size_t oracle_count_freqs_rgb(const uint8_t *rgb, size_t n, uint32_t *out_keys, uint64_t *out_counts);

/* clusterc.rs:18-52 cluster-colors front half: unique colours -> K-means (ColorCount) -> recolour.
 * out_rgb = recoloured image (3*w*h). */
int oracle_cluster_colors(const uint8_t *rgb, uint32_t w, uint32_t h, size_t k, int mode, int tie_rule, uint32_t max_iters,
                          uint8_t *out_rgb, uint8_t *out_centroids /*3k, nullable*/, oracle_kmeans_stats *stats);

/* clusterc.rs:179-186 voronoi decode fill: first centroid minimising (cx-x)^2+(cy-y)^2 in u32. */
void oracle_voronoi_fill(const uint32_t *cxy, const uint8_t *crgb, size_t k, uint32_t w, uint32_t h, uint8_t *out_rgb);

/* bench.rs:95-104 compute_error (MSE, f64 sum of sqrt(n)^2 in raster order). */
double oracle_mse(const uint8_t *a, const uint8_t *b, uint32_t w, uint32_t h);
/* exact integer sum of squared errors (what the GPU reduces) */
uint64_t oracle_sse(const uint8_t *a, const uint8_t *b, size_t npx);

/* hilbert.rs:40-43 iter(w,h): curve index i -> (x,y).  PARITY UNPINNED: the reference delegates to the
 * un-vendored crate zhang_hilbert 0.1.1 whose source is absent; this is our own deterministic pseudo-Hilbert
 * scan for arbitrary rectangles (recursive halving, classic Hilbert curve on 2^n squares, README.md:87-106
 * orientation).  out_xy = w*h x {x,y}. */
void oracle_hilbert_xy(uint32_t w, uint32_t h, uint32_t *out_xy);
void oracle_hilbert_d2xy(uint32_t w, uint32_t h, uint64_t d, uint32_t *x, uint32_t *y);

/* hilbertc.rs:449-477 DiffStream over hilbert::linearize(img): out = w*h x 3 i16. */
void oracle_delta(const uint8_t *rgb, uint32_t w, uint32_t h, int16_t *out);
/* hilbertc.rs:482-509 FromDiff + scatter (Delta::decode 417-431). */
int oracle_undelta(const int16_t *diff, uint32_t w, uint32_t h, uint8_t *out_rgb); /* 1 = the reference would panic (channel outside 0..255) */
/* hilbert.rs:34-38 linearize (gather along the curve). out = w*h x 3. */
void oracle_hilbert_gather(const uint8_t *rgb, uint32_t w, uint32_t h, uint8_t *out);

/* count_freqs over SignedColor symbols (huf.rs:30 as called from hilbertc.rs:409-414).
 * key = ((dr+255)*511 + (dg+255))*511 + (db+255); ascending key order. Returns #unique (cap entries written). */
size_t oracle_hist_delta(const int16_t *diff, size_t n, uint32_t *out_keys, uint64_t *out_counts, size_t cap);

/* hilbertc.rs:99-196 exact RLE along the Hilbert stream; records (count u8, r, g, b). Returns #records. */
size_t oracle_rle_exact(const uint8_t *stream_rgb, size_t n, uint8_t *out_counts, uint8_t *out_rgb);

/* ---- Huffman (huf.rs) + bit writer (bit.rs) + wire formats (ser.rs) ------------------------------------
 * Deterministic stand-in for BinaryHeap/HashMap order: leaves enter in ascending symbol order,
 * heap key = (freq, insertion sequence); merged nodes get the next sequence number; first popped = left.   */
/* code lengths only: syms 0..n-1 with freqs; out_len[n]. */
void oracle_huf_code_lengths(const uint64_t *freqs, size_t n, uint32_t *out_len);
/* generic symbol stream encode: symbols are ids 0..nsym-1 (ascending = canonical order), sym_bytes = serialized
 * leaf payload of each symbol (sym_size bytes each). Writes trie + payload (huf.rs:22-43). Returns bytes written
 * (or required size if cap too small). */
size_t oracle_huf_encode_ids(const uint32_t *stream, size_t n, size_t nsym, const uint8_t *sym_bytes, size_t sym_size,
                             uint8_t *out, size_t cap);
/* payload-only bit packing with given codes (huf.rs tests encode1/encode2; bit.rs:209-253). codes as '0'/'1' strings. */
size_t oracle_bitpack_codes(const uint32_t *stream, size_t n, const char *const *codes, uint8_t *out, size_t cap);

/* Whole-codec byte streams (codec.rs Codec::encode / decode). Return bytes written / required. decode returns 0 on ok. */
size_t oracle_encode_hufman(const uint8_t *rgb, uint32_t w, uint32_t h, uint8_t *out, size_t cap);        /* hufc.rs:12-17  */
int oracle_decode_hufman(const uint8_t *buf, size_t len, uint32_t *w, uint32_t *h, uint8_t *out_rgb, size_t cap_px); /* hufc.rs:19-40 */
size_t oracle_encode_delta(const uint8_t *rgb, uint32_t w, uint32_t h, uint8_t *out, size_t cap);         /* hilbertc.rs:405-415 */
int oracle_decode_delta(const uint8_t *buf, size_t len, uint32_t *w, uint32_t *h, uint8_t *out_rgb, size_t cap_px);  /* hilbertc.rs:417-431 */
size_t oracle_encode_voronoi(const uint8_t *rgb, uint32_t w, uint32_t h, size_t k, int mode, int tie_rule, uint32_t max_iters,
                             uint8_t *out, size_t cap);                                                     /* clusterc.rs:148-166 */
int oracle_decode_voronoi(const uint8_t *buf, size_t len, uint32_t *w, uint32_t *h, uint8_t *out_rgb, size_t cap_px); /* clusterc.rs:168-189 */
size_t oracle_encode_cluster_colors(const uint8_t *rgb, uint32_t w, uint32_t h, size_t k, int mode, int tie_rule,
                                    uint32_t max_iters, uint8_t *out, size_t cap);                          /* clusterc.rs:18-53 */
size_t oracle_encode_hilbert_rle(const uint8_t *rgb, uint32_t w, uint32_t h, uint8_t *out, size_t cap);   /* hilbertc.rs:26-38 */
int oracle_decode_hilbert_rle(const uint8_t *buf, size_t len, uint32_t *w, uint32_t *h, uint8_t *out_rgb, size_t cap_px);

#ifdef __cplusplus
}
#endif
#endif
