// stages.cuh -- device-level building blocks shared by stages.cu and codec.cu (all pointers are DEVICE pointers).
#pragma once
#include <vector>

#include "common.cuh"

// Unique colours of an image as weighted points, deduplicated AND Morton-sorted by one histogram + ordered compaction (stages.cu):
// d_pts = packed r | g<<8 | b<<16, d_wts = pixel count (clusterc.rs:23).  Ascending Morton code (r on the top bit of each triple) is
// also the canonical order of the unique colours of cluster-colors (header, oracle), so the list carries no permutation.
struct UniqueColours {
    uint32_t *d_pts = nullptr, *d_wts = nullptr;
    size_t u = 0;
};
int cniic_dev_unique_colours(cniic_ctx *ctx, const uint8_t *d_rgb, size_t n, UniqueColours *out);
void cniic_unique_colours_free(cniic_ctx *ctx, UniqueColours *uc);
struct cniic_kmeans;
int cniic_kmeans_open_unique(cniic_ctx *ctx, UniqueColours *uc, uint32_t k, int tie_rule, cniic_kmeans **out);  // takes over uc's arrays
void cniic_kmeans_sorted_view(cniic_kmeans *km, const uint32_t **pts_sorted, const uint16_t **assign_sorted);

int cniic_dev_dense_compact(cniic_ctx *ctx, const uint32_t *d_bins, size_t nbins, uint32_t **d_keys, unsigned long long **d_counts, size_t *n_unique);
int cniic_dev_hist_rgb_bins(cniic_ctx *ctx, const uint8_t *d_rgb, size_t n, uint32_t **d_bins);
int cniic_dev_hist_delta_bins(cniic_ctx *ctx, const uint8_t *d_rgb, uint32_t w, uint32_t h, uint32_t **d_bins, size_t *nbins);
int cniic_dev_hist_delta_bins_range(cniic_ctx *ctx, const uint8_t *d_rgb, uint32_t w, uint32_t h, unsigned long long i0, unsigned long long i1,
                                    uint32_t **d_bins, size_t *nbins);
int cniic_dev_recolor(cniic_ctx *ctx, const uint8_t *d_rgb, size_t n, const uint32_t *d_keys, const uint16_t *d_assign, size_t n_unique,
                      const int32_t *d_cen_i32, const uint8_t *d_cen_u8, uint32_t *d_lut, uint8_t *d_out);
int cniic_dev_keys_to_points(cniic_ctx *ctx, const uint32_t *d_keys, const unsigned long long *d_counts, size_t n, uint8_t *d_rgb, uint32_t *d_wts);
int cniic_dev_hilbert_gather(cniic_ctx *ctx, const uint8_t *d_rgb, uint32_t w, uint32_t h, uint8_t *d_out);
int cniic_dev_undelta(cniic_ctx *ctx, const int16_t *d_diff, uint32_t w, uint32_t h, uint8_t *d_out);
int cniic_dev_sse(cniic_ctx *ctx, const uint8_t *d_a, const uint8_t *d_b, size_t nbytes, uint64_t *out);
int cniic_dev_cluster_colors(cniic_ctx *ctx, const uint8_t *d_rgb, size_t n, uint32_t k, uint32_t max_iters, int tie_rule, uint8_t *d_out /* nullable */,
                             std::vector<int32_t> *cen_host, cniic_kmeans_stats *stats, size_t *n_unique /* nullable */);
int cniic_dev_huffman_pack(cniic_ctx *ctx, int src_kind, const void *d_src, size_t n, const uint32_t *d_keys, size_t nsym,
                           const std::vector<uint64_t> &codes, const std::vector<uint8_t> &lens, std::vector<uint8_t> *out);
// Parallel Huffman decoding (huffdec.cu; semantics of huf.rs:187-206): `payload` = HOST bytes, MSB-first bit string; trie as
// child[2*node] / child[2*node+1] (left = bit 0, right = bit 1; left < 0 marks a leaf), leaf_val = 8 bytes per node (the first
// sym_bytes are the symbol: 3 = colour bytes, 6 = [i16;3] little endian), node 0 = root.  Writes n * sym_bytes bytes to the
// DEVICE buffer d_out.  decoded == nullptr: CNIIC_ERR_DECODE when the payload holds fewer than n complete code words (hufc.rs:19-40
// returns None); decoded != nullptr: the first *decoded <= n symbols are written and the call succeeds (hilbertc.rs:417-431 zips the
// symbol iterator with the curve, a short payload just ends it).
int cniic_dev_huffman_decode(cniic_ctx *ctx, const uint8_t *payload, size_t len, const int32_t *child, const uint8_t *leaf_val, size_t nn,
                             int sym_bytes, size_t n, uint8_t *d_out, size_t *decoded);
// zero the pixels of curve indices [from, w*h) of a device image (pixels a short stream never reached: ImageBuffer::new is all zero)
int cniic_dev_zero_curve_tail(cniic_ctx *ctx, uint32_t w, uint32_t h, unsigned long long from, uint8_t *d_img);
// exact run-length coding along the Hilbert stream (hilbertc.rs:99-196 / 304-333), 12-byte records; see stages.cu
int cniic_dev_rle_encode(cniic_ctx *ctx, const uint8_t *d_lin, size_t n, std::vector<uint8_t> *out);
int cniic_dev_rle_decode(cniic_ctx *ctx, const uint8_t *recs, size_t len, uint32_t w, uint32_t h, uint8_t *d_out);
