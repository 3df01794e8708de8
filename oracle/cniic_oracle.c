/*
 * cniic_oracle.c -- CPU restatement of cniic's K-means / voronoi / pre-Huffman hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE (see cniic_oracle.h).  Plain C, single threaded,
 * compiled with -O2 -ffp-contract=off so f64 behaviour matches rustc (no FMA contraction, IEEE sqrt).
 *
 * Every function cites the reference file:line (relative to /root/reference/src) whose behaviour it
 * restates.  Nothing here is shared with, linked into or called by the product library.
 *
 * Deterministic stand-ins for behaviour the reference leaves to HashMap order / thread_rng / unstable
 * sort (unpinnable by construction):
 *   - unique colours enter K-means in ascending packed-RGB order (r<<16|g<<8|b)
 *   - neighbour lists are sorted with a STABLE sort on distance (equal distances keep previous order)
 *   - empty cluster e (the j-th empty one in ascending id order, j = 0,1,..) copies the member with the
 *     (j mod size)-th lowest point index of the heaviest cluster (largest total weight, lowest id on tie)
 *   - Huffman: leaves enter in ascending symbol order, heap ordered by (freq, creation sequence)
 */
#include "cniic_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------------ */
/* Distances and means                                                                              */
/* ------------------------------------------------------------------------------------------------ */

/* geom.rs:8-23 : i32 diff, diff*diff as f64, sum of three, sqrt */
double oracle_dist_rgb(const uint8_t a[3], const uint8_t b[3]) {
    double s = 0.0;
    for (int i = 0; i < 3; i++) {
        int32_t diff = (int32_t)a[i] - (int32_t)b[i];
        s += (double)(diff * diff);
    }
    return sqrt(s);
}

/* clusterc.rs:206-213 : wrapping u32 (a-b).pow(2) as f64 for x and y, plus (rgb dist).powi(2), sqrt */
double oracle_dist_colorpos(uint32_t ax, uint32_t ay, const uint8_t a[3], uint32_t bx, uint32_t by,
                            const uint8_t b[3]) {
    uint32_t dx = ax - bx; /* wraps like release-mode Rust */
    uint32_t dy = ay - by;
    double d = (double)(uint32_t)(dx * dx);
    d += (double)(uint32_t)(dy * dy);
    double c = oracle_dist_rgb(a, b);
    d += c * c; /* powi(2) */
    return sqrt(d);
}

/* kmeans.rs:453-461 test-only (i32,i32) distance */
static double dist_i32x2(const int32_t *a, const int32_t *b) {
    double d0 = (double)(a[0] - b[0]);
    double d1 = (double)(a[1] - b[1]);
    return sqrt(d0 * d0 + d1 * d1);
}

/* clusterc.rs:81-114 */
int oracle_mean_colorcount(const uint8_t *rgb, const uint32_t *counts, size_t n, uint8_t out[3],
                           uint32_t *out_count) {
    if (n == 0) return 0;
    if (n == 1) { /* behaves like clone() */
        memcpy(out, rgb, 3);
        if (out_count) *out_count = counts ? counts[0] : 1;
        return 1;
    }
    uint64_t s[3] = {0, 0, 0}, tw = 0;
    for (size_t i = 0; i < n; i++) {
        uint64_t c = counts ? counts[i] : 1;
        for (int j = 0; j < 3; j++) s[j] += (uint64_t)rgb[3 * i + j] * c;
        tw += c;
    }
    for (int j = 0; j < 3; j++) out[j] = (uint8_t)(s[j] / tw);
    if (out_count) *out_count = 1;
    return 1;
}

/* clusterc.rs:215-248 */
int oracle_mean_colorpos(const uint32_t *xy, const uint8_t *rgb, size_t n, uint32_t out_xy[2], uint8_t out_rgb[3]) {
    if (n == 0) return 0;
    uint64_t s[5] = {0, 0, 0, 0, 0};
    for (size_t i = 0; i < n; i++) {
        s[0] += xy[2 * i];
        s[1] += xy[2 * i + 1];
        for (int j = 0; j < 3; j++) s[2 + j] += rgb[3 * i + j];
    }
    out_xy[0] = (uint32_t)(s[0] / n);
    out_xy[1] = (uint32_t)(s[1] / n);
    for (int j = 0; j < 3; j++) out_rgb[j] = (uint8_t)(s[2 + j] / n);
    return 1;
}

/* ------------------------------------------------------------------------------------------------ */
/* Generic K-means engine over three point kinds (kmeans.rs:21-440)                                 */
/* ------------------------------------------------------------------------------------------------ */

enum { PT_I32X2 = 0, PT_RGB = 1, PT_XYRGB = 2 };

typedef struct {
    int kind;
    size_t n, k;
    const int32_t *p2;      /* PT_I32X2 */
    const uint8_t *rgb;     /* PT_RGB / PT_XYRGB */
    const uint32_t *counts; /* PT_RGB weights, may be NULL */
    uint32_t w, h;          /* PT_XYRGB */
    int dims;               /* 2, 3, 5 */
    int64_t *cen;           /* k x 5 centroid components */
    /* assignment: cluster -> vector of point ids */
    uint32_t **mem;
    size_t *mlen, *mcap;
    /* neighbour lists (verbatim mode) */
    uint32_t **nb_id;
    double **nb_d;
    size_t *nb_len, *nb_wm;
    oracle_kmeans_stats st;
} km_t;

static inline void pt_get(const km_t *K, size_t i, int64_t v[5]) {
    switch (K->kind) {
    case PT_I32X2:
        v[0] = K->p2[2 * i];
        v[1] = K->p2[2 * i + 1];
        break;
    case PT_RGB:
        v[0] = K->rgb[3 * i];
        v[1] = K->rgb[3 * i + 1];
        v[2] = K->rgb[3 * i + 2];
        break;
    default: /* clusterc.rs:150-152 : raster order, x fastest */
        v[0] = (int64_t)(i % K->w);
        v[1] = (int64_t)(i / K->w);
        v[2] = K->rgb[3 * i];
        v[3] = K->rgb[3 * i + 1];
        v[4] = K->rgb[3 * i + 2];
    }
}

/* f64 distance exactly as the reference computes it for each point kind */
static double dist_f64(const km_t *K, const int64_t a[5], const int64_t b[5]) {
    switch (K->kind) {
    case PT_I32X2: {
        int32_t x[2] = {(int32_t)a[0], (int32_t)a[1]}, y[2] = {(int32_t)b[0], (int32_t)b[1]};
        return dist_i32x2(x, y);
    }
    case PT_RGB: {
        uint8_t x[3] = {(uint8_t)a[0], (uint8_t)a[1], (uint8_t)a[2]}, y[3] = {(uint8_t)b[0], (uint8_t)b[1], (uint8_t)b[2]};
        return oracle_dist_rgb(x, y);
    }
    default: {
        uint8_t x[3] = {(uint8_t)a[2], (uint8_t)a[3], (uint8_t)a[4]}, y[3] = {(uint8_t)b[2], (uint8_t)b[3], (uint8_t)b[4]};
        return oracle_dist_colorpos((uint32_t)a[0], (uint32_t)a[1], x, (uint32_t)b[0], (uint32_t)b[1], y);
    }
    }
}

static inline int64_t dist2_int(const km_t *K, const int64_t a[5], const int64_t b[5]) {
    int64_t s = 0;
    for (int j = 0; j < K->dims; j++) {
        int64_t d = a[j] - b[j];
        s += d * d;
    }
    return s;
}

static void mem_push(km_t *K, size_t c, uint32_t id) {
    if (K->mlen[c] == K->mcap[c]) {
        size_t nc = K->mcap[c] ? K->mcap[c] * 2 : 8;
        K->mem[c] = (uint32_t *)realloc(K->mem[c], nc * sizeof(uint32_t));
        K->mcap[c] = nc;
    }
    K->mem[c][K->mlen[c]++] = id;
}

/* Point::mean over the members of cluster c. Returns 0 for an empty cluster (None). */
static int cluster_mean(const km_t *K, size_t c, int64_t out[5], uint64_t *weight) {
    size_t m = K->mlen[c];
    *weight = 0;
    if (m == 0) return 0;
    int64_t v[5];
    if (K->kind == PT_I32X2) { /* kmeans.rs:463-477 : i64 sums, truncating i64 division */
        int64_t s0 = 0, s1 = 0;
        for (size_t i = 0; i < m; i++) {
            pt_get(K, K->mem[c][i], v);
            s0 += v[0];
            s1 += v[1];
        }
        out[0] = (int32_t)(s0 / (int64_t)m);
        out[1] = (int32_t)(s1 / (int64_t)m);
        *weight = m;
        return 1;
    }
    if (K->kind == PT_RGB) { /* clusterc.rs:81-114 */
        if (m == 1) {
            pt_get(K, K->mem[c][0], out);
            *weight = K->counts ? K->counts[K->mem[c][0]] : 1;
            return 1;
        }
        uint64_t s[3] = {0, 0, 0}, tw = 0;
        for (size_t i = 0; i < m; i++) {
            uint32_t id = K->mem[c][i];
            uint64_t cnt = K->counts ? K->counts[id] : 1;
            for (int j = 0; j < 3; j++) s[j] += (uint64_t)K->rgb[3 * (size_t)id + j] * cnt;
            tw += cnt;
        }
        for (int j = 0; j < 3; j++) out[j] = (uint8_t)(s[j] / tw);
        *weight = tw;
        return 1;
    }
    /* clusterc.rs:215-248 */
    uint64_t s[5] = {0, 0, 0, 0, 0};
    for (size_t i = 0; i < m; i++) {
        pt_get(K, K->mem[c][i], v);
        for (int j = 0; j < 5; j++) s[j] += (uint64_t)v[j];
    }
    out[0] = (uint32_t)(s[0] / m);
    out[1] = (uint32_t)(s[1] / m);
    for (int j = 2; j < 5; j++) out[j] = (uint8_t)(s[j] / m);
    *weight = m;
    return 1;
}

static int cmp_u32(const void *a, const void *b) {
    uint32_t x = *(const uint32_t *)a, y = *(const uint32_t *)b;
    return (x > y) - (x < y);
}

/* kmeans.rs:110-143 compute_centroids, with the deterministic empty-cluster stand-in (header) */
static void update_centroids(km_t *K, uint64_t *weights) {
    size_t k = K->k;
    int64_t *nc = (int64_t *)malloc(k * 5 * sizeof(int64_t));
    uint8_t *empty = (uint8_t *)calloc(k, 1);
    size_t nempty = 0;
    for (size_t c = 0; c < k; c++) {
        if (!cluster_mean(K, c, nc + 5 * c, &weights[c])) {
            empty[c] = 1;
            nempty++;
        }
    }
    if (nempty) {
        size_t victim = 0;
        uint64_t best = 0;
        for (size_t c = 0; c < k; c++)
            if (weights[c] > best) {
                best = weights[c];
                victim = c;
            }
        size_t m = K->mlen[victim];
        uint32_t *sorted = (uint32_t *)malloc(m * sizeof(uint32_t));
        memcpy(sorted, K->mem[victim], m * sizeof(uint32_t));
        qsort(sorted, m, sizeof(uint32_t), cmp_u32);
        size_t j = 0;
        for (size_t c = 0; c < k; c++) {
            if (!empty[c]) continue;
            pt_get(K, sorted[j % m], nc + 5 * c); /* fake_clone(stolen), kmeans.rs:96-99,133 */
            j++;
            K->st.empty_events++;
        }
        free(sorted);
    }
    memcpy(K->cen, nc, k * 5 * sizeof(int64_t));
    free(nc);
    free(empty);
}

/* ---- neighbour lists (kmeans.rs:150-323) ---- */

static void nb_new(km_t *K, size_t s) { /* NeighbouringCentroids::new, kmeans.rs:158-172 */
    size_t k = K->k, j = 0;
    for (size_t d = 0; d < k; d++)
        if (d != s) {
            K->nb_id[s][j] = (uint32_t)d;
            K->nb_d[s][j] = 0.0;
            j++;
        }
    K->nb_len[s] = k - 1;
    K->nb_wm[s] = k - 1;
}

/* stable merge sort on distance (stand-in for sort_unstable_by, kmeans.rs:183-188) */
static void nb_sort(uint32_t *id, double *d, size_t n, uint32_t *tid, double *td) {
    if (n < 2) return;
    /* already sorted? (the common case from iteration 2 on) */
    int sorted = 1;
    for (size_t i = 1; i < n; i++)
        if (d[i] < d[i - 1]) {
            sorted = 0;
            break;
        }
    if (sorted) return;
    for (size_t width = 1; width < n; width *= 2) {
        for (size_t lo = 0; lo < n; lo += 2 * width) {
            size_t mid = lo + width < n ? lo + width : n, hi = lo + 2 * width < n ? lo + 2 * width : n;
            size_t a = lo, b = mid, o = lo;
            while (a < mid && b < hi) {
                if (d[b] < d[a]) {
                    tid[o] = id[b];
                    td[o++] = d[b++];
                } else {
                    tid[o] = id[a];
                    td[o++] = d[a++];
                }
            }
            while (a < mid) {
                tid[o] = id[a];
                td[o++] = d[a++];
            }
            while (b < hi) {
                tid[o] = id[b];
                td[o++] = d[b++];
            }
        }
        memcpy(id, tid, n * sizeof(uint32_t));
        memcpy(d, td, n * sizeof(double));
    }
}

/* kmeans.rs:299-318 compute_neighbours (update_distances 174-181, sort 183-188, resize 191-248) */
static void compute_neighbours(km_t *K) {
    size_t k = K->k;
    uint32_t *tid = (uint32_t *)malloc((k ? k : 1) * sizeof(uint32_t));
    double *td = (double *)malloc((k ? k : 1) * sizeof(double));
    size_t lower = (size_t)sqrtf((float)k);
    size_t upper = k - 1;
    for (size_t s = 0; s < k; s++) {
        for (;;) {
            for (size_t j = 0; j < K->nb_len[s]; j++)
                K->nb_d[s][j] = dist_f64(K, K->cen + 5 * s, K->cen + 5 * K->nb_id[s][j]);
            nb_sort(K->nb_id[s], K->nb_d[s], K->nb_len[s], tid, td);
            size_t dyn = 2 * K->nb_wm[s];
            size_t nn = dyn > lower ? dyn : lower;
            if (nn > upper) nn = upper;
            if (nn * 3 / 4 <= K->nb_len[s]) {
                if (nn < K->nb_len[s]) K->nb_len[s] = nn; /* truncate never grows */
                K->nb_wm[s] = 0;
                break;
            }
            nb_new(K, s);
        }
    }
    free(tid);
    free(td);
}

static inline double certainty_radius(const km_t *K, size_t c) { /* kmeans.rs:257-259 */
    return K->nb_len[c] ? K->nb_d[c][0] / 2.0 : INFINITY;
}

/* kmeans.rs:330-416 assign_points. mode exact: full integer scan. Returns #moved. */
static uint64_t assign_points(km_t *K, int mode, int tie_rule) {
    size_t k = K->k;
    uint32_t **old = K->mem;
    size_t *olen = K->mlen, *ocap = K->mcap;
    K->mem = (uint32_t **)calloc(k, sizeof(uint32_t *));
    K->mlen = (size_t *)calloc(k, sizeof(size_t));
    K->mcap = (size_t *)calloc(k, sizeof(size_t));
    uint64_t moved = 0;
    int64_t x[5];
    for (size_t cci = 0; cci < k; cci++) {
        for (size_t pi = 0; pi < olen[cci]; pi++) {
            uint32_t id = old[cci][pi];
            pt_get(K, id, x);
            size_t best = cci;
            if (mode == ORACLE_MODE_VERBATIM) {
                double min_dist = dist_f64(K, K->cen + 5 * cci, x);
                K->st.dist_evals++;
                if (!(min_dist <= certainty_radius(K, cci))) {
                    double cutoff = 2.0 * min_dist;
                    size_t next_pos = 0;
                    for (;;) {
                        /* IterAndRecord::next, kmeans.rs:270-275 */
                        size_t pos = next_pos++;
                        if (pos >= K->nb_len[cci]) break;
                        if (K->nb_d[cci][pos] > cutoff) break;
                        size_t t = K->nb_id[cci][pos];
                        double td = dist_f64(K, K->cen + 5 * t, x);
                        K->st.dist_evals++;
                        if (td < min_dist) {
                            min_dist = td;
                            best = t;
                        }
                    }
                    K->nb_wm[cci] = next_pos - 1; /* Drop, kmeans.rs:277-281 */
                }
            } else {
                int64_t bd;
                if (tie_rule == ORACLE_TIE_KEEP_CURRENT) {
                    bd = dist2_int(K, K->cen + 5 * cci, x);
                } else {
                    best = 0;
                    bd = dist2_int(K, K->cen, x);
                }
                for (size_t t = 0; t < k; t++) {
                    int64_t d = dist2_int(K, K->cen + 5 * t, x);
                    if (d < bd) {
                        bd = d;
                        best = t;
                    }
                }
                K->st.dist_evals += k;
            }
            mem_push(K, best, id);
            if (best != cci) moved++;
        }
    }
    for (size_t c = 0; c < k; c++) free(old[c]);
    free(old);
    free(olen);
    free(ocap);
    return moved;
}

static int km_run(km_t *K, int mode, int tie_rule, uint32_t max_iters, uint64_t *weights, uint32_t *out_assign,
                  double *out_radii) {
    size_t n = K->n, k = K->k;
    if (k == 0) return ORACLE_ERR_BAD_ARG;
    size_t ppc = n / k;
    if (ppc == 0) return ORACLE_ERR_TOO_FEW_POINTS; /* kmeans.rs:67-68 */
    memset(&K->st, 0, sizeof(K->st));
    K->cen = (int64_t *)calloc(k * 5, sizeof(int64_t));
    K->mem = (uint32_t **)calloc(k, sizeof(uint32_t *));
    K->mlen = (size_t *)calloc(k, sizeof(size_t));
    K->mcap = (size_t *)calloc(k, sizeof(size_t));
    /* init_assignment kmeans.rs:61-78 : split_off chunks from the tail, remainder to the last cluster */
    size_t len = n;
    for (size_t c = 0; c + 1 < k; c++) {
        size_t at = len - ppc;
        for (size_t i = at; i < len; i++) mem_push(K, c, (uint32_t)i);
        len = at;
    }
    for (size_t i = 0; i < len; i++) mem_push(K, k - 1, (uint32_t)i);
    /* init_centroids kmeans.rs:101-108 : first point of each chunk */
    for (size_t c = 0; c < k; c++) pt_get(K, K->mem[c][0], K->cen + 5 * c);
    int verbatim = (mode == ORACLE_MODE_VERBATIM);
    if (verbatim || out_radii) {
        K->nb_id = (uint32_t **)calloc(k, sizeof(uint32_t *));
        K->nb_d = (double **)calloc(k, sizeof(double *));
        K->nb_len = (size_t *)calloc(k, sizeof(size_t));
        K->nb_wm = (size_t *)calloc(k, sizeof(size_t));
        for (size_t s = 0; s < k; s++) {
            K->nb_id[s] = (uint32_t *)malloc((k > 1 ? k - 1 : 1) * sizeof(uint32_t));
            K->nb_d[s] = (double *)malloc((k > 1 ? k - 1 : 1) * sizeof(double));
            nb_new(K, s);
        }
        compute_neighbours(K); /* init_neighbours kmeans.rs:288-295 */
    }
    uint64_t *wloc = weights ? weights : (uint64_t *)calloc(k, sizeof(uint64_t));
    int changed = 1;
    uint32_t it = 0;
    while (changed && (max_iters == 0 || it < max_iters)) { /* kmeans.rs:25-32 */
        uint64_t moved = assign_points(K, mode, tie_rule);
        changed = moved != 0;
        K->st.moved_last = moved;
        K->st.moved_total += moved;
        update_centroids(K, wloc);
        if (verbatim || out_radii) compute_neighbours(K);
        it++;
    }
    K->st.iterations = it;
    if (out_assign)
        for (size_t c = 0; c < k; c++)
            for (size_t i = 0; i < K->mlen[c]; i++) out_assign[K->mem[c][i]] = (uint32_t)c;
    if (out_radii)
        for (size_t c = 0; c < k; c++) out_radii[c] = certainty_radius(K, c);
    /* check_enough_active_clusters kmeans.rs:41-57 */
    size_t active = 0;
    for (size_t c = 0; c < k; c++) active += K->mlen[c] > 0;
    size_t min_cc = (size_t)(0.99 * (double)k);
    if (n < min_cc) min_cc = n;
    int rc = active >= min_cc ? ORACLE_OK : ORACLE_ERR_TOO_FEW_ACTIVE;
    if (!weights) free(wloc);
    return rc;
}

static void km_free(km_t *K) {
    if (K->mem)
        for (size_t c = 0; c < K->k; c++) free(K->mem[c]);
    free(K->mem);
    free(K->mlen);
    free(K->mcap);
    if (K->nb_id)
        for (size_t c = 0; c < K->k; c++) {
            free(K->nb_id[c]);
            free(K->nb_d[c]);
        }
    free(K->nb_id);
    free(K->nb_d);
    free(K->nb_len);
    free(K->nb_wm);
    free(K->cen);
}

int oracle_kmeans_i32x2(const int32_t *pts, size_t n, size_t k, int mode, int tie_rule, uint32_t max_iters,
                        int32_t *out_centroids, uint32_t *out_assign, double *out_radii, oracle_kmeans_stats *stats) {
    km_t K;
    memset(&K, 0, sizeof(K));
    K.kind = PT_I32X2;
    K.n = n;
    K.k = k;
    K.p2 = pts;
    K.dims = 2;
    int rc = km_run(&K, mode, tie_rule, max_iters, NULL, out_assign, out_radii);
    if (rc != ORACLE_ERR_BAD_ARG && rc != ORACLE_ERR_TOO_FEW_POINTS && out_centroids)
        for (size_t c = 0; c < k; c++) {
            out_centroids[2 * c] = (int32_t)K.cen[5 * c];
            out_centroids[2 * c + 1] = (int32_t)K.cen[5 * c + 1];
        }
    if (stats) *stats = K.st;
    km_free(&K);
    return rc;
}

int oracle_kmeans_rgb(const uint8_t *rgb, const uint32_t *counts, size_t n, size_t k, int mode, int tie_rule,
                      uint32_t max_iters, uint8_t *out_centroids, uint64_t *out_cluster_weight, uint32_t *out_assign,
                      oracle_kmeans_stats *stats) {
    km_t K;
    memset(&K, 0, sizeof(K));
    K.kind = PT_RGB;
    K.n = n;
    K.k = k;
    K.rgb = rgb;
    K.counts = counts;
    K.dims = 3;
    int rc = km_run(&K, mode, tie_rule, max_iters, out_cluster_weight, out_assign, NULL);
    if (rc != ORACLE_ERR_BAD_ARG && rc != ORACLE_ERR_TOO_FEW_POINTS && out_centroids)
        for (size_t c = 0; c < k; c++)
            for (int j = 0; j < 3; j++) out_centroids[3 * c + j] = (uint8_t)K.cen[5 * c + j];
    if (stats) *stats = K.st;
    km_free(&K);
    return rc;
}

int oracle_kmeans_xyrgb(const uint8_t *rgb, uint32_t w, uint32_t h, size_t k, int mode, int tie_rule,
                        uint32_t max_iters, uint32_t *out_xy, uint8_t *out_rgb, uint64_t *out_cluster_weight,
                        uint32_t *out_assign, oracle_kmeans_stats *stats) {
    km_t K;
    memset(&K, 0, sizeof(K));
    K.kind = PT_XYRGB;
    K.n = (size_t)w * h;
    K.k = k;
    K.rgb = rgb;
    K.w = w;
    K.h = h;
    K.dims = 5;
    int rc = km_run(&K, mode, tie_rule, max_iters, out_cluster_weight, out_assign, NULL);
    if (rc != ORACLE_ERR_BAD_ARG && rc != ORACLE_ERR_TOO_FEW_POINTS)
        for (size_t c = 0; c < k; c++) {
            if (out_xy) {
                out_xy[2 * c] = (uint32_t)K.cen[5 * c];
                out_xy[2 * c + 1] = (uint32_t)K.cen[5 * c + 1];
            }
            if (out_rgb)
                for (int j = 0; j < 3; j++) out_rgb[3 * c + j] = (uint8_t)K.cen[5 * c + 2 + j];
        }
    if (stats) *stats = K.st;
    km_free(&K);
    return rc;
}

/* ------------------------------------------------------------------------------------------------ */
/* count_freqs, cluster-colors, voronoi fill, MSE                                                   */
/* ------------------------------------------------------------------------------------------------ */

/* utils.rs:4-16 ; ascending packed key is the canonical stand-in for HashMap order */
size_t oracle_count_freqs_rgb(const uint8_t *rgb, size_t n, uint32_t *out_keys, uint64_t *out_counts) {
    uint32_t *bins = (uint32_t *)calloc(1u << 24, sizeof(uint32_t));
    for (size_t i = 0; i < n; i++) bins[((uint32_t)rgb[3 * i] << 16) | ((uint32_t)rgb[3 * i + 1] << 8) | rgb[3 * i + 2]]++;
    size_t u = 0;
    for (uint32_t key = 0; key < (1u << 24); key++)
        if (bins[key]) {
            if (out_keys) out_keys[u] = key;
            if (out_counts) out_counts[u] = bins[key];
            u++;
        }
    free(bins);
    return u;
}

/* clusterc.rs:18-52 */
/* 8 bits -> every third bit; Morton code of a colour with r on the most significant bit of every triple */
static uint32_t spread3(uint32_t x) {
    uint32_t r = 0;
    for (int i = 0; i < 8; i++) r |= ((x >> i) & 1u) << (3 * i);
    return r;
}
static uint32_t morton_of_key(uint32_t key) { return (spread3(key >> 16) << 2) | (spread3((key >> 8) & 0xff) << 1) | spread3(key & 0xff); }

typedef struct { uint32_t morton, key; uint64_t count; } ucol_t;
static int cmp_ucol(const void *a, const void *b) {
    uint32_t x = ((const ucol_t *)a)->morton, y = ((const ucol_t *)b)->morton;
    return (x > y) - (x < y);
}

/* clusterc.rs:18-53.  The reference's point list is `HashMap::into_iter()` order (clusterc.rs:21-27: random); the
   deterministic stand-in is ASCENDING MORTON CODE of (r, g, b), r on the most significant bit of every bit triple: an order with
   colour-space locality at every scale, which is also the order the GPU path scans (its histogram bins are Morton-indexed, so
   their compaction is this list).  The chunked init (kmeans.rs:61-108) and the empty-cluster rule index this list. */
int oracle_cluster_colors(const uint8_t *rgb, uint32_t w, uint32_t h, size_t k, int mode, int tie_rule, uint32_t max_iters,
                          uint8_t *out_rgb, uint8_t *out_centroids, oracle_kmeans_stats *stats) {
    size_t n = (size_t)w * h;
    size_t cap = n < (1u << 24) ? n : (1u << 24);
    uint32_t *keys = (uint32_t *)malloc((cap ? cap : 1) * sizeof(uint32_t));
    uint64_t *cnt64 = (uint64_t *)malloc((cap ? cap : 1) * sizeof(uint64_t));
    size_t u = oracle_count_freqs_rgb(rgb, n, keys, cnt64);
    ucol_t *uc = (ucol_t *)malloc((u ? u : 1) * sizeof(ucol_t));
    for (size_t i = 0; i < u; i++) { uc[i].morton = morton_of_key(keys[i]); uc[i].key = keys[i]; uc[i].count = cnt64[i]; }
    qsort(uc, u, sizeof(ucol_t), cmp_ucol);
    uint8_t *urgb = (uint8_t *)malloc(3 * (u ? u : 1));
    uint32_t *cnt = (uint32_t *)malloc((u ? u : 1) * sizeof(uint32_t));
    for (size_t i = 0; i < u; i++) {
        keys[i] = uc[i].key;
        urgb[3 * i] = (uint8_t)(keys[i] >> 16);
        urgb[3 * i + 1] = (uint8_t)(keys[i] >> 8);
        urgb[3 * i + 2] = (uint8_t)keys[i];
        cnt[i] = (uint32_t)uc[i].count; /* clusterc.rs:23 "count as u32" */
    }
    free(uc);
    uint8_t *cen = (uint8_t *)malloc(3 * (k ? k : 1));
    uint32_t *asg = (uint32_t *)malloc((u ? u : 1) * sizeof(uint32_t));
    int rc = oracle_kmeans_rgb(urgb, cnt, u, k, mode, tie_rule, max_iters, cen, NULL, asg, stats);
    if (rc == ORACLE_OK || rc == ORACLE_ERR_TOO_FEW_ACTIVE) {
        /* colour -> centroid colour lookup (clusterc.rs:31-47) */
        uint32_t *lut = (uint32_t *)malloc((1u << 24) * sizeof(uint32_t));
        for (size_t i = 0; i < u; i++) lut[keys[i]] = asg[i];
        if (out_rgb)
            for (size_t i = 0; i < n; i++) {
                uint32_t key = ((uint32_t)rgb[3 * i] << 16) | ((uint32_t)rgb[3 * i + 1] << 8) | rgb[3 * i + 2];
                memcpy(out_rgb + 3 * i, cen + 3 * lut[key], 3);
            }
        if (out_centroids) memcpy(out_centroids, cen, 3 * k);
        free(lut);
    }
    free(keys);
    free(cnt64);
    free(urgb);
    free(cnt);
    free(cen);
    free(asg);
    return rc;
}

/* clusterc.rs:179-186 : min_by_key returns the FIRST minimum; u32 wrapping sub and pow(2) */
void oracle_voronoi_fill(const uint32_t *cxy, const uint8_t *crgb, size_t k, uint32_t w, uint32_t h, uint8_t *out_rgb) {
    for (uint32_t y = 0; y < h; y++)
        for (uint32_t x = 0; x < w; x++) {
            size_t best = 0;
            uint32_t bd = 0;
            for (size_t c = 0; c < k; c++) {
                uint32_t dx = cxy[2 * c] - x, dy = cxy[2 * c + 1] - y;
                uint32_t d = dx * dx + dy * dy;
                if (c == 0 || d < bd) {
                    bd = d;
                    best = c;
                }
            }
            memcpy(out_rgb + 3 * ((size_t)y * w + x), crgb + 3 * best, 3);
        }
}

/* bench.rs:95-104 */
double oracle_mse(const uint8_t *a, const uint8_t *b, uint32_t w, uint32_t h) {
    double s = 0.0;
    size_t n = (size_t)w * h;
    for (size_t i = 0; i < n; i++) {
        double d = oracle_dist_rgb(a + 3 * i, b + 3 * i);
        s += d * d;
    }
    return s / (double)(uint32_t)(w * h);
}

uint64_t oracle_sse(const uint8_t *a, const uint8_t *b, size_t npx) {
    uint64_t s = 0;
    for (size_t i = 0; i < 3 * npx; i++) {
        int32_t d = (int32_t)a[i] - (int32_t)b[i];
        s += (uint64_t)(d * d);
    }
    return s;
}

/* ------------------------------------------------------------------------------------------------ */
/* Pseudo-Hilbert scan for arbitrary rectangles (hilbert.rs:40-43) -- PARITY UNPINNED               */
/* ------------------------------------------------------------------------------------------------ */
/* zhang_hilbert 0.1.1 is not vendored; this is a generalized Hilbert scan by recursive halving that  */
/* reduces to the classic Hilbert curve on 2^n squares, first step along +x (README.md:87-106).       */

typedef struct {
    uint32_t *out;
    size_t pos;
} hil_sink;

static int sgn(int64_t v) { return (v > 0) - (v < 0); }
static int64_t iabs64(int64_t v) { return v < 0 ? -v : v; }
/* halve toward zero... the scan uses floor division semantics on signed extents */
static int64_t half_floor(int64_t v) { return v >= 0 ? v / 2 : -((-v + 1) / 2); }

static void hil_rec(hil_sink *s, int64_t x, int64_t y, int64_t ax, int64_t ay, int64_t bx, int64_t by) {
    int64_t w = iabs64(ax + ay), h = iabs64(bx + by);
    int dax = sgn(ax), day = sgn(ay), dbx = sgn(bx), dby = sgn(by);
    if (h == 1) {
        for (int64_t i = 0; i < w; i++) {
            s->out[2 * s->pos] = (uint32_t)x;
            s->out[2 * s->pos + 1] = (uint32_t)y;
            s->pos++;
            x += dax;
            y += day;
        }
        return;
    }
    if (w == 1) {
        for (int64_t i = 0; i < h; i++) {
            s->out[2 * s->pos] = (uint32_t)x;
            s->out[2 * s->pos + 1] = (uint32_t)y;
            s->pos++;
            x += dbx;
            y += dby;
        }
        return;
    }
    int64_t ax2 = half_floor(ax), ay2 = half_floor(ay), bx2 = half_floor(bx), by2 = half_floor(by);
    int64_t w2 = iabs64(ax2 + ay2), h2 = iabs64(bx2 + by2);
    if (2 * w > 3 * h) {
        if ((w2 % 2) && w > 2) {
            ax2 += dax;
            ay2 += day;
        }
        hil_rec(s, x, y, ax2, ay2, bx, by);
        hil_rec(s, x + ax2, y + ay2, ax - ax2, ay - ay2, bx, by);
    } else {
        if ((h2 % 2) && h > 2) {
            bx2 += dbx;
            by2 += dby;
        }
        hil_rec(s, x, y, bx2, by2, ax2, ay2);
        hil_rec(s, x + bx2, y + by2, ax, ay, bx - bx2, by - by2);
        hil_rec(s, x + (ax - dax) + (bx2 - dbx), y + (ay - day) + (by2 - dby), -bx2, -by2, -(ax - ax2), -(ay - ay2));
    }
}

void oracle_hilbert_xy(uint32_t w, uint32_t h, uint32_t *out_xy) {
    if (w == 0 || h == 0) return;
    hil_sink s = {out_xy, 0};
    if (w >= h)
        hil_rec(&s, 0, 0, (int64_t)w, 0, 0, (int64_t)h);
    else
        hil_rec(&s, 0, 0, 0, (int64_t)h, (int64_t)w, 0);
}

void oracle_hilbert_d2xy(uint32_t w, uint32_t h, uint64_t d, uint32_t *x, uint32_t *y) {
    /* small helper for spot checks: enumerate (only used on small rectangles by the tests) */
    uint32_t *xy = (uint32_t *)malloc((size_t)w * h * 2 * sizeof(uint32_t));
    oracle_hilbert_xy(w, h, xy);
    *x = xy[2 * d];
    *y = xy[2 * d + 1];
    free(xy);
}

/* hilbert.rs:34-38 */
void oracle_hilbert_gather(const uint8_t *rgb, uint32_t w, uint32_t h, uint8_t *out) {
    size_t n = (size_t)w * h;
    uint32_t *xy = (uint32_t *)malloc((n ? n : 1) * 2 * sizeof(uint32_t));
    oracle_hilbert_xy(w, h, xy);
    for (size_t i = 0; i < n; i++) memcpy(out + 3 * i, rgb + 3 * ((size_t)xy[2 * i + 1] * w + xy[2 * i]), 3);
    free(xy);
}

/* hilbertc.rs:449-477 : diff = cur - last per channel (i16), last starts at [0,0,0] */
void oracle_delta(const uint8_t *rgb, uint32_t w, uint32_t h, int16_t *out) {
    size_t n = (size_t)w * h;
    uint8_t *lin = (uint8_t *)malloc(3 * (n ? n : 1));
    oracle_hilbert_gather(rgb, w, h, lin);
    int16_t last[3] = {0, 0, 0};
    for (size_t i = 0; i < n; i++)
        for (int j = 0; j < 3; j++) {
            int16_t cur = lin[3 * i + j];
            out[3 * i + j] = (int16_t)(cur - last[j]);
            last[j] = cur;
        }
    free(lin);
}

/* hilbertc.rs:482-509 + 417-431.  FromDiff::next does `new_signed.try_into().unwrap()` (hilbertc.rs:503-506, conversion
   525-535: i16 -> u8 per channel): the reference PANICS at the first reconstructed channel outside 0..255.  Returns 1 then
   (the i16 addition itself wraps in a release build, hilbertc.rs:549-558; the first offender is reached before any wrap). */
int oracle_undelta(const int16_t *diff, uint32_t w, uint32_t h, uint8_t *out_rgb) {
    size_t n = (size_t)w * h;
    uint32_t *xy = (uint32_t *)malloc((n ? n : 1) * 2 * sizeof(uint32_t));
    oracle_hilbert_xy(w, h, xy);
    int16_t last[3] = {0, 0, 0};
    int panicked = 0;
    for (size_t i = 0; i < n && !panicked; i++) {
        uint8_t *px = out_rgb + 3 * ((size_t)xy[2 * i + 1] * w + xy[2 * i]);
        for (int j = 0; j < 3; j++) {
            last[j] = (int16_t)((uint16_t)last[j] + (uint16_t)diff[3 * i + j]);
            if (last[j] < 0 || last[j] > 255) panicked = 1;
            px[j] = (uint8_t)last[j];
        }
    }
    free(xy);
    return panicked;
}

static inline uint32_t delta_key(const int16_t *d) {
    return (uint32_t)(((d[0] + 255) * 511 + (d[1] + 255)) * 511 + (d[2] + 255));
}

static int cmp_u32p(const void *a, const void *b) { return cmp_u32(a, b); }

size_t oracle_hist_delta(const int16_t *diff, size_t n, uint32_t *out_keys, uint64_t *out_counts, size_t cap) {
    uint32_t *keys = (uint32_t *)malloc((n ? n : 1) * sizeof(uint32_t));
    for (size_t i = 0; i < n; i++) keys[i] = delta_key(diff + 3 * i);
    qsort(keys, n, sizeof(uint32_t), cmp_u32p);
    size_t u = 0;
    for (size_t i = 0; i < n;) {
        size_t j = i;
        while (j < n && keys[j] == keys[i]) j++;
        if (u < cap) {
            if (out_keys) out_keys[u] = keys[i];
            if (out_counts) out_counts[u] = j - i;
        }
        u++;
        i = j;
    }
    free(keys);
    return u;
}

/* hilbertc.rs:99-196 : run cap RepCount::MAX = 255 */
size_t oracle_rle_exact(const uint8_t *s, size_t n, uint8_t *out_counts, uint8_t *out_rgb) {
    size_t r = 0, i = 0;
    while (i < n) {
        size_t count = 1;
        size_t j = i + 1;
        while (j < n && memcmp(s + 3 * j, s + 3 * i, 3) == 0) {
            count++;
            j++;
            if (count == 255) break;
        }
        if (out_counts) out_counts[r] = (uint8_t)count;
        if (out_rgb) memcpy(out_rgb + 3 * r, s + 3 * i, 3);
        r++;
        i = j;
    }
    return r;
}

/* ------------------------------------------------------------------------------------------------ */
/* Huffman (huf.rs), bit writer (bit.rs), wire formats (ser.rs)                                     */
/* ------------------------------------------------------------------------------------------------ */

typedef struct {
    uint64_t freq;
    uint32_t seq;
    int32_t left, right; /* -1 for leaf */
    uint32_t sym;
} hnode;

typedef struct {
    hnode *nodes;
    size_t nn;
    int32_t root;
} htree;

static int heap_less(const hnode *nodes, int32_t a, int32_t b) {
    if (nodes[a].freq != nodes[b].freq) return nodes[a].freq < nodes[b].freq;
    return nodes[a].seq < nodes[b].seq;
}

static void heap_push(int32_t *heap, size_t *hn, const hnode *nodes, int32_t v) {
    size_t i = (*hn)++;
    heap[i] = v;
    while (i > 0) {
        size_t p = (i - 1) / 2;
        if (!heap_less(nodes, heap[i], heap[p])) break;
        int32_t t = heap[i];
        heap[i] = heap[p];
        heap[p] = t;
        i = p;
    }
}

static int32_t heap_pop(int32_t *heap, size_t *hn, const hnode *nodes) {
    int32_t top = heap[0];
    heap[0] = heap[--(*hn)];
    size_t i = 0;
    for (;;) {
        size_t l = 2 * i + 1, r = l + 1, m = i;
        if (l < *hn && heap_less(nodes, heap[l], heap[m])) m = l;
        if (r < *hn && heap_less(nodes, heap[r], heap[m])) m = r;
        if (m == i) break;
        int32_t t = heap[i];
        heap[i] = heap[m];
        heap[m] = t;
        i = m;
    }
    return top;
}

/* huf.rs:58-117 build: pop left (smallest), pop right, compose, push */
static htree huf_build(const uint64_t *freqs, size_t n) {
    htree T;
    T.nodes = (hnode *)malloc((2 * n) * sizeof(hnode));
    T.nn = 0;
    int32_t *heap = (int32_t *)malloc(n * sizeof(int32_t));
    size_t hn = 0;
    for (size_t i = 0; i < n; i++) {
        hnode nd = {freqs[i], (uint32_t)T.nn, -1, -1, (uint32_t)i};
        T.nodes[T.nn] = nd;
        heap_push(heap, &hn, T.nodes, (int32_t)T.nn);
        T.nn++;
    }
    while (hn > 1) {
        int32_t l = heap_pop(heap, &hn, T.nodes);
        int32_t r = heap_pop(heap, &hn, T.nodes);
        hnode nd = {T.nodes[l].freq + T.nodes[r].freq, (uint32_t)T.nn, l, r, 0};
        T.nodes[T.nn] = nd;
        heap_push(heap, &hn, T.nodes, (int32_t)T.nn);
        T.nn++;
    }
    T.root = heap_pop(heap, &hn, T.nodes);
    free(heap);
    return T;
}

/* codes: left edge = 0, right edge = 1 (huf.rs:196-197, 259-273). code bits MSB-first in a u64 + len (len<=64 asserted by tests sizes) */
typedef struct {
    uint8_t *bits; /* one byte per bit, for arbitrary lengths */
    uint32_t *len;
    size_t *off;
    size_t total;
} hcodes;

static void huf_walk(const htree *T, int32_t node, uint8_t *path, uint32_t depth, hcodes *C, int pass) {
    const hnode *nd = &T->nodes[node];
    if (nd->left < 0) {
        if (pass == 0) {
            C->len[nd->sym] = depth;
        } else {
            memcpy(C->bits + C->off[nd->sym], path, depth);
        }
        return;
    }
    path[depth] = 0;
    huf_walk(T, nd->left, path, depth + 1, C, pass);
    path[depth] = 1;
    huf_walk(T, nd->right, path, depth + 1, C, pass);
}

static hcodes huf_codes(const htree *T, size_t n) {
    hcodes C;
    C.len = (uint32_t *)calloc(n, sizeof(uint32_t));
    C.off = (size_t *)calloc(n + 1, sizeof(size_t));
    uint8_t *path = (uint8_t *)malloc(n + 1);
    huf_walk(T, T->root, path, 0, &C, 0);
    size_t tot = 0;
    for (size_t i = 0; i < n; i++) {
        C.off[i] = tot;
        tot += C.len[i];
    }
    C.off[n] = tot;
    C.total = tot;
    C.bits = (uint8_t *)malloc(tot ? tot : 1);
    huf_walk(T, T->root, path, 0, &C, 1);
    free(path);
    return C;
}

static void hcodes_free(hcodes *C) {
    free(C->bits);
    free(C->len);
    free(C->off);
}

void oracle_huf_code_lengths(const uint64_t *freqs, size_t n, uint32_t *out_len) {
    htree T = huf_build(freqs, n);
    hcodes C = huf_codes(&T, n);
    memcpy(out_len, C.len, n * sizeof(uint32_t));
    hcodes_free(&C);
    free(T.nodes);
}

/* byte sink that keeps counting past cap */
typedef struct {
    uint8_t *out;
    size_t cap, len;
    uint8_t cur;
    int nbits;
} sink;

static void put_byte(sink *s, uint8_t b) {
    if (s->len < s->cap) s->out[s->len] = b;
    s->len++;
}
static void put_u32(sink *s, uint32_t v) {
    for (int i = 0; i < 4; i++) put_byte(s, (uint8_t)(v >> (8 * i)));
}
static void put_u64(sink *s, uint64_t v) {
    for (int i = 0; i < 8; i++) put_byte(s, (uint8_t)(v >> (8 * i)));
}
/* bit.rs:209-253 IoBitWriter MSB-first, zero padded on flush */
static void put_bit(sink *s, int b) {
    s->cur = (uint8_t)(s->cur | ((b & 1) << (7 - s->nbits)));
    if (++s->nbits == 8) {
        put_byte(s, s->cur);
        s->cur = 0;
        s->nbits = 0;
    }
}
static void pad_and_flush(sink *s) {
    if (s->nbits) {
        put_byte(s, s->cur);
        s->cur = 0;
        s->nbits = 0;
    }
}

/* huf.rs:296-321 pre-order trie */
static void ser_trie(sink *s, const htree *T, int32_t node, const uint8_t *sym_bytes, size_t sym_size) {
    const hnode *nd = &T->nodes[node];
    if (nd->left < 0) {
        put_byte(s, 0);
        for (size_t i = 0; i < sym_size; i++) put_byte(s, sym_bytes[nd->sym * sym_size + i]);
        return;
    }
    put_byte(s, 1);
    ser_trie(s, T, nd->left, sym_bytes, sym_size);
    ser_trie(s, T, nd->right, sym_bytes, sym_size);
}

static void huf_encode_to(sink *s, const uint32_t *stream, size_t n, size_t nsym, const uint8_t *sym_bytes, size_t sym_size) {
    uint64_t *freqs = (uint64_t *)calloc(nsym, sizeof(uint64_t));
    for (size_t i = 0; i < n; i++) freqs[stream[i]]++;
    htree T = huf_build(freqs, nsym);
    hcodes C = huf_codes(&T, nsym);
    ser_trie(s, &T, T.root, sym_bytes, sym_size);
    for (size_t i = 0; i < n; i++) {
        uint32_t sy = stream[i];
        for (uint32_t b = 0; b < C.len[sy]; b++) put_bit(s, C.bits[C.off[sy] + b]);
    }
    pad_and_flush(s);
    hcodes_free(&C);
    free(T.nodes);
    free(freqs);
}

size_t oracle_huf_encode_ids(const uint32_t *stream, size_t n, size_t nsym, const uint8_t *sym_bytes, size_t sym_size,
                             uint8_t *out, size_t cap) {
    sink s = {out, cap, 0, 0, 0};
    huf_encode_to(&s, stream, n, nsym, sym_bytes, sym_size);
    return s.len;
}

size_t oracle_bitpack_codes(const uint32_t *stream, size_t n, const char *const *codes, uint8_t *out, size_t cap) {
    sink s = {out, cap, 0, 0, 0};
    for (size_t i = 0; i < n; i++)
        for (const char *c = codes[stream[i]]; *c; c++) put_bit(&s, *c == '1');
    pad_and_flush(&s);
    return s.len;
}

/* ---- symbol tables ---- */

/* map an RGB image to dense ids in ascending packed-key order; returns nsym; sym_bytes = 11-byte Rgb slices (ser.rs:210-214) */
static size_t rgb_symbols(const uint8_t *rgb, size_t n, uint32_t *ids, uint8_t **sym_bytes) {
    size_t cap = n < (1u << 24) ? n : (1u << 24);
    uint32_t *keys = (uint32_t *)malloc((cap ? cap : 1) * sizeof(uint32_t));
    size_t u = oracle_count_freqs_rgb(rgb, n, keys, NULL);
    uint32_t *lut = (uint32_t *)malloc((1u << 24) * sizeof(uint32_t));
    *sym_bytes = (uint8_t *)malloc((u ? u : 1) * 11);
    for (size_t i = 0; i < u; i++) {
        lut[keys[i]] = (uint32_t)i;
        uint8_t *p = *sym_bytes + 11 * i;
        memset(p, 0, 8);
        p[0] = 3;
        p[8] = (uint8_t)(keys[i] >> 16);
        p[9] = (uint8_t)(keys[i] >> 8);
        p[10] = (uint8_t)keys[i];
    }
    for (size_t i = 0; i < n; i++)
        ids[i] = lut[((uint32_t)rgb[3 * i] << 16) | ((uint32_t)rgb[3 * i + 1] << 8) | rgb[3 * i + 2]];
    free(lut);
    free(keys);
    return u;
}

/* hufc.rs:12-17 */
size_t oracle_encode_hufman(const uint8_t *rgb, uint32_t w, uint32_t h, uint8_t *out, size_t cap) {
    size_t n = (size_t)w * h;
    sink s = {out, cap, 0, 0, 0};
    put_u32(&s, w);
    put_u32(&s, h);
    uint32_t *ids = (uint32_t *)malloc((n ? n : 1) * sizeof(uint32_t));
    uint8_t *sb;
    size_t u = rgb_symbols(rgb, n, ids, &sb);
    if (u) huf_encode_to(&s, ids, n, u, sb, 11);
    free(ids);
    free(sb);
    return s.len;
}

/* ---- decoding ---- */
typedef struct {
    const uint8_t *buf;
    size_t len, pos;
} rdr;

static int rd_byte(rdr *r, uint8_t *b) {
    if (r->pos >= r->len) return 0;
    *b = r->buf[r->pos++];
    return 1;
}
static int rd_u32(rdr *r, uint32_t *v) {
    *v = 0;
    for (int i = 0; i < 4; i++) {
        uint8_t b;
        if (!rd_byte(r, &b)) return 0;
        *v |= (uint32_t)b << (8 * i);
    }
    return 1;
}
static int rd_u64(rdr *r, uint64_t *v) {
    *v = 0;
    for (int i = 0; i < 8; i++) {
        uint8_t b;
        if (!rd_byte(r, &b)) return 0;
        *v |= (uint64_t)b << (8 * i);
    }
    return 1;
}

typedef struct {
    int32_t left, right;
    uint8_t val[11];
} dnode;
typedef struct {
    dnode *nodes;
    size_t nn, cap;
} dtrie;

static int32_t deser_trie(rdr *r, dtrie *T, size_t sym_size) { /* huf.rs:330-350 */
    uint8_t tag;
    if (!rd_byte(r, &tag)) return -1;
    if (T->nn == T->cap) {
        T->cap = T->cap ? T->cap * 2 : 64;
        T->nodes = (dnode *)realloc(T->nodes, T->cap * sizeof(dnode));
    }
    int32_t me = (int32_t)T->nn++;
    if (tag == 0) {
        T->nodes[me].left = T->nodes[me].right = -1;
        for (size_t i = 0; i < sym_size; i++) {
            uint8_t b;
            if (!rd_byte(r, &b)) return -1;
            T->nodes[me].val[i] = b;
        }
        return me;
    }
    if (tag != 1) return -1;
    int32_t l = deser_trie(r, T, sym_size);
    if (l < 0) return -1;
    int32_t rr = deser_trie(r, T, sym_size);
    if (rr < 0) return -1;
    T->nodes[me].left = l;
    T->nodes[me].right = rr;
    return me;
}

/* decode nsyms symbols (huf.rs:187-206 trie walk over MSB-first bits); returns 0 ok */
static int huf_decode_n_partial(rdr *r, size_t sym_size, size_t nsyms, uint8_t *out_vals, size_t *decoded) {
    dtrie T = {NULL, 0, 0};
    int32_t root = deser_trie(r, &T, sym_size);
    if (root < 0) {
        free(T.nodes);
        return 1;
    }
    if (sym_size == 11) /* ser.rs:210-214: Rgb = a serialised slice that must hold exactly 3 elements (vec.try_into().ok()?); the
                           trie is deserialised as a whole (huf.rs:330-350), so ANY leaf with another length fails the decode */
        for (size_t i = 0; i < T.nn; i++)
            if (T.nodes[i].left < 0 && (T.nodes[i].val[0] != 3 || T.nodes[i].val[1] | T.nodes[i].val[2] | T.nodes[i].val[3] |
                                                                      T.nodes[i].val[4] | T.nodes[i].val[5] | T.nodes[i].val[6] | T.nodes[i].val[7])) {
                free(T.nodes);
                return 1;
            }
    size_t bitpos = r->pos * 8, bitend = r->len * 8;
    if (decoded) *decoded = nsyms;
    for (size_t i = 0; i < nsyms; i++) {
        int32_t nd = root;
        while (T.nodes[nd].left >= 0) {
            if (bitpos >= bitend) {
                free(T.nodes);
                if (decoded) { *decoded = i; return 0; } /* the symbol iterator just ends (huf.rs lookup -> None) */
                return 2;
            }
            int bit = (r->buf[bitpos >> 3] >> (7 - (bitpos & 7))) & 1;
            bitpos++;
            nd = bit ? T.nodes[nd].right : T.nodes[nd].left;
        }
        memcpy(out_vals + i * sym_size, T.nodes[nd].val, sym_size);
    }
    free(T.nodes);
    return 0;
}

static int huf_decode_n(rdr *r, size_t sym_size, size_t nsyms, uint8_t *out_vals) {
    return huf_decode_n_partial(r, sym_size, nsyms, out_vals, NULL);
}

int oracle_decode_hufman(const uint8_t *buf, size_t len, uint32_t *w, uint32_t *h, uint8_t *out_rgb, size_t cap_px) {
    rdr r = {buf, len, 0};
    if (!rd_u32(&r, w) || !rd_u32(&r, h)) return 1;
    size_t n = (size_t)*w * *h;
    if (n > cap_px) return 3;
    if (n == 0) return 0;
    uint8_t *vals = (uint8_t *)malloc(n * 11);
    int rc = huf_decode_n(&r, 11, n, vals);
    if (rc == 0)
        for (size_t i = 0; i < n; i++) memcpy(out_rgb + 3 * i, vals + 11 * i + 8, 3);
    free(vals);
    return rc;
}

/* hilbertc.rs:405-415 */
size_t oracle_encode_delta(const uint8_t *rgb, uint32_t w, uint32_t h, uint8_t *out, size_t cap) {
    size_t n = (size_t)w * h;
    sink s = {out, cap, 0, 0, 0};
    put_u32(&s, w);
    put_u32(&s, h);
    if (n == 0) return s.len;
    int16_t *diff = (int16_t *)malloc(n * 3 * sizeof(int16_t));
    oracle_delta(rgb, w, h, diff);
    uint32_t *keys = (uint32_t *)malloc(n * sizeof(uint32_t));
    uint64_t *cnts = (uint64_t *)malloc(n * sizeof(uint64_t));
    size_t u = oracle_hist_delta(diff, n, keys, cnts, n);
    /* ids by binary search in the ascending key table */
    uint32_t *ids = (uint32_t *)malloc(n * sizeof(uint32_t));
    for (size_t i = 0; i < n; i++) {
        uint32_t key = delta_key(diff + 3 * i);
        size_t lo = 0, hi = u;
        while (lo + 1 < hi) {
            size_t mid = (lo + hi) / 2;
            if (keys[mid] <= key) lo = mid;
            else hi = mid;
        }
        ids[i] = (uint32_t)lo;
    }
    uint8_t *sb = (uint8_t *)malloc(u * 6);
    for (size_t i = 0; i < u; i++) { /* ser.rs:188-195 : [i16;3] little endian, no length */
        uint32_t key = keys[i];
        int16_t d[3];
        d[2] = (int16_t)(key % 511) - 255;
        d[1] = (int16_t)((key / 511) % 511) - 255;
        d[0] = (int16_t)(key / (511 * 511)) - 255;
        for (int j = 0; j < 3; j++) {
            sb[6 * i + 2 * j] = (uint8_t)((uint16_t)d[j] & 0xff);
            sb[6 * i + 2 * j + 1] = (uint8_t)((uint16_t)d[j] >> 8);
        }
    }
    huf_encode_to(&s, ids, n, u, sb, 6);
    free(sb);
    free(ids);
    free(keys);
    free(cnts);
    free(diff);
    return s.len;
}

/* hilbertc.rs:417-431 */
int oracle_decode_delta(const uint8_t *buf, size_t len, uint32_t *w, uint32_t *h, uint8_t *out_rgb, size_t cap_px) {
    rdr r = {buf, len, 0};
    if (!rd_u32(&r, w) || !rd_u32(&r, h)) return 1;
    size_t n = (size_t)*w * *h;
    if (n > cap_px) return 3;
    if (n == 0) return 0;
    /* hilbertc.rs:417-431: decode_all(reader).unwrap() -- a bad trie panics (reported as failure); the colour stream is then
       ZIPPED with the curve over a zero-initialised image, so a payload that ends early leaves the remaining pixels ZERO and
       decode still returns Some(img). */
    uint8_t *vals = (uint8_t *)calloc(n, 6);
    size_t got = 0;
    int rc = huf_decode_n_partial(&r, 6, n, vals, &got);
    if (rc == 0) {
        int16_t *diff = (int16_t *)malloc(n * 3 * sizeof(int16_t));
        for (size_t i = 0; i < 3 * n; i++) diff[i] = (int16_t)((uint16_t)vals[2 * i] | ((uint16_t)vals[2 * i + 1] << 8));
        if (oracle_undelta(diff, *w, *h, out_rgb)) rc = 2; /* FromDiff unwrap panics (symbols behind `got` are zero: no effect) */
        free(diff);
        if (rc == 0 && got < n) {
            uint32_t *xy = (uint32_t *)malloc(n * 2 * sizeof(uint32_t));
            oracle_hilbert_xy(*w, *h, xy);
            for (size_t i = got; i < n; i++) memset(out_rgb + 3 * ((size_t)xy[2 * i + 1] * *w + xy[2 * i]), 0, 3);
            free(xy);
        }
    }
    free(vals);
    return rc;
}

/* clusterc.rs:148-166 ; wire: w u32, h u32, k as u64, k x (x u32, y u32, Rgb slice = u64 3 + 3 bytes) */
size_t oracle_encode_voronoi(const uint8_t *rgb, uint32_t w, uint32_t h, size_t k, int mode, int tie_rule, uint32_t max_iters,
                             uint8_t *out, size_t cap) {
    uint32_t *cxy = (uint32_t *)malloc((k ? k : 1) * 2 * sizeof(uint32_t));
    uint8_t *crgb = (uint8_t *)malloc((k ? k : 1) * 3);
    int rc = oracle_kmeans_xyrgb(rgb, w, h, k, mode, tie_rule, max_iters, cxy, crgb, NULL, NULL, NULL);
    sink s = {out, cap, 0, 0, 0};
    if (rc == ORACLE_OK) {
        put_u32(&s, w);
        put_u32(&s, h);
        put_u64(&s, (uint64_t)k);
        for (size_t c = 0; c < k; c++) {
            put_u32(&s, cxy[2 * c]);
            put_u32(&s, cxy[2 * c + 1]);
            put_u64(&s, 3);
            for (int j = 0; j < 3; j++) put_byte(&s, crgb[3 * c + j]);
        }
    }
    free(cxy);
    free(crgb);
    return s.len;
}

/* clusterc.rs:168-189 */
int oracle_decode_voronoi(const uint8_t *buf, size_t len, uint32_t *w, uint32_t *h, uint8_t *out_rgb, size_t cap_px) {
    rdr r = {buf, len, 0};
    uint64_t k;
    if (!rd_u32(&r, w) || !rd_u32(&r, h) || !rd_u64(&r, &k)) return 1;
    if ((size_t)*w * *h > cap_px) return 3;
    if (k > (len - r.pos) / 19) return 1;
    uint32_t *cxy = (uint32_t *)malloc((k ? k : 1) * 2 * sizeof(uint32_t));
    uint8_t *crgb = (uint8_t *)malloc((k ? k : 1) * 3);
    int rc = 0;
    for (uint64_t c = 0; c < k && !rc; c++) {
        uint64_t l;
        if (!rd_u32(&r, &cxy[2 * c]) || !rd_u32(&r, &cxy[2 * c + 1]) || !rd_u64(&r, &l) || l != 3) rc = 1;
        for (int j = 0; j < 3 && !rc; j++)
            if (!rd_byte(&r, &crgb[3 * c + j])) rc = 1;
    }
    if (!rc && k == 0 && (size_t)*w * *h > 0) rc = 4; /* min_by_key(...).unwrap() panics on k = 0 */
    if (!rc) oracle_voronoi_fill(cxy, crgb, (size_t)k, *w, *h, out_rgb);
    free(cxy);
    free(crgb);
    return rc;
}

/* clusterc.rs:18-53 */
size_t oracle_encode_cluster_colors(const uint8_t *rgb, uint32_t w, uint32_t h, size_t k, int mode, int tie_rule,
                                    uint32_t max_iters, uint8_t *out, size_t cap) {
    size_t n = (size_t)w * h;
    uint8_t *red = (uint8_t *)malloc(3 * (n ? n : 1));
    int rc = oracle_cluster_colors(rgb, w, h, k, mode, tie_rule, max_iters, red, NULL, NULL);
    size_t len = 0;
    if (rc == ORACLE_OK) len = oracle_encode_hufman(red, w, h, out, cap);
    free(red);
    return len;
}

/* hilbertc.rs:26-38 exact RLE codec: dims, then (u8 count, Rgb slice) records */
size_t oracle_encode_hilbert_rle(const uint8_t *rgb, uint32_t w, uint32_t h, uint8_t *out, size_t cap) {
    size_t n = (size_t)w * h;
    sink s = {out, cap, 0, 0, 0};
    put_u32(&s, w);
    put_u32(&s, h);
    if (n == 0) return s.len;
    uint8_t *lin = (uint8_t *)malloc(3 * n);
    oracle_hilbert_gather(rgb, w, h, lin);
    uint8_t *cnt = (uint8_t *)malloc(n);
    uint8_t *col = (uint8_t *)malloc(3 * n);
    size_t r = oracle_rle_exact(lin, n, cnt, col);
    for (size_t i = 0; i < r; i++) {
        put_byte(&s, cnt[i]);
        put_u64(&s, 3);
        for (int j = 0; j < 3; j++) put_byte(&s, col[3 * i + j]);
    }
    free(lin);
    free(cnt);
    free(col);
    return s.len;
}

/* hilbertc.rs:55-79, 304-333 */
int oracle_decode_hilbert_rle(const uint8_t *buf, size_t len, uint32_t *w, uint32_t *h, uint8_t *out_rgb, size_t cap_px) {
    rdr r = {buf, len, 0};
    if (!rd_u32(&r, w) || !rd_u32(&r, h)) return 1;
    size_t n = (size_t)*w * *h;
    if (n > cap_px) return 3;
    if (n == 0) return 0;
    uint32_t *xy = (uint32_t *)malloc(n * 2 * sizeof(uint32_t));
    oracle_hilbert_xy(*w, *h, xy);
    size_t i = 0;
    int rc = 0;
    /* hilbertc.rs:55-79 + 322-333: the image starts all zero (ImageBuffer::new); the RleDecoder iterator is zipped with the curve.
       - the byte stream ends exactly where a record would start: the iterator ends, zip stops, the rest of the image STAYS ZERO
         and decode still returns Some(img)  (RepCount::deserialize(..)? -> None ends the iterator, not the decode);
       - a record with count 0: assert!(self.count > 0) panics;  a truncated / malformed colour: .unwrap() panics.
       Panics are reported as failures here (rc 2), like the C ABI does (no unwinding across FFI). */
    memset(out_rgb, 0, n * 3);
    while (i < n) {
        uint8_t cnt, c[3];
        uint64_t l;
        if (!rd_byte(&r, &cnt)) break; /* clean end of the stream: partial image */
        if (cnt == 0 || !rd_u64(&r, &l) || l != 3 || !rd_byte(&r, &c[0]) || !rd_byte(&r, &c[1]) || !rd_byte(&r, &c[2])) {
            rc = 2;
            break;
        }
        for (uint8_t j = 0; j < cnt && i < n; j++, i++) memcpy(out_rgb + 3 * ((size_t)xy[2 * i + 1] * *w + xy[2 * i]), c, 3);
    }
    free(xy);
    return rc;
}
