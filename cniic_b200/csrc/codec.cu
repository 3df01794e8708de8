// codec.cu -- host side of the reference's Codec trait (src/codec.rs:14-19) for the codecs on the hot path:
//   ClusterColors (src/codec/clusterc.rs:14-62), VoronoiCluster (clusterc.rs:143-194), Delta (src/codec/hilbertc.rs:402-439),
//   Hufman (src/codec/hufc.rs), Hilbert RLE exact (hilbertc.rs:12-96).
// Wire formats follow src/ser.rs, src/huf.rs and src/bit.rs byte for byte (DESIGN.md "Wire formats").
// The per-pixel work (K-means, histograms, recolour, Hilbert gather, delta, fill, scatter, Huffman bit packing) runs on the
// GPU; the Huffman tree (<= #symbols nodes), the decoders' bit-serial trie walk and the run-length pass run on the host.
#include <cctype>
#include <cstring>
#include <algorithm>
#include <string>
#include <unordered_map>
#include <vector>

#include "common.cuh"
#include "stages.cuh"

namespace {


// ---- codec expression parsing (clusterc.rs:116-141, 274-297; hilbertc.rs:337-395, 574-582; hufc.rs:51-63) ----
enum CodecKind { CK_NONE, CK_CLUSTER_COLORS, CK_VORONOI, CK_DELTA, CK_HUFMAN, CK_HILBERT_RLE };
struct CodecSpec { CodecKind kind = CK_NONE; uint32_t arg = 0; };

// finds name '(' digits ')' anywhere in s (the reference's regexes are unanchored)
bool find_call(const std::string &s, const std::vector<std::string> &names, uint32_t *arg) {
    for (const std::string &nm : names) {
        size_t pos = 0;
        while ((pos = s.find(nm + "(", pos)) != std::string::npos) {
            size_t i = pos + nm.size() + 1, j = i;
            unsigned long long v = 0;
            while (j < s.size() && isdigit((unsigned char)s[j]) && j - i < 10) v = v * 10 + (s[j++] - '0');
            if (j > i && j < s.size() && s[j] == ')' && v <= 0xffffffffull) { *arg = (uint32_t)v; return true; }
            pos++;
        }
    }
    return false;
}

CodecSpec parse_codec(const char *expr) {
    CodecSpec sp;
    if (!expr) return sp;
    const std::string s(expr);
    std::string lower;
    for (char c : s) lower.push_back((char)tolower((unsigned char)c));
    if (lower == "hufman") { sp.kind = CK_HUFMAN; return sp; }
    // c(?:luster)?-?col(?:ors)?\((\d+)\)
    std::vector<std::string> cc;
    for (const char *a : {"cluster", "c"})
        for (const char *b : {"-", ""})
            for (const char *c : {"colors", "col"}) cc.push_back(std::string(a) + b + c);
    uint32_t arg = 0;
    if (find_call(s, cc, &arg)) { sp.kind = CK_CLUSTER_COLORS; sp.arg = arg; return sp; }
    if (find_call(s, {"voronoi"}, &arg)) { sp.kind = CK_VORONOI; sp.arg = arg; return sp; }
    if (s == "delta") { sp.kind = CK_DELTA; return sp; }
    if (s == "hilbert(rle)" || s == "hilbert(rle(0))") { sp.kind = CK_HILBERT_RLE; return sp; }
    return sp;
}

// ---- byte sink / source (ser.rs: little endian, usize as u64, Rgb as a slice = u64 length + 3 bytes) ----
struct Sink {
    std::vector<uint8_t> v;
    void u8(uint8_t b) { v.push_back(b); }
    void u32(uint32_t x) { for (int i = 0; i < 4; i++) v.push_back((uint8_t)(x >> (8 * i))); }
    void u64(uint64_t x) { for (int i = 0; i < 8; i++) v.push_back((uint8_t)(x >> (8 * i))); }
    void rgb(uint32_t key) { u64(3); u8(key >> 16); u8(key >> 8); u8(key); }
};

struct Source {
    const uint8_t *p;
    size_t len, pos = 0;
    bool u8(uint8_t *b) { if (pos >= len) return false; *b = p[pos++]; return true; }
    bool u32(uint32_t *x) { if (len - pos < 4) return false; *x = 0; for (int i = 0; i < 4; i++) *x |= (uint32_t)p[pos++] << (8 * i); return true; }
    bool u64(uint64_t *x) { if (len - pos < 8) return false; *x = 0; for (int i = 0; i < 8; i++) *x |= (uint64_t)p[pos++] << (8 * i); return true; }
};

// ---- Huffman (huf.rs:58-117): deterministic heap order (freq, creation sequence); first pop = left = bit 0 ----
struct HufTree {
    struct Node { int left, right; uint32_t sym; };
    std::vector<Node> nodes;
    int root = -1;
    std::vector<uint64_t> code;   // per symbol, MSB-first in the low `len` bits
    std::vector<uint8_t> len;
};

bool huf_build(const std::vector<uint64_t> &freq, HufTree *T) {
    const size_t n = freq.size();
    if (n == 0) return false;
    // Two-queue construction.  It reproduces the min-heap ordered by (freq, creation sequence) exactly: leaves are
    // taken in (freq, symbol id) order, internal nodes are created with non-decreasing freq and increasing sequence
    // numbers, and on equal freq a leaf (smaller sequence number) precedes any internal node.
    std::vector<uint32_t> order(n);
    for (size_t i = 0; i < n; i++) order[i] = (uint32_t)i;
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return freq[a] < freq[b]; });
    T->nodes.reserve(2 * n);
    std::vector<uint64_t> nf;
    nf.reserve(2 * n);
    for (size_t i = 0; i < n; i++) {
        T->nodes.push_back({-1, -1, (uint32_t)i});
        nf.push_back(freq[i]);
    }
    size_t li = 0, ii = n;  // next leaf (index into order), next internal node (node id)
    auto pop_min = [&]() -> int {
        const bool has_leaf = li < n, has_int = ii < T->nodes.size();
        if (has_leaf && (!has_int || nf[order[li]] <= nf[ii])) return (int)order[li++];
        return (int)ii++;
    };
    size_t remaining = n;
    while (remaining > 1) {
        const int l = pop_min();
        const int r = pop_min();
        T->nodes.push_back({l, r, 0});
        nf.push_back(nf[l] + nf[r]);
        remaining--;
    }
    T->root = (n == 1) ? 0 : (int)T->nodes.size() - 1;
    T->code.assign(n, 0);
    T->len.assign(n, 0);
    // iterative DFS assigning codes
    struct Fr { int node; uint64_t code; uint32_t len; };
    std::vector<Fr> st{{T->root, 0, 0}};
    while (!st.empty()) {
        const Fr f = st.back();
        st.pop_back();
        const HufTree::Node &nd = T->nodes[f.node];
        if (nd.left < 0) {
            if (f.len > 64) return false;
            T->code[nd.sym] = f.code;
            T->len[nd.sym] = (uint8_t)f.len;
        } else {
            st.push_back({nd.right, (f.code << 1) | 1, f.len + 1});
            st.push_back({nd.left, f.code << 1, f.len + 1});
        }
    }
    return true;
}

// huf.rs:296-321 pre-order: 0x00 + leaf payload | 0x01 + left + right
template <class LeafFn>
void huf_serialize(const HufTree &T, Sink &s, LeafFn leaf) {
    std::vector<int> st{T.root};
    while (!st.empty()) {
        const int id = st.back();
        st.pop_back();
        const HufTree::Node &nd = T.nodes[id];
        if (nd.left < 0) { s.u8(0); leaf(nd.sym); }
        else { s.u8(1); st.push_back(nd.right); st.push_back(nd.left); }
    }
}

struct DecTrie {
    struct Node { int left, right; uint8_t val[11]; };
    std::vector<Node> nodes;
};

// huf.rs:330-350 ; iterative to be safe against deep (malformed) tries
bool huf_deserialize(Source &src, size_t sym_size, DecTrie *T) {
    // stack of nodes waiting for children: (node id, number of children attached)
    std::vector<std::pair<int, int>> st;
    int root = -1;
    for (;;) {
        uint8_t tag;
        if (!src.u8(&tag)) return false;
        if (tag > 1) return false;
        if (T->nodes.size() > (size_t(1) << 26)) return false;
        const int id = (int)T->nodes.size();
        T->nodes.push_back({-1, -1, {0}});
        if (root < 0) root = id;
        if (!st.empty()) {
            if (st.back().second == 0) T->nodes[st.back().first].left = id;
            else T->nodes[st.back().first].right = id;
            st.back().second++;
        }
        if (tag == 0) {
            for (size_t i = 0; i < sym_size; i++)
                if (!src.u8(&T->nodes[id].val[i])) return false;
            while (!st.empty() && st.back().second == 2) st.pop_back();
            if (st.empty()) return true;
        } else {
            st.push_back({id, 0});
        }
    }
}

// huf.rs:187-206: the payload after the trie is decoded on the GPU (huffdec.cu: self-synchronising parallel decoder with the
// sequential decoder's results and error behaviour).  val_off = offset of the symbol bytes inside a leaf's serialised value.
int huf_decode_device(cniic_ctx *ctx, const Source &src, const DecTrie &T, size_t val_off, int sym_bytes, size_t n, uint8_t *d_out,
                      size_t *decoded = nullptr) {
    const size_t nn = T.nodes.size();
    std::vector<int32_t> child(2 * nn);
    std::vector<uint8_t> leaf(8 * nn, 0);
    for (size_t i = 0; i < nn; i++) {
        child[2 * i] = T.nodes[i].left;
        child[2 * i + 1] = T.nodes[i].right;
        if (T.nodes[i].left < 0) memcpy(&leaf[8 * i], T.nodes[i].val + val_off, (size_t)sym_bytes);
    }
    return cniic_dev_huffman_decode(ctx, src.p + src.pos, src.len - src.pos, child.data(), leaf.data(), nn, sym_bytes, n, d_out, decoded);
}

int finish(cniic_ctx *ctx, Sink &s, uint8_t *out, size_t cap, size_t *out_len) {
    *out_len = s.v.size();
    if (s.v.size() > cap || (!out && !s.v.empty())) {
        // keep the finished stream: the caller fetches it with cniic_codec_encode_fetch instead of encoding a second time
        ctx->pending_stream.swap(s.v);
        ctx->has_pending_stream = true;
        return cniic_set_error(ctx, CNIIC_ERR_BUFFER_TOO_SMALL, "need %zu bytes", *out_len);
    }
    if (!s.v.empty()) memcpy(out, s.v.data(), s.v.size());
    return CNIIC_OK;
}

// Hufman codec body over a DEVICE-resident image (hufc.rs:12-17, huf.rs:22-43)
int encode_hufman_body(cniic_ctx *ctx, const uint8_t *d_rgb, size_t n, Sink &s) {
    if (n == 0) return cniic_set_error(ctx, CNIIC_ERR_BAD_ARG, "empty image (huf.rs:95 asserts a non-empty alphabet)");
    uint32_t *d_bins = nullptr, *d_keys = nullptr;
    unsigned long long *d_counts = nullptr;
    size_t u = 0;
    int rc = cniic_dev_hist_rgb_bins(ctx, d_rgb, n, &d_bins);  // pass 1: count_freqs on the GPU
    if (rc == CNIIC_OK) rc = cniic_dev_dense_compact(ctx, d_bins, size_t(1) << 24, &d_keys, &d_counts, &u);
    std::vector<uint32_t> keys(u);
    std::vector<uint64_t> counts(u);
    if (rc == CNIIC_OK && u) {
        cudaMemcpyAsync(keys.data(), d_keys, u * 4, cudaMemcpyDeviceToHost, ctx->stream);
        cudaMemcpyAsync(counts.data(), d_counts, u * 8, cudaMemcpyDeviceToHost, ctx->stream);
        if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = cniic_set_error(ctx, CNIIC_ERR_CUDA, "histogram copy failed");
    }
    cniic_cache_free(ctx, d_bins);
    cniic_cache_free(ctx, d_counts);
    HufTree T;
    if (rc == CNIIC_OK && !huf_build(counts, &T)) rc = cniic_set_error(ctx, CNIIC_ERR_UNSUPPORTED, "Huffman code longer than 64 bits");
    if (rc == CNIIC_OK) {
        huf_serialize(T, s, [&](uint32_t sym) { s.rgb(keys[sym]); });
        // pass 2: bit packing on the GPU (code lookup, scan of the lengths, MSB-first word assembly)
        rc = cniic_dev_huffman_pack(ctx, 0, d_rgb, n, d_keys, u, T.code, T.len, &s.v);
    }
    cniic_cache_free(ctx, d_keys);
    return rc;
}

int decode_hufman_body(cniic_ctx *ctx, Source &src, size_t n, uint8_t *out_rgb) {
    if (n == 0) return CNIIC_OK;
    DecTrie T;
    if (!huf_deserialize(src, 11, &T)) return cniic_set_error(ctx, CNIIC_ERR_DECODE, "bad Huffman trie");
    for (const DecTrie::Node &nd : T.nodes)
        if (nd.left < 0 && (nd.val[0] != 3 || memcmp(nd.val + 1, "\0\0\0\0\0\0\0", 7) != 0)) return cniic_set_error(ctx, CNIIC_ERR_DECODE, "bad Rgb leaf");
    DevBuf d_out(ctx);
    CU_TRY(ctx, d_out.alloc(n * 3));
    ST_TRY(huf_decode_device(ctx, src, T, 8, 3, n, d_out.as<uint8_t>()));  // leaf = u64 length (= 3) + the colour bytes
    CU_TRY(ctx, cudaMemcpyAsync(out_rgb, d_out.p, n * 3, cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return CNIIC_OK;
}

}  // namespace

extern "C" int cniic_codec_name(const char *codec, char *out, size_t cap) {
    const CodecSpec sp = parse_codec(codec);
    std::string nm;
    switch (sp.kind) {
    case CK_CLUSTER_COLORS: nm = "cluster-colors_" + std::to_string(sp.arg); break;  // clusterc.rs:59-61
    case CK_VORONOI: nm = "voronoi_" + std::to_string(sp.arg); break;                 // clusterc.rs:191-193
    case CK_DELTA: nm = "delta"; break;                                               // hilbertc.rs:433-435
    case CK_HUFMAN: nm = "Hufman"; break;                                             // hufc.rs:42-44
    case CK_HILBERT_RLE: nm = "hilbert-rle"; break;                                   // hilbertc.rs:81-87
    default: return CNIIC_ERR_BAD_ARG;
    }
    if (!out || cap < nm.size() + 1) return CNIIC_ERR_BUFFER_TOO_SMALL;
    memcpy(out, nm.c_str(), nm.size() + 1);
    return CNIIC_OK;
}

extern "C" int cniic_codec_encode_fetch(cniic_ctx *ctx, uint8_t *out, size_t cap, size_t *out_len) {
    if (!ctx || !out_len) return CNIIC_ERR_BAD_ARG;
    if (!ctx->has_pending_stream) return cniic_set_error(ctx, CNIIC_ERR_BAD_ARG, "no encoded stream is pending on this context");
    *out_len = ctx->pending_stream.size();
    if (*out_len > cap || (!out && *out_len)) return cniic_set_error(ctx, CNIIC_ERR_BUFFER_TOO_SMALL, "need %zu bytes", *out_len);
    if (*out_len) memcpy(out, ctx->pending_stream.data(), *out_len);
    std::vector<uint8_t>().swap(ctx->pending_stream);
    ctx->has_pending_stream = false;
    return CNIIC_OK;
}

extern "C" int cniic_codec_encode(cniic_ctx *ctx, const char *codec, const uint8_t *rgb, uint32_t w, uint32_t h, uint8_t *out,
                                  size_t cap, size_t *out_len) {
    if (!ctx || !out_len) return CNIIC_ERR_BAD_ARG;
    ctx->has_pending_stream = false;
    ctx->pending_stream.clear();
    const CodecSpec sp = parse_codec(codec);
    if (sp.kind == CK_NONE) return cniic_set_error(ctx, CNIIC_ERR_BAD_ARG, "unknown codec expression '%s'", codec ? codec : "(null)");
    const size_t n = (size_t)w * h;
    if (!rgb && n) return cniic_set_error(ctx, CNIIC_ERR_BAD_ARG, "null image");
    if (n >= (size_t(1) << 31)) return cniic_set_error(ctx, CNIIC_ERR_BAD_ARG, "image too large");
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    Sink s;
    DevBuf din(ctx);
    CU_TRY(ctx, din.alloc(n * 3));
    if (n) CU_TRY(ctx, cudaMemcpyAsync(din.p, rgb, n * 3, cudaMemcpyHostToDevice, ctx->stream));
    switch (sp.kind) {
    case CK_HUFMAN: {  // hufc.rs:12-17
        s.u32(w); s.u32(h);
        ST_TRY(encode_hufman_body(ctx, din.as<uint8_t>(), n, s));
        break;
    }
    case CK_CLUSTER_COLORS: {  // clusterc.rs:18-53
        if (sp.arg == 0 || sp.arg > CNIIC_MAX_K) return cniic_set_error(ctx, CNIIC_ERR_BAD_ARG, "k must be in 1..%d", CNIIC_MAX_K);
        DevBuf dred(ctx);
        CU_TRY(ctx, dred.alloc(n * 3));
        ST_TRY(cniic_dev_cluster_colors(ctx, din.as<uint8_t>(), n, sp.arg, ctx->codec_max_iters, CNIIC_TIE_KEEP_CURRENT, dred.as<uint8_t>(), nullptr, nullptr, nullptr));
        s.u32(w); s.u32(h);
        ST_TRY(encode_hufman_body(ctx, dred.as<uint8_t>(), n, s));  // the recoloured image never leaves HBM
        break;
    }
    case CK_VORONOI: {  // clusterc.rs:148-166
        const uint32_t k = sp.arg;
        if (k == 0 || k > CNIIC_MAX_K) return cniic_set_error(ctx, CNIIC_ERR_BAD_ARG, "k must be in 1..%d", CNIIC_MAX_K);
        std::vector<uint32_t> cxy(2 * (size_t)k);
        std::vector<uint8_t> crgb(3 * (size_t)k);
        ST_TRY(cniic_kmeans_xyrgb(ctx, rgb, w, h, k, ctx->codec_max_iters, CNIIC_TIE_KEEP_CURRENT, cxy.data(), crgb.data(), nullptr, nullptr, nullptr));
        s.u32(w); s.u32(h);
        s.u64(k);
        for (uint32_t c = 0; c < k; c++) {  // clusterc.rs:250-257
            s.u32(cxy[2 * c]); s.u32(cxy[2 * c + 1]);
            s.u64(3); s.u8(crgb[3 * c]); s.u8(crgb[3 * c + 1]); s.u8(crgb[3 * c + 2]);
        }
        break;
    }
    case CK_DELTA: {  // hilbertc.rs:405-415
        s.u32(w); s.u32(h);
        if (n == 0) return cniic_set_error(ctx, CNIIC_ERR_BAD_ARG, "empty image (huf.rs:95 asserts a non-empty alphabet)");
        // pass 1 on the GPU: fused Hilbert gather + diff + histogram (no stream materialised)
        uint32_t *d_bins = nullptr, *d_keys = nullptr;
        unsigned long long *d_counts = nullptr;
        size_t nbins = 0, u = 0;
        int rc = cniic_dev_hist_delta_bins(ctx, din.as<uint8_t>(), w, h, &d_bins, &nbins);
        if (rc == CNIIC_OK) rc = cniic_dev_dense_compact(ctx, d_bins, nbins, &d_keys, &d_counts, &u);
        std::vector<uint32_t> keys(u);
        std::vector<uint64_t> counts(u);
        if (rc == CNIIC_OK) {
            cudaMemcpyAsync(keys.data(), d_keys, u * 4, cudaMemcpyDeviceToHost, ctx->stream);
            cudaMemcpyAsync(counts.data(), d_counts, u * 8, cudaMemcpyDeviceToHost, ctx->stream);
            if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = cniic_set_error(ctx, CNIIC_ERR_CUDA, "histogram copy failed");
        }
        cniic_cache_free(ctx, d_bins);
        cniic_cache_free(ctx, d_counts);
        HufTree T;
        if (rc == CNIIC_OK && !huf_build(counts, &T)) rc = cniic_set_error(ctx, CNIIC_ERR_UNSUPPORTED, "Huffman code longer than 64 bits");
        DevBuf dd(ctx);
        if (rc == CNIIC_OK && dd.alloc(n * 6) != cudaSuccess) rc = CNIIC_ERR_CUDA;
        if (rc == CNIIC_OK) {
            huf_serialize(T, s, [&](uint32_t sym) {  // ser.rs:188-195 : [i16;3] LE
                const uint32_t key = keys[sym];
                const int16_t d[3] = {(int16_t)(int(key / (511 * 511)) - 255), (int16_t)(int((key / 511) % 511) - 255), (int16_t)(int(key % 511) - 255)};
                for (int j = 0; j < 3; j++) { s.u8((uint8_t)((uint16_t)d[j] & 0xff)); s.u8((uint8_t)((uint16_t)d[j] >> 8)); }
            });
            // pass 2: the delta stream itself, then bit packing, both on the GPU
            rc = cniic_delta_i16_device(ctx, din.as<uint8_t>(), w, h, dd.as<int16_t>());
        }
        if (rc == CNIIC_OK) rc = cniic_dev_huffman_pack(ctx, 1, dd.p, n, d_keys, u, T.code, T.len, &s.v);
        cniic_cache_free(ctx, d_keys);
        ST_TRY(rc);
        break;
    }
    case CK_HILBERT_RLE: {  // hilbertc.rs:26-38, 99-196 ; records = u8 count (1..=255) + Rgb slice
        s.u32(w); s.u32(h);
        if (n) {
            DevBuf dl(ctx);
            CU_TRY(ctx, dl.alloc(n * 3));
            ST_TRY(cniic_dev_hilbert_gather(ctx, din.as<uint8_t>(), w, h, dl.as<uint8_t>()));
            ST_TRY(cniic_dev_rle_encode(ctx, dl.as<uint8_t>(), n, &s.v));  // segmented scans + record scatter on the GPU
        }
        break;
    }
    default: return CNIIC_ERR_BAD_ARG;
    }
    return finish(ctx, s, out, cap, out_len);
}

extern "C" int cniic_codec_decode(cniic_ctx *ctx, const char *codec, const uint8_t *data, size_t len, uint32_t *w, uint32_t *h,
                                  uint8_t *out_rgb, size_t cap_pixels) {
    if (!ctx || !data || !w || !h) return CNIIC_ERR_BAD_ARG;
    const CodecSpec sp = parse_codec(codec);
    if (sp.kind == CK_NONE) return cniic_set_error(ctx, CNIIC_ERR_BAD_ARG, "unknown codec expression '%s'", codec ? codec : "(null)");
    Source src{data, len};
    if (!src.u32(w) || !src.u32(h)) return cniic_set_error(ctx, CNIIC_ERR_DECODE, "truncated header");
    const size_t n = (size_t)*w * *h;
    if (!out_rgb) return CNIIC_OK;  // dimension query
    if (n > cap_pixels) return cniic_set_error(ctx, CNIIC_ERR_BUFFER_TOO_SMALL, "need room for %zu pixels", n);
    if (n >= (size_t(1) << 31)) return cniic_set_error(ctx, CNIIC_ERR_DECODE, "image too large");
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    switch (sp.kind) {
    case CK_HUFMAN:
    case CK_CLUSTER_COLORS:  // clusterc.rs:55-57 -> hufc.rs:19-40
        return decode_hufman_body(ctx, src, n, out_rgb);
    case CK_VORONOI: {  // clusterc.rs:168-189
        uint64_t k;
        if (!src.u64(&k)) return cniic_set_error(ctx, CNIIC_ERR_DECODE, "truncated header");
        if (k > (len - src.pos) / 19) return cniic_set_error(ctx, CNIIC_ERR_DECODE, "truncated centroid list");
        if (n == 0) return CNIIC_OK;
        if (k == 0) return cniic_set_error(ctx, CNIIC_ERR_DECODE, "no centroids (clusterc.rs:182-184 unwraps None)");
        if (k > CNIIC_MAX_K) return cniic_set_error(ctx, CNIIC_ERR_UNSUPPORTED, "more than %d centroids", CNIIC_MAX_K);
        std::vector<uint32_t> cxy(2 * k);
        std::vector<uint8_t> crgb(3 * k);
        for (uint64_t c = 0; c < k; c++) {
            uint64_t l;
            if (!src.u32(&cxy[2 * c]) || !src.u32(&cxy[2 * c + 1]) || !src.u64(&l) || l != 3 || !src.u8(&crgb[3 * c]) ||
                !src.u8(&crgb[3 * c + 1]) || !src.u8(&crgb[3 * c + 2]))
                return cniic_set_error(ctx, CNIIC_ERR_DECODE, "bad centroid record");
        }
        return cniic_voronoi_fill(ctx, cxy.data(), crgb.data(), (uint32_t)k, *w, *h, out_rgb);
    }
    case CK_DELTA: {  // hilbertc.rs:417-431
        if (n == 0) return CNIIC_OK;
        DecTrie T;
        if (!huf_deserialize(src, 6, &T)) return cniic_set_error(ctx, CNIIC_ERR_DECODE, "bad Huffman trie");
        // symbols = little-endian i16 triples (ser.rs:188-195) == the layout the device un-delta pass reads: the difference
        // stream never visits the host (decode -> prefix sums along the curve -> scatter, all in HBM)
        DevBuf d_diff(ctx), d_img(ctx);
        CU_TRY(ctx, d_diff.alloc(n * 6));
        CU_TRY(ctx, d_img.alloc(n * 3));
        // hilbertc.rs:425-428: the colour stream is zipped with the curve over a zero image -- a payload that ends early paints the
        // pixels it reaches and leaves the rest zero (Hufman::decode, in contrast, returns None: hufc.rs:24-36)
        size_t got = 0;
        ST_TRY(huf_decode_device(ctx, src, T, 0, 6, n, d_diff.as<uint8_t>(), &got));
        if (got < n) CU_TRY(ctx, cudaMemsetAsync(d_diff.as<uint8_t>() + got * 6, 0, (n - got) * 6, ctx->stream));
        ST_TRY(cniic_dev_undelta(ctx, d_diff.as<int16_t>(), *w, *h, d_img.as<uint8_t>()));
        if (got < n) ST_TRY(cniic_dev_zero_curve_tail(ctx, *w, *h, got, d_img.as<uint8_t>()));
        CU_TRY(ctx, cudaMemcpyAsync(out_rgb, d_img.p, n * 3, cudaMemcpyDeviceToHost, ctx->stream));
        CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        return CNIIC_OK;
    }
    case CK_HILBERT_RLE: {  // hilbertc.rs:55-79, 304-333
        if (n == 0) return CNIIC_OK;
        DevBuf d_img(ctx);
        CU_TRY(ctx, d_img.alloc(n * 3));
        ST_TRY(cniic_dev_rle_decode(ctx, src.p + src.pos, src.len - src.pos, *w, *h, d_img.as<uint8_t>()));
        CU_TRY(ctx, cudaMemcpyAsync(out_rgb, d_img.p, n * 3, cudaMemcpyDeviceToHost, ctx->stream));
        CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        return CNIIC_OK;
    }
    default: return CNIIC_ERR_BAD_ARG;
    }
}
