// huffdec.cu -- parallel Huffman decoding on the GPU (replaces the bit-serial trie walk of huf.rs:187-206 that the
// Hufman / cluster-colors / delta decoders run, hufc.rs:19-40 and hilbertc.rs:417-431).
//
// The reference's stream has no synchronisation points (one MSB-first bit string, huf.rs:33-41), so the decoder relies on the
// self-synchronisation of Huffman codes (Weissenberger & Schmidt, "Massively Parallel Huffman Decoding on GPUs", ICPP 2018):
//   1. the bit string is cut into subsequences of SUB bits, one per thread.  Every thread decodes from the start of its
//      subsequence -- a guess, the true first code word starts a few bits later -- and records where it left the subsequence;
//   2. a thread whose predecessor left at a position other than the one it started from decodes again from there.  This is
//      repeated (inside the CTA through shared memory, across CTAs by re-launching) until nothing changes.  Thread 0 starts at
//      the true position, so after j steps threads 0..j are right: the fixpoint IS the sequential decoding, whatever the codes;
//      in practice wrong starts re-synchronise within a subsequence or two and two or three sweeps suffice;
//   3. an exclusive scan of the per-subsequence symbol counts gives every thread its output index;
//   4. every thread decodes its subsequence once more and writes the symbols.
// Work is 2-3 decoding passes over the payload; the output is identical to the sequential decoder, including its error
// behaviour: the stream is rejected when it holds fewer than n complete code words (huf.rs:190-204 returns None).
#include <algorithm>
#include <cstdlib>
#include <vector>

#include "common.cuh"
#include "stages.cuh"

namespace {

constexpr int HD_SUB = 256;      // bits per subsequence (>= 64 = longest code, so a code word spans at most two subsequences)
constexpr int HD_THREADS = 256;  // subsequences per CTA
constexpr int HD_LUT_BITS = 12;

struct HdArgs {
    const uint32_t *words;       // payload, big-endian bit order inside bytes, padded with >= 8 zero bytes
    unsigned long long end_bit;  // payload bits
    unsigned long long nsub;     // subsequences
    const int2 *child;           // per trie node: (left, right); left < 0 = leaf
    const uint2 *leaf;           // per trie node: symbol bytes (first sym_bytes meaningful)
    const uint32_t *lut;         // 2^HD_LUT_BITS entries: node reached | bits consumed << 28
    unsigned long long *sub_start, *sub_exit;  // per subsequence: start position used, position of the first code word at/after its end
    uint32_t *sub_cnt;                         // per subsequence: code words starting in [start, end of subsequence)
    unsigned long long *chunk_off;             // per CTA chunk: exclusive prefix of the counts (nchunks + 1 entries)
    uint32_t *changed;
    uint8_t *out;
    unsigned long long n;  // symbols wanted
    int sym_bytes;
};

// 32 payload bits starting at bit `at` (MSB first)
__device__ __forceinline__ uint32_t hd_window(const uint32_t *words, unsigned long long at) {
    const unsigned long long w = at >> 5;
    const uint32_t hi = __byte_perm(__ldg(words + w), 0, 0x0123), lo = __byte_perm(__ldg(words + w + 1), 0, 0x0123);
    const uint32_t sh = uint32_t(at & 31);
    return sh ? (hi << sh) | (lo >> (32 - sh)) : hi;
}

// Decodes the code words that START in [pos, hi) and lie completely inside the payload.  Returns the position after the last
// one (>= hi unless the payload ends inside a code word) and their number.  WRITE: also stores the symbols from index `idx` on.
template <bool WRITE>
__device__ __forceinline__ void hd_decode_range(const HdArgs &a, unsigned long long pos, unsigned long long hi, unsigned long long *exit_pos,
                                                uint32_t *count, unsigned long long idx) {
    uint32_t cnt = 0;
    while (pos < hi) {
        uint32_t win = hd_window(a.words, pos);
        const uint32_t ent = __ldg(a.lut + (win >> (32 - HD_LUT_BITS)));
        uint32_t used = ent >> 28;
        int nd = int(ent & 0x0fffffffu);
        int2 c = __ldg(a.child + nd);
        unsigned long long p = pos + used;
        win <<= used;
        uint32_t avail = 32 - used;
        while (c.x >= 0) {  // codes longer than the table: one bit per step
            if (p >= a.end_bit) { p = a.end_bit + 1; break; }  // the payload ends inside the code word (also bounds the reads)
            if (!avail) { win = hd_window(a.words, p); avail = 32; }
            nd = (win >> 31) ? c.y : c.x;
            win <<= 1; avail--; p++;
            c = __ldg(a.child + nd);
        }
        if (p > a.end_bit) break;  // the payload ends inside this code word (huf.rs:190-204: None)
        if (WRITE && idx + cnt < a.n) {
            const uint2 v = __ldg(a.leaf + nd);
            uint8_t *o = a.out + (idx + cnt) * (unsigned long long)a.sym_bytes;
            if (a.sym_bytes == 6) {  // [i16; 3] little endian (ser.rs:188-195): three aligned 16-bit stores
                uint16_t *o2 = reinterpret_cast<uint16_t *>(o);
                o2[0] = uint16_t(v.x); o2[1] = uint16_t(v.x >> 16); o2[2] = uint16_t(v.y);
            } else {
                o[0] = uint8_t(v.x); o[1] = uint8_t(v.x >> 8); o[2] = uint8_t(v.x >> 16);
            }
        }
        cnt++;
        pos = p;
    }
    *exit_pos = pos;
    *count = cnt;
}

// steps 1 + 2: (re)synchronise the subsequences of one chunk; `round` 0 starts from the guesses
__global__ void __launch_bounds__(HD_THREADS) hd_sync_kernel(HdArgs a, int round) {
    __shared__ unsigned long long s_exit[HD_THREADS];
    const int tid = threadIdx.x;
    const unsigned long long g = (unsigned long long)blockIdx.x * HD_THREADS + tid;
    const bool active = g < a.nsub;
    const unsigned long long lo = g * HD_SUB, hi = min(lo + HD_SUB, a.end_bit);
    unsigned long long start = lo, my_exit = a.end_bit;
    uint32_t my_cnt = 0;
    bool dirty = false;
    if (active) {
        if (round == 0) { hd_decode_range<false>(a, start, hi, &my_exit, &my_cnt, 0); dirty = true; }
        else { start = a.sub_start[g]; my_exit = a.sub_exit[g]; my_cnt = a.sub_cnt[g]; }
    }
    s_exit[tid] = my_exit;
    __syncthreads();
    for (;;) {
        unsigned long long want = start;
        if (active && g > 0) {
            if (tid) want = s_exit[tid - 1];
            else if (round > 0) want = __ldcv(a.sub_exit + g - 1);  // previous chunk: its value of the last sweep (or newer)
        }
        const bool redo = active && want != start;
        __syncthreads();  // everybody has read its predecessor before anybody overwrites
        if (redo) {
            start = want;
            hd_decode_range<false>(a, start, hi, &my_exit, &my_cnt, 0);
            s_exit[tid] = my_exit;
            dirty = true;
        }
        if (!__syncthreads_or(redo)) break;
    }
    if (active && dirty) {
        a.sub_start[g] = start;
        a.sub_exit[g] = my_exit;
        a.sub_cnt[g] = my_cnt;
        if (round > 0) atomicOr(a.changed, 1u);  // somebody downstream may have read my old exit: sweep again
    }
}

// step 3a: per-chunk totals
__global__ void __launch_bounds__(HD_THREADS) hd_chunk_sums_kernel(HdArgs a) {
    __shared__ uint32_t s_w[HD_THREADS / 32];
    const unsigned long long g = (unsigned long long)blockIdx.x * HD_THREADS + threadIdx.x;
    uint32_t v = g < a.nsub ? a.sub_cnt[g] : 0u;
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (int i = 0; i < HD_THREADS / 32; i++) t += s_w[i];
        a.chunk_off[blockIdx.x + 1] = t;  // scanned in place by the next kernel; entry 0 is the leading zero
    }
}

// step 3b: single-CTA inclusive scan of the chunk totals (chunk_off[0] = 0, chunk_off[c + 1] = total of chunks 0..c)
__global__ void __launch_bounds__(1024) hd_scan_chunks_kernel(unsigned long long *chunk_off, unsigned long long nchunks) {
    __shared__ unsigned long long s_w[32];
    __shared__ unsigned long long s_carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) { s_carry = 0; chunk_off[0] = 0; }
    __syncthreads();
    for (unsigned long long base = 0; base < nchunks; base += 1024) {
        const unsigned long long i = base + tid;
        unsigned long long x = i < nchunks ? chunk_off[i + 1] : 0ull;
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) s_w[warp] = x;
        __syncthreads();
        unsigned long long before = s_carry;
        for (int j = 0; j < warp; j++) before += s_w[j];
        if (i < nchunks) chunk_off[i + 1] = before + x;
        __syncthreads();
        if (tid == 1023) s_carry = before + x;
        __syncthreads();
    }
}

// step 4: decode every subsequence once more from its synchronised start and write the symbols
__global__ void __launch_bounds__(HD_THREADS) hd_write_kernel(HdArgs a) {
    __shared__ uint32_t s_w[HD_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned long long g = (unsigned long long)blockIdx.x * HD_THREADS + tid;
    const bool active = g < a.nsub;
    const uint32_t cnt = active ? a.sub_cnt[g] : 0u;
    uint32_t x = cnt;  // inclusive scan inside the CTA
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) s_w[warp] = x;
    __syncthreads();
    unsigned long long idx = a.chunk_off[blockIdx.x] + (x - cnt);
    for (int j = 0; j < warp; j++) idx += s_w[j];
    if (!active || idx >= a.n) return;
    const unsigned long long lo = g * HD_SUB, hi = min(lo + HD_SUB, a.end_bit);
    unsigned long long e;
    uint32_t c;
    hd_decode_range<true>(a, a.sub_start[g], hi, &e, &c, idx);
}

// single-symbol alphabet: zero-length code, no payload bits (huf.rs:139-142)
__global__ void hd_fill_kernel(uint8_t *out, unsigned long long n, int sym_bytes, uint2 v) {
    const uint8_t b[8] = {uint8_t(v.x), uint8_t(v.x >> 8), uint8_t(v.x >> 16), uint8_t(v.x >> 24), uint8_t(v.y), uint8_t(v.y >> 8), 0, 0};
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * blockDim.x)
        for (int j = 0; j < sym_bytes; j++) out[i * sym_bytes + j] = b[j];
}

}  // namespace

int cniic_dev_huffman_decode(cniic_ctx *ctx, const uint8_t *payload, size_t len, const int32_t *child, const uint8_t *leaf_val, size_t nn,
                             int sym_bytes, size_t n, uint8_t *d_out, size_t *decoded) {
    if (decoded) *decoded = n;
    if (n == 0) return CNIIC_OK;
    if (nn == 0 || nn >= (size_t(1) << 28) || (sym_bytes != 3 && sym_bytes != 6)) return cniic_set_error(ctx, CNIIC_ERR_DECODE, "bad Huffman trie");
    uint32_t launched = 0;
    if (child[0] < 0) {  // the root is a leaf
        uint2 v = make_uint2(0, 0);
        memcpy(&v, leaf_val, 8);
        hd_fill_kernel<<<(int)std::max<size_t>(1, std::min<size_t>((n + 255) / 256, (size_t)ctx->sm_count * 8)), 256, 0, ctx->stream>>>(d_out, n, sym_bytes, v);
        CU_TRY(ctx, cudaGetLastError());
        ctx->launches += 1;
        return CNIIC_OK;
    }
    const unsigned long long end_bit = (unsigned long long)len * 8;
    if (end_bit == 0) {
        if (decoded) { *decoded = 0; return CNIIC_OK; }
        return cniic_set_error(ctx, CNIIC_ERR_DECODE, "truncated Huffman payload");
    }
    // prefix table (same construction as the sequential decoder's): node reached after <= HD_LUT_BITS bits
    std::vector<uint32_t> lut(size_t(1) << HD_LUT_BITS);
    for (uint32_t pre = 0; pre < (1u << HD_LUT_BITS); pre++) {
        int nd = 0, used = 0;
        while (used < HD_LUT_BITS && child[2 * nd] >= 0) {
            nd = ((pre >> (HD_LUT_BITS - 1 - used)) & 1) ? child[2 * nd + 1] : child[2 * nd];
            used++;
        }
        lut[pre] = uint32_t(nd) | (uint32_t(used) << 28);
    }
    const unsigned long long nsub = (end_bit + HD_SUB - 1) / HD_SUB;
    const unsigned long long nchunks = (nsub + HD_THREADS - 1) / HD_THREADS;
    const size_t padded = ((len + 3) & ~size_t(3)) + 16;  // a 32-bit window may start up to 32 bits behind the last payload bit
    DevBuf d_words(ctx), d_child(ctx), d_leaf(ctx), d_lut(ctx), d_start(ctx), d_exit(ctx), d_cnt(ctx), d_off(ctx), d_changed(ctx);
    CU_TRY(ctx, d_words.alloc(padded));
    CU_TRY(ctx, d_child.alloc(nn * 8));
    CU_TRY(ctx, d_leaf.alloc(nn * 8));
    CU_TRY(ctx, d_lut.alloc(lut.size() * 4));
    CU_TRY(ctx, d_start.alloc(nsub * 8));
    CU_TRY(ctx, d_exit.alloc(nsub * 8));
    CU_TRY(ctx, d_cnt.alloc(nsub * 4));
    CU_TRY(ctx, d_off.alloc((nchunks + 1) * 8));
    CU_TRY(ctx, d_changed.alloc(256));
    CU_TRY(ctx, cudaMemsetAsync(static_cast<uint8_t *>(d_words.p) + (padded - 20), 0, 20, ctx->stream));  // zero padding behind the payload
    CU_TRY(ctx, cudaMemcpyAsync(d_words.p, payload, len, cudaMemcpyHostToDevice, ctx->stream));
    CU_TRY(ctx, cudaMemcpyAsync(d_child.p, child, nn * 8, cudaMemcpyHostToDevice, ctx->stream));
    CU_TRY(ctx, cudaMemcpyAsync(d_leaf.p, leaf_val, nn * 8, cudaMemcpyHostToDevice, ctx->stream));
    CU_TRY(ctx, cudaMemcpyAsync(d_lut.p, lut.data(), lut.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    HdArgs a{};
    a.words = d_words.as<uint32_t>();
    a.end_bit = end_bit;
    a.nsub = nsub;
    a.child = d_child.as<int2>();
    a.leaf = d_leaf.as<uint2>();
    a.lut = d_lut.as<uint32_t>();
    a.sub_start = d_start.as<unsigned long long>();
    a.sub_exit = d_exit.as<unsigned long long>();
    a.sub_cnt = d_cnt.as<uint32_t>();
    a.chunk_off = d_off.as<unsigned long long>();
    a.changed = d_changed.as<uint32_t>();
    a.out = d_out;
    a.n = n;
    a.sym_bytes = sym_bytes;
    // sweeps until no chunk changes (the fixpoint is reached after at most nchunks sweeps; typically two)
    for (unsigned long long round = 0;; round++) {
        CU_TRY(ctx, cudaMemsetAsync(d_changed.p, 0, 4, ctx->stream));
        hd_sync_kernel<<<(unsigned)nchunks, HD_THREADS, 0, ctx->stream>>>(a, round ? 1 : 0);
        launched++;
        CU_TRY(ctx, cudaGetLastError());
        if (round == 0) continue;  // the first sweep never reads across chunks: always sweep once more
        uint32_t changed = 0;
        CU_TRY(ctx, cudaMemcpyAsync(&changed, d_changed.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
        CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        if (!changed) {
            if (getenv("CNIIC_DEBUG")) fprintf(stderr, "cniic huffdec: %llu chunks, fixpoint after %llu sweeps\n", nchunks, round + 1);
            break;
        }
        if (round > nchunks + 1) return cniic_set_error(ctx, CNIIC_ERR_CUDA, "Huffman decoder did not reach its fixpoint");
    }
    hd_chunk_sums_kernel<<<(unsigned)nchunks, HD_THREADS, 0, ctx->stream>>>(a);
    hd_scan_chunks_kernel<<<1, 1024, 0, ctx->stream>>>(a.chunk_off, nchunks);
    launched += 2;
    CU_TRY(ctx, cudaGetLastError());
    unsigned long long total = 0;
    CU_TRY(ctx, cudaMemcpyAsync(&total, a.chunk_off + nchunks, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->launches += launched;
    if (total < n) {
        if (!decoded) return cniic_set_error(ctx, CNIIC_ERR_DECODE, "truncated Huffman payload (%llu of %zu symbols)", total, n);
        *decoded = (size_t)total;  // the caller's consumer is zipped with the symbol iterator and simply stops (hilbertc.rs:425-428)
        if (total == 0) return CNIIC_OK;
    }
    hd_write_kernel<<<(unsigned)nchunks, HD_THREADS, 0, ctx->stream>>>(a);
    ctx->launches += 1;
    CU_TRY(ctx, cudaGetLastError());
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));  // the scratch buffers go back to the cache with this scope
    return CNIIC_OK;
}
