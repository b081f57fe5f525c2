// zhuff_core.h -- zstd frames whose blocks carry Huffman-coded literals only, built on the device.
//
// SURVEY.md §8(f) row 3 (the make_fastq writer).  The reference writes R1.fq.zst / R2.fq.zst through the zstd crate
// (python/src/lib.rs:136-146, seqspec/src/utils.rs:33-36: level 9, 8 worker threads), which is the floor of an
// API-level run once everything else is on the GPU.  This file produces standard zstd blocks (RFC 8878) on the
// device instead: every block of up to 128 KB is either
//   * a Compressed_Block whose Literals_Section is Huffman-coded in four streams with a directly encoded weight table
//     (RFC 8878 §3.1.1.3.1, §4.2.1) and whose Sequences_Section is empty (Number_of_Sequences = 0), or
//   * a Raw_Block, when the block is tiny, holds a byte above 128, a single distinct byte, or would not shrink.
// Any zstd decoder reads the result; there are no matches, so the ratio is that of an order-0 entropy coder.
//
// The functions below are the per-thread phases of one CTA working on one block, written so that tests/emul can run
// them on the host (phase by phase, thread by thread) and hand the output to libzstd's decoder in the CPU tier.
//   zh_rank          position of every present byte value in (count, value) order  (thread per byte value)
//   zh_build_table   Huffman code lengths limited to 11 bits, canonical codes, weight table    (one thread)
//   zh_span_bits     bits the thread's span of a stream needs                       (thread per span)
//   zh_span_encode   the span's codes OR-ed into the stream's bit buffer            (thread per span)
#ifndef PCL_ZHUFF_CORE_H
#define PCL_ZHUFF_CORE_H

#include <stdint.h>

#ifdef __CUDACC__
#define ZH_HD __host__ __device__ __forceinline__
#define ZH_HD_NOINLINE __host__ __device__ __noinline__
#else
#define ZH_HD inline
#define ZH_HD_NOINLINE inline
#endif

#define ZH_BLOCK_BYTES 131072u                    // Block_Maximum_Size of the zstd format
#define ZH_THREADS 256u
#define ZH_MAX_BITS 11u                           // longest Huffman code the format allows for literals
#define ZH_MIN_COMPRESS 1024u                     // smaller blocks are stored raw
#define ZH_STREAM_CAP_WORDS 8208u                 // bit buffer of one stream: 32 KB + 64 bytes
#define ZH_SLOT_BYTES (ZH_BLOCK_BYTES + 16u)      // worst case of a block in the output: 3-byte header + raw bytes
#define ZH_FRAME_HEADER_BYTES 14u                 // magic 4, descriptor 1, window 1, content size 8

struct ZhTable {
    uint32_t sym[256];           // code | number of bits << 16 (one shared-memory lookup per byte in the coding loops)
    uint16_t code[256];
    uint8_t nbits[256];
    uint8_t desc[72];            // Huffman_Tree_Description: header byte + packed 4-bit weights
    uint32_t desc_len;
    uint32_t ok;                 // 0: store the block raw
    uint64_t total_bits;         // of the whole block with this table
};

// rank of byte value t among the present values in ascending (count, value) order; 0xFFFF when absent
ZH_HD uint32_t zh_rank(const uint32_t *hist, uint32_t t) {
    const uint32_t c = hist[t];
    if (c == 0) return 0xFFFFu;
    uint32_t r = 0;
    for (uint32_t u = 0; u < 256u; ++u) {
        const uint32_t cu = hist[u];
        r += (cu != 0 && (cu < c || (cu == c && u < t))) ? 1u : 0u;
    }
    return r;
}

// hist[256]: byte counts of the block; rank[256] from zh_rank.  One thread.
ZH_HD_NOINLINE void zh_build_table(const uint32_t *hist, const uint16_t *rank, ZhTable *t) {
    uint8_t sym[256];            // present byte values in ascending (count, value) order
    uint32_t n = 0, max_sym = 0;
    for (uint32_t s = 0; s < 256u; ++s) {
        t->code[s] = 0;
        t->nbits[s] = 0;
        t->sym[s] = 0;
        if (rank[s] != 0xFFFFu) { sym[rank[s]] = (uint8_t)s; ++n; max_sym = s; }
    }
    t->ok = 0;
    t->desc_len = 0;
    t->total_bits = 0;
    if (n < 2 || max_sym > 128u) return;          // a direct weight table lists at most 128 weights (values 0..127) + the deduced one
    // Huffman tree over the sorted leaves with the two-queue method: leaves 0..n-1, internal nodes n..2n-2 in creation order
    uint32_t weight[512];
    uint16_t parent[512];
    for (uint32_t k = 0; k < n; ++k) weight[k] = hist[sym[k]];
    uint32_t leaf = 0, inner = n, made = n;
    while (made < 2 * n - 1) {
        uint32_t pick[2];
        for (int q = 0; q < 2; ++q) {
            if (leaf < n && (inner >= made || weight[leaf] <= weight[inner])) pick[q] = leaf++;
            else pick[q] = inner++;
        }
        weight[made] = weight[pick[0]] + weight[pick[1]];
        parent[pick[0]] = (uint16_t)made;
        parent[pick[1]] = (uint16_t)made;
        ++made;
    }
    uint8_t depth[512];
    depth[2 * n - 2] = 0;
    for (int k = (int)(2 * n - 3); k >= 0; --k) {
        const uint32_t d = depth[parent[k]] + 1u;
        depth[k] = (uint8_t)(d > 255u ? 255u : d);
    }
    // limit the lengths to ZH_MAX_BITS: clamp, then repair the Kraft sum (in units of 2^-ZH_MAX_BITS) to exactly one
    uint8_t len[256];
    uint32_t kraft = 0;
    for (uint32_t k = 0; k < n; ++k) {
        len[k] = depth[k] > ZH_MAX_BITS ? (uint8_t)ZH_MAX_BITS : depth[k];
        kraft += 1u << (ZH_MAX_BITS - len[k]);
    }
    const uint32_t full = 1u << ZH_MAX_BITS;
    while (kraft > full) {                        // overfull: lengthen the rarest symbol among the longest codes below the limit
        int found = -1;
        for (uint32_t L = ZH_MAX_BITS - 1; L >= 1 && found < 0; --L)
            for (uint32_t k = 0; k < n; ++k)
                if (len[k] == L) { found = (int)k; break; }
        if (found < 0) return;
        kraft -= 1u << (ZH_MAX_BITS - len[found] - 1u);
        len[found] += 1;
    }
    for (int pass = 0; pass < 64 && kraft < full; ++pass) {      // underfull: shorten the most frequent symbols that fit the gap
        bool moved = false;
        for (int k = (int)n - 1; k >= 0 && kraft < full; --k) {
            const uint32_t gain = 1u << (ZH_MAX_BITS - len[k]);
            if (len[k] > 1 && gain <= full - kraft) { len[k] -= 1; kraft += gain; moved = true; }
        }
        if (!moved) break;
    }
    if (kraft != full) return;
    uint32_t max_bits = 0;
    for (uint32_t k = 0; k < n; ++k) max_bits = len[k] > max_bits ? len[k] : max_bits;
    // canonical codes (RFC 8878 §4.2.1.3): within a length in ascending byte value, the longest codes first from 0
    uint32_t per_len[ZH_MAX_BITS + 2] = {0}, next_val[ZH_MAX_BITS + 2] = {0};
    for (uint32_t k = 0; k < n; ++k) { t->nbits[sym[k]] = len[k]; per_len[len[k]] += 1; }
    uint32_t v = 0;
    for (uint32_t L = max_bits; L >= 1; --L) { next_val[L] = v; v = (v + per_len[L]) >> 1; }
    uint64_t bits = 0;
    for (uint32_t s = 0; s < 256u; ++s) {
        const uint32_t L = t->nbits[s];
        if (L) { t->code[s] = (uint16_t)next_val[L]++; bits += (uint64_t)hist[s] * L; t->sym[s] = t->code[s] | (L << 16); }
    }
    // weights of the byte values 0 .. max_sym-1, four bits each; the weight of max_sym is deduced by the decoder
    const uint32_t n_listed = max_sym;
    t->desc[0] = (uint8_t)(127u + n_listed);
    for (uint32_t k = 0; k < (n_listed + 1u) / 2u; ++k) t->desc[1 + k] = 0;
    for (uint32_t s = 0; s < n_listed; ++s) {
        const uint32_t w = t->nbits[s] ? max_bits + 1u - t->nbits[s] : 0u;
        t->desc[1 + s / 2] |= (uint8_t)((s & 1u) ? w : (w << 4));
    }
    t->desc_len = 1u + (n_listed + 1u) / 2u;
    t->total_bits = bits;
    t->ok = 1;
}

// The four streams of a block of n bytes: streams 0..2 hold (n + 3) / 4 bytes each, the last one the rest.
ZH_HD uint32_t zh_stream_begin(uint32_t n, uint32_t s) { return s * ((n + 3u) / 4u); }
ZH_HD uint32_t zh_stream_len(uint32_t n, uint32_t s) { return s < 3u ? (n + 3u) / 4u : n - 3u * ((n + 3u) / 4u); }

// A stream is written back to front (the decoder reads the bit buffer from its end): position p of the coding order
// is byte src[len - 1 - p].  Thread t owns positions [t * per, min((t + 1) * per, len)), i.e. the bytes
// [len - min((t + 1) * per, len), len - t * per) walked downwards.  Spans that are whole 16-byte vectors (every span of
// a full 128 KB block out of an aligned buffer) are read with 128-bit loads.
#ifdef __CUDA_ARCH__
#define ZH_OR(ptr, val) atomicOr((ptr), (val))
#else
#define ZH_OR(ptr, val) (*(ptr) |= (val))
#endif

ZH_HD bool zh_span_is_vector(const uint8_t *lo_ptr, uint32_t n) { return n && (n & 15u) == 0 && (((uintptr_t)lo_ptr) & 15u) == 0; }

struct ZhWord4 { uint32_t w[4]; };
ZH_HD ZhWord4 zh_load16(const uint8_t *p) {
#ifdef __CUDA_ARCH__
    const uint4 v = *reinterpret_cast<const uint4 *>(p);
    return ZhWord4{{v.x, v.y, v.z, v.w}};
#else
    ZhWord4 o;
    for (int k = 0; k < 4; ++k) o.w[k] = (uint32_t)p[4 * k] | ((uint32_t)p[4 * k + 1] << 8) | ((uint32_t)p[4 * k + 2] << 16) | ((uint32_t)p[4 * k + 3] << 24);
    return o;
#endif
}

ZH_HD uint32_t zh_span_bits(const uint8_t *src, uint32_t len, uint32_t per, uint32_t tid, const uint32_t *sym) {
    const uint32_t lo = tid * per, hi = lo + per < len ? lo + per : len;
    if (lo >= hi) return 0;
    const uint8_t *base = src + (len - hi);                      // lowest byte of the span
    const uint32_t n = hi - lo;
    uint32_t b = 0;
    if (zh_span_is_vector(base, n)) {
        for (uint32_t v = 0; v < n; v += 16u) {
            const ZhWord4 q = zh_load16(base + v);
            for (int k = 0; k < 4; ++k)
                b += (sym[q.w[k] & 255u] >> 16) + (sym[(q.w[k] >> 8) & 255u] >> 16) + (sym[(q.w[k] >> 16) & 255u] >> 16) + (sym[q.w[k] >> 24] >> 16);
        }
        return b;
    }
    for (uint32_t k = 0; k < n; ++k) b += sym[base[k]] >> 16;
    return b;
}

// OR the codes of the thread's span into the zeroed bit buffer `out`, starting at bit `bitpos`
ZH_HD void zh_span_encode(const uint8_t *src, uint32_t len, uint32_t per, uint32_t tid, const uint32_t *sym, uint32_t bitpos, uint32_t *out) {
    const uint32_t lo = tid * per, hi = lo + per < len ? lo + per : len;
    if (lo >= hi) return;
    const uint8_t *base = src + (len - hi);
    const uint32_t n = hi - lo;
    uint32_t w = bitpos >> 5, fill = bitpos & 31u;
    uint64_t acc = 0;
#define ZH_PUT(byte_)                                                                         \
    do {                                                                                      \
        const uint32_t e_ = sym[(byte_)];                                                     \
        acc |= (uint64_t)(e_ & 0xFFFFu) << fill;                                              \
        fill += e_ >> 16;                                                                     \
        if (fill >= 32u) { ZH_OR(out + w, (uint32_t)acc); acc >>= 32; fill -= 32u; ++w; }     \
    } while (0)
    if (zh_span_is_vector(base, n)) {
        for (uint32_t v = n; v > 0; v -= 16u) {                  // downwards: the span's last byte is coded first
            const ZhWord4 q = zh_load16(base + v - 16u);
            for (int k = 3; k >= 0; --k) {
                ZH_PUT(q.w[k] >> 24); ZH_PUT((q.w[k] >> 16) & 255u); ZH_PUT((q.w[k] >> 8) & 255u); ZH_PUT(q.w[k] & 255u);
            }
        }
    } else {
        for (uint32_t k = n; k > 0; --k) ZH_PUT(base[k - 1u]);
    }
#undef ZH_PUT
    if (fill && (uint32_t)acc) ZH_OR(out + w, (uint32_t)acc);
}

// 3-byte block header: Last_Block, Block_Type (0 raw, 2 compressed), Block_Size
ZH_HD void zh_put_block_header(uint8_t *dst, uint32_t last, uint32_t type, uint32_t size) {
    const uint32_t h = (last & 1u) | (type << 1) | (size << 3);
    dst[0] = (uint8_t)h; dst[1] = (uint8_t)(h >> 8); dst[2] = (uint8_t)(h >> 16);
}

// 5-byte Literals_Section_Header: Compressed_Literals_Block (2), Size_Format 3 (four streams, 18-bit sizes)
ZH_HD void zh_put_literals_header(uint8_t *dst, uint32_t regenerated, uint32_t compressed) {
    const uint64_t h = 2ull | (3ull << 2) | ((uint64_t)regenerated << 4) | ((uint64_t)compressed << 22);
    for (int k = 0; k < 5; ++k) dst[k] = (uint8_t)(h >> (8 * k));
}

// Frame header written by the host in front of the blocks: magic, no checksum / dictionary, 128 KB window, content size
inline void zh_put_frame_header(uint8_t *dst, uint64_t content_bytes) {
    dst[0] = 0x28; dst[1] = 0xB5; dst[2] = 0x2F; dst[3] = 0xFD;
    dst[4] = 0xC0;               // Frame_Content_Size_flag 3 (8 bytes), Single_Segment 0
    dst[5] = 0x38;               // Window_Descriptor: exponent 7 -> 2^17 bytes
    for (int k = 0; k < 8; ++k) dst[6 + k] = (uint8_t)(content_bytes >> (8 * k));
}

// Shared state of one CTA (device: __shared__; host emulation: a plain object)
struct ZhShared {
    uint32_t hist[256];
    uint32_t whist[ZH_THREADS / 32u][256];       // one histogram per warp: the bytes of FASTQ text collide on a handful of values
    uint16_t rank[256];
    ZhTable table;
    uint32_t span[ZH_THREADS];
    uint32_t out[ZH_STREAM_CAP_WORDS];
    uint32_t stream_bits;
    uint32_t fail;
};

#ifdef __CUDA_ARCH__
#define ZH_ADD(ptr, val) atomicAdd((ptr), (val))
#else
#define ZH_ADD(ptr, val) (*(ptr) += (val))
#endif

// One block of n <= ZH_BLOCK_BYTES bytes -> block header + content in `slot` (ZH_SLOT_BYTES); returns the bytes used.
// `ex.each(f)` runs f(tid) for the ZH_THREADS threads of the CTA and separates the phases (device: __syncthreads()).
template <class Exec>
ZH_HD uint32_t zh_compress_block(Exec &ex, ZhShared &sh, const uint8_t *src, uint32_t n, uint32_t last, uint8_t *slot) {
    ex.each([&](uint32_t tid) {
        for (uint32_t k = 0; k < ZH_THREADS / 32u; ++k) sh.whist[k][tid] = 0;
        if (tid == 0) { sh.fail = 0; sh.table.desc_len = 0; }
    });
    ex.each([&](uint32_t tid) {
        uint32_t *h = sh.whist[tid >> 5];
        const uint32_t head = (uint32_t)((16u - (((uintptr_t)src) & 15u)) & 15u);      // bytes in front of the first aligned vector
        const uint32_t n_head = head < n ? head : n, n_vec = (n - n_head) / 16u;
        for (uint32_t v = tid; v < n_vec; v += ZH_THREADS) {
            const ZhWord4 q = zh_load16(src + n_head + 16u * v);
            for (int k = 0; k < 4; ++k) {
                ZH_ADD(&h[q.w[k] & 255u], 1u); ZH_ADD(&h[(q.w[k] >> 8) & 255u], 1u);
                ZH_ADD(&h[(q.w[k] >> 16) & 255u], 1u); ZH_ADD(&h[q.w[k] >> 24], 1u);
            }
        }
        for (uint32_t k = tid; k < n_head; k += ZH_THREADS) ZH_ADD(&h[src[k]], 1u);
        for (uint32_t k = n_head + 16u * n_vec + tid; k < n; k += ZH_THREADS) ZH_ADD(&h[src[k]], 1u);
    });
    ex.each([&](uint32_t tid) {
        uint32_t c = 0;
        for (uint32_t k = 0; k < ZH_THREADS / 32u; ++k) c += sh.whist[k][tid];
        sh.hist[tid] = c;
    });
    ex.each([&](uint32_t tid) { sh.rank[tid] = (uint16_t)zh_rank(sh.hist, tid); });
    ex.each([&](uint32_t tid) {
        if (tid) return;
        sh.fail = 1;
        if (n < ZH_MIN_COMPRESS) return;
        zh_build_table(sh.hist, sh.rank, &sh.table);
        if (!sh.table.ok) return;
        const uint64_t bound = 3u + 5u + sh.table.desc_len + 6u + sh.table.total_bits / 8u + 4u + 1u;
        if (bound < 3u + (uint64_t)n) sh.fail = 0;
    });
    uint32_t pos = 3u + 5u + sh.table.desc_len + 6u;
    uint32_t stream_bytes[4] = {0, 0, 0, 0};
    for (uint32_t s = 0; s < 4u && !sh.fail; ++s) {
        const uint8_t *ssrc = src + zh_stream_begin(n, s);
        const uint32_t slen = zh_stream_len(n, s), per = (slen + ZH_THREADS - 1u) / ZH_THREADS;
        ex.each([&](uint32_t tid) {
            for (uint32_t k = tid; k < ZH_STREAM_CAP_WORDS; k += ZH_THREADS) sh.out[k] = 0;
            sh.span[tid] = zh_span_bits(ssrc, slen, per, tid, sh.table.sym);
        });
        ex.each([&](uint32_t tid) {
            if (tid) return;
            uint32_t run = 0;
            for (uint32_t k = 0; k < ZH_THREADS; ++k) { const uint32_t b = sh.span[k]; sh.span[k] = run; run += b; }
            sh.stream_bits = run;
            if (run / 8u + 1u > 4u * ZH_STREAM_CAP_WORDS - 8u) sh.fail = 1;
        });
        if (sh.fail) break;
        const uint32_t bits = sh.stream_bits, bytes = bits / 8u + 1u;
        ex.each([&](uint32_t tid) { zh_span_encode(ssrc, slen, per, tid, sh.table.sym, sh.span[tid], sh.out); });
        ex.each([&](uint32_t tid) { if (tid == 0) sh.out[bits >> 5] |= 1u << (bits & 31u); });     // end mark above the last code
        ex.each([&](uint32_t tid) {
            const uint8_t *ob = reinterpret_cast<const uint8_t *>(sh.out);
            for (uint32_t k = tid; k < bytes; k += ZH_THREADS) slot[pos + k] = ob[k];
        });
        stream_bytes[s] = bytes;
        pos += bytes;
    }
    if (sh.fail) {                                 // Raw_Block
        ex.each([&](uint32_t tid) {
            if (tid == 0) zh_put_block_header(slot, last, 0u, n);
            for (uint32_t k = tid; k < n; k += ZH_THREADS) slot[3u + k] = src[k];
        });
        return 3u + n;
    }
    ex.each([&](uint32_t tid) {
        if (tid) return;
        const uint32_t dl = sh.table.desc_len;
        zh_put_block_header(slot, last, 2u, pos + 1u - 3u);
        zh_put_literals_header(slot + 3, n, pos - 8u);                      // tree description + jump table + streams
        for (uint32_t k = 0; k < dl; ++k) slot[8u + k] = sh.table.desc[k];
        for (uint32_t k = 0; k < 3u; ++k) {
            slot[8u + dl + 2u * k] = (uint8_t)stream_bytes[k];
            slot[8u + dl + 2u * k + 1u] = (uint8_t)(stream_bytes[k] >> 8);
        }
        slot[pos] = 0;                                                      // Number_of_Sequences = 0
    });
    return pos + 1u;
}

#endif  // PCL_ZHUFF_CORE_H
