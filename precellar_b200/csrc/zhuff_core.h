// zhuff_core.h -- standard zstd frames built on the device: Huffman-coded literals plus FSE-coded sequences.
//
// SURVEY.md §8(f) row 3 (the make_fastq writer).  The reference writes R1.fq.zst / R2.fq.zst through the zstd crate
// (python/src/lib.rs:136-146, seqspec/src/utils.rs:33-36: level 9, 8 worker threads), which is the floor of an
// API-level run once everything else is on the GPU.  This file produces standard zstd blocks (RFC 8878) on the
// device instead: every block of up to 128 KB is either
//   * a Compressed_Block: the matches of every line against the line four lines earlier (the same line of the previous
//     FASTQ record: a common prefix and a common suffix) as a Sequences_Section coded with the format's predefined
//     FSE distributions (RFC 8878 §3.1.1.3.2), and the remaining bytes as a Literals_Section, Huffman-coded in four
//     streams with a directly encoded weight table (§3.1.1.3.1, §4.2.1), or
//   * a Raw_Block, when the block is tiny or would not shrink.
// Any zstd decoder reads the result.
//
// The functions below are the per-thread phases of one CTA working on one block, written so that tests/emul can run
// them on the host (phase by phase, thread by thread) and hand the output to libzstd's decoder in the CPU tier.
//   zh_find_matches       newline index, prefix / suffix matches per line, compaction, literal buffer
//   zh_literals_section   histogram, zh_rank / zh_build_table (length-limited Huffman code), zh_span_bits / zh_span_encode
//   zh_sequences_section  codes, the three FSE state chains (one thread), bit counts, bits
//   zh_compress_block     the block: sections, headers, raw fallback
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

// 16 bytes -> 16-bit mask of the positions holding '\n' (bit k = byte k)
ZH_HD uint32_t zh_nl_mask16(const ZhWord4 q) {
    uint32_t m = 0;
    for (int k = 0; k < 4; ++k) {
        const uint32_t x = q.w[k] ^ 0x0A0A0A0Au;
        const uint32_t z = ~(((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x) & 0x80808080u;     // 0x80 where the byte was '\n'
        m |= ((((z >> 7) * 0x01020408u) >> 24) & 0xFu) << (4 * k);
    }
    return m;
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

// ------------------------------------------------------------------------------------------
// Sequences (RFC 8878 §3.1.1.3.2): matches between a line and the line four lines earlier - in FASTQ text the same
// line of the previous record - coded with the format's predefined FSE distributions.
// ------------------------------------------------------------------------------------------
#define ZH_MAX_LINES 4096u                        // lines of a block the match finder indexes (more: the block is coded without matches)
#define ZH_MAX_SEQ (2u * ZH_MAX_LINES)            // a line yields at most a prefix and a suffix match
#define ZH_MIN_MATCH 8u                           // shorter matches cost more (about 4.5 bytes per sequence) than their literals

// Per-block scratch in global memory (the device keeps one per block of a launch; the host emulation one)
struct ZhScratch {
    uint32_t nl[ZH_MAX_LINES];                    // position of every '\n'
    uint32_t cand_pos[ZH_MAX_SEQ], cand_off[ZH_MAX_SEQ];
    uint16_t cand_len[ZH_MAX_SEQ];                // 0: no match
    uint32_t seq_pos[ZH_MAX_SEQ + 1], seq_off[ZH_MAX_SEQ + 1];     // compacted, in text order; entry n_seq closes the block (length 0)
    uint16_t seq_len[ZH_MAX_SEQ + 1];
    uint32_t seq_lit[ZH_MAX_SEQ + 1];             // offset of the sequence's literal run in lit[]
    uint8_t lit[ZH_BLOCK_BYTES + 64];             // the bytes no match covers, in text order
};

// FSE encoding table of one predefined distribution (the construction of FSE_buildCTable, RFC 8878 §4.1.1)
struct ZhFse {
    uint16_t state[64];          // next state, indexed by the symbol's cumulative rank
    uint32_t delta_bits[53];     // (bits to flush << 16) - threshold
    int32_t delta_state[53];
    uint32_t log;
};

ZH_HD uint32_t zh_highbit(uint32_t v) {          // position of the highest set bit (v > 0)
#ifdef __CUDA_ARCH__
    return 31u - (uint32_t)__clz((int)v);
#else
    uint32_t r = 0;
    while (v >>= 1) ++r;
    return r;
#endif
}

ZH_HD uint32_t zh_popc(uint32_t v) {
#ifdef __CUDA_ARCH__
    return (uint32_t)__popc(v);
#else
    uint32_t c = 0;
    for (; v; v &= v - 1u) ++c;
    return c;
#endif
}
ZH_HD uint32_t zh_ctz(uint32_t v) {               // v > 0
#ifdef __CUDA_ARCH__
    return (uint32_t)__ffs((int)v) - 1u;
#else
    uint32_t k = 0;
    while (!((v >> k) & 1u)) ++k;
    return k;
#endif
}

// length of the common run of x[0..lim) and y[0..lim) walking upwards (dir = +1) or downwards (dir = -1) from x / y;
// eight bytes are fetched per round so that a round costs one memory latency, not eight
ZH_HD uint32_t zh_common_run(const uint8_t *x, const uint8_t *y, uint32_t lim, int dir) {
    uint32_t L = 0;
    while (L + 8u <= lim) {
        uint32_t diff = 0;
        for (int k = 7; k >= 0; --k) {
            const int o = dir * (int)(L + (uint32_t)k);
            diff = (diff << 1) | (x[o] != y[o] ? 1u : 0u);
        }
        if (diff) {
            uint32_t k = 0;
            while (!((diff >> k) & 1u)) ++k;
            return L + k;
        }
        L += 8u;
    }
    while (L < lim && x[dir * (int)L] == y[dir * (int)L]) ++L;
    return L;
}

ZH_HD_NOINLINE void zh_fse_build(const int8_t *norm, uint32_t n_sym, uint32_t log, ZhFse *t) {
    const uint32_t size = 1u << log, mask = size - 1u, step = (size >> 1) + (size >> 3) + 3u;
    uint8_t spread[64];
    uint32_t cumul[54];
    uint32_t high = size - 1u;
    cumul[0] = 0;
    for (uint32_t u = 1; u <= n_sym; ++u) {
        if (norm[u - 1] == -1) { cumul[u] = cumul[u - 1] + 1u; spread[high--] = (uint8_t)(u - 1); }
        else cumul[u] = cumul[u - 1] + (uint32_t)norm[u - 1];
    }
    uint32_t pos = 0;
    for (uint32_t sy = 0; sy < n_sym; ++sy)
        for (int k = 0; k < norm[sy]; ++k) {
            spread[pos] = (uint8_t)sy;
            pos = (pos + step) & mask;
            while (pos > high) pos = (pos + step) & mask;
        }
    for (uint32_t u = 0; u < size; ++u) { const uint32_t sy = spread[u]; t->state[cumul[sy]++] = (uint16_t)(size + u); }
    uint32_t total = 0;
    for (uint32_t sy = 0; sy < n_sym; ++sy) {
        const int c = norm[sy];
        if (c == 0) { t->delta_bits[sy] = ((log + 1u) << 16) - size; t->delta_state[sy] = 0; }
        else if (c == -1 || c == 1) { t->delta_bits[sy] = (log << 16) - size; t->delta_state[sy] = (int32_t)total - 1; total += 1; }
        else {
            const uint32_t max_out = log - zh_highbit((uint32_t)c - 1u), min_plus = (uint32_t)c << max_out;
            t->delta_bits[sy] = (max_out << 16) - min_plus;
            t->delta_state[sy] = (int32_t)total - c;
            total += (uint32_t)c;
        }
    }
    t->log = log;
}

// predefined distributions of literal lengths (which = 0), match lengths (1) and offsets (2) (RFC 8878 §3.1.1.3.2.2)
ZH_HD_NOINLINE void zh_fse_predefined(uint32_t which, ZhFse *t) {
    const int8_t nll[36] = {4, 3, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 2, 2, 3, 2, 1, 1, 1, 1, 1, -1, -1, -1, -1};
    const int8_t nml[53] = {1, 4, 3, 2, 2, 2, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1,
                            1, 1, 1, 1, 1, 1, 1, 1, 1, -1, -1, -1, -1, -1, -1, -1};
    const int8_t nof[29] = {1, 1, 1, 1, 1, 1, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, -1, -1, -1, -1, -1};
    if (which == 0) zh_fse_build(nll, 36, 6, t);
    else if (which == 1) zh_fse_build(nml, 53, 6, t);
    else zh_fse_build(nof, 29, 5, t);
}

// code and number of extra bits of a literal length / of a match length minus 3 (ZSTD_LLcode / ZSTD_MLcode)
ZH_HD uint32_t zh_ll_code(uint32_t v, uint32_t *bits) {
    uint32_t c;
    if (v < 16u) c = v;
    else if (v < 24u) c = 16u + ((v - 16u) >> 1);
    else if (v < 32u) c = 20u + ((v - 24u) >> 2);
    else if (v < 48u) c = 22u + ((v - 32u) >> 3);
    else if (v < 64u) c = 24u;
    else c = zh_highbit(v) + 19u;
    *bits = c < 16u ? 0u : c < 20u ? 1u : c < 22u ? 2u : c < 24u ? 3u : c == 24u ? 4u : c - 19u;
    return c;
}
ZH_HD uint32_t zh_ml_code(uint32_t v, uint32_t *bits) {
    uint32_t c;
    if (v < 32u) c = v;
    else if (v < 40u) c = 32u + ((v - 32u) >> 1);
    else if (v < 48u) c = 36u + ((v - 40u) >> 2);
    else if (v < 64u) c = 38u + ((v - 48u) >> 3);
    else if (v < 96u) c = 40u + ((v - 64u) >> 4);
    else if (v < 128u) c = 42u;
    else c = zh_highbit(v) + 36u;
    *bits = c < 32u ? 0u : c < 36u ? 1u : c < 38u ? 2u : c < 40u ? 3u : c < 42u ? 4u : c == 42u ? 5u : c - 36u;
    return c;
}

// one step of an FSE encoder: the bits the state gives up for `sym` and the state that follows (FSE_encodeSymbol)
ZH_HD uint32_t zh_fse_nbits(const ZhFse &t, uint32_t state, uint32_t sym) { return (state + t.delta_bits[sym]) >> 16; }
ZH_HD uint32_t zh_fse_next(const ZhFse &t, uint32_t state, uint32_t sym) {
    return t.state[(int32_t)(state >> zh_fse_nbits(t, state, sym)) + t.delta_state[sym]];
}
ZH_HD uint32_t zh_fse_first(const ZhFse &t, uint32_t sym) {       // FSE_initCState2
    const uint32_t nb = (t.delta_bits[sym] + (1u << 15)) >> 16;
    const uint32_t value = (nb << 16) - t.delta_bits[sym];
    return t.state[(int32_t)(value >> nb) + t.delta_state[sym]];
}

// Shared state of one CTA (device: __shared__; host emulation: a plain object)
struct ZhShared {
    uint32_t hist[256];
    uint32_t whist[ZH_THREADS / 32u][256];       // one histogram per warp: the bytes of FASTQ text collide on a handful of values
    uint16_t rank[256];
    ZhTable table;
    uint32_t span[ZH_THREADS];
    uint32_t out[ZH_STREAM_CAP_WORDS];
    ZhFse fse[3];                                // literal lengths, match lengths, offsets
    // per sequence and code: first the code (written by all threads), then - in place, by the thread that walks that
    // code's state chain - the state the code meets (minus the table size)
    uint8_t seq_code[3][ZH_MAX_SEQ];
    uint32_t end_state[3];                       // states left after the last (text-order first) sequence
    uint32_t stream_bits;
    uint32_t fail;
    uint32_t n_lines, n_seq, n_lit, seq_bytes, lit_bytes;
};

#ifdef __CUDA_ARCH__
#define ZH_ADD(ptr, val) atomicAdd((ptr), (val))
#else
#define ZH_ADD(ptr, val) (*(ptr) += (val))
#endif

// exclusive scan of sh.span[] by one thread; returns nothing: sh.stream_bits receives the total
template <class Exec>
ZH_HD void zh_scan_spans(Exec &ex, ZhShared &sh) {
    ex.each([&](uint32_t tid) {
        if (tid) return;
        uint32_t run = 0;
        for (uint32_t k = 0; k < ZH_THREADS; ++k) { const uint32_t b = sh.span[k]; sh.span[k] = run; run += b; }
        sh.stream_bits = run;
    });
}

// Literals_Section of `n` bytes at `lit` -> dst; returns its size.  Huffman-coded in four streams when that pays, else raw.
template <class Exec>
ZH_HD uint32_t zh_literals_section(Exec &ex, ZhShared &sh, const uint8_t *lit, uint32_t n, uint8_t *dst) {
    ex.each([&](uint32_t tid) {
        for (uint32_t k = 0; k < ZH_THREADS / 32u; ++k) sh.whist[k][tid] = 0;
        if (tid == 0) { sh.fail = 0; sh.table.desc_len = 0; }
    });
    ex.each([&](uint32_t tid) {
        uint32_t *h = sh.whist[tid >> 5];
        const uint32_t head = (uint32_t)((16u - (((uintptr_t)lit) & 15u)) & 15u);      // bytes in front of the first aligned vector
        const uint32_t n_head = head < n ? head : n, n_vec = (n - n_head) / 16u;
        for (uint32_t v = tid; v < n_vec; v += ZH_THREADS) {
            const ZhWord4 q = zh_load16(lit + n_head + 16u * v);
            for (int k = 0; k < 4; ++k) {
                ZH_ADD(&h[q.w[k] & 255u], 1u); ZH_ADD(&h[(q.w[k] >> 8) & 255u], 1u);
                ZH_ADD(&h[(q.w[k] >> 16) & 255u], 1u); ZH_ADD(&h[q.w[k] >> 24], 1u);
            }
        }
        for (uint32_t k = tid; k < n_head; k += ZH_THREADS) ZH_ADD(&h[lit[k]], 1u);
        for (uint32_t k = n_head + 16u * n_vec + tid; k < n; k += ZH_THREADS) ZH_ADD(&h[lit[k]], 1u);
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
        const uint64_t bound = 5u + sh.table.desc_len + 6u + sh.table.total_bits / 8u + 4u;
        if (bound < 3u + (uint64_t)n) sh.fail = 0;
    });
    uint32_t pos = 5u + sh.table.desc_len + 6u;
    uint32_t stream_bytes[4] = {0, 0, 0, 0};
    for (uint32_t s = 0; s < 4u && !sh.fail; ++s) {
        const uint8_t *ssrc = lit + zh_stream_begin(n, s);
        const uint32_t slen = zh_stream_len(n, s), per = (slen + ZH_THREADS - 1u) / ZH_THREADS;
        ex.each([&](uint32_t tid) {
            for (uint32_t k = tid; k < ZH_STREAM_CAP_WORDS; k += ZH_THREADS) sh.out[k] = 0;
            sh.span[tid] = zh_span_bits(ssrc, slen, per, tid, sh.table.sym);
        });
        zh_scan_spans(ex, sh);
        ex.each([&](uint32_t tid) { if (tid == 0 && sh.stream_bits / 8u + 1u > 4u * ZH_STREAM_CAP_WORDS - 8u) sh.fail = 1; });
        if (sh.fail) break;
        const uint32_t bits = sh.stream_bits, bytes = bits / 8u + 1u;
        ex.each([&](uint32_t tid) { zh_span_encode(ssrc, slen, per, tid, sh.table.sym, sh.span[tid], sh.out); });
        ex.each([&](uint32_t tid) { if (tid == 0) sh.out[bits >> 5] |= 1u << (bits & 31u); });     // end mark above the last code
        ex.each([&](uint32_t tid) {
            const uint8_t *ob = reinterpret_cast<const uint8_t *>(sh.out);
            for (uint32_t k = tid; k < bytes; k += ZH_THREADS) dst[pos + k] = ob[k];
        });
        stream_bytes[s] = bytes;
        pos += bytes;
    }
    if (sh.fail) {                                 // Raw_Literals_Block: 3-byte header (20-bit size), the bytes
        ex.each([&](uint32_t tid) {
            if (tid == 0) {
                const uint32_t h = 0u | (3u << 2) | (n << 4);
                dst[0] = (uint8_t)h; dst[1] = (uint8_t)(h >> 8); dst[2] = (uint8_t)(h >> 16);
            }
            for (uint32_t k = tid; k < n; k += ZH_THREADS) dst[3u + k] = lit[k];
        });
        return 3u + n;
    }
    ex.each([&](uint32_t tid) {
        if (tid) return;
        const uint32_t dl = sh.table.desc_len;
        zh_put_literals_header(dst, n, pos - 5u);                           // tree description + jump table + streams
        for (uint32_t k = 0; k < dl; ++k) dst[5u + k] = sh.table.desc[k];
        for (uint32_t k = 0; k < 3u; ++k) {
            dst[5u + dl + 2u * k] = (uint8_t)stream_bytes[k];
            dst[5u + dl + 2u * k + 1u] = (uint8_t)(stream_bytes[k] >> 8);
        }
    });
    return pos;
}

// Matches of every line against the line four lines earlier (a common prefix and a common suffix of at least ZH_MIN_MATCH
// bytes, inside the line and its '\n'), compacted into sc.seq_*; the bytes outside the matches go to sc.lit.
// Afterwards sh.n_seq / sh.n_lit hold the counts (n_seq = 0: lit is not filled - the caller codes the block from src).
template <class Exec>
ZH_HD void zh_find_matches(Exec &ex, ZhShared &sh, ZhScratch &sc, const uint8_t *src, uint32_t n) {
    // every thread owns a slice of whole 16-byte vectors (the block starts on a 16-byte boundary whenever its buffer does)
    const uint32_t slice = ((n + ZH_THREADS - 1u) / ZH_THREADS + 15u) & ~15u;
    const bool vec = ((((uintptr_t)src) & 15u) == 0);
    ex.each([&](uint32_t tid) {                                             // newlines of the thread's slice
        const uint32_t lo = tid * slice < n ? tid * slice : n, hi = lo + slice < n ? lo + slice : n;
        uint32_t c = 0, k = lo;
        if (vec) for (; k + 16u <= hi; k += 16u) c += zh_popc(zh_nl_mask16(zh_load16(src + k)));
        for (; k < hi; ++k) c += src[k] == (uint8_t)'\n';
        sh.span[tid] = c;
        if (tid == 0) { sh.n_seq = 0; sh.n_lit = 0; }
    });
    zh_scan_spans(ex, sh);
    const uint32_t n_lines = sh.stream_bits;
    if (n_lines < 5u || n_lines > ZH_MAX_LINES) return;
    ex.each([&](uint32_t tid) {
        const uint32_t lo = tid * slice < n ? tid * slice : n, hi = lo + slice < n ? lo + slice : n;
        uint32_t at = sh.span[tid], k = lo;
        if (vec) for (; k + 16u <= hi; k += 16u) {
            uint32_t m = zh_nl_mask16(zh_load16(src + k));
            while (m) { const uint32_t bit = zh_ctz(m); m &= m - 1u; sc.nl[at++] = k + bit; }
        }
        for (; k < hi; ++k) if (src[k] == (uint8_t)'\n') sc.nl[at++] = k;
    });
    const uint32_t per_line = (n_lines + ZH_THREADS - 1u) / ZH_THREADS;
    ex.each([&](uint32_t tid) {                                             // candidates of the thread's lines, and how many
        uint32_t c = 0;
        const uint32_t k0 = tid * per_line, k1 = k0 + per_line < n_lines ? k0 + per_line : n_lines;
        for (uint32_t k = k0; k < k1; ++k) {
            uint32_t L = 0, S = 0;
            if (k >= 4u) {
                const uint32_t a = sc.nl[k - 1u] + 1u, e = sc.nl[k];        // line k: [a, e], e is its '\n'
                const uint32_t b = k >= 5u ? sc.nl[k - 5u] + 1u : 0u, eb = sc.nl[k - 4u];
                const uint32_t len_a = e - a + 1u, len_b = eb - b + 1u;
                const uint32_t lim = (len_a < len_b ? len_a : len_b) < 65535u ? (len_a < len_b ? len_a : len_b) : 65535u;     // lengths are kept in 16 bits
                L = zh_common_run(src + a, src + b, lim, 1);
                S = zh_common_run(src + e, src + eb, lim - L, -1);
                if (L < ZH_MIN_MATCH) L = 0;
                if (S < ZH_MIN_MATCH) S = 0;
                sc.cand_pos[2u * k] = a; sc.cand_off[2u * k] = a - b;
                sc.cand_pos[2u * k + 1u] = e + 1u - S; sc.cand_off[2u * k + 1u] = e - eb;
            }
            sc.cand_len[2u * k] = (uint16_t)L;
            sc.cand_len[2u * k + 1u] = (uint16_t)S;
            c += (L != 0) + (S != 0);
        }
        sh.span[tid] = c;
    });
    zh_scan_spans(ex, sh);
    const uint32_t n_seq = sh.stream_bits;
    if (n_seq == 0) return;
    ex.each([&](uint32_t tid) {                                             // compaction in text order
        const uint32_t k0 = tid * per_line, k1 = k0 + per_line < n_lines ? k0 + per_line : n_lines;
        uint32_t at = sh.span[tid];
        for (uint32_t c = 2u * k0; c < 2u * k1; ++c)
            if (sc.cand_len[c]) { sc.seq_pos[at] = sc.cand_pos[c]; sc.seq_off[at] = sc.cand_off[c]; sc.seq_len[at] = sc.cand_len[c]; ++at; }
        if (tid == 0) { sc.seq_pos[n_seq] = n; sc.seq_off[n_seq] = 0; sc.seq_len[n_seq] = 0; }     // closes the last literal run
    });
    const uint32_t per_seq = (n_seq + 1u + ZH_THREADS - 1u) / ZH_THREADS;
    ex.each([&](uint32_t tid) {                                             // literal bytes in front of the thread's sequences
        const uint32_t j0 = tid * per_seq, j1 = j0 + per_seq < n_seq + 1u ? j0 + per_seq : n_seq + 1u;
        uint32_t c = 0;
        for (uint32_t j = j0; j < j1; ++j) c += sc.seq_pos[j] - (j ? sc.seq_pos[j - 1u] + sc.seq_len[j - 1u] : 0u);
        sh.span[tid] = c;
    });
    zh_scan_spans(ex, sh);
    ex.each([&](uint32_t tid) {                                             // where every literal run goes
        const uint32_t j0 = tid * per_seq, j1 = j0 + per_seq < n_seq + 1u ? j0 + per_seq : n_seq + 1u;
        uint32_t at = sh.span[tid];
        for (uint32_t j = j0; j < j1; ++j) {
            sc.seq_lit[j] = at;
            at += sc.seq_pos[j] - (j ? sc.seq_pos[j - 1u] + sc.seq_len[j - 1u] : 0u);
        }
        if (tid == 0) { sh.n_seq = n_seq; sh.n_lit = sh.stream_bits; }
    });
    // the bytes: batches of ZH_THREADS runs - every thread looks one run up, then each warp copies 32 of them, lanes side by side
    for (uint32_t base = 0; base < n_seq + 1u; base += ZH_THREADS) {
        ex.each([&](uint32_t tid) {
            const uint32_t j = base + tid;
            uint32_t from = 0, run = 0, at = 0;
            if (j < n_seq + 1u) {
                from = j ? sc.seq_pos[j - 1u] + sc.seq_len[j - 1u] : 0u;
                run = sc.seq_pos[j] - from;
                at = sc.seq_lit[j];
            }
            sh.whist[0][tid] = from; sh.whist[1][tid] = run; sh.whist[2][tid] = at;     // (the histograms are not in use yet)
        });
        ex.each([&](uint32_t tid) {
            const uint32_t lane = tid & 31u, first = tid & ~31u;
            for (uint32_t r = first; r < first + 32u; ++r) {
                const uint32_t from = sh.whist[0][r], run = sh.whist[1][r], at = sh.whist[2][r];
                for (uint32_t k = lane; k < run; k += 32u) sc.lit[at + k] = src[from + k];
            }
        });
    }
}

// literal length, match length - 3 and offset + 3 of sequence j
struct ZhSeqVal { uint32_t ll, mlb, ofb; };
ZH_HD ZhSeqVal zh_seq_val(const ZhScratch &sc, uint32_t j) {
    ZhSeqVal v;
    v.ll = sc.seq_pos[j] - (j ? sc.seq_pos[j - 1u] + sc.seq_len[j - 1u] : 0u);
    v.mlb = (uint32_t)sc.seq_len[j] - 3u;
    v.ofb = sc.seq_off[j] + 3u;
    return v;
}

// OR `nbits` low bits of `value` into the zeroed bit buffer at bit `at` (nbits <= 17)
ZH_HD void zh_or_bits(uint32_t *out, uint32_t at, uint32_t value, uint32_t nbits) {
    if (!nbits) return;
    const uint64_t v = (uint64_t)(value & ((1u << nbits) - 1u)) << (at & 31u);
    ZH_OR(out + (at >> 5), (uint32_t)v);
    if ((uint32_t)(v >> 32)) ZH_OR(out + (at >> 5) + 1u, (uint32_t)(v >> 32));
}

// Sequences_Section of sc.seq_[0, n_seq) -> dst (ZSTD_encodeSequences with the predefined tables); returns its size, 0
// when it does not fit the bit buffer.  The sequences are coded last to first; everything but the walk along the three
// FSE state chains (a table lookup per sequence and chain, one thread) is done by all threads: the codes, the number
// of bits every sequence contributes, and - after a scan - the bits themselves.
template <class Exec>
ZH_HD uint32_t zh_sequences_section(Exec &ex, ZhShared &sh, const ZhScratch &sc, uint32_t n_seq, uint8_t *dst, uint32_t dst_cap) {
    const uint32_t hdr = n_seq < 128u ? 2u : 3u;                             // Number_of_Sequences + Symbol_Compression_Modes
    const uint32_t per = (n_seq + ZH_THREADS - 1u) / ZH_THREADS;
    ex.each([&](uint32_t tid) {
        if (tid < 3u) zh_fse_predefined(tid, &sh.fse[tid]);
        for (uint32_t k = tid; k < ZH_STREAM_CAP_WORDS; k += ZH_THREADS) sh.out[k] = 0;
        for (uint32_t j = tid; j < n_seq; j += ZH_THREADS) {
            const ZhSeqVal v = zh_seq_val(sc, j);
            uint32_t eb;
            sh.seq_code[0][j] = (uint8_t)zh_ll_code(v.ll, &eb);
            sh.seq_code[1][j] = (uint8_t)zh_ml_code(v.mlb, &eb);
            sh.seq_code[2][j] = (uint8_t)zh_highbit(v.ofb);
        }
    });
    ex.each([&](uint32_t tid) {                                              // the three state chains, a thread each
        if (tid >= 3u) return;
        const ZhFse &t = sh.fse[tid];
        uint8_t *code = sh.seq_code[tid];
        const uint32_t size = 1u << t.log;
        uint32_t j = n_seq - 1u;
        uint32_t st = zh_fse_first(t, code[j]);
        // the table words of the next code are fetched while the current step's lookup is in flight
        uint32_t c = j ? code[j - 1u] : 0u, db = t.delta_bits[c];
        int32_t ds = t.delta_state[c];
        while (j-- > 0) {
            const uint32_t db_now = db;
            const int32_t ds_now = ds;
            if (j) { c = code[j - 1u]; db = t.delta_bits[c]; ds = t.delta_state[c]; }
            code[j] = (uint8_t)(st - size);
            st = t.state[(int32_t)(st >> ((st + db_now) >> 16)) + ds_now];
        }
        sh.end_state[tid] = st;
    });
    // coding order position q <-> sequence j = n_seq - 1 - q; thread t owns positions [t * per, (t + 1) * per)
    ex.each([&](uint32_t tid) {
        const uint32_t q0 = tid * per < n_seq ? tid * per : n_seq, q1 = q0 + per < n_seq ? q0 + per : n_seq;
        uint32_t bits = 0;
        for (uint32_t q = q0; q < q1; ++q) {
            const uint32_t j = n_seq - 1u - q;
            const ZhSeqVal v = zh_seq_val(sc, j);
            uint32_t ll_bits, ml_bits;
            const uint32_t c_ll = zh_ll_code(v.ll, &ll_bits), c_ml = zh_ml_code(v.mlb, &ml_bits), c_of = zh_highbit(v.ofb);
            bits += ll_bits + ml_bits + c_of;
            if (q) bits += zh_fse_nbits(sh.fse[2], sh.seq_code[2][j] + 32u, c_of) + zh_fse_nbits(sh.fse[1], sh.seq_code[1][j] + 64u, c_ml) +
                           zh_fse_nbits(sh.fse[0], sh.seq_code[0][j] + 64u, c_ll);
        }
        sh.span[tid] = bits;
    });
    zh_scan_spans(ex, sh);
    const uint32_t body_bits = sh.stream_bits, total_bits = body_bits + sh.fse[1].log + sh.fse[2].log + sh.fse[0].log;
    const uint32_t bytes = hdr + total_bits / 8u + 1u;
    if (bytes + 8u > 4u * ZH_STREAM_CAP_WORDS || bytes > dst_cap) return 0;
    ex.each([&](uint32_t tid) {
        const uint32_t q0 = tid * per < n_seq ? tid * per : n_seq, q1 = q0 + per < n_seq ? q0 + per : n_seq;
        uint32_t at = 8u * hdr + sh.span[tid];
        for (uint32_t q = q0; q < q1; ++q) {
            const uint32_t j = n_seq - 1u - q;
            const ZhSeqVal v = zh_seq_val(sc, j);
            uint32_t ll_bits, ml_bits;
            const uint32_t c_ll = zh_ll_code(v.ll, &ll_bits), c_ml = zh_ml_code(v.mlb, &ml_bits), c_of = zh_highbit(v.ofb);
            if (q) {
                const uint32_t s_of = sh.seq_code[2][j] + 32u, s_ml = sh.seq_code[1][j] + 64u, s_ll = sh.seq_code[0][j] + 64u;
                uint32_t nb = zh_fse_nbits(sh.fse[2], s_of, c_of);
                zh_or_bits(sh.out, at, s_of, nb); at += nb;
                nb = zh_fse_nbits(sh.fse[1], s_ml, c_ml);
                zh_or_bits(sh.out, at, s_ml, nb); at += nb;
                nb = zh_fse_nbits(sh.fse[0], s_ll, c_ll);
                zh_or_bits(sh.out, at, s_ll, nb); at += nb;
            }
            zh_or_bits(sh.out, at, v.ll, ll_bits); at += ll_bits;
            zh_or_bits(sh.out, at, v.mlb, ml_bits); at += ml_bits;
            zh_or_bits(sh.out, at, v.ofb, c_of); at += c_of;
        }
    });
    ex.each([&](uint32_t tid) {
        if (tid) return;
        uint32_t at = 8u * hdr + body_bits;                                  // the final states, then the end mark
        zh_or_bits(sh.out, at, sh.end_state[1], sh.fse[1].log); at += sh.fse[1].log;
        zh_or_bits(sh.out, at, sh.end_state[2], sh.fse[2].log); at += sh.fse[2].log;
        zh_or_bits(sh.out, at, sh.end_state[0], sh.fse[0].log); at += sh.fse[0].log;
        zh_or_bits(sh.out, at, 1u, 1u);
        if (n_seq < 128u) sh.out[0] |= n_seq;                                // byte 1 (the modes) stays 0: predefined tables
        else sh.out[0] |= ((n_seq >> 8) + 128u) | ((n_seq & 255u) << 8);     // n_seq <= ZH_MAX_SEQ < 0x7F00; byte 2 stays 0
    });
    ex.each([&](uint32_t tid) {
        const uint8_t *ob = reinterpret_cast<const uint8_t *>(sh.out);
        for (uint32_t k = tid; k < bytes; k += ZH_THREADS) dst[k] = ob[k];
    });
    return bytes;
}

// One block of n <= ZH_BLOCK_BYTES bytes -> block header + content in `slot` (ZH_SLOT_BYTES); returns the bytes used.
// `ex.each(f)` runs f(tid) for the ZH_THREADS threads of the CTA and separates the phases (device: __syncthreads()).
// `sc`: scratch of this block; NULL codes the block without matches.
template <class Exec>
ZH_HD uint32_t zh_compress_block(Exec &ex, ZhShared &sh, ZhScratch *sc, const uint8_t *src, uint32_t n, uint32_t last, uint8_t *slot) {
    uint32_t n_seq = 0, n_lit = n;
    const uint8_t *lit = src;
    if (sc && n >= ZH_MIN_COMPRESS) {
        zh_find_matches(ex, sh, *sc, src, n);
        n_seq = sh.n_seq;
        if (n_seq) { n_lit = sh.n_lit; lit = sc->lit; }
    }
    uint32_t total = 0;
    if (n >= ZH_MIN_COMPRESS) {
        const uint32_t lit_bytes = zh_literals_section(ex, sh, lit, n_lit, slot + 3u);
        uint32_t seq_bytes = 1;
        if (n_seq) seq_bytes = zh_sequences_section(ex, sh, *sc, n_seq, slot + 3u + lit_bytes, ZH_SLOT_BYTES - 3u - lit_bytes);
        if (seq_bytes && 3u + lit_bytes + seq_bytes < 3u + n) {
            ex.each([&](uint32_t tid) {
                if (tid) return;
                if (!n_seq) slot[3u + lit_bytes] = 0;                        // Number_of_Sequences = 0
                zh_put_block_header(slot, last, 2u, lit_bytes + seq_bytes);
            });
            total = 3u + lit_bytes + seq_bytes;
        }
    }
    if (!total) {                                  // Raw_Block
        ex.each([&](uint32_t tid) {
            if (tid == 0) zh_put_block_header(slot, last, 0u, n);
            for (uint32_t k = tid; k < n; k += ZH_THREADS) slot[3u + k] = src[k];
        });
        total = 3u + n;
    }
    return total;
}

#endif  // PCL_ZHUFF_CORE_H
