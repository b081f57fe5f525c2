// hd_core.h -- per-record logic of the barcode hot path as __host__ __device__ functions.
//
// The CUDA kernels in kernels.cu are thin wrappers around these functions.  They are also
// compiled for the host by tests/emul (a CPU-side unit-test harness that checks this logic
// against the oracle without a GPU).  The product library only calls them from device code.
//
// Reference semantics implemented here (paths relative to the precellar repository):
//   pcl_split            SegmentInfo::split_with_tolerance(read, 1.0, 0.2) + consume_buf +
//                        find_best_pattern_match      seqspec/src/read/segment.rs:180-368,496-535
//   pcl_pack_key         the byte string handed to WhitelistBuilder::add / correct_barcode, i.e.
//                        the segment or precellar::utils::rev_compl(segment)
//                        precellar/src/barcode.rs:104-115, precellar/src/utils.rs:138-149
//   pcl_probe            OligoFrequncy: HashMap<Vec<u8>, usize> lookup   precellar/src/barcode.rs:189
//   pcl_correct_one      correct_barcode + likelihood1 + update_best_option
//                        precellar/src/barcode.rs:160-184,311-358
#ifndef PCL_HD_CORE_H
#define PCL_HD_CORE_H

#include <stdint.h>

#include "../../include/precellar_b200.h"

#ifdef __CUDACC__
#define PCL_HD __host__ __device__ __forceinline__
#define PCL_HD_COLD __host__ __device__ __noinline__      // rare paths: kept out of the hot loops' instruction stream
#else
#define PCL_HD inline
#define PCL_HD_COLD inline
#endif

// ------------------------------------------------------------------------------------------
// device-side layout tables
// ------------------------------------------------------------------------------------------
struct DElem {
    uint16_t min_len, max_len;
    uint8_t kind;        // pcl_kind
    uint8_t is_anchor;   // sequence_type == Fixed
    uint8_t seq_len;     // strlen(sequence)
    uint8_t max_mis;     // largest k with (double)k <= 0.2 * (double)seq_len   (segment.rs:267)
    int8_t bc_j;         // index among the read's barcode elements, else -1
    int8_t umi_j;        // index among the read's UMI elements, else -1
    uint8_t pad[2];
    uint8_t seq[PCL_MAX_ANCHOR];
    uint64_t a_code;     // anchors of <= 32 upper-case ACGT bases: 2-bit codes, base i at bits 2i ...
    uint64_t a_rc;       // ... and the same for the reverse complement of the anchor
};

// Register-resident path for the common short barcode reads (10x v2/v3 R1, ATAC / multiome I2 ...):
// every element fixed-length (an optional trailing variable target), the whole record within 16 or 32 bytes,
// exactly one barcode of <= 16 bases, at most one UMI of <= 15 bases.
// Field geometries with compile-time instances of the extraction (a 16-base barcode and an optional UMI on word
// boundaries): the byte masks, the 64-bit code string and its variable shifts fold away.
enum { PCL_FAST_SPEC_NONE = 0, PCL_FAST_SPEC_B0_U16x12 = 1 /* 10x v3 R1 */, PCL_FAST_SPEC_B0 = 2 /* 16-nt index read */,
       PCL_FAST_SPEC_B8 = 3 /* 8-nt linker + 16-nt barcode */ };

struct FastLayout {
    uint8_t enabled;
    uint8_t bc_start, bc_len, umi_start, umi_len;   // umi_len == 0: no UMI
    uint8_t has_tail_target;                        // trailing variable target element
    uint8_t spec;                                   // word-aligned field geometry the fast kernel is specialised for (PCL_FAST_SPEC_*)
    uint16_t t_start, t_min, t_max;
    uint32_t bc_mask[8], umi_mask[8];               // byte masks of the two regions over the 8 record words
};

// Layouts of the shape  [fixed ...] [ONE variable element] [anchor] [fixed ...] [optional trailing variable element]
// with <= 16-base barcodes / anchors and <= 15-base UMIs (sci-RNA-seq3 R1, dscATAC R1 ...): everything but the anchor
// position is known when the layout is compiled, so split() collapses to a handful of comparisons (pcl_anchor_split)
// and every field is cut out of the record's code string at a warp-uniform offset (k_count_anchor).
struct AnchorPlan {
    uint8_t enabled;
    uint8_t v;                       // index of the variable element; the anchor is element v + 1
    uint8_t has_t;                   // the last element is variable too
    uint8_t t_elem;                  // the (only) target element, 0xFF if none
    uint8_t ustart[PCL_MAX_ELEMS];   // start of element e when the variable element has its minimal length
    // straight-line path for records long enough to hold every element at every anchor position (pcl_plan_record_fast)
    uint8_t n_pos;                   // anchor positions inside its window: vmax - vmin + 1
    uint8_t a_start;                 // first window position: ustart[v] + vmin
    uint8_t a_len, a_maxmis;         // anchor length and tolerated mismatches
    uint16_t full_len;               // such records have L >= full_len
    uint8_t n_words;                 // record words the fields can touch: ceil((full_len or stride) / 4), <= 16
    uint8_t bc_elem[PCL_MAX_BC_PER_READ], umi_elem[PCL_MAX_UMI_PER_READ];
};

struct DRead {
    uint16_t n_elems;
    uint8_t is_reverse;
    uint8_t fixed_layout;
    uint32_t max_len;    // reader truncation, 0 = none
    uint32_t stride;     // bytes per record slot
    uint8_t n_bc, n_umi, has_target, needs_payload;
    uint8_t qual_only, codes_ok;            // qual_only: QC-mode insert read; codes_ok: pcl_split_codes / pcl_pack_codes apply
    uint16_t payload_bytes;                 // record bytes the path can touch, rounded up to 4 (<= stride): the compact upload row
    uint32_t qoff, qstride;       // the quality plane holds record bytes [qoff, qoff + qstride) per record
    uint32_t check_linkers;       // Assay::verify mode: split_with_tolerance(read, 0.2, 0.2) -- fixed sequences that are not
                                  // searched for are compared with the read as well (segment.rs:325-333)
    DElem elems[PCL_MAX_ELEMS];
    FastLayout fast;
    AnchorPlan plan;
};

// Whitelist hash table of one barcode region: buckets of 32 bytes (one DRAM/L2 sector) holding
// 8 x 32-bit or 4 x 64-bit keys; the slot position is the identity of the entry (counts[slot]).
enum { PCL_TAB_K32_MARKER = 0, PCL_TAB_K32_L16 = 1, PCL_TAB_K64 = 2 };
struct DTable {
    const void *keys;
    uint32_t *counts;                   // [n_buckets * slots_per_bucket]
    unsigned long long *total;          // total_count
    unsigned long long *mismatch;       // mismatch_count
    uint64_t empty;                     // key value marking an empty slot (never a whitelist key)
    uint32_t n_buckets;
    uint32_t mode;                      // PCL_TAB_*
    // 32-bit tables only: the same keys in four more tables, table q bucketed by hash(key & ~(0xFF << 8q)).
    // All keys that differ from a query only inside byte q of the packed key (4 bases) share the query's
    // bucket there, so the 1-Hamming neighbourhood of a barcode is found with 4 bucket reads instead of 3*L.
    const uint32_t *nkeys[4];
    // ... and one bitmap per neighbourhood table over the 24-bit index of (key with byte q removed): bit set iff some
    // whitelist key has that partial key.  4 x 2 MB, L2-resident: a query whose partial key is absent (two of three
    // "cold" partials of a typical miss) never touches the neighbourhood table in DRAM.
    const uint32_t *nbits[4];
};

#define PCL_NBITS_WORDS (1u << 19)      // 2^24 bits per bitmap

// 24-bit index of a 32-bit key with byte q taken out
PCL_HD uint32_t pcl_partial_index(uint32_t key32, uint32_t q) {
    const uint32_t lo = key32 & ((1u << (8u * q)) - 1u);
    const uint32_t hi = q >= 3u ? 0u : ((key32 >> (8u * q + 8u)) << (8u * q));
    return hi | lo;
}

// ------------------------------------------------------------------------------------------
// base coding
// ------------------------------------------------------------------------------------------
// Forward reads: only upper-case ACGT are bases; any other byte is "invalid" (it can never equal a
// whitelist byte, and likelihood1 tries all four substitutions there).  Reverse reads go through
// precellar::utils::rev_compl which also maps lower-case acgt to their upper-case complement
// (precellar/src/utils.rs:138-149), so there lower-case letters are bases too.
// Returns 0..3 (A C G T) or 4 (invalid); `rev` also complements.
PCL_HD uint32_t pcl_base_code(uint8_t b, bool rev) {
    if (rev) b &= 0xDF;                                  // fold case: only 'a','c','g','t' reach ACGT
    uint32_t c = ((b >> 1) ^ (b >> 2)) & 3u;             // A->0 C->1 G->2 T->3
    uint32_t expect = (0x54474341u >> (8u * c)) & 0xFFu; // "ACGT"[c]
    if (b != expect) return 4u;
    return rev ? 3u - c : c;
}

// Pack the barcode/UMI that the reference would see for segment [start, start+len) of `rec`.
// key = (1 << 2L) | bases (first base most significant); positions with an invalid byte contribute 0.
// *n_invalid = number of invalid positions, *inv_pos = position (in reported orientation) of the last one.
// len > PCL_MAX_KEY_LEN is not packable: returns 0 and n_invalid = len.
PCL_HD uint64_t pcl_pack_key(const uint8_t *rec, uint32_t start, uint32_t len, bool rev, uint32_t *n_invalid,
                             uint32_t *inv_pos) {
    if (len > PCL_MAX_KEY_LEN) { *n_invalid = len; *inv_pos = 0; return 0; }
    uint64_t key = 1;
    uint32_t ninv = 0, ip = 0;
    for (uint32_t p = 0; p < len; ++p) {
        uint8_t b = rev ? rec[start + len - 1 - p] : rec[start + p];
        uint32_t c = pcl_base_code(b, rev);
        if (c > 3u) { ++ninv; ip = p; c = 0; }
        key = (key << 2) | c;
    }
    *n_invalid = ninv;
    *inv_pos = ip;
    return key;
}

// ------------------------------------------------------------------------------------------
// miss worklist entries: rec | start << 32 | len << 48 | j << 56 | one_invalid << 58 | inv_pos << 59
// (inv_pos: position, in reported orientation, of the single non-ACGT base; barcodes with two or more such bases
// never enter the list -- no single substitution reaches the whitelist, they stay PCL_BC_NO_MATCH)
// ------------------------------------------------------------------------------------------
PCL_HD unsigned long long pcl_wl_pack(uint64_t rec, uint32_t start, uint32_t len, uint32_t j, uint32_t n_invalid, uint32_t inv_pos) {
    return (unsigned long long)rec | ((unsigned long long)(start & 0xFFFFu) << 32) | ((unsigned long long)(len & 0xFFu) << 48) |
           ((unsigned long long)(j & 3u) << 56) | ((unsigned long long)(n_invalid == 1u ? 1u : 0u) << 58) |
           ((unsigned long long)(n_invalid == 1u ? (inv_pos & 31u) : 0u) << 59);
}
#define PCL_WL_REC(e) ((uint64_t)((e) & 0xFFFFFFFFull))
#define PCL_WL_START(e) ((uint32_t)((e) >> 32) & 0xFFFFu)
#define PCL_WL_LEN(e) ((uint32_t)((e) >> 48) & 0xFFu)
#define PCL_WL_J(e) ((uint32_t)((e) >> 56) & 3u)
#define PCL_WL_NINV(e) ((uint32_t)((e) >> 58) & 1u)
#define PCL_WL_IPOS(e) ((uint32_t)((e) >> 59) & 31u)
// k_count_fast / k_count_spec do not locate the non-ACGT bases of a barcode (that would tax every warp for the sake of
// a few lanes): such a barcode enters the list with this marker and k_filter re-derives key, count and position of its
// invalid bases from the record bytes (positions are <= 15 on that path, so 31 is free).
#define PCL_WL_REPACK_POS 31u
#define PCL_WL_IS_REPACK(e) (PCL_WL_NINV(e) == 1u && PCL_WL_IPOS(e) == PCL_WL_REPACK_POS)

// ------------------------------------------------------------------------------------------
// whitelist table probe
// ------------------------------------------------------------------------------------------
PCL_HD uint32_t pcl_hash_bucket(uint64_t key, uint32_t n_buckets) {
    uint64_t h = key * 0x9E3779B97F4A7C15ull;
    h ^= h >> 32;
    h *= 0xD6E8FEB86659FD93ull;
    h ^= h >> 32;
    return (uint32_t)(((h & 0xFFFFFFFFull) * (uint64_t)n_buckets) >> 32);
}

// 32-bit tables: lowbias32 mix + fastrange
PCL_HD uint32_t pcl_hash32_bucket(uint32_t k, uint32_t n_buckets) {
    k ^= k >> 16;
    k *= 0x7feb352du;
    k ^= k >> 15;
    k *= 0x846ca68bu;
    k ^= k >> 16;
#ifdef __CUDA_ARCH__
    return __umulhi(k, n_buckets);
#else
    return (uint32_t)(((uint64_t)k * (uint64_t)n_buckets) >> 32);
#endif
}

#ifdef __CUDA_ARCH__
#define PCL_LDG128(p) __ldg(reinterpret_cast<const uint4 *>(p))
#else
struct pcl_u4 { uint32_t x, y, z, w; };
#define PCL_LDG128(p) (*reinterpret_cast<const pcl_u4 *>(p))
#endif

// One 32-byte bucket of a 32-bit table: slot of `k` or 0xFFFFFFFF; *full = no free slot in the bucket.
// The host fills the slots of a bucket in order, so the bucket is full iff its last slot is occupied.
PCL_HD uint32_t pcl_bucket32(const uint32_t *keys, uint32_t b, uint32_t k, uint32_t empty, bool *full) {
    auto v0 = PCL_LDG128(keys + (uint64_t)b * 8);
    auto v1 = PCL_LDG128(keys + (uint64_t)b * 8 + 4);
    uint32_t j = 8;
    j = (v1.w == k) ? 7u : j;
    j = (v1.z == k) ? 6u : j;
    j = (v1.y == k) ? 5u : j;
    j = (v1.x == k) ? 4u : j;
    j = (v0.w == k) ? 3u : j;
    j = (v0.z == k) ? 2u : j;
    j = (v0.y == k) ? 1u : j;
    j = (v0.x == k) ? 0u : j;
    *full = v1.w != empty;
    return j < 8u ? b * 8u + j : 0xFFFFFFFFu;
}

// Probe of a 32-bit table (PCL_TAB_K32_MARKER / PCL_TAB_K32_L16); k != empty is the caller's business.
PCL_HD uint32_t pcl_probe32(const DTable &t, uint32_t k) {
    const uint32_t *keys = reinterpret_cast<const uint32_t *>(t.keys);
    const uint32_t e = (uint32_t)t.empty;
    if (k == e) return 0xFFFFFFFFu;
    uint32_t b = pcl_hash32_bucket(k, t.n_buckets);
    for (;;) {
        bool full;
        uint32_t slot = pcl_bucket32(keys, b, k, e, &full);
        if (slot != 0xFFFFFFFFu || !full) return slot;
        b = (b + 1u == t.n_buckets) ? 0u : b + 1u;
    }
}

// Look `key` (marker form, `len` bases) up in the table.  Returns the slot or 0xFFFFFFFF.
PCL_HD uint32_t pcl_probe(const DTable &t, uint64_t key, uint32_t len) {
    const uint32_t MISS = 0xFFFFFFFFu;
    if (t.n_buckets == 0) return MISS;
    if (t.mode == PCL_TAB_K64) {
        if (key == t.empty) return MISS;
        uint32_t b = pcl_hash_bucket(key, t.n_buckets);
        const uint64_t *keys = reinterpret_cast<const uint64_t *>(t.keys);
        for (;;) {
            auto v0 = PCL_LDG128(keys + (uint64_t)b * 4);
            auto v1 = PCL_LDG128(keys + (uint64_t)b * 4 + 2);
            uint64_t k0 = ((uint64_t)v0.y << 32) | v0.x, k1 = ((uint64_t)v0.w << 32) | v0.z;
            uint64_t k2 = ((uint64_t)v1.y << 32) | v1.x, k3 = ((uint64_t)v1.w << 32) | v1.z;
            if (k0 == key) return b * 4u;
            if (k1 == key) return b * 4u + 1u;
            if (k2 == key) return b * 4u + 2u;
            if (k3 == key) return b * 4u + 3u;
            if (k3 == t.empty) return MISS;                 // slots fill in order: a free last slot ends the chain
            b = (b + 1u == t.n_buckets) ? 0u : b + 1u;
        }
    } else {
        if (t.mode == PCL_TAB_K32_L16) { if (len != 16u) return MISS; }
        else if (len > 15u) return MISS;
        return pcl_probe32(t, (uint32_t)key);
    }
}

// ------------------------------------------------------------------------------------------
// split
// ------------------------------------------------------------------------------------------
struct SplitOut {
    uint32_t seg[PCL_MAX_ELEMS];   // (start << 16) | len per element, PCL_COORD_ABSENT if the element has no Single segment
    uint32_t tcoord;               // the target segment (Single target or a Joined run containing one)
    uint32_t status;               // PCL_SPLIT_*
};

// haystack byte k of the anchor search: the window itself, or seqspec::utils::rev_compl(window)
// (upper-case only, seqspec/src/utils.rs:176-187) for neg-strand reads.
PCL_HD uint8_t pcl_window_byte(const uint8_t *w, uint32_t n, uint32_t k, bool rev) {
    if (!rev) return w[k];
    uint8_t x = w[n - 1u - k];
    switch (x) {
        case 'A': return 'T';
        case 'T': return 'A';
        case 'C': return 'G';
        case 'G': return 'C';
        default: return x;
    }
}

// find_best_pattern_match (segment.rs:496-535).  The KMP pre-pass returns the first exact occurrence,
// which is also what the leftmost-strict-minimum scan returns when the minimum is 0, so one scan suffices.
// Returns false for (None, _).
PCL_HD bool pcl_best_match(const uint8_t *w, uint32_t n, const uint8_t *needle, uint32_t m, bool rev, uint32_t *pos,
                           uint32_t *mis) {
    if (m == 0u || n == 0u || m > n) return false;
    uint32_t best = 0xFFFFFFFFu, bpos = 0;
    for (uint32_t i = 0; i + m <= n; ++i) {
        uint32_t d = 0;
        for (uint32_t j = 0; j < m; ++j) {
            d += pcl_window_byte(w, n, i + j, rev) != needle[j];
            if (d >= best) break;                          // cannot become a strict improvement
        }
        if (d < best) { best = d; bpos = i; if (d == 0u) break; }
    }
    if (m == 1u && best != 0u) return false;               // (None, 1)  (segment.rs:508-512)
    *pos = bpos;
    *mis = best;
    return true;
}

// consume_buf (segment.rs:186-232) over buffered elements [b0, b1] and record range [lo, hi).
PCL_HD void pcl_consume(const DRead &rd, SplitOut &o, int b0, int b1, uint32_t lo, uint32_t hi, uint32_t *n_target) {
    int e = b1;
    while (e >= b0 && rd.elems[e].min_len == rd.elems[e].max_len) {
        uint32_t l = hi - lo;
        uint32_t ml = rd.elems[e].max_len;
        uint32_t i = l - ml;                               // never underflows for derived layouts (range >= sum of min_len)
        uint32_t c = ((lo + i) << 16) | ml;
        o.seg[e] = c;
        if (rd.elems[e].kind == PCL_KIND_TARGET) { o.tcoord = c; ++*n_target; }
        hi = lo + i;
        --e;
    }
    if (e >= b0) {
        uint32_t c = (lo << 16) | (hi - lo);
        if (e == b0) {
            o.seg[b0] = c;
            if (rd.elems[b0].kind == PCL_KIND_TARGET) { o.tcoord = c; ++*n_target; }
        } else {
            // Joined(region types): never a barcode / UMI; a target if any member is (segment.rs:71-90)
            bool has_t = false;
            for (int k = b0; k <= e; ++k) has_t |= rd.elems[k].kind == PCL_KIND_TARGET;
            if (has_t) { o.tcoord = c; ++*n_target; }
        }
    }
}

// Matcher: bool operator()(const DElem &el, uint32_t wstart, uint32_t wn, uint32_t *pos, uint32_t *mis)
template <class Matcher>
PCL_HD void pcl_split_t(const DRead &rd, uint32_t len, SplitOut &o, const Matcher &match) {
    for (int e = 0; e < PCL_MAX_ELEMS; ++e) o.seg[e] = PCL_COORD_ABSENT;
    o.tcoord = PCL_COORD_ABSENT;
    o.status = PCL_SPLIT_OK;
    uint32_t min_off = 0, max_off = 0, buf_sum_min = 0, n_target = 0;
    int buf_start = -1, buf_end = -1;
    const int n_elems = rd.n_elems;
    for (int e = 0; e < n_elems; ++e) {
        const DElem &el = rd.elems[e];
        const uint32_t mn = el.min_len, mx = el.max_len;
        if (max_off >= len || min_off + mn > len) break;                         // segment.rs:245-248
        if (mn == mx) {
            if (buf_start >= 0) {
                if (el.is_anchor) {                                              // segment.rs:254-311
                    uint32_t wend = max_off + mx < len ? max_off + mx : len;
                    uint32_t pos = 0, mis = 0;
                    if (!match(el, min_off, wend - min_off, &pos, &mis)) {
                        o.status = PCL_SPLIT_ANCHOR_NOT_FOUND;
                        return;
                    }
                    if (mis <= el.max_mis) {
                        uint32_t off_left = min_off - buf_sum_min, off_right = min_off + pos;
                        pcl_consume(rd, o, buf_start, buf_end, off_left, off_right, &n_target);
                        buf_start = -1; buf_sum_min = 0;
                        min_off = off_right; max_off = off_right;
                        uint32_t c = (off_right << 16) | mx;
                        o.seg[e] = c;
                        if (el.kind == PCL_KIND_TARGET) { o.tcoord = c; ++n_target; }
                    } else if (max_off + mx > len) {
                        break;                                                   // segment.rs:303-305
                    } else {
                        o.status = PCL_SPLIT_PATTERN_MISMATCH;                   // segment.rs:306-308
                        return;
                    }
                } else {
                    buf_end = e; buf_sum_min += mn;                              // segment.rs:312-314
                }
            } else {
                if (rd.check_linkers && el.is_anchor && el.seq_len == mx) {      // linker_tolerance < 1.0 (segment.rs:325-333)
                    const uint32_t d = match.fixed_distance(el, max_off);
                    if ((double)d > 0.2 * (double)mx) { o.status = PCL_SPLIT_PATTERN_MISMATCH; return; }
                }
                uint32_t c = (max_off << 16) | mx;                               // segment.rs:315-344
                o.seg[e] = c;
                if (el.kind == PCL_KIND_TARGET) { o.tcoord = c; ++n_target; }
            }
        } else {
            if (buf_start < 0) buf_start = e;                                    // segment.rs:345-347
            buf_end = e; buf_sum_min += mn;
        }
        min_off += mn;                                                           // segment.rs:349-350
        max_off += mx;
    }
    if (buf_start >= 0) {                                                        // segment.rs:353-365
        uint32_t hi = max_off < len ? max_off : len;
        pcl_consume(rd, o, buf_start, buf_end, min_off - buf_sum_min, hi, &n_target);
    }
    if (n_target > 1u) o.status = PCL_SPLIT_MULTI_TARGET;                        // fastq.rs:505-506 panics
}

struct ByteMatcher {            // anchor search on the record bytes (any anchor, any record length)
    const uint8_t *rec;
    bool rev;
    PCL_HD bool operator()(const DElem &el, uint32_t wstart, uint32_t wn, uint32_t *pos, uint32_t *mis) const {
        return pcl_best_match(rec + wstart, wn, el.seq, el.seq_len, rev, pos, mis);
    }
    // hamming(segment or its seqspec::utils::rev_compl, sequence) of a fixed element found at `start`
    PCL_HD uint32_t fixed_distance(const DElem &el, uint32_t start) const {
        uint32_t d = 0;
        for (uint32_t k = 0; k < el.seq_len; ++k) d += pcl_window_byte(rec + start, el.seq_len, k, rev) != el.seq[k];
        return d;
    }
};

PCL_HD void pcl_split(const DRead &rd, const uint8_t *rec, uint32_t len, SplitOut &o) {
    ByteMatcher m{rec, rd.is_reverse != 0};
    pcl_split_t(rd, len, o, m);
}

// ------------------------------------------------------------------------------------------
// correction
// ------------------------------------------------------------------------------------------
#ifdef __CUDA_ARCH__
#define PCL_DMUL(a, b) __dmul_rn((a), (b))
#define PCL_DADD(a, b) __dadd_rn((a), (b))
#define PCL_DDIV(a, b) __ddiv_rn((a), (b))
#else
// host build of the test harness uses -ffp-contract=off
#define PCL_DMUL(a, b) ((a) * (b))
#define PCL_DADD(a, b) ((a) + (b))
#define PCL_DDIV(a, b) ((a) / (b))
#endif

// correct_barcode (barcode.rs:160-184) for one barcode that is NOT an exact whitelist member unless
// `is_exact`.  `qual` points at the segment's quality bytes in record orientation; position p of the
// reported barcode uses qual[start + (rev ? len-1-p : p)].  lut_cap[q] = error_probability(min(q, 66)),
// lut_raw[q] = error_probability(q).  Returns a PCL_BC_* code; *out_key = chosen whitelist key.
PCL_HD uint32_t pcl_correct_one(const DTable &t, uint64_t key, uint32_t len, uint32_t n_invalid, uint32_t inv_pos,
                                const uint8_t *qual, uint32_t start, bool rev, bool is_exact, double threshold,
                                double max_expected_errors, bool check_ee, const double *lut_cap,
                                const double *lut_raw, uint64_t *out_key) {
    if (check_ee) {                                                   // barcode.rs:167-170, uncapped q
        double expected = 0.0;
        for (uint32_t p = 0; p < len; ++p) {
            uint8_t q = qual[start + (rev ? len - 1u - p : p)];
            expected = PCL_DADD(expected, lut_raw[q]);
        }
        if (expected >= max_expected_errors) return PCL_BC_EXCEED_EE;
    }
    if (is_exact) { *out_key = key; return PCL_BC_EXACT; }            // (query, 1.0); 1.0 >= thr decided by the caller
    if (len > PCL_MAX_KEY_LEN || n_invalid > 1u) return PCL_BC_NO_MATCH;   // no single substitution reaches the list
    bool have = false;
    double best_like = 0.0, total = 0.0;
    uint64_t best_key = 0;
    for (uint32_t pos = 0; pos < len; ++pos) {                        // barcode.rs:319-334
        if (n_invalid == 1u && pos != inv_pos) continue;              // another position holds a non-ACGT byte
        uint8_t q = qual[start + (rev ? len - 1u - pos : pos)];
        uint32_t sh = 2u * (len - 1u - pos);
        uint32_t existing = (n_invalid == 1u) ? 4u : (uint32_t)((key >> sh) & 3u);
        uint64_t base = key & ~(3ull << sh);
        for (uint32_t val = 0; val < 4u; ++val) {                     // BASE_OPTS order A C G T
            if (val == existing) continue;
            uint64_t cand = base | ((uint64_t)val << sh);
            uint32_t slot = pcl_probe(t, cand, len);
            if (slot != 0xFFFFFFFFu) {
                uint32_t c = t.counts[slot];
                double like = PCL_DMUL((double)(1ull + (uint64_t)c), lut_cap[q]);   // barcode.rs:326-327
                if (!have || best_like < like) { have = true; best_like = like; best_key = cand; }   // :345-358
                total = PCL_DADD(total, like);
            }
        }
    }
    if (!have) return PCL_BC_NO_MATCH;                                // (query, 0.0) -> prob <= 0
    double prob = PCL_DDIV(best_like, total);
    if (prob <= 0.0) return PCL_BC_NO_MATCH;
    if (prob >= threshold) { *out_key = best_key; return PCL_BC_CORRECTED; }
    return PCL_BC_LOW_CONF;
}

// ------------------------------------------------------------------------------------------
// register-resident fixed-layout path
// ------------------------------------------------------------------------------------------
PCL_HD uint32_t pcl_brev32(uint32_t x) {
#ifdef __CUDA_ARCH__
    return __brev(x);
#else
    x = ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);
    x = ((x >> 2) & 0x33333333u) | ((x & 0x33333333u) << 2);
    x = ((x >> 4) & 0x0F0F0F0Fu) | ((x & 0x0F0F0F0Fu) << 4);
    x = ((x >> 8) & 0x00FF00FFu) | ((x & 0x00FF00FFu) << 8);
    return (x >> 16) | (x << 16);
#endif
}

// Four ASCII bytes -> four 2-bit codes packed into the low 8 bits (byte k -> bits 2k..2k+1), plus a word
// whose byte k is non-zero iff byte k of w is not one of A C G T.
//   codes : c = ((w>>1) ^ (w>>2)) & 3 per byte gives A0 C1 G2 T3; one multiply gathers the four fields
//           (c * 0x01041040 puts c0|c1<<2|c2<<4|c3<<6 into the top byte without carries).
//   valid : A 0100_0001  C 0100_0011  G 0100_0111  T 0101_0100.  Bits 7,6,5,3 must read 0,1,0,0; of the
//           free bits (4,2,1,0) the four bases are exactly those with  b0 == !b4  and  (b2 & !b1) == b4.
PCL_HD uint32_t pcl_word_codes(uint32_t w, uint32_t *bad) {
#ifdef __CUDA_ARCH__
    // the kernel is bound by the ALU pipe (shifts / logic); constant right shifts go to the FMA pipe as
    // multiply-high, and the gather of the four 2-bit codes is one dot product
    const uint32_t w1 = __umulhi(w, 0x80000000u), w2 = __umulhi(w, 0x40000000u), w4 = __umulhi(w, 0x10000000u);
#else
    const uint32_t w1 = w >> 1, w2 = w >> 2, w4 = w >> 4;
#endif
    const uint32_t c = (w1 ^ w2) & 0x03030303u;
    const uint32_t t = (w2 & ~w1) ^ w4;                                   // bit 0 of each byte must be 0
    const uint32_t u = (~(w4 ^ w)) | t;                                   // ... and b0 != b4
    *bad = ((w & 0xE8E8E8E8u) ^ 0x40404040u) | (u & 0x01010101u);
#ifdef __CUDA_ARCH__
    return (uint32_t)__dp4a(c, 0x40100401u, 0u);                          // c0 + 4 c1 + 16 c2 + 64 c3
#else
    return (c * 0x01041040u) >> 24;
#endif
}

// 2-bit little-endian-by-position field (base i at bits 2i) -> reported key:
//   forward: first base most significant; reverse: reverse complement, which is the complement of the
//   little-endian field read as big-endian.
PCL_HD uint32_t pcl_field_to_key(uint64_t codes, uint32_t start, uint32_t len, bool rev) {
    const uint32_t mask = len >= 16u ? 0xFFFFFFFFu : ((1u << (2u * len)) - 1u);
    const uint32_t f = (uint32_t)(codes >> (2u * start)) & mask;
    if (rev) return (~f) & mask;
    uint32_t r = pcl_brev32(f);
    r = ((r >> 1) & 0x55555555u) | ((r & 0x55555555u) << 1);
    return len >= 16u ? r : r >> (32u - 2u * len);
}

// Four bytes -> flags at bits 0, 2, 4, 6: byte k non-zero.  The high bit of ((x & 0x7F) + 0x7F) | x tells "non-zero";
// one multiply-high by 2^25 + 2^19 + 2^13 + 2^7 moves bits 7, 15, 23, 31 to bits 32, 34, 36, 38 of the 64-bit product
// (no two partial products share a bit, so there are no carries).
PCL_HD uint32_t pcl_flags8(uint32_t x) {
    const uint32_t hi = (((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x) & 0x80808080u;
#ifdef __CUDA_ARCH__
    return __umulhi(hi, 0x02082080u) & 0x55u;
#else
    return (uint32_t)(((uint64_t)hi * 0x02082080ull) >> 32) & 0x55u;
#endif
}

PCL_HD uint32_t pcl_popc32(uint32_t x) {
#ifdef __CUDA_ARCH__
    return (uint32_t)__popc(x);
#else
    return (uint32_t)__builtin_popcount(x);
#endif
}

PCL_HD uint32_t pcl_clz32(uint32_t x) {
#ifdef __CUDA_ARCH__
    return (uint32_t)__clz((int)x);
#else
    return x ? (uint32_t)__builtin_clz(x) : 32u;
#endif
}

// Flag word of a field (bit 2i set: base i of the field, in RECORD order, is not a base) -> what pcl_pack_key reports:
// the number of invalid positions and the position of the last one in REPORTED orientation.
PCL_HD void pcl_flags_to_inv(uint32_t flags, uint32_t len, bool rev, uint32_t *n_invalid, uint32_t *inv_pos) {
    *n_invalid = pcl_popc32(flags);
    if (!flags) { *inv_pos = 0; return; }
    // reported position p of record-order base i: forward p = i (last = highest set bit), reverse p = len-1-i (last = lowest)
    const uint32_t hi = (31u - pcl_clz32(flags)) >> 1, lo = (31u - pcl_clz32(flags & (0u - flags))) >> 1;
    *inv_pos = rev ? len - 1u - lo : hi;
}

// 2-bit field (base i at bits 2i, record order, already cut to `len` bases) -> reported key without marker; bases flagged
// in `flags` contribute 0, like pcl_pack_key.
PCL_HD uint32_t pcl_field32_to_key(uint32_t f, uint32_t len, bool rev, uint32_t flags) {
    const uint32_t mask = len >= 16u ? 0xFFFFFFFFu : ((1u << (2u * len)) - 1u);
    const uint32_t keep = ~(flags | (flags << 1));
    if (rev) return (~f) & mask & keep;
    uint32_t r = pcl_brev32(f & mask & keep);
    r = ((r >> 1) & 0x55555555u) | ((r & 0x55555555u) << 1);
    return len >= 16u ? r : r >> (32u - 2u * len);
}

// w[8]: the 32 record bytes.  Keys come back WITHOUT marker for 16-base fields and with it otherwise
// (i.e. in their 4-byte device form); *bc_flags: bit 2i set when base i (record order) of the barcode is not a base
// (0 for a clean barcode); *umi_bad = the UMI holds a byte that is not a base.
// want_flags == false (the count kernels): *bc_flags is only zero / non-zero and the key is not masked.
PCL_HD void pcl_fast_extract(const FastLayout &f, const uint32_t *w, bool rev, uint32_t *bc_key, uint32_t *bc_flags,
                             uint32_t *umi_key, bool *umi_bad, bool want_flags = true) {
    uint64_t codes = 0;
    uint32_t bbad = 0, ubad = 0;
    uint32_t dd[8];
    const uint32_t fold = rev ? 0xDFDFDFDFu : 0xFFFFFFFFu;               // precellar::utils::rev_compl accepts lower case
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int k = 0; k < 8; ++k) {
        dd[k] = 0u;
        if ((f.bc_mask[k] | f.umi_mask[k]) == 0u) continue;              // uniform: this word holds no barcode / UMI byte
        uint32_t d;
        const uint32_t c8 = pcl_word_codes(w[k] & fold, &d);
        codes |= (uint64_t)c8 << (8 * k);
        dd[k] = d & f.bc_mask[k];
        bbad |= dd[k];
        ubad |= d & f.umi_mask[k];
    }
    uint32_t flags = want_flags ? 0u : bbad;
    if (bbad && want_flags) {                                             // locate the bytes that are not bases
        uint64_t fl = 0;
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
        for (int k = 0; k < 8; ++k) fl |= (uint64_t)pcl_flags8(dd[k]) << (8 * k);
        flags = (uint32_t)(fl >> (2u * f.bc_start));
        if (f.bc_len < 16u) flags &= (1u << (2u * f.bc_len)) - 1u;
    }
    uint32_t bk = pcl_field32_to_key((uint32_t)(codes >> (2u * f.bc_start)), f.bc_len, rev, want_flags ? flags : 0u);
    if (f.bc_len < 16u) bk |= 1u << (2u * f.bc_len);
    *bc_key = bk;
    *bc_flags = flags;
    uint32_t uk = 0;
    if (f.umi_len) {
        uk = pcl_field32_to_key((uint32_t)(codes >> (2u * f.umi_start)), f.umi_len, rev, 0u) | (1u << (2u * f.umi_len));
        if (ubad) uk = 0;
    }
    *umi_key = uk;
    *umi_bad = ubad != 0u;
}

// pcl_fast_extract for a 16-base barcode in words [kBcW, kBcW + 4) and a UMI of 4 * kUmiW bases in words
// [kUmiW0, kUmiW0 + kUmiW) (kUmiW == 0: none): same results, everything about the geometry known at compile time.
template <int kBcW, int kUmiW0, int kUmiW>
PCL_HD void pcl_fast_extract_aligned(const uint32_t *w, bool rev, uint32_t *bc_key, uint32_t *bc_flags, uint32_t *umi_key, bool *umi_bad,
                                     bool want_flags = true) {
    const uint32_t fold = rev ? 0xDFDFDFDFu : 0xFFFFFFFFu;
    uint32_t bcodes = 0, bbad = 0;
    uint32_t d[4];
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int k = 0; k < 4; ++k) {
        bcodes |= pcl_word_codes(w[kBcW + k] & fold, &d[k]) << (8 * k);
        bbad |= d[k];
    }
    uint32_t flags = want_flags ? 0u : bbad;
    if (bbad && want_flags) flags = pcl_flags8(d[0]) | (pcl_flags8(d[1]) << 8) | (pcl_flags8(d[2]) << 16) | (pcl_flags8(d[3]) << 24);
    *bc_key = pcl_field32_to_key(bcodes, 16u, rev, want_flags ? flags : 0u);
    *bc_flags = flags;
    uint32_t uk = 0, ubad = 0;
    if (kUmiW > 0) {
        uint32_t ucodes = 0;
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
        for (int k = 0; k < kUmiW; ++k) {
            uint32_t du;
            ucodes |= pcl_word_codes(w[kUmiW0 + k] & fold, &du) << (8 * k);
            ubad |= du;
        }
        uk = pcl_field32_to_key(ucodes, 4u * kUmiW, rev, 0u) | (1u << (8u * kUmiW));
        if (ubad) uk = 0;
    }
    *umi_key = uk;
    *umi_bad = ubad != 0u;
}

PCL_HD void pcl_fast_extract_spec(const FastLayout &f, const uint32_t *w, bool rev, uint32_t *bc_key, uint32_t *bc_flags,
                                  uint32_t *umi_key, bool *umi_bad, bool want_flags = true) {
    switch (f.spec) {
        case PCL_FAST_SPEC_B0_U16x12: pcl_fast_extract_aligned<0, 4, 3>(w, rev, bc_key, bc_flags, umi_key, umi_bad, want_flags); break;
        case PCL_FAST_SPEC_B0: pcl_fast_extract_aligned<0, 0, 0>(w, rev, bc_key, bc_flags, umi_key, umi_bad, want_flags); break;
        case PCL_FAST_SPEC_B8: pcl_fast_extract_aligned<2, 0, 0>(w, rev, bc_key, bc_flags, umi_key, umi_bad, want_flags); break;
        default: pcl_fast_extract(f, w, rev, bc_key, bc_flags, umi_key, umi_bad, want_flags); break;
    }
}

// What k_filter does with a PCL_WL_REPACK entry: key (invalid bases read as 0), number of invalid bases and the position
// of the last one in reported orientation.  pcl_repack16: a 16-base barcode in four aligned words (PCL_FAST_SPEC_*);
// pcl_fast_repack: any FastLayout, from the record's words.
PCL_HD uint32_t pcl_repack16(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3, bool rev, uint32_t *n_invalid, uint32_t *inv_pos) {
    const uint32_t fold = rev ? 0xDFDFDFDFu : 0xFFFFFFFFu;
    uint32_t d0, d1, d2, d3;
    const uint32_t codes = pcl_word_codes(w0 & fold, &d0) | (pcl_word_codes(w1 & fold, &d1) << 8) |
                           (pcl_word_codes(w2 & fold, &d2) << 16) | (pcl_word_codes(w3 & fold, &d3) << 24);
    const uint32_t flags = pcl_flags8(d0) | (pcl_flags8(d1) << 8) | (pcl_flags8(d2) << 16) | (pcl_flags8(d3) << 24);
    pcl_flags_to_inv(flags, 16u, rev, n_invalid, inv_pos);
    return pcl_field32_to_key(codes, 16u, rev, flags);
}

PCL_HD_COLD uint32_t pcl_fast_repack(const FastLayout &f, const uint32_t *w, bool rev, uint32_t *n_invalid, uint32_t *inv_pos) {
    if (f.spec != PCL_FAST_SPEC_NONE) {
        const uint32_t k = f.bc_start >> 2;
        return pcl_repack16(w[k], w[k + 1], w[k + 2], w[k + 3], rev, n_invalid, inv_pos);
    }
    uint32_t key, flags, uk;
    bool ub;
    pcl_fast_extract(f, w, rev, &key, &flags, &uk, &ub, true);
    pcl_flags_to_inv(flags, f.bc_len, rev, n_invalid, inv_pos);
    return key;
}

// Any of the 8 keys of bucket b equal to k?  *full = the bucket has no free slot (a longer probe chain is possible).
PCL_HD bool pcl_bucket32_any(const uint32_t *keys, uint32_t b, uint32_t k, uint32_t empty, bool *full) {
    auto v0 = PCL_LDG128(keys + (uint64_t)b * 8);
    auto v1 = PCL_LDG128(keys + (uint64_t)b * 8 + 4);
    *full = v1.w != empty;
    return (v0.x == k) | (v0.y == k) | (v0.z == k) | (v0.w == k) | (v1.x == k) | (v1.y == k) | (v1.z == k) | (v1.w == k);
}

PCL_HD uint32_t pcl_ctz64(uint64_t x) {
#ifdef __CUDA_ARCH__
    return (uint32_t)(__ffsll((long long)x) - 1);
#else
    return (uint32_t)__builtin_ctzll(x);
#endif
}

// correct_barcode / likelihood1 for barcodes of a 32-bit table (see pcl_correct_one for the semantics).
// `key32` is the 4-byte device key (marker kept for < 16 bases, dropped for 16; an invalid base reads as 0).
//
// Two stages, so that the common case (a neighbour that is simply absent) is straight-line code and so that the
// stage that needs no counts can run before the counts of all ranks are known:
//   1. pcl_filter32: the 1-Hamming neighbours of the query that may be whitelist members, as bits of a 64-bit mask
//      indexed (pos << 2 | base), i.e. already in the reference's (pos, ACGT) order.
//   2. pcl_resolve32: the few survivors are probed in the main table, in mask order, and enter the fp64 posterior in
//      that order (barcode.rs:319-334: sum order and first-maximum tie-break).
// Is table q worth a look?  Byte q must hold a base of this key, be the byte of the N when there is one, and some
// whitelist key must share the query's other three bytes (bitmap; L2-resident).
PCL_HD bool pcl_filter_live(const DTable &t, uint32_t key32, uint32_t len, uint32_t n_invalid, uint32_t inv_pos, uint32_t q) {
    if (n_invalid == 1u && ((2u * (len - 1u - inv_pos)) >> 3) != q) return false;
    if (len < 16u && ((0xFFu << (8u * q)) & ((1u << (2u * len)) - 1u)) == 0u) return false;   // byte q holds no base of this key
    if (!t.nbits[q]) return true;
    const uint32_t idx = pcl_partial_index(key32, q);
#ifdef __CUDA_ARCH__
    return (__ldg(t.nbits[q] + (idx >> 5)) >> (idx & 31u)) & 1u;
#else
    return (t.nbits[q][idx >> 5] >> (idx & 31u)) & 1u;
#endif
}

// (bytes << 8) | byte q of e
PCL_HD uint32_t pcl_push_byte(uint32_t bytes, uint32_t e, uint32_t q) {
#ifdef __CUDA_ARCH__
    return __byte_perm(bytes, e, 0x2100u | (4u + q));             // result bytes: [e.q, bytes.0, bytes.1, bytes.2]
#else
    return (bytes << 8) | ((e >> (8u * q)) & 0xFFu);
#endif
}

// One candidate byte `eb` of table q against the query byte `qb`: its bit in the table's 16-bit sub-mask
// (bit 4 * (3 - base_in_byte) + base_value), or 0 when the two bytes are not exactly one substitution apart
// (with an N in the query: when they differ anywhere but at the N).
PCL_HD uint32_t pcl_byte_candidate(uint32_t eb, uint32_t qb, bool has_n, uint32_t n_p2) {
    const uint32_t xb = eb ^ qb;
    const uint32_t d = (xb | (xb >> 1)) & 0x55u;                     // one bit per differing base of the byte
    const bool one = has_n ? (d & ~(1u << n_p2)) == 0u : (d != 0u && (d & (d - 1u)) == 0u);
    const uint32_t p2 = has_n ? n_p2 : 31u - pcl_clz32(d | 1u);      // bit index (0, 2, 4, 6) of the substituted base
    return one ? 1u << (12u - 2u * p2 + ((eb >> p2) & 3u)) : 0u;
}

PCL_HD_COLD uint32_t pcl_bucket_candidates(uint32_t e0, uint32_t e1, uint32_t e2, uint32_t e3, uint32_t e4, uint32_t e5, uint32_t e6,
                                           uint32_t e7, uint32_t key32, uint32_t keep, uint32_t sq, uint32_t qb, bool has_n, uint32_t n_p2) {
    const uint32_t e[8] = {e0, e1, e2, e3, e4, e5, e6, e7};
    uint32_t sub = 0;
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int k = 0; k < 8; ++k)
        if (((e[k] ^ key32) & keep) == 0u) sub |= pcl_byte_candidate((e[k] >> sq) & 0xFFu, qb, has_n, n_p2);
    return sub;
}

// One bucket of neighbourhood table q against the query: the table's 16-bit sub-mask of 1-Hamming neighbours found in
// it (they differ from the query inside byte q only); *full = the chain continues in the next bucket.  The entries that
// share the query's other three bytes are found with one compare each and their q-th bytes queued in a register (an
// empty slot that happens to share them queues a byte too: the main-table probe of pcl_resolve32 drops that candidate);
// the queued bytes -- one, seldom two -- go through the 1-Hamming test.  has_n: the query holds one non-ACGT base, at bit
// n_p2 (0, 2, 4, 6) of byte q.
PCL_HD uint32_t pcl_filter_bucket(const uint32_t *nk, uint32_t b, uint32_t key32, uint32_t q, uint32_t empty, bool has_n,
                                  uint32_t n_p2, bool *full) {
    const uint32_t sq = 8u * q, keep = ~(0xFFu << sq);
    const uint32_t qb = (key32 >> sq) & 0xFFu;
    auto v0 = PCL_LDG128(nk + (uint64_t)b * 8);
    auto v1 = PCL_LDG128(nk + (uint64_t)b * 8 + 4);
    const uint32_t e[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
    uint32_t n_match = 0, bytes = 0;
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int k = 0; k < 8; ++k) {
        const bool mk = ((e[k] ^ key32) & keep) == 0u;
        bytes = mk ? pcl_push_byte(bytes, e[k], q) : bytes;
        n_match += mk;
    }
    // (queue slots beyond n_match hold zeros: tested, counted only when filled)
    const uint32_t c0 = pcl_byte_candidate(bytes & 0xFFu, qb, has_n, n_p2);
    const uint32_t c1 = pcl_byte_candidate((bytes >> 8) & 0xFFu, qb, has_n, n_p2);
    uint32_t sub = (n_match >= 1u ? c0 : 0u) | (n_match >= 2u ? c1 : 0u);
    if (n_match > 2u) sub |= pcl_bucket_candidates(e[0], e[1], e[2], e[3], e[4], e[5], e[6], e[7], key32, keep, sq, qb, has_n, n_p2);   // rare
    *full = v1.w != empty;                                 // the chain ends at a bucket with a free slot
    return sub;
}

// Sub-mask of table q (bit 4 * (3 - base_in_byte) + base_value) -> bits (pos << 2 | base) of the candidate mask:
// base p2/2 of byte q is position len - 1 - (4q + p2/2), so sub-mask bit 0 belongs to position len - 4 - 4q.  For keys
// shorter than 16 bases the top byte holds fewer than four bases (and the marker): positions < 0 fall off.
PCL_HD uint64_t pcl_filter_place(uint32_t sub, uint32_t len, uint32_t q) {
    const int sh4 = 4 * ((int)len - 4 - 4 * (int)q);
    return sh4 >= 0 ? (uint64_t)sub << sh4 : (uint64_t)(sub >> (-sh4));
}

PCL_HD uint64_t pcl_filter_scan(const DTable &t, uint32_t key32, uint32_t len, uint32_t n_invalid, uint32_t inv_pos, uint32_t q) {
    const uint32_t nb = t.n_buckets;
    const uint32_t n_p2 = (2u * (len - 1u - (n_invalid == 1u ? inv_pos : 0u))) & 7u;
    uint32_t b = pcl_hash32_bucket(key32 & ~(0xFFu << (8u * q)), nb), sub = 0;
    bool full = true;
    while (full) {
        sub |= pcl_filter_bucket(t.nkeys[q], b, key32, q, (uint32_t)t.empty, n_invalid == 1u, n_p2, &full);
        b = (b + 1u == nb) ? 0u : b + 1u;
    }
    return pcl_filter_place(sub, len, q);
}

// Without neighbourhood tables (PCL_NO_NEIGHBOUR_TABLES, kept for tests): one bucket of the main table per neighbour,
// branch-free; a neighbour survives when it is found or when its bucket is full (the probe chain may continue).
PCL_HD_COLD uint64_t pcl_filter_direct(const DTable &t, uint32_t key32, uint32_t len, uint32_t n_invalid, uint32_t inv_pos) {
    const uint32_t *keys = reinterpret_cast<const uint32_t *>(t.keys);
    const uint32_t empty = (uint32_t)t.empty, nb = t.n_buckets;
    uint64_t pend = 0;
    const uint32_t p0 = n_invalid == 1u ? inv_pos : 0u, p1 = n_invalid == 1u ? inv_pos + 1u : len;
    for (uint32_t pos = p0; pos < p1; ++pos) {
        const uint32_t sh = 2u * (len - 1u - pos);
        const uint32_t existing = n_invalid == 1u ? 4u : ((key32 >> sh) & 3u);
        const uint32_t base = key32 & ~(3u << sh);
        uint32_t m4 = 0;
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
        for (uint32_t val = 0; val < 4u; ++val) {
            const uint32_t cand = base | (val << sh);
            const bool act = (val != existing) & (cand != empty);
            // inactive lanes read bucket 0 (one broadcast sector) instead of branching around the loads
            const uint32_t b = act ? pcl_hash32_bucket(cand, nb) : 0u;
            bool full;
            const bool any = pcl_bucket32_any(keys, b, cand, empty, &full);
            m4 |= (uint32_t)(act & (any | full)) << val;
        }
        pend |= (uint64_t)m4 << (4u * pos);
    }
    return pend;
}

// no entry of that length in the table, or no single substitution can reach the whitelist
PCL_HD bool pcl_filter_hopeless(const DTable &t, uint32_t len, uint32_t n_invalid) {
    return n_invalid > 1u || (t.mode == PCL_TAB_K32_L16 ? len != 16u : len > 15u);
}

PCL_HD uint64_t pcl_filter32(const DTable &t, uint32_t key32, uint32_t len, uint32_t n_invalid, uint32_t inv_pos) {
    if (pcl_filter_hopeless(t, len, n_invalid)) return 0;
    if (!t.nkeys[0]) return pcl_filter_direct(t, key32, len, n_invalid, inv_pos);
    uint64_t pend = 0;
    for (uint32_t q = 0; q < 4u; ++q)
        if (pcl_filter_live(t, key32, len, n_invalid, inv_pos, q)) pend |= pcl_filter_scan(t, key32, len, n_invalid, inv_pos, q);
    return pend;
}

// `qual` points at the record's quality row in record-relative indexing (see pcl_correct_one).
PCL_HD uint32_t pcl_resolve32(const DTable &t, uint32_t key32, uint32_t len, uint64_t pend, const uint8_t *qual, uint32_t start,
                              bool rev, double threshold, const double *lut_cap, uint32_t *out_key) {
    bool have = false;
    double best_like = 0.0, total = 0.0;
    uint32_t best_key = 0;
    while (pend) {
        const uint32_t bit = pcl_ctz64(pend);
        pend &= pend - 1;
        const uint32_t pos = bit >> 2, val = bit & 3u;
        const uint32_t sh = 2u * (len - 1u - pos);
        const uint32_t cand = (key32 & ~(3u << sh)) | (val << sh);
        const uint32_t slot = pcl_probe32(t, cand);
        if (slot == 0xFFFFFFFFu) continue;
        const uint32_t c = t.counts[slot];
        const double perr = lut_cap[qual[start + (rev ? len - 1u - pos : pos)]];
        const double like = PCL_DMUL((double)(1ull + (uint64_t)c), perr);          // barcode.rs:326-327
        if (!have || best_like < like) { have = true; best_like = like; best_key = cand; }   // :345-358
        total = PCL_DADD(total, like);
    }
    if (!have) return PCL_BC_NO_MATCH;
    const double prob = PCL_DDIV(best_like, total);
    if (prob <= 0.0) return PCL_BC_NO_MATCH;
    if (prob >= threshold) { *out_key = best_key; return PCL_BC_CORRECTED; }
    return PCL_BC_LOW_CONF;
}

PCL_HD uint32_t pcl_correct_one32(const DTable &t, uint32_t key32, uint32_t len, uint32_t n_invalid, uint32_t inv_pos,
                                  const uint8_t *qual, uint32_t start, bool rev, double threshold, const double *lut_cap,
                                  uint32_t *out_key) {
    const uint64_t pend = pcl_filter32(t, key32, len, n_invalid, inv_pos);
    if (!pend) return PCL_BC_NO_MATCH;
    return pcl_resolve32(t, key32, len, pend, qual, start, rev, threshold, lut_cap, out_key);
}

// ------------------------------------------------------------------------------------------
// whole-record 2-bit representation (records of up to 64 bytes): general layouts without byte loops
// ------------------------------------------------------------------------------------------
struct RecCodes {
    uint64_t lo, hi;        // 2-bit code of base i at bits 2i (i < 32: lo, else hi)
    uint64_t bad_lo, bad_hi;   // bit 2i set: byte i is not a base for barcodes / UMIs (case folded on neg-strand reads)
    uint64_t sb_lo, sb_hi;     // bit 2i set: byte i is not upper-case ACGT (anchor comparison, seqspec::utils::rev_compl)
};

// w: n_words 32-bit words of the record (n_words <= 16).  kRev: neg-strand read (case folding for barcodes / UMIs,
// a separate strict mask for anchors); forward reads share one mask.
template <bool kRev>
PCL_HD void pcl_rec_codes_t(const uint32_t *w, int n_words, RecCodes &rc) {
    uint32_t c32[4] = {0, 0, 0, 0}, b32[4] = {0, 0, 0, 0}, s32[4] = {0, 0, 0, 0};   // 16 bases per word
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int k = 0; k < 16; ++k) {
        if (k >= n_words) break;
        uint32_t bad, sbad = 0;
        const uint32_t c8 = pcl_word_codes(kRev ? (w[k] & 0xDFDFDFDFu) : w[k], &bad);
        // non-zero byte -> one flag bit per base at the even bit positions of an 8-bit field
        const uint32_t nb = ((((bad & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | bad) & 0x80808080u) >> 7;
        c32[k >> 2] |= c8 << (8 * (k & 3));
        b32[k >> 2] |= ((nb * 0x01041040u) >> 24) << (8 * (k & 3));
        if (kRev) {
            (void)pcl_word_codes(w[k], &sbad);
            const uint32_t ns = ((((sbad & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | sbad) & 0x80808080u) >> 7;
            s32[k >> 2] |= ((ns * 0x01041040u) >> 24) << (8 * (k & 3));
        }
    }
    rc.lo = ((uint64_t)c32[1] << 32) | c32[0];
    rc.hi = ((uint64_t)c32[3] << 32) | c32[2];
    rc.bad_lo = ((uint64_t)b32[1] << 32) | b32[0];
    rc.bad_hi = ((uint64_t)b32[3] << 32) | b32[2];
    if (kRev) { rc.sb_lo = ((uint64_t)s32[1] << 32) | s32[0]; rc.sb_hi = ((uint64_t)s32[3] << 32) | s32[2]; }
    else { rc.sb_lo = rc.bad_lo; rc.sb_hi = rc.bad_hi; }
}

PCL_HD void pcl_rec_codes(const uint32_t *w, int n_words, bool rev, RecCodes &rc) {
    if (rev) pcl_rec_codes_t<true>(w, n_words, rc); else pcl_rec_codes_t<false>(w, n_words, rc);
}

// 2*len bits starting at base `start` of a 128-bit field (len <= 32, start + len <= 64)
PCL_HD uint64_t pcl_bits128(uint64_t lo, uint64_t hi, uint32_t start, uint32_t len) {
    const uint32_t bp = 2u * start;
    uint64_t v;
    if (bp == 0u) v = lo;
    else if (bp < 64u) v = (lo >> bp) | (hi << (64u - bp));
    else v = hi >> (bp - 64u);
    return len >= 32u ? v : (v & ((1ull << (2u * len)) - 1ull));
}

PCL_HD uint32_t pcl_popc64(uint64_t x) {
#ifdef __CUDA_ARCH__
    return (uint32_t)__popcll(x);
#else
    return (uint32_t)__builtin_popcountll(x);
#endif
}

PCL_HD uint64_t pcl_brev64(uint64_t x) {
#ifdef __CUDA_ARCH__
    return __brevll(x);
#else
    return ((uint64_t)pcl_brev32((uint32_t)x) << 32) | pcl_brev32((uint32_t)(x >> 32));
#endif
}

// pcl_pack_key on the code representation: key in marker form, *n_invalid = bases that are not bases.
PCL_HD uint64_t pcl_pack_codes(const RecCodes &rc, uint32_t start, uint32_t len, bool rev, uint32_t *n_invalid,
                               uint32_t *inv_pos = nullptr) {
    if (inv_pos) *inv_pos = 0;
    if (len > PCL_MAX_KEY_LEN) { *n_invalid = len; return 0; }
    if (len == 0u) { *n_invalid = 0; return 1; }
    const uint64_t bad = pcl_bits128(rc.bad_lo, rc.bad_hi, start, len);
    *n_invalid = pcl_popc64(bad);
    if (inv_pos && bad) {                                            // the last invalid position in reported orientation
        const uint32_t lo = pcl_ctz64(bad) >> 1, hi = (63u - pcl_ctz64(pcl_brev64(bad))) >> 1;
        *inv_pos = rev ? len - 1u - lo : hi;
    }
    uint64_t f = pcl_bits128(rc.lo, rc.hi, start, len);
    f &= ~(bad | (bad << 1));                                        // invalid positions contribute 0, like pcl_pack_key
    const uint64_t mask = (1ull << (2u * len)) - 1ull;
    uint64_t key;
    if (rev) {
        key = (~f) & mask & ~(bad | (bad << 1));                     // reverse complement == complement read big-endian
    } else {
        uint64_t r = pcl_brev64(f);
        r = ((r >> 1) & 0x5555555555555555ull) | ((r & 0x5555555555555555ull) << 1);
        key = r >> (64u - 2u * len);
    }
    return key | (1ull << (2u * len));
}

// find_best_pattern_match in code space (anchors of <= 32 upper-case ACGT bases)
struct CodeMatcher {
    const RecCodes *rc;
    bool rev;
    PCL_HD bool operator()(const DElem &el, uint32_t wstart, uint32_t wn, uint32_t *pos, uint32_t *mis) const {
        const uint32_t m = el.seq_len;
        if (m == 0u || wn == 0u || m > wn) return false;
        const uint64_t needle = rev ? el.a_rc : el.a_code;
        const uint64_t even = m >= 32u ? 0x5555555555555555ull : (0x5555555555555555ull & ((1ull << (2u * m)) - 1ull));
        uint32_t best = 0xFFFFFFFFu, bpos = 0;
        for (uint32_t i = 0; i + m <= wn; ++i) {
            // neg-strand reads search rev_compl(window): haystack position i is the forward position wn - m - i
            const uint32_t p = wstart + (rev ? wn - m - i : i);
            const uint64_t x = pcl_bits128(rc->lo, rc->hi, p, m) ^ needle;
            const uint64_t d = ((x | (x >> 1)) | pcl_bits128(rc->sb_lo, rc->sb_hi, p, m)) & even;
            const uint32_t c = pcl_popc64(d);
            if (c < best) { best = c; bpos = i; if (c == 0u) break; }
        }
        if (m == 1u && best != 0u) return false;
        *pos = bpos;
        *mis = best;
        return true;
    }
    PCL_HD uint32_t fixed_distance(const DElem &el, uint32_t start) const {      // anchors of <= 32 upper-case ACGT bases
        const uint32_t m = el.seq_len;
        const uint64_t even = m >= 32u ? 0x5555555555555555ull : (0x5555555555555555ull & ((1ull << (2u * m)) - 1ull));
        const uint64_t x = pcl_bits128(rc->lo, rc->hi, start, m) ^ (rev ? el.a_rc : el.a_code);
        return pcl_popc64(((x | (x >> 1)) | pcl_bits128(rc->sb_lo, rc->sb_hi, start, m)) & even);
    }
};

PCL_HD void pcl_split_codes(const DRead &rd, const RecCodes &rc, uint32_t len, SplitOut &o) {
    CodeMatcher m{&rc, rd.is_reverse != 0};
    pcl_split_t(rd, len, o, m);
}

// ------------------------------------------------------------------------------------------
// AnchorPlan path: 32-bit code strings (16 bases per word) and the collapsed split
// ------------------------------------------------------------------------------------------
struct Codes32 {
    uint32_t c[4];      // 2-bit code of base i at bits 2(i % 16) of word i / 16
    uint32_t bd[4];     // bit 2(i % 16) set: byte i is not a base for barcodes / UMIs (case folded on neg-strand reads)
    uint32_t sb[4];     // the same with the strict upper-case test used for anchors (== bd on forward reads)
};

template <bool kRev>
PCL_HD void pcl_codes32_t(const uint32_t *w, int n_words, Codes32 &k) {
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int g = 0; g < 4; ++g) {
        uint32_t c = 0, b = 0, sb = 0;
        if (4 * g < n_words) {                                    // warp-uniform: the record's words come in groups of four
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
            for (int q = 0; q < 4; ++q) {
                const uint32_t x = w[4 * g + q];
                uint32_t bad;
                const uint32_t c8 = pcl_word_codes(kRev ? (x & 0xDFDFDFDFu) : x, &bad);
                c |= c8 << (8 * q);
                b |= pcl_flags8(bad) << (8 * q);
                if (kRev) {
                    uint32_t sbad;
                    (void)pcl_word_codes(x, &sbad);
                    sb |= pcl_flags8(sbad) << (8 * q);
                }
            }
        }
        k.c[g] = c; k.bd[g] = b; k.sb[g] = kRev ? sb : b;
    }
}

// The same, converting exactly n_words words (warp-uniform): the fields of an AnchorPlan layout seldom end on a
// 16-byte boundary, and the conversion is a quarter of that kernel's instructions.
template <bool kRev>
PCL_HD void pcl_codes32_exact(const uint32_t *w, int n_words, Codes32 &k) {
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int g = 0; g < 4; ++g) {
        uint32_t c = 0, b = 0, sb = 0;
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
        for (int q = 0; q < 4; ++q) {
            if (4 * g + q >= n_words) break;                          // warp-uniform
            const uint32_t x = w[4 * g + q];
            uint32_t bad;
            const uint32_t c8 = pcl_word_codes(kRev ? (x & 0xDFDFDFDFu) : x, &bad);
            c |= c8 << (8 * q);
            b |= pcl_flags8(bad) << (8 * q);
            if (kRev) {
                uint32_t sbad;
                (void)pcl_word_codes(x, &sbad);
                sb |= pcl_flags8(sbad) << (8 * q);
            }
        }
        k.c[g] = c; k.bd[g] = b; k.sb[g] = kRev ? sb : b;
    }
}

// bases [start, start + len) of a code string as a 32-bit field (len <= 16, start + len <= 64)
PCL_HD uint32_t pcl_win32(const uint32_t *c, uint32_t start, uint32_t len) {
    const uint32_t wi = start >> 4, sh = (start & 15u) * 2u;
    const uint32_t a = wi == 0u ? c[0] : wi == 1u ? c[1] : wi == 2u ? c[2] : c[3];
    const uint32_t b = wi == 0u ? c[1] : wi == 1u ? c[2] : wi == 2u ? c[3] : 0u;
#ifdef __CUDA_ARCH__
    const uint32_t v = __funnelshift_r(a, b, sh);
#else
    const uint32_t v = sh ? (a >> sh) | (b << (32u - sh)) : a;
#endif
    return len >= 16u ? v : (v & ((1u << (2u * len)) - 1u));
}

// pcl_win32 for a start that is the same in every lane of the warp: the word pair is picked by a (divergence-free)
// branch instead of per-lane selects
PCL_HD uint32_t pcl_win32u(const uint32_t *c, uint32_t start, uint32_t len) {
    const uint32_t wi = start >> 4, sh = (start & 15u) * 2u;
    uint32_t a, b;
    switch (wi) {
        case 0u: a = c[0]; b = c[1]; break;
        case 1u: a = c[1]; b = c[2]; break;
        case 2u: a = c[2]; b = c[3]; break;
        default: a = c[3]; b = 0u; break;
    }
#ifdef __CUDA_ARCH__
    const uint32_t v = __funnelshift_r(a, b, sh);
#else
    const uint32_t v = sh ? (a >> sh) | (b << (32u - sh)) : a;
#endif
    return len >= 16u ? v : (v & ((1u << (2u * len)) - 1u));
}

// the whole string shifted down by `bases` (< 16) bases
PCL_HD void pcl_shr_bases(const uint32_t *in, uint32_t bases, uint32_t *out) {
    const uint32_t sh = 2u * bases;
#ifdef __CUDA_ARCH__
    out[0] = __funnelshift_r(in[0], in[1], sh); out[1] = __funnelshift_r(in[1], in[2], sh);
    out[2] = __funnelshift_r(in[2], in[3], sh); out[3] = in[3] >> sh;
#else
    for (int k = 0; k < 4; ++k) out[k] = sh ? (in[k] >> sh) | ((k < 3 ? in[k + 1] : 0u) << (32u - sh)) : in[k];
#endif
}

struct AnchorSplit {
    uint32_t status;     // PCL_SPLIT_OK / ANCHOR_NOT_FOUND / PATTERN_MISMATCH
    uint32_t reach;      // elements [0, reach) have a segment
    uint32_t pos;        // anchor position inside its window (0 unless the anchor matched)
    uint32_t v_len;      // length of the variable element's segment
    uint32_t t_len;      // length of the trailing variable element's segment
};

// split_with_tolerance (segment.rs:180-368) for an AnchorPlan layout; same control flow as pcl_split_t, element by element
PCL_HD void pcl_anchor_split(const DRead &rd, const Codes32 &k, uint32_t len, AnchorSplit &s) {
    const AnchorPlan &pl = rd.plan;
    const uint32_t v = pl.v, a = v + 1u, n = rd.n_elems;
    const bool rev = rd.is_reverse != 0;
    s.status = PCL_SPLIT_OK; s.reach = 0; s.pos = 0; s.v_len = 0; s.t_len = 0;
    uint32_t off = 0;
    for (uint32_t e = 0; e < v; ++e) {                                            // fixed elements, empty buffer (segment.rs:315-344)
        const uint32_t l = rd.elems[e].max_len;
        if (off >= len || off + l > len) return;                                  // :245-248
        s.reach = e + 1u;
        off += l;
    }
    const uint32_t vmin = rd.elems[v].min_len, vmax = rd.elems[v].max_len;
    if (off >= len || off + vmin > len) return;
    const uint32_t min_off = off + vmin, max_off = off + vmax;                    // buffer = [v]
    const uint32_t m = rd.elems[a].max_len;                                       // == seq_len
    const uint32_t tail_hi = max_off < len ? max_off : len;
    if (max_off >= len || min_off + m > len) {                                    // the loop breaks at the anchor: tail consume_buf (:353-365)
        s.reach = v + 1u; s.v_len = tail_hi - off;
        return;
    }
    const uint32_t wend = max_off + m < len ? max_off + m : len, wn = wend - min_off;      // :258-262, wn >= m
    const uint32_t needle = (uint32_t)(rev ? rd.elems[a].a_rc : rd.elems[a].a_code);
    const uint32_t even = m >= 16u ? 0x55555555u : (0x55555555u & ((1u << (2u * m)) - 1u));
    uint32_t best = 0xFFFFFFFFu, bpos = 0;
    for (uint32_t i = 0; i + m <= wn; ++i) {                                      // find_best_pattern_match (:496-535)
        const uint32_t p = min_off + (rev ? wn - m - i : i);                      // neg strand: position i of rev_compl(window)
        // forward reads: p is the same in every lane; neg strand: it depends on the record length through wn
        const uint32_t x = (rev ? pcl_win32(k.c, p, m) : pcl_win32u(k.c, p, m)) ^ needle;
        const uint32_t d = ((x | (x >> 1)) | (rev ? pcl_win32(k.sb, p, m) : pcl_win32u(k.sb, p, m))) & even;
#ifdef __CUDA_ARCH__
        const uint32_t cnt = (uint32_t)__popc(d);
#else
        const uint32_t cnt = (uint32_t)__builtin_popcount(d);
#endif
        if (cnt < best) { best = cnt; bpos = i; if (cnt == 0u) break; }
    }
    if (m == 1u && best != 0u) { s.status = PCL_SPLIT_ANCHOR_NOT_FOUND; return; } // (None, 1) (:508-512)
    if (best > rd.elems[a].max_mis) {
        if (max_off + m > len) { s.reach = v + 1u; s.v_len = tail_hi - off; return; }       // :303-305, then the tail consume
        s.status = PCL_SPLIT_PATTERN_MISMATCH;                                    // :306-308
        return;
    }
    s.pos = bpos;
    s.v_len = vmin + bpos;                                                        // [offset_left, offset_right)
    s.reach = a + 1u;
    off = min_off + bpos + m;
    const uint32_t n_fixed_end = pl.has_t ? n - 1u : n;
    for (uint32_t e = a + 1u; e < n_fixed_end; ++e) {
        const uint32_t l = rd.elems[e].max_len;
        if (off >= len || off + l > len) return;
        s.reach = e + 1u;
        off += l;
    }
    if (pl.has_t) {
        const uint32_t tmin = rd.elems[n - 1u].min_len, tmax = rd.elems[n - 1u].max_len;
        if (off >= len || off + tmin > len) return;
        s.reach = n;
        s.t_len = (off + tmax < len ? off + tmax : len) - off;                    // tail consume_buf
    }
}

// (start << 16 | len) of element e after pcl_anchor_split, or PCL_COORD_ABSENT
PCL_HD uint32_t pcl_anchor_coord(const DRead &rd, const AnchorSplit &s, uint32_t e) {
    if (s.status != PCL_SPLIT_OK || e >= s.reach) return PCL_COORD_ABSENT;
    const AnchorPlan &pl = rd.plan;
    const uint32_t start = pl.ustart[e] + (e > pl.v ? s.pos : 0u);
    const uint32_t l = e == pl.v ? s.v_len : (pl.has_t && e + 1u == rd.n_elems) ? s.t_len : rd.elems[e].max_len;
    return (start << 16) | l;
}

// Device-form 32-bit key (marker kept below 16 bases) of the field at coordinate `coord`; `k2` holds the strings
// shifted down by s.pos bases, so that elements behind the anchor sit at the warp-uniform offset ustart[e].
PCL_HD uint32_t pcl_anchor_key(const DRead &rd, const Codes32 &k, const uint32_t *c2, const uint32_t *bd2, uint32_t e, uint32_t coord,
                               uint32_t *n_invalid, uint32_t *inv_pos = nullptr) {
    const uint32_t len = coord & 0xFFFFu, us = rd.plan.ustart[e];
    const bool behind = e > rd.plan.v;
    const uint32_t wlen = len > 16u ? 16u : len;
    if (inv_pos) *inv_pos = 0;
    if (wlen == 0u) { *n_invalid = 0; return 1u; }                                // an empty field: marker only, like pcl_pack_key
    uint32_t f = pcl_win32u(behind ? c2 : k.c, us, wlen);                         // `us` and `behind` are warp-uniform
    const uint32_t bad = pcl_win32u(behind ? bd2 : k.bd, us, wlen);
    if (inv_pos) pcl_flags_to_inv(bad, wlen, rd.is_reverse != 0, n_invalid, inv_pos);
    else *n_invalid = pcl_popc32(bad);
    const uint32_t bad2 = bad | (bad << 1);
    const bool rev = rd.is_reverse != 0;
    uint32_t key;
    if (rev) {
        const uint32_t mask = wlen >= 16u ? 0xFFFFFFFFu : ((1u << (2u * wlen)) - 1u);
        key = (~f) & mask & ~bad2;                                               // reverse complement == complement read big-endian
    } else {
        f &= ~bad2;                                                               // an invalid base reads as 0, like pcl_pack_key
        key = pcl_field_to_key(f, 0u, wlen, false);
    }
    if (wlen < 16u) key |= 1u << (2u * wlen);
    return key;
}

// What pass 1 derives from one record of an AnchorPlan layout (keys in their 4-byte device form).
struct PlanOut {
    uint32_t status;                                   // PCL_SPLIT_*
    uint32_t tcoord;                                   // (start << 16) | len of the target, or PCL_COORD_ABSENT
    uint32_t bcoord[PCL_MAX_BC_PER_READ], bkey[PCL_MAX_BC_PER_READ], bninv[PCL_MAX_BC_PER_READ], bipos[PCL_MAX_BC_PER_READ];
    uint32_t ucoord[PCL_MAX_UMI_PER_READ], ukey[PCL_MAX_UMI_PER_READ];
};

// The general route: element-by-element walk (pcl_anchor_split) and per-element coordinates, any record length.
template <bool kRev>
PCL_HD void pcl_plan_record(const DRead &rd, const uint32_t *w, uint32_t L, PlanOut &o) {
    const uint32_t nbytes = L < rd.stride ? L : rd.stride;
    Codes32 k;
    pcl_codes32_t<kRev>(w, (int)((nbytes + 3u) >> 2), k);
    AnchorSplit sp;
    pcl_anchor_split(rd, k, L, sp);
    uint32_t c2[4], bd2[4];                                 // the strings behind the anchor, at uniform offsets
    pcl_shr_bases(k.c, sp.pos, c2);
    pcl_shr_bases(k.bd, sp.pos, bd2);
    o.status = sp.status;
    o.tcoord = (rd.has_target && rd.plan.t_elem != 0xFFu) ? pcl_anchor_coord(rd, sp, rd.plan.t_elem) : PCL_COORD_ABSENT;
    for (uint32_t j = 0; j < rd.n_bc; ++j) {
        const uint32_t e = rd.plan.bc_elem[j], c = pcl_anchor_coord(rd, sp, e);
        o.bcoord[j] = c; o.bkey[j] = 0; o.bninv[j] = 0; o.bipos[j] = 0;
        if (c != PCL_COORD_ABSENT) o.bkey[j] = pcl_anchor_key(rd, k, c2, bd2, e, c, &o.bninv[j], &o.bipos[j]);
    }
    for (uint32_t j = 0; j < rd.n_umi; ++j) {
        const uint32_t e = rd.plan.umi_elem[j], c = pcl_anchor_coord(rd, sp, e);
        o.ucoord[j] = c; o.ukey[j] = PCL_KEY_ABSENT32;
        if (c != PCL_COORD_ABSENT) {
            uint32_t ninv;
            const uint32_t key = pcl_anchor_key(rd, k, c2, bd2, e, c, &ninv);
            o.ukey[j] = ninv ? 0u : key;
        }
    }
}

// Key of bases [start, start + len) of a code string (len <= 16) with its flag string: device form, invalid bases read
// as 0; `start` must be the same in every lane of the warp.
PCL_HD uint32_t pcl_plan_key(const uint32_t *c, const uint32_t *bd, uint32_t start, uint32_t len, bool rev, uint32_t *n_invalid, uint32_t *inv_pos) {
    if (len == 0u) { *n_invalid = 0; *inv_pos = 0; return 1u; }
    const uint32_t f = pcl_win32u(c, start, len), bad = pcl_win32u(bd, start, len);
    pcl_flags_to_inv(bad, len, rev, n_invalid, inv_pos);
    uint32_t key = pcl_field32_to_key(f, len, rev, bad);
    if (len < 16u) key |= 1u << (2u * len);
    return key;
}

// ------------------------------------------------------------------------------------------
// Geometry of an AnchorPlan layout as seen by the straight-line route: either read from the DRead at run time
// (PlanGeoRt: any layout the plan accepts) or a set of compile-time constants (PlanGeoSci3 ...: element offsets, anchor
// code, window width and field widths fold into immediates, the loops over positions / barcodes / UMIs unroll, and the
// words of the code strings are picked without selects or uniform branches).  pcl_plan_geo_matches tells the host which
// instance a layout may use; the results are the same by construction (one function body, pcl_plan_fast_g).
// ------------------------------------------------------------------------------------------
struct PlanGeoRt {
    static constexpr bool kConst = false;
    const DRead &rd;
    PCL_HD explicit PlanGeoRt(const DRead &r) : rd(r) {}
    PCL_HD uint32_t n_elems() const { return rd.n_elems; }
    PCL_HD uint32_t v() const { return rd.plan.v; }
    PCL_HD uint32_t vmin() const { return rd.elems[rd.plan.v].min_len; }
    PCL_HD uint32_t n_pos() const { return rd.plan.n_pos; }
    PCL_HD uint32_t a_start() const { return rd.plan.a_start; }
    PCL_HD uint32_t a_len() const { return rd.plan.a_len; }
    PCL_HD uint32_t a_maxmis() const { return rd.plan.a_maxmis; }
    PCL_HD uint32_t needle(bool rev) const { return (uint32_t)(rev ? rd.elems[rd.plan.v + 1u].a_rc : rd.elems[rd.plan.v + 1u].a_code); }
    PCL_HD uint32_t full_len() const { return rd.plan.full_len; }
    PCL_HD uint32_t n_words() const { return rd.plan.n_words; }
    PCL_HD uint32_t stride() const { return rd.stride; }
    PCL_HD uint32_t n_bc() const { return rd.n_bc; }
    PCL_HD uint32_t n_umi() const { return rd.n_umi; }
    PCL_HD uint32_t bc_elem(uint32_t j) const { return rd.plan.bc_elem[j]; }
    PCL_HD uint32_t umi_elem(uint32_t j) const { return rd.plan.umi_elem[j]; }
    PCL_HD uint32_t ustart(uint32_t e) const { return rd.plan.ustart[e]; }
    PCL_HD uint32_t max_len(uint32_t e) const { return rd.elems[e].max_len; }
    PCL_HD bool has_t() const { return rd.plan.has_t != 0; }
    PCL_HD bool has_target() const { return rd.has_target != 0 && rd.plan.t_elem != 0xFFu; }
    PCL_HD uint32_t t_elem() const { return rd.plan.t_elem; }
};

// 2-bit code string of an upper-case ACGT literal, base i at bits 2i (DElem::a_code) / of its reverse complement (a_rc)
constexpr uint32_t pcl_lit_code(const char *s, uint32_t n, bool rc) {
    uint32_t out = 0;
    for (uint32_t t = 0; t < n; ++t) {
        const char b = s[t];
        const uint32_t c = b == 'A' ? 0u : b == 'C' ? 1u : b == 'G' ? 2u : 3u;
        out |= rc ? (3u - c) << (2u * (n - 1u - t)) : c << (2u * t);
    }
    return out;
}

// sci-RNA-seq3 R1 (seqspec_templates/sci_rna_seq3.yaml): hairpin barcode 9-10 | CAGAGC | UMI 8 | RT barcode 10
struct PlanGeoSci3 {
    static constexpr bool kConst = true;
    PCL_HD explicit PlanGeoSci3(const DRead &) {}
    PCL_HD static constexpr uint32_t n_elems() { return 4u; }
    PCL_HD static constexpr uint32_t v() { return 0u; }
    PCL_HD static constexpr uint32_t vmin() { return 9u; }
    PCL_HD static constexpr uint32_t n_pos() { return 2u; }
    PCL_HD static constexpr uint32_t a_start() { return 9u; }
    PCL_HD static constexpr uint32_t a_len() { return 6u; }
    PCL_HD static constexpr uint32_t a_maxmis() { return 1u; }                      // floor(0.2 * 6)
    PCL_HD static constexpr uint32_t needle(bool rev) { return pcl_lit_code("CAGAGC", 6u, rev); }
    PCL_HD static constexpr uint32_t full_len() { return 34u; }
    PCL_HD static constexpr uint32_t n_words() { return 9u; }
    PCL_HD static constexpr uint32_t stride() { return 48u; }                        // 34 record bytes in 16-byte vectors
    PCL_HD static constexpr uint32_t n_bc() { return 2u; }
    PCL_HD static constexpr uint32_t n_umi() { return 1u; }
    PCL_HD static constexpr uint32_t bc_elem(uint32_t j) { return j == 0u ? 0u : 3u; }
    PCL_HD static constexpr uint32_t umi_elem(uint32_t) { return 2u; }
    PCL_HD static constexpr uint32_t ustart(uint32_t e) { return e == 0u ? 0u : e == 1u ? 9u : e == 2u ? 15u : 23u; }
    PCL_HD static constexpr uint32_t max_len(uint32_t e) { return e == 0u ? 10u : e == 1u ? 6u : e == 2u ? 8u : 10u; }
    PCL_HD static constexpr bool has_t() { return false; }
    PCL_HD static constexpr bool has_target() { return false; }
    PCL_HD static constexpr uint32_t t_elem() { return 0xFFu; }
};

// Host side: may layout `rd` run the instance G?  Every constant G holds is compared with the compiled plan.
template <class G>
inline bool pcl_plan_geo_matches(const DRead &rd) {
    const PlanGeoRt r(rd);
    const G g(rd);
    if (!rd.plan.enabled || rd.plan.full_len == 0xFFFFu) return false;
    bool ok = r.n_elems() == g.n_elems() && r.v() == g.v() && r.vmin() == g.vmin() && r.n_pos() == g.n_pos() && r.a_start() == g.a_start() &&
              r.a_len() == g.a_len() && r.a_maxmis() == g.a_maxmis() && r.needle(false) == g.needle(false) && r.needle(true) == g.needle(true) &&
              r.full_len() == g.full_len() && r.n_words() == g.n_words() && r.stride() == g.stride() && r.n_bc() == g.n_bc() && r.n_umi() == g.n_umi() &&
              r.has_t() == g.has_t() && r.has_target() == g.has_target() && (!r.has_target() || r.t_elem() == g.t_elem());
    for (uint32_t e = 0; ok && e < r.n_elems(); ++e) ok = r.ustart(e) == g.ustart(e) && r.max_len(e) == g.max_len(e);
    for (uint32_t j = 0; ok && j < r.n_bc(); ++j) ok = r.bc_elem(j) == g.bc_elem(j);
    for (uint32_t j = 0; ok && j < r.n_umi(); ++j) ok = r.umi_elem(j) == g.umi_elem(j);
    return ok;
}

// The same results as pcl_plan_record for a record that holds every element whatever the anchor position
// (L >= full_len): no walk, no per-element bounds, only the words the fields can touch are converted.  Returns false
// for shorter records.
template <bool kRev, class G>
PCL_HD bool pcl_plan_fast_g(const G &g, const uint32_t *w, uint32_t L, PlanOut &o) {
    if (L < g.full_len()) return false;
    Codes32 k;
    pcl_codes32_exact<kRev>(w, (int)g.n_words(), k);
    // find_best_pattern_match over the n_pos window positions (leftmost strict minimum; segment.rs:496-535)
    const uint32_t m = g.a_len(), v = g.v();
    const uint32_t needle = g.needle(kRev);
    const uint32_t even = m >= 16u ? 0x55555555u : (0x55555555u & ((1u << (2u * m)) - 1u));
    uint32_t best = 0xFFFFFFFFu, bpos = 0;
#ifdef __CUDA_ARCH__
#pragma unroll(G::kConst ? 16 : 1)
#endif
    for (uint32_t i = 0; i < 16u; ++i) {
        if (i >= g.n_pos()) break;
        const uint32_t p = g.a_start() + (kRev ? g.n_pos() - 1u - i : i);            // neg strand: position i of rev_compl(window)
        const uint32_t x = pcl_win32u(k.c, p, m) ^ needle;
        const uint32_t cnt = pcl_popc32(((x | (x >> 1)) | pcl_win32u(k.sb, p, m)) & even);
        if (cnt < best) { best = cnt; bpos = i; }
    }
    uint32_t status = PCL_SPLIT_OK;
    if (m == 1u && best != 0u) status = PCL_SPLIT_ANCHOR_NOT_FOUND;                   // (None, 1) (segment.rs:508-512)
    else if (best > g.a_maxmis()) status = PCL_SPLIT_PATTERN_MISMATCH;               // the window lies inside the record (L >= full_len)
    const bool ok = status == PCL_SPLIT_OK;
    const uint32_t pos = ok ? bpos : 0u;
    uint32_t c2[4], bd2[4];
    pcl_shr_bases(k.c, pos, c2);
    pcl_shr_bases(k.bd, pos, bd2);
    o.status = status;
    const uint32_t n = g.n_elems(), vmin = g.vmin();
    o.tcoord = PCL_COORD_ABSENT;
    if (g.has_target()) {
        const uint32_t e = g.t_elem();
        const uint32_t start = g.ustart(e) + (e > v ? pos : 0u);
        uint32_t len = e == v ? vmin + pos : g.max_len(e);
        if (g.has_t() && e + 1u == n) len = (start + g.max_len(e) < L ? start + g.max_len(e) : L) - start;   // tail consume_buf
        if (ok) o.tcoord = (start << 16) | len;
    }
#ifdef __CUDA_ARCH__
#pragma unroll(G::kConst ? 4 : 1)
#endif
    for (uint32_t j = 0; j < PCL_MAX_BC_PER_READ; ++j) {
        if (j >= g.n_bc()) break;
        const uint32_t e = g.bc_elem(j);
        const bool behind = e > v;
        const uint32_t us = g.ustart(e);
        uint32_t len = e == v ? vmin + pos : g.max_len(e);
        if (g.has_t() && e + 1u == n) len = (us + pos + g.max_len(e) < L ? us + pos + g.max_len(e) : L) - (us + pos);
        o.bcoord[j] = ok ? (((us + (behind ? pos : 0u)) << 16) | len) : PCL_COORD_ABSENT;
        o.bkey[j] = 0; o.bninv[j] = 0; o.bipos[j] = 0;
        if (ok) o.bkey[j] = pcl_plan_key(behind ? c2 : k.c, behind ? bd2 : k.bd, us, len > 16u ? 16u : len, kRev, &o.bninv[j], &o.bipos[j]);
    }
#ifdef __CUDA_ARCH__
#pragma unroll(G::kConst ? 4 : 1)
#endif
    for (uint32_t j = 0; j < PCL_MAX_UMI_PER_READ; ++j) {
        if (j >= g.n_umi()) break;
        const uint32_t e = g.umi_elem(j);
        const bool behind = e > v;
        const uint32_t us = g.ustart(e);
        uint32_t len = e == v ? vmin + pos : g.max_len(e);
        if (g.has_t() && e + 1u == n) len = (us + pos + g.max_len(e) < L ? us + pos + g.max_len(e) : L) - (us + pos);
        o.ucoord[j] = ok ? (((us + (behind ? pos : 0u)) << 16) | len) : PCL_COORD_ABSENT;
        o.ukey[j] = PCL_KEY_ABSENT32;
        if (ok) {
            uint32_t ninv, ipos;
            const uint32_t key = pcl_plan_key(behind ? c2 : k.c, behind ? bd2 : k.bd, us, len > 16u ? 16u : len, kRev, &ninv, &ipos);
            o.ukey[j] = ninv ? 0u : key;
        }
    }
    return true;
}

template <bool kRev>
PCL_HD bool pcl_plan_record_fast(const DRead &rd, const uint32_t *w, uint32_t L, PlanOut &o) {
    return pcl_plan_fast_g<kRev>(PlanGeoRt(rd), w, L, o);
}

// Store a key in its device width (see precellar_b200.h).
PCL_HD void pcl_store_key(void *arr, uint32_t key_bytes, uint64_t i, uint64_t key) {
    if (key_bytes == 4u) reinterpret_cast<uint32_t *>(arr)[i] = (uint32_t)key;
    else reinterpret_cast<uint64_t *>(arr)[i] = key;
}

#endif  // PCL_HD_CORE_H
