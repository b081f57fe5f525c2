// inflate_core.h -- BGZF members inflated on the device: one warp per member.
//
// SURVEY.md §8(f) row 1 (ingest).  The reference opens every FASTQ file through open_file's MultiGzDecoder
// (seqspec/src/utils.rs:44-56) and inflates it on one host thread, twice per run (seqspec/src/read.rs:83-155).  A BGZF
// file (bgzip, bcl-convert, samtools) is a chain of independent gzip members of at most 64 KB whose compressed size sits
// in the header and whose inflated size and CRC-32 sit in the trailer, so the host only has to walk the headers: the
// compressed bytes go over PCIe as they are (a fifth of the text) and every member is inflated by one warp, straight
// into its place in the text block the FASTQ parser (textparse.cuh) indexes next.
//
// DEFLATE (RFC 1951) is a serial bit stream, so one lane of the warp walks it: it reads the block headers, decodes the
// literal / length and distance symbols through lookup tables in shared memory, stores literal bytes itself and queues
// the matches.  The whole warp then copies the queued matches: consecutive matches whose sources all lie below the
// first destination byte of the group form a "run" that is copied in one step, a lane per byte, so the loads of
// a run are independent of its stores.  After the last block the lanes compute the CRC-32 of 1/32 of the member each
// and lane 0 joins the parts (multiplication by x^(8 n) modulo the CRC polynomial), as the gzip trailer demands.
//
// The functions below are the phases of one warp working on one member, written so that tests/emul can run them on the
// host (phase by phase, lane by lane) against zlib in the CPU tier.
#ifndef PCL_INFLATE_CORE_H
#define PCL_INFLATE_CORE_H

#include <stdint.h>
#include <string.h>

#ifdef __CUDACC__
#define INF_HD __host__ __device__ __forceinline__
#define INF_HD_NOINLINE __host__ __device__ __noinline__
#else
#define INF_HD inline
#define INF_HD_NOINLINE inline
#endif

#define INF_LANES 32u
#define INF_FAST_BITS 10u                         // literal / length codes up to this length are one table lookup
#define INF_DFAST_BITS 8u                         // distance codes
#define INF_QUEUE 32u                             // matches queued between two copy steps
#define INF_PAD_BYTES 16u                         // readable bytes the caller keeps behind the compressed data
#define INF_MAX_MEMBER 65536u                     // BGZF: a member inflates to at most 64 KB

// what the host finds in a BGZF member's header and trailer
struct InfMember {
    uint32_t src_off;            // first byte of the DEFLATE stream in the compressed buffer
    uint32_t clen;               // its length
    uint32_t dst_off;            // where the member's bytes go in the text block
    uint32_t isize;              // ISIZE of the trailer
    uint32_t crc;                // CRC32 of the trailer
};

enum {
    INF_OK = 0,
    INF_ERR_BTYPE = 1,           // reserved block type
    INF_ERR_STORED = 2,          // LEN / NLEN of a stored block disagree, or the block leaves the input
    INF_ERR_HEADER = 3,          // bad counts in a dynamic block header
    INF_ERR_LENS = 4,            // bad code length sequence / over-subscribed code / no end-of-block code
    INF_ERR_SYMBOL = 5,          // a bit pattern that is no code, or a reserved symbol
    INF_ERR_DIST = 6,            // a match reaches before the start of the member
    INF_ERR_OUTPUT = 7,          // more bytes than ISIZE says
    INF_ERR_INPUT = 8,           // the stream runs past the member's compressed bytes
    INF_ERR_SIZE = 9,            // fewer bytes than ISIZE says
    INF_ERR_CRC = 10,
};

// status of the decoding lane after a phase
enum { INF_ST_BLOCK_END = 0, INF_ST_QUEUE_FULL = 1, INF_ST_STORED = 2, INF_ST_CODES = 3, INF_ST_ERROR = 4 };

struct InfWarp {
    uint16_t lut[1u << INF_FAST_BITS];            // literal / length: symbol | bits << 9; 0 = a longer code (or none)
    uint16_t dlut[1u << INF_DFAST_BITS];          // distance
    uint16_t code[288 + 32];                      // canonical code of every symbol (filled by lane 0, used by the table fill)
    uint16_t sorted[288 + 32];                    // symbols in canonical order, for codes longer than the tables
    uint16_t cnt[2][16];                          // codes per length
    uint8_t lens[288 + 32];                       // code lengths: literal / length symbols, then distance symbols
    uint8_t cl_lut[128];                          // code length code: symbol | bits << 5
    // the queue of matches and its runs
    uint32_t qpos[INF_QUEUE];                     // destination of the match inside the member
    uint32_t qpre[INF_QUEUE];                     // bytes of the run before this match
    uint16_t qlen[INF_QUEUE], qdist[INF_QUEUE];
    uint32_t rtotal[INF_QUEUE];                   // bytes of run r
    uint8_t rstart[INF_QUEUE + 4];                // first match of run r; rstart[n_runs] = n_matches
    uint32_t n_matches, n_runs;
    // the decoding lane's registers between phases
    uint64_t bitbuf;
    uint32_t bitcnt, in_pos, out_pos, next_word;
    uint32_t status, error, bfinal;
    uint32_t st_src, st_dst, st_len;              // stored block: the copy the warp makes
    uint32_t n_lit, n_dist;                       // symbols of the current block's two codes
    uint32_t pcrc[INF_LANES];                     // CRC-32 of the lanes' slices
};

// ---------------------------------------------------------------------------------------------------------------
// bit input: a 64-bit buffer refilled 32 bits at a time
// ---------------------------------------------------------------------------------------------------------------
INF_HD uint32_t inf_load32(const uint8_t *base, uint32_t pos) {
#ifdef __CUDA_ARCH__
    const uint32_t *p = reinterpret_cast<const uint32_t *>(base + (pos & ~3u));       // base is 4-byte aligned
    return __funnelshift_r(p[0], p[1], (pos & 3u) * 8u);
#else
    uint32_t v;
    memcpy(&v, base + pos, 4);
    return v;
#endif
}

struct InfBits {
    uint64_t bb;
    uint32_t bc, ip;
    const uint8_t *in;
    uint32_t lim;                // last position a refill may read from (8 bytes behind the member's data)
    uint32_t over;               // a refill was asked for behind it: zeros were fed instead, the stream is truncated
    uint32_t nw;                 // the 32 bits at ip, loaded one refill ahead so that the load's latency is not waited for
    INF_HD void prime() { nw = ip <= lim ? inf_load32(in, ip) : 0u; }
    INF_HD void refill() {
        if (bc <= 32u) {
            if (ip > lim) over = 1u;
            bb |= (uint64_t)nw << bc;
            ip += 4u;
            bc += 32u;
            prime();
        }
    }
    INF_HD uint32_t peek(uint32_t n) const { return (uint32_t)bb & ((1u << n) - 1u); }
    INF_HD void drop(uint32_t n) { bb >>= n; bc -= n; }
    INF_HD uint32_t take(uint32_t n) { const uint32_t v = peek(n); drop(n); return v; }
    INF_HD uint32_t byte_pos() const { return ip - (bc >> 3); }      // next unread byte (whole buffered bytes given back)
};

INF_HD uint32_t inf_bitrev(uint32_t v, uint32_t n) {
#ifdef __CUDA_ARCH__
    return __brev(v) >> (32u - n);
#else
    uint32_t r = 0;
    for (uint32_t i = 0; i < n; ++i) r |= ((v >> i) & 1u) << (n - 1u - i);
    return r;
#endif
}

// ---------------------------------------------------------------------------------------------------------------
// code tables
// ---------------------------------------------------------------------------------------------------------------
// lens[base .. base+n): code lengths.  Fills cnt[which], code[] and sorted[] (one lane).  false: over-subscribed.
INF_HD_NOINLINE bool inf_canonical(InfWarp &w, uint32_t which, uint32_t base, uint32_t n) {
    uint16_t *cnt = w.cnt[which];
    for (uint32_t l = 0; l < 16u; ++l) cnt[l] = 0;
    for (uint32_t s = 0; s < n; ++s) cnt[w.lens[base + s]] += 1;
    cnt[0] = 0;
    int32_t left = 1;
    for (uint32_t l = 1; l < 16u; ++l) {
        left = (left << 1) - (int32_t)cnt[l];
        if (left < 0) return false;
    }
    uint16_t next[16], offs[16];
    uint32_t c = 0, o = 0;
    next[0] = 0; offs[0] = 0;
    for (uint32_t l = 1; l < 16u; ++l) {
        c = (c + cnt[l - 1]) << 1;
        next[l] = (uint16_t)c;
        offs[l] = (uint16_t)o;
        o += cnt[l];
    }
    for (uint32_t s = 0; s < n; ++s) {
        const uint32_t l = w.lens[base + s];
        if (l) {
            w.code[base + s] = next[l]++;
            w.sorted[base + offs[l]++] = (uint16_t)s;
        }
    }
    return true;
}

// one lane's share of a table: entries of the symbols lane, lane + 32, ...
INF_HD void inf_fill_lut(const InfWarp &w, uint16_t *lut, uint32_t fast_bits, uint32_t base, uint32_t n, uint32_t lane) {
    for (uint32_t s = lane; s < n; s += INF_LANES) {
        const uint32_t l = w.lens[base + s];
        if (l == 0 || l > fast_bits) continue;
        const uint32_t entry = s | (l << 9);
        for (uint32_t e = inf_bitrev(w.code[base + s], l); e < (1u << fast_bits); e += 1u << l) lut[e] = (uint16_t)entry;
    }
}

// a code longer than the table (or no code at all): canonical decoding bit by bit.  symbol | bits << 9, or 0xFFFFFFFF
INF_HD_NOINLINE uint32_t inf_slow(const InfWarp &w, uint32_t which, uint32_t base, uint64_t bb) {
    int32_t code = 0, first = 0, index = 0;
    for (uint32_t l = 1; l < 16u; ++l) {
        code |= (int32_t)((bb >> (l - 1u)) & 1u);
        const int32_t count = (int32_t)w.cnt[which][l];
        if (code - count < first) return (uint32_t)w.sorted[base + (uint32_t)(index + (code - first))] | (l << 9);
        index += count;
        first += count;
        first <<= 1;
        code <<= 1;
    }
    return 0xFFFFFFFFu;
}

// order in which a dynamic header lists the lengths of the code length code: 16 17 18 0 8 7 9 6 10 5 11 4 12 3 13 2 14 1 15
INF_HD uint32_t inf_cl_order(uint32_t i) {
    const uint64_t a = 16ull | 17ull << 5 | 18ull << 10 | 0ull << 15 | 8ull << 20 | 7ull << 25 | 9ull << 30 | 6ull << 35 |
                       10ull << 40 | 5ull << 45 | 11ull << 50 | 4ull << 55;
    const uint64_t b = 12ull | 3ull << 5 | 13ull << 10 | 2ull << 15 | 14ull << 20 | 1ull << 25 | 15ull << 30;
    return (uint32_t)((i < 12u ? a >> (5u * i) : b >> (5u * (i - 12u))) & 31u);
}

// ---------------------------------------------------------------------------------------------------------------
// phase: block header (one lane).  Leaves status = STORED (the warp copies st_len bytes), CODES (the warp fills the
// tables from lens[]) or ERROR.
// ---------------------------------------------------------------------------------------------------------------
INF_HD_NOINLINE void inf_block_header(InfWarp &w, const uint8_t *in, uint32_t in_end, uint32_t out_cap) {
    InfBits b{w.bitbuf, w.bitcnt, w.in_pos, in, in_end + 8u, 0u, w.next_word};
    uint32_t err = 0, status = INF_ST_CODES;
    b.refill();
    w.bfinal = b.take(1);
    const uint32_t btype = b.take(2);
    if (btype == 3u) {
        err = INF_ERR_BTYPE;
    } else if (btype == 0u) {
        b.drop(b.bc & 7u);
        b.refill();
        const uint32_t len = b.take(16), nlen = b.take(16);
        const uint32_t src = b.byte_pos();
        if (len != (~nlen & 0xFFFFu) || src + len > in_end) err = INF_ERR_STORED;
        else if (w.out_pos + len > out_cap) err = INF_ERR_OUTPUT;
        else {
            w.st_src = src; w.st_dst = w.out_pos; w.st_len = len;
            w.out_pos += len;
            b.bb = 0; b.bc = 0; b.ip = src + len;
            b.prime();
            status = INF_ST_STORED;
        }
    } else if (btype == 1u) {
        for (uint32_t s = 0; s < 288u; ++s) w.lens[s] = (uint8_t)(s < 144u ? 8 : s < 256u ? 9 : s < 280u ? 7 : 8);
        for (uint32_t s = 0; s < 32u; ++s) w.lens[288u + s] = 5;
        w.n_lit = 288u; w.n_dist = 32u;
    } else {
        const uint32_t hlit = b.take(5) + 257u, hdist = b.take(5) + 1u, hclen = b.take(4) + 4u;
        if (hlit > 286u || hdist > 30u) err = INF_ERR_HEADER;
        else {
            // the code length code: 19 symbols of up to 7 bits, decoded through a 128-entry table
            uint8_t cl[19];
            for (uint32_t i = 0; i < 19u; ++i) cl[i] = 0;
            for (uint32_t i = 0; i < hclen; ++i) { b.refill(); cl[inf_cl_order(i)] = (uint8_t)b.take(3); }
            uint32_t cnt[8], next[8];
            for (uint32_t l = 0; l < 8u; ++l) cnt[l] = 0;
            for (uint32_t i = 0; i < 19u; ++i) cnt[cl[i]] += 1;
            cnt[0] = 0;
            int32_t left = 1;
            uint32_t c = 0;
            next[0] = 0;
            for (uint32_t l = 1; l < 8u; ++l) {
                if (left >= 0) left = left * 2 - (int32_t)cnt[l];
                c = (c + cnt[l - 1]) << 1;
                next[l] = c;
            }
            if (left < 0) err = INF_ERR_LENS;
            for (uint32_t e = 0; e < 128u; ++e) w.cl_lut[e] = 0xFF;
            for (uint32_t i = 0; i < 19u && !err; ++i) {
                const uint32_t l = cl[i];
                if (!l) continue;
                for (uint32_t e = inf_bitrev(next[l]++, l); e < 128u; e += 1u << l) w.cl_lut[e] = (uint8_t)(i | (l << 5));
            }
            // the lengths of the two codes, run-length coded
            const uint32_t total = hlit + hdist;
            uint32_t k = 0, prev = 0;
            uint8_t tmp[320];
            while (k < total && !err) {
                b.refill();
                const uint32_t e = w.cl_lut[b.peek(7)];
                if (e == 0xFFu) { err = INF_ERR_LENS; break; }
                b.drop(e >> 5);
                const uint32_t sym = e & 31u;
                if (sym < 16u) { tmp[k++] = (uint8_t)sym; prev = sym; continue; }
                uint32_t rep, val = 0;
                if (sym == 16u) { if (k == 0) { err = INF_ERR_LENS; break; } val = prev; rep = 3u + b.take(2); }
                else if (sym == 17u) rep = 3u + b.take(3);
                else rep = 11u + b.take(7);
                if (k + rep > total) { err = INF_ERR_LENS; break; }
                for (uint32_t r = 0; r < rep; ++r) tmp[k++] = (uint8_t)val;
                prev = val;
            }
            if (!err) {
                for (uint32_t s = 0; s < 288u; ++s) w.lens[s] = s < hlit ? tmp[s] : 0;
                for (uint32_t s = 0; s < 32u; ++s) w.lens[288u + s] = s < hdist ? tmp[hlit + s] : 0;
                w.n_lit = hlit; w.n_dist = hdist;
                if (w.lens[256] == 0) err = INF_ERR_LENS;
            }
        }
    }
    if (!err && status == INF_ST_CODES) {
        if (!inf_canonical(w, 0, 0, w.n_lit) || !inf_canonical(w, 1, 288u, w.n_dist)) err = INF_ERR_LENS;
    }
    if (!err && (b.over || b.byte_pos() > in_end)) err = INF_ERR_INPUT;
    w.bitbuf = b.bb; w.bitcnt = b.bc; w.in_pos = b.ip; w.next_word = b.nw;
    w.error = err;
    w.status = err ? INF_ST_ERROR : status;
}

// ---------------------------------------------------------------------------------------------------------------
// phase: symbols (one lane).  Decodes until the end of the block, an error, or a full match queue; literal bytes are
// stored on the way, matches are queued and grouped into runs.
// ---------------------------------------------------------------------------------------------------------------
INF_HD void inf_decode(InfWarp &w, const uint8_t *in, uint32_t in_end, uint8_t *out, uint32_t out_cap) {
    InfBits b{w.bitbuf, w.bitcnt, w.in_pos, in, in_end + 8u, 0u, w.next_word};
    uint32_t op = w.out_pos, nq = 0, nr = 0, run_pos = 0, pre = 0;
    uint32_t err = 0, status = INF_ST_BLOCK_END;
    for (;;) {
        if (b.over) { err = INF_ERR_INPUT; break; }
        b.refill();
        uint32_t e = w.lut[b.peek(INF_FAST_BITS)];
        if (e == 0) {
            e = inf_slow(w, 0, 0, b.bb);
            if (e == 0xFFFFFFFFu) { err = INF_ERR_SYMBOL; break; }
        }
        b.drop(e >> 9);
        const uint32_t sym = e & 511u;
        if (sym < 256u) {
            if (op >= out_cap) { err = INF_ERR_OUTPUT; break; }
            out[op++] = (uint8_t)sym;
            continue;
        }
        if (sym == 256u) break;
        if (sym > 285u) { err = INF_ERR_SYMBOL; break; }
        const uint32_t c = sym - 257u;
        uint32_t len;
        if (c < 8u) len = 3u + c;
        else if (c == 28u) len = 258u;
        else { const uint32_t x = (c >> 2) - 1u; len = 3u + ((4u + (c & 3u)) << x) + b.take(x); }
        b.refill();
        uint32_t de = w.dlut[b.peek(INF_DFAST_BITS)];
        if (de == 0) {
            de = inf_slow(w, 1, 288u, b.bb);
            if (de == 0xFFFFFFFFu) { err = INF_ERR_SYMBOL; break; }
        }
        b.drop(de >> 9);
        const uint32_t d = de & 511u;
        if (d > 29u) { err = INF_ERR_SYMBOL; break; }
        uint32_t dist;
        if (d < 4u) dist = 1u + d;
        else { const uint32_t x = (d >> 1) - 1u; dist = 1u + ((2u + (d & 1u)) << x) + b.take(x); }
        if (dist > op) { err = INF_ERR_DIST; break; }
        if (op + len > out_cap) { err = INF_ERR_OUTPUT; break; }
        // a match whose source ends below the first destination byte of the current run joins it; any other starts a run
        const uint32_t src_end = op - dist + (len < dist ? len : dist);
        if (nq == 0 || src_end > run_pos) {
            w.rstart[nr] = (uint8_t)nq;
            w.rtotal[nr] = 0;
            ++nr;
            run_pos = op;
            pre = 0;
        }
        w.qpos[nq] = op; w.qpre[nq] = pre; w.qlen[nq] = (uint16_t)len; w.qdist[nq] = (uint16_t)dist;
        pre += len;
        w.rtotal[nr - 1u] = pre;
        op += len;
        if (++nq == INF_QUEUE) { status = INF_ST_QUEUE_FULL; break; }
    }
    if (!err && (b.over || b.byte_pos() > in_end)) err = INF_ERR_INPUT;
    w.rstart[nr] = (uint8_t)nq;
    w.n_matches = nq; w.n_runs = nr;
    w.bitbuf = b.bb; w.bitcnt = b.bc; w.in_pos = b.ip; w.next_word = b.nw; w.out_pos = op;
    w.error = err;
    w.status = err ? INF_ST_ERROR : status;
}

// phase: one lane's bytes of run r (all lanes; the sources of a run lie below its first destination byte)
INF_HD void inf_copy_run(const InfWarp &w, uint32_t r, uint8_t *out, uint32_t lane) {
    const uint32_t lo = w.rstart[r], hi = w.rstart[r + 1u], total = w.rtotal[r];
    for (uint32_t f = lane; f < total; f += INF_LANES) {
        uint32_t a = lo, z = hi;                  // last match of the run with qpre <= f
        while (z - a > 1u) {
            const uint32_t m = (a + z) >> 1;
            if (w.qpre[m] <= f) a = m; else z = m;
        }
        const uint32_t j = f - w.qpre[a], pos = w.qpos[a], dist = w.qdist[a];
        const uint32_t src = pos - dist + (j < dist ? j : j % dist);
        out[pos + j] = out[src];
    }
}

// ---------------------------------------------------------------------------------------------------------------
// CRC-32 (the gzip polynomial, reflected): a slice per lane, joined by lane 0
// ---------------------------------------------------------------------------------------------------------------
#define INF_CRC_POLY 0xEDB88320u

INF_HD uint32_t inf_crc_table_entry(uint32_t t) {
    uint32_t c = t;
    for (int k = 0; k < 8; ++k) c = (c & 1u) ? (c >> 1) ^ INF_CRC_POLY : c >> 1;
    return c;
}

INF_HD uint32_t inf_slice_len(uint32_t isize) { return (isize + INF_LANES - 1u) / INF_LANES; }

INF_HD uint32_t inf_crc_slice(const uint32_t *tab, const uint8_t *out, uint32_t isize, uint32_t lane) {
    const uint32_t L = inf_slice_len(isize);
    const uint32_t lo = lane * L < isize ? lane * L : isize, hi = lo + L < isize ? lo + L : isize;
    uint32_t c = 0xFFFFFFFFu;
    for (uint32_t i = lo; i < hi; ++i) c = tab[(c ^ out[i]) & 0xFFu] ^ (c >> 8);
    return c ^ 0xFFFFFFFFu;
}

// a(x) * b(x) modulo the CRC polynomial, bit-reflected operands (bit 31 is x^0)
INF_HD uint32_t inf_multmodp(uint32_t a, uint32_t b) {
    uint32_t p = 0;
    for (uint32_t m = 0x80000000u; m; m >>= 1) {
        if (a & m) p ^= b;
        b = (b & 1u) ? (b >> 1) ^ INF_CRC_POLY : b >> 1;
    }
    return p;
}

// x^(8 n) modulo the CRC polynomial
INF_HD uint32_t inf_x8n(uint32_t n) {
    uint32_t sq = 0x40000000u;                   // x^1
    for (int k = 0; k < 3; ++k) sq = inf_multmodp(sq, sq);       // x^8
    uint32_t r = 0x80000000u;                    // x^0
    for (; n; n >>= 1) {
        if (n & 1u) r = inf_multmodp(sq, r);
        sq = inf_multmodp(sq, sq);
    }
    return r;
}

// CRC-32 of the member from the lanes' slices (one lane)
INF_HD_NOINLINE uint32_t inf_crc_join(const InfWarp &w, uint32_t isize) {
    const uint32_t L = inf_slice_len(isize);
    if (L == 0) return 0;
    const uint32_t shift_full = inf_x8n(L);
    uint32_t total = w.pcrc[0];
    for (uint32_t lane = 1; lane < INF_LANES; ++lane) {
        const uint32_t lo = lane * L;
        if (lo >= isize) break;
        const uint32_t n = isize - lo < L ? isize - lo : L;
        total = inf_multmodp(n == L ? shift_full : inf_x8n(n), total) ^ w.pcrc[lane];
    }
    return total;
}

// ---------------------------------------------------------------------------------------------------------------
// one member, all phases.  `ex.each(f)` runs f(lane) for the 32 lanes and orders their memory accesses before the
// next phase (the kernel: __syncwarp; the emulation: a loop).  `in` is the compressed buffer (4-byte aligned, with
// INF_PAD_BYTES readable bytes behind it), `out` the member's first byte in the text block.  Returns INF_OK or the error.
// ---------------------------------------------------------------------------------------------------------------
template <class Exec>
INF_HD uint32_t inf_member(Exec &ex, InfWarp &w, const uint32_t *crc_tab, const uint8_t *in, const InfMember &m, uint8_t *out) {
    const uint32_t in_end = m.src_off + m.clen;
    ex.each([&](uint32_t lane) {
        if (lane == 0) {
            w.bitbuf = 0; w.bitcnt = 0; w.in_pos = m.src_off; w.out_pos = 0;
            w.next_word = inf_load32(in, m.src_off);
            w.error = 0; w.bfinal = 0; w.status = INF_ST_BLOCK_END;
        }
    });
    for (;;) {
        ex.each([&](uint32_t lane) { if (lane == 0) inf_block_header(w, in, in_end, m.isize); });
        if (w.status == INF_ST_ERROR) return w.error;
        if (w.status == INF_ST_STORED) {
            ex.each([&](uint32_t lane) {
                for (uint32_t j = lane; j < w.st_len; j += INF_LANES) out[w.st_dst + j] = in[w.st_src + j];
            });
        } else {
            ex.each([&](uint32_t lane) {
                uint32_t *l32 = reinterpret_cast<uint32_t *>(w.lut), *d32 = reinterpret_cast<uint32_t *>(w.dlut);
                for (uint32_t e = lane; e < (1u << INF_FAST_BITS) / 2u; e += INF_LANES) l32[e] = 0;
                for (uint32_t e = lane; e < (1u << INF_DFAST_BITS) / 2u; e += INF_LANES) d32[e] = 0;
            });
            ex.each([&](uint32_t lane) {
                inf_fill_lut(w, w.lut, INF_FAST_BITS, 0, w.n_lit, lane);
                inf_fill_lut(w, w.dlut, INF_DFAST_BITS, 288u, w.n_dist, lane);
            });
            for (;;) {
                ex.each([&](uint32_t lane) { if (lane == 0) inf_decode(w, in, in_end, out, m.isize); });
                const uint32_t n_runs = w.n_runs;
                for (uint32_t r = 0; r < n_runs; ++r) ex.each([&](uint32_t lane) { inf_copy_run(w, r, out, lane); });
                if (w.status == INF_ST_ERROR) return w.error;
                if (w.status == INF_ST_BLOCK_END) break;
            }
        }
        if (w.bfinal) break;
    }
    if (w.out_pos != m.isize) return INF_ERR_SIZE;
    ex.each([&](uint32_t lane) { w.pcrc[lane] = inf_crc_slice(crc_tab, out, m.isize, lane); });
    ex.each([&](uint32_t lane) { if (lane == 0) w.pcrc[0] = inf_crc_join(w, m.isize); });
    return w.pcrc[0] == m.crc ? INF_OK : INF_ERR_CRC;
}

#endif  // PCL_INFLATE_CORE_H
