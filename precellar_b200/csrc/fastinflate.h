// fastinflate.h -- DEFLATE (RFC 1951) decoder of the host byte source for ordinary gzip files (gzsrc.cpp).
//
// SURVEY.md §8(f) row 1.  The reference reads every FASTQ(.gz) file through flate2's MultiGzDecoder on one thread
// (seqspec/src/utils.rs:44-56).  An ordinary gzip file is ONE DEFLATE stream, so it cannot be spread over threads or
// warps the way BGZF members are (inflate.cuh); what is left is making the one thread fast.  Against zlib's inflate this
// decoder keeps a 64-bit bit buffer refilled without a branch, resolves a literal / length code of up to 11 bits with
// one lookup of a packed entry (kind, code length, extra bits, base value), emits up to three literals per refill, goes
// from a literal run into the match behind it without a second lookup, and copies the first 16 bytes of a match without
// a length test (most matches of FASTQ text are shorter; the loop-exit mispredictions were the largest single cost).
// It decodes into a sliding window of its own (32 KB of history + 512 KB) that stays in the L2 cache; the caller copies
// finished bytes out and takes the CRC-32 of them.
//
// Host only.  Header-only so that tests/emul can compile it for the CPU tier (against zlib).
#ifndef PCL_FASTINFLATE_H
#define PCL_FASTINFLATE_H

#include <stdint.h>
#include <string.h>

#include <vector>

namespace pcl_fi {

enum { FI_LBITS = 11, FI_DBITS = 8, FI_HIST = 32768, FI_SPACE = 512 * 1024, FI_SLACK = 320, FI_IN_GUARD = 2048 };
enum { FI_LIT = 0x10u, FI_LEN = 0x20u, FI_EOB = 0x40u };      // entry: code length | kind flag | extra bits << 8 | base value << 16
enum Status { FI_FLUSH = 0, FI_NEED_INPUT = 1, FI_END = 2, FI_ERROR = 3 };

struct Code {                                     // canonical description for codes longer than the table
    uint16_t cnt[16];
    uint16_t sorted[288];
};

inline uint32_t bitrev(uint32_t v, uint32_t n) {
    uint32_t r = 0;
    for (uint32_t i = 0; i < n; ++i) r |= ((v >> i) & 1u) << (n - 1u - i);
    return r;
}

inline uint32_t len_entry(uint32_t sym) {         // literal / length symbol -> entry without the code length
    if (sym < 256u) return FI_LIT | (sym << 16);
    if (sym == 256u) return FI_EOB;
    const uint32_t c = sym - 257u;
    if (c < 8u) return FI_LEN | ((3u + c) << 16);
    if (c == 28u) return FI_LEN | (258u << 16);
    const uint32_t x = (c >> 2) - 1u;
    return FI_LEN | (x << 8) | ((3u + ((4u + (c & 3u)) << x)) << 16);
}

inline uint32_t dist_entry(uint32_t d) {
    if (d < 4u) return FI_LEN | ((1u + d) << 16);
    const uint32_t x = (d >> 1) - 1u;
    return FI_LEN | (x << 8) | ((1u + ((2u + (d & 1u)) << x)) << 16);
}

// lens[n] -> table of 2^bits entries (0 = longer code or none) + canonical description.  false: over-subscribed.
inline bool build(const uint8_t *lens, uint32_t n, uint32_t *table, uint32_t bits, Code &code, bool is_dist) {
    for (int l = 0; l < 16; ++l) code.cnt[l] = 0;
    for (uint32_t s = 0; s < n; ++s) code.cnt[lens[s]] += 1;
    code.cnt[0] = 0;
    int left = 1;
    for (int l = 1; l < 16; ++l) {
        left = left * 2 - (int)code.cnt[l];
        if (left < 0) return false;
    }
    uint32_t next[16], offs[16], c = 0, o = 0;
    next[0] = offs[0] = 0;
    for (int l = 1; l < 16; ++l) {
        c = (c + code.cnt[l - 1]) << 1;
        next[l] = c;
        offs[l] = o;
        o += code.cnt[l];
    }
    memset(table, 0, sizeof(uint32_t) << bits);
    for (uint32_t s = 0; s < n; ++s) {
        const uint32_t l = lens[s];
        if (!l) continue;
        code.sorted[offs[l]++] = (uint16_t)s;
        const uint32_t cd = next[l]++;
        if (l > bits) continue;
        if (is_dist ? s > 29u : s > 285u) continue;                 // reserved symbols: left as "no code" (an error when met)
        const uint32_t e = (is_dist ? dist_entry(s) : len_entry(s)) | l;
        for (uint32_t k = bitrev(cd, l); k < (1u << bits); k += 1u << l) table[k] = e;
    }
    return true;
}

// a code longer than the table: bit by bit through the canonical description.  Returns symbol | length << 16, or ~0u.
inline uint32_t slow(const Code &code, uint64_t bb) {
    int c = 0, first = 0, index = 0;
    for (int l = 1; l < 16; ++l) {
        c |= (int)((bb >> (l - 1)) & 1u);
        const int count = code.cnt[l];
        if (c - count < first) return (uint32_t)code.sorted[index + (c - first)] | ((uint32_t)l << 16);
        index += count;
        first += count;
        first <<= 1;
        c <<= 1;
    }
    return ~0u;
}

struct Inflater {
    uint32_t lt[1u << FI_LBITS], dt[1u << FI_DBITS];
    Code lcode, dcode;
    uint64_t bb = 0;
    uint32_t bc = 0;
    enum { ST_HEADER, ST_STORED, ST_CODES, ST_DONE } state = ST_HEADER;
    bool bfinal = false;
    uint32_t stored_left = 0;
    uint64_t produced = 0;                        // bytes of this stream so far (bounds the match distance)
    std::vector<uint8_t> win;
    uint8_t *out = nullptr;                       // next byte to write
    uint8_t *rd = nullptr;                        // next byte the caller has not taken yet
    const char *msg = "";

    Inflater() : win(FI_HIST + FI_SPACE + FI_SLACK + 64) { out = rd = win.data() + FI_HIST; }

    void reset() {                                // a new DEFLATE stream (the next gzip member); pending bytes stay readable
        bb = 0; bc = 0; state = ST_HEADER; bfinal = false; stored_left = 0; produced = 0;
    }
    size_t pending() const { return (size_t)(out - rd); }
    // all pending bytes were taken: keep the last 32 KB as history and start over behind them
    void slide() {
        uint8_t *base = win.data();
        if (out - base > (ptrdiff_t)FI_HIST) {
            memmove(base, out - FI_HIST, FI_HIST);
            out = rd = base + FI_HIST;
        }
    }
    bool window_full() const { return out > win.data() + FI_HIST + FI_SPACE; }

    // Decode from *pin (at least 64 readable bytes behind in_end; `last`: no more input will come).  Stops with
    // FI_FLUSH (window full: take pending(), slide()), FI_NEED_INPUT (fewer than FI_IN_GUARD bytes left and !last),
    // FI_END (final block done; *pin = first byte behind the stream) or FI_ERROR (msg).
    Status run(const uint8_t **pin, const uint8_t *in_end, bool last) {
        const uint8_t *in = *pin;
        uint64_t b = bb;
        uint32_t n = bc;
        uint8_t *o = out;
        uint8_t *const o_lim = win.data() + FI_HIST + FI_SPACE;
        // !last: stop while FI_IN_GUARD bytes are left (a block header can take several hundred); last: run into the padding
        const uint8_t *const in_lim = last ? in_end + 8 : ((size_t)(in_end - in) > (size_t)FI_IN_GUARD ? in_end - FI_IN_GUARD : in - 1);
        Status st = FI_ERROR;
#define FI_REFILL() do { uint64_t v_; memcpy(&v_, in, 8); b |= v_ << n; in += (63u - n) >> 3; n |= 56u; } while (0)
#define FI_TAKE(k) do { b >>= (k); n -= (k); } while (0)
#define FI_FAIL(m) do { msg = (m); st = FI_ERROR; goto done; } while (0)
        for (;;) {
            if (state == ST_DONE) { st = FI_END; goto done; }
            if (o > o_lim) { st = FI_FLUSH; goto done; }
            if (in > in_lim) {
                if (!last) { st = FI_NEED_INPUT; goto done; }
                FI_FAIL("compressed data ends early");
            }
            if (state == ST_HEADER) {
                FI_REFILL();
                bfinal = b & 1u;
                const uint32_t type = (uint32_t)(b >> 1) & 3u;
                FI_TAKE(3);
                if (type == 3u) FI_FAIL("reserved block type");
                if (type == 0u) {
                    FI_TAKE(n & 7u);
                    if (n < 32u) FI_REFILL();
                    const uint32_t len = (uint32_t)b & 0xFFFFu, nlen = (uint32_t)(b >> 16) & 0xFFFFu;
                    FI_TAKE(32);
                    if (len != (~nlen & 0xFFFFu)) FI_FAIL("bad stored block");
                    in -= n >> 3;                                   // whole buffered bytes go back to the input
                    b = 0; n = 0;
                    stored_left = len;
                    state = ST_STORED;
                    continue;
                }
                uint8_t lens[320];
                uint32_t hlit = 288u, hdist = 32u;
                if (type == 1u) {
                    for (uint32_t s = 0; s < 288u; ++s) lens[s] = (uint8_t)(s < 144u ? 8 : s < 256u ? 9 : s < 280u ? 7 : 8);
                    for (uint32_t s = 0; s < 32u; ++s) lens[288u + s] = 5;
                } else {
                    hlit = ((uint32_t)b & 31u) + 257u;
                    hdist = ((uint32_t)(b >> 5) & 31u) + 1u;
                    const uint32_t hclen = ((uint32_t)(b >> 10) & 15u) + 4u;
                    FI_TAKE(14);
                    if (hlit > 286u || hdist > 30u) FI_FAIL("bad dynamic block header");
                    static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
                    uint8_t cl[19] = {0};
                    for (uint32_t i = 0; i < hclen; ++i) {
                        if (n < 3u) FI_REFILL();
                        cl[order[i]] = (uint8_t)(b & 7u);
                        FI_TAKE(3);
                    }
                    uint32_t clt[128];
                    Code clc;
                    // the code length code through the same builder: "literals" 0..18 (entries are FI_LIT | symbol << 16)
                    if (!build(cl, 19, clt, 7, clc, false)) FI_FAIL("bad code lengths");
                    const uint32_t total = hlit + hdist;
                    uint32_t k = 0, prev = 0;
                    uint8_t tmp[320];
                    while (k < total) {
                        FI_REFILL();
                        if (in > in_end + 8) FI_FAIL("compressed data ends early");
                        const uint32_t e = clt[b & 127u];
                        if (!e) FI_FAIL("bad code lengths");
                        FI_TAKE(e & 15u);
                        const uint32_t sym = e >> 16;
                        if (sym < 16u) { tmp[k++] = (uint8_t)sym; prev = sym; continue; }
                        uint32_t rep, val = 0;
                        if (sym == 16u) { if (!k) FI_FAIL("bad code lengths"); val = prev; rep = 3u + ((uint32_t)b & 3u); FI_TAKE(2); }
                        else if (sym == 17u) { rep = 3u + ((uint32_t)b & 7u); FI_TAKE(3); }
                        else { rep = 11u + ((uint32_t)b & 127u); FI_TAKE(7); }
                        if (k + rep > total) FI_FAIL("bad code lengths");
                        for (uint32_t r = 0; r < rep; ++r) tmp[k++] = (uint8_t)val;
                        prev = val;
                    }
                    if (tmp[256] == 0) FI_FAIL("bad code lengths");
                    for (uint32_t s = 0; s < 288u; ++s) lens[s] = s < hlit ? tmp[s] : 0;
                    for (uint32_t s = 0; s < 32u; ++s) lens[288u + s] = s < hdist ? tmp[hlit + s] : 0;
                }
                if (!build(lens, hlit, lt, FI_LBITS, lcode, false) || !build(lens + 288, hdist, dt, FI_DBITS, dcode, true))
                    FI_FAIL("bad code lengths");
                state = ST_CODES;
                continue;
            }
            if (state == ST_STORED) {
                size_t k = stored_left;
                if (k > (size_t)(o_lim + FI_SLACK - o)) k = (size_t)(o_lim + FI_SLACK - o);
                if (in + k > in_end) {
                    if (in >= in_end && last) FI_FAIL("compressed data ends early");
                    k = in < in_end ? (size_t)(in_end - in) : 0;
                }
                memcpy(o, in, k);
                o += k; in += k; produced += k;
                stored_left -= (uint32_t)k;
                if (stored_left == 0) state = bfinal ? ST_DONE : ST_HEADER;
                else if (in >= in_end) {
                    if (last) FI_FAIL("compressed data ends early");
                    st = FI_NEED_INPUT;
                    goto done;
                }
                continue;
            }
            // ST_CODES: the symbols of a block
            {
                uint8_t *const o0 = o;
                for (;;) {
                    if (o > o_lim || in > in_lim) break;
                    FI_REFILL();
                    uint32_t e = lt[b & ((1u << FI_LBITS) - 1u)];
                    if (e & FI_LIT) {                               // up to three literals per refill
                        FI_TAKE(e & 15u); *o++ = (uint8_t)(e >> 16);
                        e = lt[b & ((1u << FI_LBITS) - 1u)];
                        if (e & FI_LIT) {
                            FI_TAKE(e & 15u); *o++ = (uint8_t)(e >> 16);
                            e = lt[b & ((1u << FI_LBITS) - 1u)];
                            if (e & FI_LIT) { FI_TAKE(e & 15u); *o++ = (uint8_t)(e >> 16); continue; }
                        }
                        FI_REFILL();                                // e is the entry of the bits that follow: a refill leaves them where they are
                    }
                    if (!e) {
                        const uint32_t r = slow(lcode, b);
                        if (r == ~0u || (r & 0xFFFFu) > 285u) { produced += (uint64_t)(o - o0); FI_FAIL("invalid code"); }
                        e = len_entry(r & 0xFFFFu) | (r >> 16);
                        if (e & FI_LIT) { FI_TAKE(e & 15u); *o++ = (uint8_t)(e >> 16); continue; }
                    }
                    if (e & FI_EOB) { FI_TAKE(e & 15u); state = bfinal ? ST_DONE : ST_HEADER; break; }
                    uint32_t cl = e & 15u, x = (e >> 8) & 15u;
                    const uint32_t len = (e >> 16) + (((uint32_t)b >> cl) & ((1u << x) - 1u));
                    FI_TAKE(cl + x);
                    uint32_t de = dt[b & ((1u << FI_DBITS) - 1u)];
                    if (!de) {
                        const uint32_t r = slow(dcode, b);
                        if (r == ~0u || (r & 0xFFFFu) > 29u) { produced += (uint64_t)(o - o0); FI_FAIL("invalid code"); }
                        de = dist_entry(r & 0xFFFFu) | (r >> 16);
                    }
                    cl = de & 15u; x = (de >> 8) & 15u;
                    const uint32_t dist = (de >> 16) + (((uint32_t)b >> cl) & ((1u << x) - 1u));
                    FI_TAKE(cl + x);
                    if ((uint64_t)dist > produced + (uint64_t)(o - o0)) { produced += (uint64_t)(o - o0); FI_FAIL("match before the start of the stream"); }
                    const uint8_t *s = o - dist;
                    uint8_t *d = o;
                    o += len;
                    if (dist >= 8u) {                               // 16 bytes without a test (most matches are shorter), then the rest
                        uint64_t v;
                        memcpy(&v, s, 8); memcpy(d, &v, 8);
                        memcpy(&v, s + 8, 8); memcpy(d + 8, &v, 8);
                        if (len > 16u) {
                            s += 16; d += 16;
                            do { memcpy(&v, s, 8); memcpy(d, &v, 8); s += 8; d += 8; } while (d < o);
                        }
                    } else if (dist == 1u) {
                        const uint64_t v = 0x0101010101010101ull * *s;
                        memcpy(d, &v, 8); memcpy(d + 8, &v, 8);
                        if (len > 16u) { d += 16; do { memcpy(d, &v, 8); d += 8; } while (d < o); }
                    } else {
                        do { *d++ = *s++; } while (d < o);
                    }
                }
                produced += (uint64_t)(o - o0);
            }
        }
    done:
#undef FI_REFILL
#undef FI_TAKE
#undef FI_FAIL
        if (st == FI_END) {                                          // whole buffered bytes go back to the input
            in -= n >> 3;
            b = 0; n = 0;
        }
        bb = b; bc = n; out = o;
        *pin = in;
        return st;
    }
};

}  // namespace pcl_fi

#endif  // PCL_FASTINFLATE_H
