// gzsrc.cpp -- decompressed byte stream of one gzip file for the device text parser (textparse.cuh).
//
// Replaces, for this path: open_file's MultiGzDecoder (seqspec/src/utils.rs:44-56) -- the reference inflates every
// FASTQ file on one thread, twice per run.  Here the stream only has to be turned into bytes (records are cut on the
// GPU), so the inflate is all the host does, and it is parallel where the format allows it:
//   * BGZF (bgzip, and what several demultiplexers write): every <= 64 KB block is an independent gzip member whose
//     compressed size is in its header ('BC' extra field) and whose inflated size is in its trailer, so a batch of
//     blocks is inflated by several threads straight into place in the caller's (pinned) buffer;
//   * any other gzip file (single or multi-member) is one DEFLATE stream per member and inherently sequential: it is
//     decoded by fastinflate.h (1.6-1.7 x zlib's inflate on FASTQ text), header, CRC-32 and ISIZE of every member checked
//     here like MultiGzDecoder does.  PCL_ZLIB_INFLATE=1 (or a zlib-wrapped stream) takes zlib instead.
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/precellar_b200.h"
#include "fastinflate.h"

namespace {
struct Block { uint64_t src_off; uint32_t csize, isize; uint64_t dst_off; };

inline uint32_t rd16(const uint8_t *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8); }
inline uint32_t rd32(const uint8_t *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }

// total size of the BGZF block that starts at p (n bytes available), 0 if p is not a BGZF block header
uint32_t bgzf_block_size(const uint8_t *p, size_t n) {
    if (n < 18 || p[0] != 0x1f || p[1] != 0x8b || p[2] != 8 || !(p[3] & 4)) return 0;
    const uint32_t xlen = rd16(p + 10);
    if (12 + (size_t)xlen > n) return 0;
    for (uint32_t o = 0; o + 4 <= xlen;) {
        const uint8_t *f = p + 12 + o;
        const uint32_t slen = rd16(f + 2);
        if (f[0] == 'B' && f[1] == 'C' && slen == 2 && o + 6 <= xlen) return rd16(f + 4) + 1;
        o += 4 + slen;
    }
    return 0;
}
}  // namespace

struct pcl_gzsrc {
    FILE *f = nullptr;
    std::string path, err;
    int n_threads = 1;
    bool bgzf = false, eof = false;
    // BGZF: compressed bytes not yet consumed
    std::vector<uint8_t> cbuf;
    size_t cpos = 0, cend = 0;
    std::vector<uint8_t> carry;          // BGZF: inflated bytes of a block that did not fit the caller's buffer
    size_t carry_pos = 0;
    // plain gzip: one zlib stream
    z_stream zs;
    bool zs_open = false, mid_member = false;
    std::vector<uint8_t> zin;
    // plain gzip through fastinflate.h
    pcl_fi::Inflater *fi = nullptr;
    enum { GZ_HEADER, GZ_BODY, GZ_TRAILER } gz_state = GZ_HEADER;
    uint32_t crc = 0;
    uint64_t member_bytes = 0, members = 0;
    bool file_eof = false;

    size_t grow_to = 0;
    size_t ccap() const { return cbuf.size() - 64; }     // 64 zero bytes behind the data: the decoder reads ahead
    bool fill_c() {             // keep the unconsumed tail, read more
        if (cpos && cpos < cend) memmove(cbuf.data(), cbuf.data() + cpos, cend - cpos);
        cend -= cpos;
        cpos = 0;
        if (ccap() < grow_to) cbuf.resize(grow_to + 64);
        const size_t got = fread(cbuf.data() + cend, 1, ccap() - cend, f);
        cend += got;
        memset(cbuf.data() + cend, 0, 64);
        return got > 0;
    }
    // the gzip member header at cbuf[cpos..): its length, 0 if more bytes are needed, -1 if it is no gzip header
    long gz_header_len() const {
        const uint8_t *p = cbuf.data() + cpos;
        const size_t n = cend - cpos;
        if (n < 10) return 0;
        if (p[0] != 0x1f || p[1] != 0x8b || p[2] != 8 || (p[3] & 0xE0)) return -1;
        size_t k = 10;
        if (p[3] & 4) {
            if (n < k + 2) return 0;
            k += 2 + (size_t)rd16(p + k);
            if (n < k) return 0;
        }
        for (int flag = 8; flag <= 16; flag <<= 1)                                       // FNAME, FCOMMENT: zero-terminated
            if (p[3] & flag) {
                while (k < n && p[k]) ++k;
                if (k >= n) return 0;
                ++k;
            }
        if (p[3] & 2) k += 2;
        return n < k ? 0 : (long)k;
    }
};

extern "C" {

int pcl_gzsrc_open(const char *path, int n_threads, pcl_gzsrc **out) {
    if (!path || !out) return PCL_ERR_ARG;
    *out = nullptr;
    FILE *f = fopen(path, "rb");
    if (!f) return PCL_ERR_INPUT;
    pcl_gzsrc *s = new pcl_gzsrc();
    s->f = f;
    s->path = path;
    s->n_threads = n_threads < 1 ? 1 : n_threads;
    // the input buffer starts small (a reader that only looks at the first records - infer_length - must not pay for
    // 32 MB of zeroed memory) and takes its working size once the format is known: BGZF batches feed several threads,
    // a single stream only needs to keep fread busy
    const char *bufenv = getenv("PCL_GZSRC_BUF");                                        // tests: many refills
    const size_t forced = (bufenv && *bufenv) ? std::max<size_t>(8192, (size_t)strtoull(bufenv, nullptr, 10)) : 0;
    s->cbuf.resize((forced ? forced : (size_t)(256u << 10)) + 64);
    s->cend = fread(s->cbuf.data(), 1, s->ccap(), f);
    memset(s->cbuf.data() + s->cend, 0, 64);
    s->bgzf = bgzf_block_size(s->cbuf.data(), s->cend) != 0;
    s->grow_to = forced ? forced : (s->bgzf ? (size_t)(32u << 20) : (size_t)(4u << 20));
    const char *force_zlib = getenv("PCL_ZLIB_INFLATE");
    if (!s->bgzf && !(force_zlib && force_zlib[0] == '1') && s->cend >= 3 && s->cbuf[0] == 0x1f && s->cbuf[1] == 0x8b && s->cbuf[2] == 8) {
        s->fi = new pcl_fi::Inflater();
    } else if (!s->bgzf) {
        memset(&s->zs, 0, sizeof s->zs);
        if (inflateInit2(&s->zs, 15 + 32) != Z_OK) { fclose(f); delete s; return PCL_ERR_INPUT; }   // gzip or zlib header
        s->zs_open = true;
        s->zs.next_in = s->cbuf.data();              // the bytes read for the format probe
        s->zs.avail_in = (uInt)s->cend;
    }
    *out = s;
    return PCL_OK;
}

void pcl_gzsrc_close(pcl_gzsrc *s) {
    if (!s) return;
    if (s->zs_open) inflateEnd(&s->zs);
    if (s->f) fclose(s->f);
    delete s->fi;
    delete s;
}

const char *pcl_gzsrc_error(const pcl_gzsrc *s) { return s ? s->err.c_str() : ""; }
int pcl_gzsrc_is_bgzf(const pcl_gzsrc *s) { return s && s->bgzf; }

// Inflate up to `cap` bytes into dst; returns the bytes produced (0 at the end of the file) or a negative pcl_status.
int64_t pcl_gzsrc_read(pcl_gzsrc *s, uint8_t *dst, uint64_t cap) {
    if (!s || (!dst && cap)) return PCL_ERR_ARG;
    if (s->eof || cap == 0) return 0;
    if (s->fi) {
        pcl_fi::Inflater &fi = *s->fi;
        uint64_t produced = 0;
        for (;;) {
            if (fi.pending()) {                                      // finished bytes: copy out, checksum
                const size_t k = std::min<uint64_t>(fi.pending(), cap - produced);
                memcpy(dst + produced, fi.rd, k);
                s->crc = (uint32_t)crc32(s->crc, fi.rd, (uInt)k);
                fi.rd += k;
                produced += k;
                s->member_bytes += k;
                if (produced == cap) break;
            }
            if (fi.window_full()) fi.slide();
            if (s->gz_state == pcl_gzsrc::GZ_HEADER) {
                if (s->cpos == s->cend && !s->file_eof && !s->fill_c()) s->file_eof = true;
                if (s->cpos == s->cend && s->file_eof) {
                    if (s->members == 0) { s->err = "empty gzip file: " + s->path; return PCL_ERR_INPUT; }
                    s->eof = true;
                    break;
                }
                const long h = s->gz_header_len();
                if (h == 0) {
                    if (s->file_eof || !s->fill_c()) { s->err = "truncated gzip stream: " + s->path; return PCL_ERR_INPUT; }
                    continue;
                }
                if (h < 0) { s->err = "gzip error in " + s->path + ": incorrect header check"; return PCL_ERR_INPUT; }
                s->cpos += (size_t)h;
                fi.reset();
                s->crc = 0;
                s->member_bytes = 0;
                s->members += 1;
                s->gz_state = pcl_gzsrc::GZ_BODY;
            } else if (s->gz_state == pcl_gzsrc::GZ_BODY) {
                const uint8_t *in = s->cbuf.data() + s->cpos;
                const pcl_fi::Status st = fi.run(&in, s->cbuf.data() + s->cend, s->file_eof);
                s->cpos = std::min<size_t>((size_t)(in - s->cbuf.data()), s->cend);
                if (st == pcl_fi::FI_ERROR) {
                    s->err = (s->file_eof && s->cpos >= s->cend ? "truncated gzip stream: " : "gzip error in ") + s->path + ": " + fi.msg;
                    return PCL_ERR_INPUT;
                }
                if (st == pcl_fi::FI_NEED_INPUT && !s->fill_c()) s->file_eof = true;
                if (st == pcl_fi::FI_END) s->gz_state = pcl_gzsrc::GZ_TRAILER;
            } else {                                                 // GZ_TRAILER, once every byte of the member is out
                if (fi.pending()) continue;
                if (s->cend - s->cpos < 8) {
                    if (s->file_eof || !s->fill_c()) { s->err = "truncated gzip stream: " + s->path; return PCL_ERR_INPUT; }
                    continue;
                }
                const uint8_t *t = s->cbuf.data() + s->cpos;
                if (rd32(t) != s->crc) { s->err = "gzip error in " + s->path + ": incorrect data check"; return PCL_ERR_INPUT; }
                if (rd32(t + 4) != (uint32_t)s->member_bytes) { s->err = "gzip error in " + s->path + ": incorrect length check"; return PCL_ERR_INPUT; }
                s->cpos += 8;
                s->gz_state = pcl_gzsrc::GZ_HEADER;
            }
        }
        return (int64_t)produced;
    }
    if (!s->bgzf) {
        s->zs.next_out = dst;
        uint64_t produced = 0;
        while (produced < cap) {
            if (s->zs.avail_in == 0) {
                s->cpos = 0;
                s->cend = fread(s->cbuf.data(), 1, s->ccap(), s->f);
                if (s->cend == 0) {
                    if (s->mid_member) { s->err = "truncated gzip stream: " + s->path; return PCL_ERR_INPUT; }
                    s->eof = true;
                    break;
                }
                s->zs.next_in = s->cbuf.data();
                s->zs.avail_in = (uInt)s->cend;
            }
            const uint64_t room = cap - produced;
            s->zs.avail_out = (uInt)(room > (1u << 30) ? (1u << 30) : room);
            const uInt before = s->zs.avail_out;
            const int rc = inflate(&s->zs, Z_NO_FLUSH);
            produced += before - s->zs.avail_out;
            s->mid_member = rc != Z_STREAM_END;
            if (rc == Z_STREAM_END) {                               // next member of a multi-member file, if any
                if (s->zs.avail_in == 0) {
                    s->cend = fread(s->cbuf.data(), 1, s->ccap(), s->f);
                    if (s->cend == 0) { s->eof = true; break; }
                    s->zs.next_in = s->cbuf.data();
                    s->zs.avail_in = (uInt)s->cend;
                }
                if (inflateReset2(&s->zs, 15 + 32) != Z_OK) { s->err = "zlib reset failed: " + s->path; return PCL_ERR_INPUT; }
            } else if (rc != Z_OK && rc != Z_BUF_ERROR) {
                s->err = std::string("gzip error in ") + s->path + ": " + (s->zs.msg ? s->zs.msg : "corrupt stream");
                return PCL_ERR_INPUT;
            } else if (rc == Z_BUF_ERROR && s->zs.avail_in == 0 && feof(s->f)) {
                s->err = "truncated gzip stream: " + s->path;
                return PCL_ERR_INPUT;
            }
        }
        return (int64_t)produced;
    }
    // BGZF: collect whole blocks whose output fits, then inflate them in parallel; a block that straddles the end of
    // the caller's buffer is inflated into `carry` and handed out in two parts, so a read is short only at the end
    uint64_t produced = 0;
    auto inflate_block = [&](const uint8_t *hp, uint32_t csize, uint32_t isize, uint8_t *out, z_stream &z) -> bool {
        const uint32_t hdr = 12 + rd16(hp + 10);
        if (csize < hdr + 8) return false;
        if (isize == 0) return true;                                                    // the empty end-of-file marker block
        inflateReset(&z);
        z.next_in = const_cast<uint8_t *>(hp + hdr);
        z.avail_in = csize - hdr - 8;
        z.next_out = out;
        z.avail_out = isize;
        const int rc = inflate(&z, Z_FINISH);
        return rc == Z_STREAM_END && z.avail_out == 0 && (uint32_t)crc32(0L, out, isize) == rd32(hp + csize - 8);
    };
    while (produced < cap) {
        if (s->carry_pos < s->carry.size()) {
            const size_t n = std::min<size_t>(s->carry.size() - s->carry_pos, cap - produced);
            memcpy(dst + produced, s->carry.data() + s->carry_pos, n);
            s->carry_pos += n;
            produced += n;
            continue;
        }
        if (s->eof) break;
        std::vector<Block> blocks;
        uint64_t out_off = produced;
        size_t p = s->cpos;
        bool straddle = false;
        for (;;) {
            const uint32_t bs = bgzf_block_size(s->cbuf.data() + p, s->cend - p);
            if (bs == 0 || p + bs > s->cend) break;                                     // need more compressed bytes
            const uint32_t isize = rd32(s->cbuf.data() + p + bs - 4);
            if (out_off + isize > cap) { straddle = true; break; }
            blocks.push_back({p, bs, isize, out_off});
            out_off += isize;
            p += bs;
        }
        if (blocks.empty() && !straddle) {
            const size_t before = s->cend - s->cpos;
            if (!s->fill_c()) {
                if (before == 0) { s->eof = true; break; }
                s->err = "truncated or non-BGZF block in " + s->path;
                return PCL_ERR_INPUT;
            }
            continue;
        }
        std::atomic<size_t> next{0};
        std::atomic<int> bad{0};
        const uint8_t *src = s->cbuf.data();
        auto work = [&]() {
            z_stream z;
            memset(&z, 0, sizeof z);
            if (inflateInit2(&z, -15) != Z_OK) { bad = 1; return; }
            for (;;) {
                const size_t k = next.fetch_add(1);
                if (k >= blocks.size()) break;
                const Block &b = blocks[k];
                if (!inflate_block(src + b.src_off, b.csize, b.isize, dst + b.dst_off, z)) { bad = 1; break; }
            }
            inflateEnd(&z);
        };
        if (!blocks.empty()) {
            const int nt = (int)std::min<size_t>((size_t)s->n_threads, blocks.size());
            std::vector<std::thread> th;
            for (int t = 1; t < nt; ++t) th.emplace_back(work);
            work();
            for (auto &t : th) t.join();
            if (bad) { s->err = "corrupt BGZF block in " + s->path; return PCL_ERR_INPUT; }
            produced = out_off;
            s->cpos = p;
        }
        if (straddle) {
            const uint32_t bs = bgzf_block_size(s->cbuf.data() + s->cpos, s->cend - s->cpos);
            const uint32_t isize = rd32(s->cbuf.data() + s->cpos + bs - 4);
            s->carry.resize(isize);
            s->carry_pos = 0;
            z_stream z;
            memset(&z, 0, sizeof z);
            bool ok = inflateInit2(&z, -15) == Z_OK;
            ok = ok && inflate_block(s->cbuf.data() + s->cpos, bs, isize, s->carry.data(), z);
            inflateEnd(&z);
            if (!ok) { s->err = "corrupt BGZF block in " + s->path; return PCL_ERR_INPUT; }
            s->cpos += bs;
        }
    }
    return (int64_t)produced;
}

// The member table of the device inflate (inflate.cuh): whole BGZF members at the front of buf.
int pcl_bgzf_walk(const uint8_t *buf, uint64_t n_bytes, uint64_t want_out, pcl_bgzf_member *members, uint32_t cap,
                  uint32_t *n_members, uint64_t *consumed, uint64_t *out_bytes) {
    if ((!buf && n_bytes) || (!members && cap) || !n_members || !consumed || !out_bytes) return PCL_ERR_ARG;
    uint64_t p = 0, out = 0;
    uint32_t n = 0;
    int rc = PCL_OK;
    while (out < want_out && n < cap && p < n_bytes) {
        const uint64_t left = n_bytes - p;
        const uint32_t bs = bgzf_block_size(buf + p, (size_t)left);
        if (bs == 0) {
            // too few bytes to tell (a header cut off by the end of buf) is "incomplete"; anything else is no BGZF member
            const bool cut = left < 18 || (buf[p] == 0x1f && buf[p + 1] == 0x8b && buf[p + 2] == 8 && (buf[p + 3] & 4) &&
                                           12 + (uint64_t)rd16(buf + p + 10) > left);
            if (!cut) rc = PCL_ERR_INPUT;
            break;
        }
        if (bs > left) break;                                                           // the member is cut off
        const uint32_t hdr = 12 + rd16(buf + p + 10);
        if (bs < hdr + 8) { rc = PCL_ERR_INPUT; break; }
        const uint32_t isize = rd32(buf + p + bs - 4);
        if (isize) {
            if (p + hdr >= (1ull << 32) || out + isize >= (1ull << 32)) break;          // 32-bit offsets: the caller walks again from here
            members[n].src_off = (uint32_t)(p + hdr);
            members[n].clen = bs - hdr - 8;
            members[n].dst_off = (uint32_t)out;
            members[n].isize = isize;
            members[n].crc = rd32(buf + p + bs - 8);
            ++n;
            out += isize;
        }
        p += bs;
    }
    *n_members = n;
    *consumed = p;
    *out_bytes = out;
    return rc;
}

}  // extern "C"
