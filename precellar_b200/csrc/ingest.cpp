// ingest.cpp -- host ingest: FASTQ (plain / gzip, concatenated files) -> SoA planes for the device path.
//
// Replaces, for this path (paths relative to the precellar repository):
//   Read::open + FastqReader::read_record     seqspec/src/read.rs:98-108,137-155 (files of a read are concatenated;
//                                              gzip is detected by its magic, multi-member streams are fine)
//   BatchedFqReader::next / strip_fq_suffix    precellar/src/align/fastq.rs:400-448,612-621 (one record per file per
//                                              group; names must agree after stripping a trailing /1 or /2)
// Records land directly in caller-provided (pinned) planes: `stride` bytes of sequence and quality per record
// plus a u16 length, so a batch can be handed to pcl_chunk_upload without another copy.  Optionally the full
// text of every record (definition, sequence, quality) is kept in an arena for writers (make_fastq).
//
// FASTQ dialect: what noodles-fastq reads -- four lines per record, single-line sequence and quality,
// "@name[ description]"; \r\n tolerated.
#include <dlfcn.h>
#include <zlib.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/precellar_b200.h"

// libzstd has no headers in this image: the few stable entry points are resolved at run time.
namespace {
struct ZBuf { void *p; size_t size, pos; };
struct ZstdApi {
    void *h = nullptr;
    void *(*createDStream)() = nullptr;
    size_t (*freeDStream)(void *) = nullptr;
    size_t (*initDStream)(void *) = nullptr;
    size_t (*decompressStream)(void *, ZBuf *, ZBuf *) = nullptr;
    unsigned (*isError)(size_t) = nullptr;
    bool ok = false;
};
ZstdApi &zstd() {
    static ZstdApi a;
    static bool tried = false;
    if (tried) return a;
    tried = true;
    a.h = dlopen("libzstd.so.1", RTLD_NOW);
    if (!a.h) return a;
    a.createDStream = (void *(*)())dlsym(a.h, "ZSTD_createDStream");
    a.freeDStream = (size_t(*)(void *))dlsym(a.h, "ZSTD_freeDStream");
    a.initDStream = (size_t(*)(void *))dlsym(a.h, "ZSTD_initDStream");
    a.decompressStream = (size_t(*)(void *, ZBuf *, ZBuf *))dlsym(a.h, "ZSTD_decompressStream");
    a.isError = (unsigned (*)(size_t))dlsym(a.h, "ZSTD_isError");
    a.ok = a.createDStream && a.freeDStream && a.initDStream && a.decompressStream && a.isError;
    return a;
}
}  // namespace

struct pcl_fastq {
    std::vector<std::string> paths;
    size_t cur = 0;
    gzFile gz = nullptr;                // uncompressed input (zlib's transparent mode)
    pcl_gzsrc *gs = nullptr;            // gzip input: gzsrc.cpp (BGZF members in parallel, any other gzip through fastinflate.h)
    FILE *zf = nullptr;                 // .zst input
    void *zds = nullptr;
    std::vector<uint8_t> zin;
    size_t zin_pos = 0, zin_end = 0;
    std::vector<uint8_t> buf;
    size_t pos = 0, end = 0;
    bool eof_all = false;
    std::string err;
    uint64_t n_records = 0;
    std::string line[4];
    bool text_full = false;             // the last pcl_fastq_read stopped because the caller's text arena was full
    bool pending = false;               // ... with the record in line[0..3] read but not delivered yet

    static bool is_zst(const std::string &p) { return p.size() > 4 && p.compare(p.size() - 4, 4, ".zst") == 0; }
    void close_cur() {
        if (gz) { gzclose(gz); gz = nullptr; }
        if (gs) { pcl_gzsrc_close(gs); gs = nullptr; }
        if (zf) { fclose(zf); zf = nullptr; }
        if (zds) { zstd().freeDStream(zds); zds = nullptr; }
    }
    bool open_next() {
        close_cur();
        if (cur >= paths.size()) { eof_all = true; return false; }
        const std::string &path = paths[cur];
        // open_file (seqspec/src/utils.rs:44-56): gzip by magic, zstd by the .zst extension, else plain
        bool gzip_magic = false;
        if (FILE *f = fopen(path.c_str(), "rb")) {
            unsigned char m[2] = {0, 0};
            gzip_magic = fread(m, 1, 2, f) == 2 && m[0] == 0x1f && m[1] == 0x8b;
            fclose(f);
        }
        if (!gzip_magic && is_zst(path)) {
            if (!zstd().ok) { err = "libzstd.so.1 is not available: cannot read " + path; eof_all = true; return false; }
            zf = fopen(path.c_str(), "rb");
            if (!zf) { err = "cannot open file: " + path; eof_all = true; return false; }
            zds = zstd().createDStream();
            zstd().initDStream(zds);
            zin.resize(1u << 20);
            zin_pos = zin_end = 0;
        } else if (gzip_magic) {
            if (pcl_gzsrc_open(path.c_str(), 4, &gs) != PCL_OK) { err = "cannot open file: " + path; eof_all = true; return false; }
        } else {
            gz = gzopen(path.c_str(), "rb");       // transparent for uncompressed files
            if (!gz) { err = "cannot open file: " + path; eof_all = true; return false; }
            gzbuffer(gz, 1u << 20);
        }
        ++cur;
        return true;
    }
    bool fill() {
        for (;;) {
            if (!gz && !gs && !zf && !open_next()) return false;
            if (zf) {
                if (zin_pos == zin_end) {
                    zin_end = fread(zin.data(), 1, zin.size(), zf);
                    zin_pos = 0;
                    if (zin_end == 0) { close_cur(); continue; }              // next file of the same read
                }
                ZBuf in{zin.data(), zin_end, zin_pos}, out{buf.data(), buf.size(), 0};
                size_t rc = zstd().decompressStream(zds, &out, &in);
                if (zstd().isError(rc)) { err = "zstd error in " + paths[cur - 1]; eof_all = true; return false; }
                zin_pos = in.pos;
                if (out.pos == 0) continue;
                pos = 0; end = out.pos;
                return true;
            }
            if (gs) {
                const int64_t n = pcl_gzsrc_read(gs, buf.data(), buf.size());
                if (n < 0) { err = pcl_gzsrc_error(gs); eof_all = true; return false; }
                if (n == 0) { close_cur(); continue; }
                pos = 0; end = (size_t)n;
                return true;
            }
            int n = gzread(gz, buf.data(), (unsigned)buf.size());
            if (n < 0) { int e; err = std::string("gzip error: ") + gzerror(gz, &e); eof_all = true; return false; }
            if (n == 0) { close_cur(); continue; }    // next file of the same read (read.rs:148-149)
            pos = 0; end = (size_t)n;
            return true;
        }
    }
    // one line without its terminator; false at end of input
    bool getline(std::string &out) {
        out.clear();
        bool any = false;
        for (;;) {
            if (pos == end && !fill()) return any;
            any = true;
            const uint8_t *p = buf.data() + pos;
            const uint8_t *nl = (const uint8_t *)memchr(p, '\n', end - pos);
            if (nl) {
                out.append((const char *)p, nl - p);
                pos = (nl - buf.data()) + 1;
                if (!out.empty() && out.back() == '\r') out.pop_back();
                return true;
            }
            out.append((const char *)p, end - pos);
            pos = end;
        }
    }
};

static inline uint64_t fnv1a(const char *p, size_t n) {
    uint64_t h = 1469598103934665603ull;
    for (size_t i = 0; i < n; ++i) { h ^= (uint8_t)p[i]; h *= 1099511628211ull; }
    return h;
}

extern "C" {

int pcl_fastq_open(const char *const *paths, int n_paths, pcl_fastq **out) {
    if (!out || !paths || n_paths < 1) return PCL_ERR_ARG;
    pcl_fastq *r = new pcl_fastq();
    for (int i = 0; i < n_paths; ++i) r->paths.push_back(paths[i]);
    r->buf.resize(4u << 20);
    // fail early on an unreadable first file, like File::open(..).unwrap()
    FILE *f = fopen(paths[0], "rb");
    if (!f) { delete r; *out = nullptr; return PCL_ERR_INPUT; }
    fclose(f);
    *out = r;
    return PCL_OK;
}

void pcl_fastq_close(pcl_fastq *r) {
    if (!r) return;
    r->close_cur();
    delete r;
}

const char *pcl_fastq_error(const pcl_fastq *r) { return r ? r->err.c_str() : ""; }

// Read up to max_records records.
//   seq/qual : [max_records][stride] planes (may be NULL when the read needs lengths only); bytes beyond a
//              record's length are filled with 'N' / '!'.
//   len      : [max_records] sequence length (before any Read.max_len truncation; records beyond 65535 bases are an error)
//   name_hash: [max_records] FNV-1a of the read name (definition up to the first blank, trailing /1 or /2 removed)
//   text/off : optional arena keeping definition, sequence and quality of every record:
//              off[3*i], off[3*i+1], off[3*i+2], off[3*i+3] delimit them; *text_used is updated.
// Returns the number of records read (0 at end of input) or a negative pcl_status.
int64_t pcl_fastq_read(pcl_fastq *r, uint64_t max_records, uint32_t stride, uint32_t qual_offset, uint32_t qual_stride,
                       uint8_t *seq, uint8_t *qual, uint16_t *len, uint64_t *name_hash, uint8_t *text, uint64_t text_cap,
                       uint64_t *off, uint64_t *text_used) {
    if (!r || !len) return PCL_ERR_ARG;
    uint64_t n = 0, used = text_used ? *text_used : 0;
    if (off && n == 0) off[0] = used;
    r->text_full = false;
    while (n < max_records) {
        // Fast route: the four lines of the record lie inside the decode buffer -> no per-line string copies.
        // (Blank lines between records, "\r\n" and records that straddle a refill take the general route below.)
        if (!r->pending && r->pos < r->end && r->buf[r->pos] == '@') {
            const uint8_t *p0 = r->buf.data() + r->pos, *pe = r->buf.data() + r->end;
            const uint8_t *e0 = (const uint8_t *)memchr(p0, '\n', pe - p0);
            const uint8_t *e1 = e0 ? (const uint8_t *)memchr(e0 + 1, '\n', pe - (e0 + 1)) : nullptr;
            const uint8_t *e2 = e1 ? (const uint8_t *)memchr(e1 + 1, '\n', pe - (e1 + 1)) : nullptr;
            const uint8_t *e3 = e2 ? (const uint8_t *)memchr(e2 + 1, '\n', pe - (e2 + 1)) : nullptr;
            if (e3 && e0 > p0 && e0[-1] != '\r' && e1[-1] != '\r' && e3[-1] != '\r' && e1 + 1 < e2 && e1[1] == '+' &&
                (size_t)(e1 - (e0 + 1)) == (size_t)(e3 - (e2 + 1))) {
                const uint8_t *sq = e0 + 1, *ql = e2 + 1;
                const size_t L = (size_t)(e1 - sq), dlen = (size_t)(e0 - p0);        // dlen counts the '@'
                if (L > 65535) { r->err = "FASTQ record " + std::to_string(r->n_records) + ": sequences longer than 65535 bases are not supported"; return PCL_ERR_UNSUPPORTED; }
                len[n] = (uint16_t)L;
                if (seq && stride) {
                    const size_t c = L < stride ? L : stride;
                    uint8_t *ds = seq + n * (uint64_t)stride;
                    memcpy(ds, sq, c);
                    if (c < stride) memset(ds + c, 'N', stride - c);
                }
                if (qual && qual_stride) {
                    const size_t avail = L > qual_offset ? L - qual_offset : 0;
                    const size_t c = avail < qual_stride ? avail : qual_stride;
                    uint8_t *dq = qual + n * (uint64_t)qual_stride;
                    if (c) memcpy(dq, ql + qual_offset, c);
                    if (c < qual_stride) memset(dq + c, '!', qual_stride - c);
                }
                size_t e = 1;
                while (e < dlen && p0[e] != ' ' && p0[e] != '\t') ++e;
                size_t name_len = e - 1;
                if (name_len > 2 && p0[e - 2] == '/' && (p0[e - 1] == '1' || p0[e - 1] == '2')) name_len -= 2;   // strip_fq_suffix
                if (name_hash) name_hash[n] = fnv1a((const char *)p0 + 1, name_len);
                if (text && off) {
                    const uint64_t need = (dlen - 1) + 2 * L;
                    if (used + need > text_cap) { r->text_full = true; break; }      // the record stays unread: the caller grows the arena
                    memcpy(text + used, p0 + 1, name_len); used += name_len;
                    memcpy(text + used, p0 + e, dlen - e); used += dlen - e;
                    off[3 * n + 1] = used;
                    memcpy(text + used, sq, L); used += L; off[3 * n + 2] = used;
                    memcpy(text + used, ql, L); used += L; off[3 * n + 3] = used;
                }
                r->pos = (size_t)(e3 - r->buf.data()) + 1;
                ++n;
                ++r->n_records;
                continue;
            }
        }
        if (!r->pending) {
            if (!r->getline(r->line[0])) break;
            if (r->line[0].empty()) continue;                     // stray blank line between records
            if (r->line[0][0] != '@') { r->err = "FASTQ record " + std::to_string(r->n_records) + " does not start with '@'"; return PCL_ERR_INPUT; }
            if (!r->getline(r->line[1]) || !r->getline(r->line[2]) || !r->getline(r->line[3])) {
                r->err = "truncated FASTQ record " + std::to_string(r->n_records);
                return PCL_ERR_INPUT;
            }
        }
        r->pending = false;
        if (r->line[2].empty() || r->line[2][0] != '+') { r->err = "FASTQ record " + std::to_string(r->n_records) + ": missing '+' line"; return PCL_ERR_INPUT; }
        const std::string &def = r->line[0], &s = r->line[1], &q = r->line[3];
        if (s.size() != q.size()) { r->err = "FASTQ record " + std::to_string(r->n_records) + ": sequence and quality lengths differ"; return PCL_ERR_INPUT; }
        const size_t L = s.size();
        if (text && off && used + (def.size() - 1) + 2 * L > text_cap) {      // arena full: the record waits for the next call
            r->text_full = true;
            r->pending = true;
            break;
        }
        if (L > 65535) { r->err = "FASTQ record " + std::to_string(r->n_records) + ": sequences longer than 65535 bases are not supported"; return PCL_ERR_UNSUPPORTED; }
        len[n] = (uint16_t)L;
        if (seq && stride) {
            const size_t c = L < stride ? L : stride;
            uint8_t *ds = seq + n * (uint64_t)stride;
            memcpy(ds, s.data(), c);
            if (c < stride) memset(ds + c, 'N', stride - c);
        }
        if (qual && qual_stride) {
            const size_t avail = L > qual_offset ? L - qual_offset : 0;
            const size_t c = avail < qual_stride ? avail : qual_stride;
            uint8_t *dq = qual + n * (uint64_t)qual_stride;
            if (c) memcpy(dq, q.data() + qual_offset, c);
            if (c < qual_stride) memset(dq + c, '!', qual_stride - c);
        }
        if (name_hash) {
            size_t e = 1;
            while (e < def.size() && def[e] != ' ' && def[e] != '\t') ++e;
            size_t nl = e - 1;                                    // name = def[1..e)
            if (nl > 2 && def[e - 2] == '/' && (def[e - 1] == '1' || def[e - 1] == '2')) nl -= 2;   // strip_fq_suffix
            name_hash[n] = fnv1a(def.data() + 1, nl);
        }
        if (text && off) {
            // BatchedFqReader strips a trailing /1 or /2 from the NAME of every record before anything downstream sees
            // it (strip_fq_suffix, fastq.rs:421,612-621), so AnnotatedFastq and the make_fastq files carry the stripped
            // name; the description stays.
            size_t e = 1;
            while (e < def.size() && def[e] != ' ' && def[e] != '\t') ++e;
            size_t name_len = e - 1;
            if (name_len > 2 && def[e - 2] == '/' && (def[e - 1] == '1' || def[e - 1] == '2')) name_len -= 2;
            memcpy(text + used, def.data() + 1, name_len); used += name_len;
            memcpy(text + used, def.data() + e, def.size() - e); used += def.size() - e;
            off[3 * n + 1] = used;
            memcpy(text + used, s.data(), L); used += L; off[3 * n + 2] = used;
            memcpy(text + used, q.data(), L); used += L; off[3 * n + 3] = used;
        }
        ++n;
        ++r->n_records;
    }
    if (text_used) *text_used = used;
    if (n == 0 && !r->err.empty()) return PCL_ERR_INPUT;
    return (int64_t)n;
}

// 1 when the last pcl_fastq_read returned early because the text arena was full (the next record is still to come:
// call again with a larger arena, `text_used` carried over, and the planes advanced by the records already delivered).
int pcl_fastq_text_full(const pcl_fastq *r) { return r && r->text_full ? 1 : 0; }

}  // extern "C"
