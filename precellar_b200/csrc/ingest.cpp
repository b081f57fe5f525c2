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
#include <zlib.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/precellar_b200.h"

struct pcl_fastq {
    std::vector<std::string> paths;
    size_t cur = 0;
    gzFile gz = nullptr;
    std::vector<uint8_t> buf;
    size_t pos = 0, end = 0;
    bool eof_all = false;
    std::string err;
    uint64_t n_records = 0;
    std::string line[4];

    bool open_next() {
        if (gz) { gzclose(gz); gz = nullptr; }
        if (cur >= paths.size()) { eof_all = true; return false; }
        gz = gzopen(paths[cur].c_str(), "rb");       // transparent for uncompressed files
        if (!gz) { err = "cannot open file: " + paths[cur]; eof_all = true; return false; }
        gzbuffer(gz, 1u << 20);
        ++cur;
        return true;
    }
    bool fill() {
        for (;;) {
            if (!gz && !open_next()) return false;
            int n = gzread(gz, buf.data(), (unsigned)buf.size());
            if (n < 0) { int e; err = std::string("gzip error: ") + gzerror(gz, &e); eof_all = true; return false; }
            if (n == 0) { gzclose(gz); gz = nullptr; continue; }    // next file of the same read (read.rs:148-149)
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
    if (r->gz) gzclose(r->gz);
    delete r;
}

const char *pcl_fastq_error(const pcl_fastq *r) { return r ? r->err.c_str() : ""; }

// Read up to max_records records.
//   seq/qual : [max_records][stride] planes (may be NULL when the read needs lengths only); bytes beyond a
//              record's length are filled with 'N' / '!'.
//   len      : [max_records] sequence length (before any Read.max_len truncation; clamped to 65535)
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
    while (n < max_records) {
        if (!r->getline(r->line[0])) break;
        if (r->line[0].empty()) continue;                         // stray blank line between records
        if (r->line[0][0] != '@') { r->err = "FASTQ record " + std::to_string(r->n_records) + " does not start with '@'"; return PCL_ERR_INPUT; }
        if (!r->getline(r->line[1]) || !r->getline(r->line[2]) || !r->getline(r->line[3])) {
            r->err = "truncated FASTQ record " + std::to_string(r->n_records);
            return PCL_ERR_INPUT;
        }
        if (r->line[2].empty() || r->line[2][0] != '+') { r->err = "FASTQ record " + std::to_string(r->n_records) + ": missing '+' line"; return PCL_ERR_INPUT; }
        const std::string &def = r->line[0], &s = r->line[1], &q = r->line[3];
        if (s.size() != q.size()) { r->err = "FASTQ record " + std::to_string(r->n_records) + ": sequence and quality lengths differ"; return PCL_ERR_INPUT; }
        const size_t L = s.size();
        len[n] = (uint16_t)(L > 65535 ? 65535 : L);
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
            const uint64_t need = (def.size() - 1) + 2 * L;
            if (used + need > text_cap) {                         // arena full: push the record back is not possible -> error
                r->err = "text arena too small";
                return PCL_ERR_NOMEM;
            }
            memcpy(text + used, def.data() + 1, def.size() - 1); used += def.size() - 1; off[3 * n + 1] = used;
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

}  // extern "C"
