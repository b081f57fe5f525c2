// oracle.cpp -- CPU parity oracle: a literal restatement of precellar's barcode hot path.
//
// TEST INFRASTRUCTURE ONLY (see oracle.h).  Byte strings, std::unordered_map and scalar
// fp64, on purpose: the point is to follow the reference line by line, not to be fast.
// All citations are relative to the precellar repository root.
//
// Parity status: split() pinned by the reference's 11 known-answer vectors and, end to end, by its golden
// output files test{2,3,4}.out.gz; counting and correction are "parity unpinned" (the reference has no
// test for them; a second independent restatement must agree, tests/test_oracle_independent.py).

#include "oracle.h"

#include <dlfcn.h>
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

namespace {

// ---------------------------------------------------------------------------------------
// seqspec region model, reduced to what SegmentInfoElem carries (segment.rs:371-435)
// ---------------------------------------------------------------------------------------
struct Elem {
    uint8_t rtype = ORC_OTHER;
    bool seq_fixed = false;       // SequenceType::Fixed (region.rs:436-439)
    size_t min_len = 0, max_len = 0;
    int region = -1;
    std::string sequence;
    bool is_fixed_len() const { return min_len == max_len; }   // segment.rs:424-426
    bool is_barcode() const { return rtype == ORC_BARCODE; }    // region.rs:342-344
    bool is_umi() const { return rtype == ORC_UMI; }            // region.rs:346-348
    bool is_target() const { return rtype == ORC_TARGET; }      // region.rs:351-353
};

// RegionType Display (region.rs:321-335)
char type_letter(uint8_t rtype) {
    switch (rtype) {
        case ORC_BARCODE: return 'B';
        case ORC_UMI: return 'U';
        case ORC_PRIMER: return 'P';
        case ORC_TARGET: return 'T';
        default: return 'O';
    }
}

// Segment (segment.rs:8-38): slices are kept as (start,len) into the record.
struct Seg {
    bool joined = false;
    int elem = -1;                 // Single: element index
    std::vector<int> members;      // Joined: element indices
    size_t start = 0, len = 0;
};

struct SegInfo {                   // SegmentInfo (segment.rs:93-97)
    std::vector<Elem> elems;
    bool is_reverse = false;
};

bool seg_is_barcode(const SegInfo &info, const Seg &s) { return !s.joined && info.elems[s.elem].is_barcode(); } // segment.rs:71-76
bool seg_is_umi(const SegInfo &info, const Seg &s) { return !s.joined && info.elems[s.elem].is_umi(); }         // segment.rs:78-83
bool seg_contains_target(const SegInfo &info, const Seg &s) {                                                    // segment.rs:85-90
    if (!s.joined) return info.elems[s.elem].is_target();
    for (int m : s.members) if (info.elems[m].is_target()) return true;
    return false;
}

// seqspec::utils::rev_compl (seqspec/src/utils.rs:176-187): upper-case only.
std::string rev_compl_seqspec(const uint8_t *p, size_t n) {
    std::string out(n, '\0');
    for (size_t i = 0; i < n; ++i) {
        uint8_t x = p[n - 1 - i];
        switch (x) {
            case 'A': x = 'T'; break;
            case 'T': x = 'A'; break;
            case 'C': x = 'G'; break;
            case 'G': x = 'C'; break;
            default: break;
        }
        out[i] = (char)x;
    }
    return out;
}

// precellar::utils::rev_compl (precellar/src/utils.rs:138-149): case-insensitive.
std::string rev_compl_precellar(const uint8_t *p, size_t n) {
    std::string out(n, '\0');
    for (size_t i = 0; i < n; ++i) {
        uint8_t x = p[n - 1 - i];
        switch (x) {
            case 'A': case 'a': x = 'T'; break;
            case 'T': case 't': x = 'A'; break;
            case 'C': case 'c': x = 'G'; break;
            case 'G': case 'g': x = 'C'; break;
            default: break;
        }
        out[i] = (char)x;
    }
    return out;
}

// hamming_distance (seqspec/src/utils.rs:273-286); lengths are equal at the call sites.
size_t hamming_distance(const uint8_t *a, const uint8_t *b, size_t n) {
    size_t d = 0;
    for (size_t i = 0; i < n; ++i) d += a[i] != b[i];
    return d;
}

// find_pattern: KMP first exact occurrence (segment.rs:438-483)
bool find_pattern(const uint8_t *hay, size_t n, const uint8_t *needle, size_t m, size_t *pos) {
    if (m == 0 || n == 0 || m > n) return false;
    std::vector<size_t> lps(m, 0);
    size_t len = 0, i = 1;
    while (i < m) {
        if (needle[i] == needle[len]) { len += 1; lps[i] = len; i += 1; }
        else if (len != 0) { len = lps[len - 1]; }
        else { lps[i] = 0; i += 1; }
    }
    i = 0;
    size_t j = 0;
    while (i < n) {
        if (needle[j] == hay[i]) { i += 1; j += 1; }
        if (j == m) { *pos = i - j; return true; }
        else if (i < n && needle[j] != hay[i]) {
            if (j != 0) j = lps[j - 1]; else i += 1;
        }
    }
    return false;
}

// find_best_pattern_match (segment.rs:496-535): returns has_pos, pos, mismatches.
struct Match { bool has_pos; size_t pos; size_t mis; };
Match find_best_pattern_match(const uint8_t *hay, size_t n, const uint8_t *needle, size_t m) {
    if (m == 0 || n == 0 || m > n) return {false, 0, SIZE_MAX};
    size_t p;
    if (find_pattern(hay, n, needle, m, &p)) return {true, p, 0};
    if (m == 1) return {false, 0, 1};
    size_t min_mis = SIZE_MAX, best = 0;
    bool has = false;
    for (size_t i = 0; i + m <= n; ++i) {            // 0..=n.saturating_sub(m), m <= n here
        size_t mis = 0;
        for (size_t j = 0; j < m; ++j) {
            if (hay[i + j] != needle[j]) {
                mis += 1;
                if (mis >= min_mis && min_mis != SIZE_MAX) break;
            }
        }
        if (mis < min_mis) { min_mis = mis; best = i; has = true; }
    }
    return {has, best, min_mis};
}

// SegmentInfo::truncate_max (segment.rs:144-164)
size_t truncate_max(const std::vector<Elem> &elems, size_t len) {
    size_t sum = 0, kept = 0;
    for (const Elem &e : elems) {
        if (sum + e.max_len > len) {
            if (sum + e.min_len <= len) kept += 1;
            break;
        } else {
            sum += e.max_len;
            kept += 1;
        }
    }
    return kept;
}

const int SPLIT_PANIC = -100;  // a Rust slice/usize panic would have happened (never on valid layouts)

// SegmentInfo::split_with_tolerance (segment.rs:180-368)
int split_with_tolerance(const SegInfo &info, const uint8_t *seq, size_t len, double linker_tolerance,
                         double anchor_tolerance, std::vector<Seg> &result, size_t *mis_out) {
    const std::vector<Elem> &elems = info.elems;
    bool panic = false;

    // consume_buf (segment.rs:186-232); the slice is [lo, hi) of the record
    auto consume_buf = [&](std::vector<int> &buf, size_t lo, size_t hi) {
        size_t n1 = result.size();
        while (!buf.empty()) {
            int e = buf.back();
            buf.pop_back();
            if (!elems[e].is_fixed_len()) { buf.push_back(e); break; }
            size_t l = hi - lo;
            if (elems[e].max_len > l) { panic = true; return; }   // `l - elem.max_len()` underflow
            size_t i = l - elems[e].max_len;
            Seg s; s.elem = e; s.start = lo + i; s.len = l - i;
            hi = lo + i;
            result.push_back(s);
        }
        if (!buf.empty()) {
            Seg s;
            if (buf.size() == 1) { s.elem = buf[0]; }
            else { s.joined = true; s.members = buf; }
            s.start = lo; s.len = hi - lo;
            buf.clear();
            result.push_back(s);
        }
        size_t n2 = result.size();
        if (n2 > n1) std::reverse(result.begin() + n1, result.begin() + n2);
    };

    result.clear();
    size_t min_offset = 0, max_offset = 0;
    std::vector<int> buffer;

    for (int ei = 0; ei < (int)elems.size(); ++ei) {
        const Elem &segment = elems[ei];
        size_t min_len = segment.min_len, max_len = segment.max_len;

        if (max_offset >= len || min_offset + min_len > len) break;          // segment.rs:245-248

        if (min_len == max_len) {
            if (!buffer.empty()) {
                if (segment.seq_fixed) {                                        // segment.rs:254-311
                    size_t wend = std::min(len, max_offset + max_len);
                    if (min_offset > wend) return SPLIT_PANIC;                  // .get(..).unwrap()
                    const uint8_t *w = seq + min_offset;
                    size_t wn = wend - min_offset;
                    const uint8_t *needle = (const uint8_t *)segment.sequence.data();
                    size_t m = segment.sequence.size();
                    Match r;
                    if (info.is_reverse) {
                        std::string rc = rev_compl_seqspec(w, wn);
                        r = find_best_pattern_match((const uint8_t *)rc.data(), wn, needle, m);
                    } else {
                        r = find_best_pattern_match(w, wn, needle, m);
                    }
                    if (r.has_pos) {
                        if ((double)r.mis <= anchor_tolerance * (double)m) {
                            size_t sum_min = 0;
                            for (int b : buffer) sum_min += elems[b].min_len;
                            size_t offset_left = min_offset - sum_min;
                            size_t offset_right = min_offset + r.pos;
                            consume_buf(buffer, offset_left, offset_right);
                            if (panic) return SPLIT_PANIC;
                            min_offset = offset_right;
                            max_offset = offset_right;
                            if (offset_right + max_len > len) return SPLIT_PANIC;   // .get(..).unwrap()
                            Seg s; s.elem = ei; s.start = offset_right; s.len = max_len;
                            result.push_back(s);
                        } else if (max_offset + max_len > len) {
                            break;                                              // segment.rs:303-305
                        } else {
                            if (mis_out) *mis_out = r.mis;
                            return ORC_SPLIT_PATTERN_MISMATCH;                 // segment.rs:306-308
                        }
                    } else {
                        return ORC_SPLIT_ANCHOR_NOT_FOUND;                      // segment.rs:309-311
                    }
                } else {
                    buffer.push_back(ei);                                       // segment.rs:312-314
                }
            } else {                                                            // segment.rs:315-344
                if (max_offset + max_len > len) return SPLIT_PANIC;             // .get(..).unwrap()
                if (linker_tolerance < 1.0 && segment.seq_fixed) {
                    size_t d;
                    const uint8_t *needle = (const uint8_t *)segment.sequence.data();
                    if (segment.sequence.size() != max_len) return SPLIT_PANIC; // hamming_distance(..).unwrap()
                    if (info.is_reverse) {
                        std::string rc = rev_compl_seqspec(seq + max_offset, max_len);
                        d = hamming_distance((const uint8_t *)rc.data(), needle, max_len);
                    } else {
                        d = hamming_distance(seq + max_offset, needle, max_len);
                    }
                    if ((double)d > linker_tolerance * (double)max_len) {
                        if (mis_out) *mis_out = d;
                        return ORC_SPLIT_PATTERN_MISMATCH;
                    }
                }
                Seg s; s.elem = ei; s.start = max_offset; s.len = max_len;
                result.push_back(s);
            }
        } else {
            buffer.push_back(ei);                                               // segment.rs:345-347
        }
        min_offset += min_len;                                                  // segment.rs:349-350
        max_offset += max_len;
    }

    if (!buffer.empty()) {                                                      // segment.rs:353-365
        size_t sum_min = 0;
        for (int b : buffer) sum_min += elems[b].min_len;
        size_t offset_left = min_offset - sum_min;
        size_t hi = std::min(max_offset, len);
        if (offset_left > hi) return SPLIT_PANIC;
        consume_buf(buffer, offset_left, hi);
        if (panic) return SPLIT_PANIC;
    }
    return ORC_SPLIT_OK;
}

// ---------------------------------------------------------------------------------------
// barcode.rs
// ---------------------------------------------------------------------------------------
const uint8_t BC_MAX_QV = 66;                          // barcode.rs:18
const uint8_t BASE_OPTS[4] = {'A', 'C', 'G', 'T'};     // barcode.rs:19

// error_probability (barcode.rs:464-468): 10f64.powf(-((q as f64 - 33.0) / 10.0)).
// Rust's powf is llvm.pow.f64 -> libm pow on glibc targets.
double error_probability(uint8_t qual) {
    double offset = 33.0;
    return std::pow(10.0, -(((double)qual - offset) / 10.0));
}

struct CountEntry { uint64_t count; uint64_t index; };
typedef std::unordered_map<std::string, CountEntry> OligoFrequency;   // barcode.rs:189

// WhitelistBuilder / Whitelist (barcode.rs:362-460)
struct Whitelist {
    bool whitelist_exists = false;
    OligoFrequency barcode_counts;
    std::vector<const std::string *> by_index;    // file order (IndexSet order, region.rs:459-469)
    uint64_t mismatch_count = 0, total_count = 0;
    bool configured = false;

    void add(const std::string &barcode) {          // barcode.rs:410-426
        if (whitelist_exists) {
            auto it = barcode_counts.find(barcode);
            if (it != barcode_counts.end()) it->second.count += 1;
            else mismatch_count += 1;
        } else {
            bool all_valid = barcode.size() > 1;
            for (char c : barcode) {
                bool ok = false;
                for (uint8_t b : BASE_OPTS) ok |= ((uint8_t)c == b);
                all_valid &= ok;
            }
            if (all_valid) {
                auto it = barcode_counts.find(barcode);
                if (it == barcode_counts.end())
                    it = barcode_counts.emplace(barcode, CountEntry{0, (uint64_t)barcode_counts.size()}).first;
                it->second.count += 1;
            } else {
                mismatch_count += 1;
            }
        }
        total_count += 1;
    }
};

// update_best_option (barcode.rs:345-358): strict '<' -> the first maximum wins.
struct Best { bool some = false; double like = 0.0; const std::string *key = nullptr; uint64_t index = 0; };
void update_best_option(Best &best, double likelihood, const std::string *key, uint64_t index) {
    if (!best.some) { best = {true, likelihood, key, index}; }
    else if (best.like < likelihood) { best = {true, likelihood, key, index}; }
}

// OligoFrequncy::likelihood1 (barcode.rs:311-342).  Returns posterior; *index = matched entry
// (or -1 when the returned key is the query itself without a table entry).
double likelihood1(const OligoFrequency &table, const std::string &query, const uint8_t *qual, size_t nqual,
                   int64_t *index, std::string *key_out) {
    auto it0 = table.find(query);
    if (it0 != table.end()) { *index = (int64_t)it0->second.index; *key_out = query; return 1.0; }

    Best best;
    double total_likelihood = 0.0;
    std::string query_bytes = query;
    for (size_t pos = 0; pos < nqual; ++pos) {            // qual.iter().enumerate()
        uint8_t qv = std::min(qual[pos], BC_MAX_QV);
        if (pos >= query_bytes.size()) break;             // (index panic in Rust; lengths are equal at call sites)
        char existing = query_bytes[pos];
        for (uint8_t val : BASE_OPTS) {
            if ((char)val != existing) {
                query_bytes[pos] = (char)val;
                auto it = table.find(query_bytes);
                if (it != table.end()) {
                    uint64_t bc_count = 1 + it->second.count;
                    double likelihood = (double)bc_count * error_probability(qv);
                    update_best_option(best, likelihood, &it->first, it->second.index);
                    total_likelihood += likelihood;
                }
            }
        }
        query_bytes[pos] = existing;
    }
    if (best.some) { *index = (int64_t)best.index; *key_out = *best.key; return best.like / total_likelihood; }
    *index = -1; *key_out = query; return 0.0;
}

struct CorrectOptions { bool some = false; double max_expected_errors = 0, threshold = 0; };

// BarcodeAnalyzer::correct_barcode (barcode.rs:160-184)
int correct_barcode(const Whitelist &wl, const CorrectOptions &opt, const std::string &barcode, const uint8_t *qual,
                    size_t nqual, int64_t *index, std::string *out) {
    if (opt.some) {
        double expected_errors = 0.0;                       // qual.iter().map(error_probability).sum()
        for (size_t i = 0; i < nqual; ++i) expected_errors += error_probability(qual[i]);
        if (expected_errors >= opt.max_expected_errors) { *index = -1; return ORC_BC_EXCEED_EE; }
        double prob = likelihood1(wl.barcode_counts, barcode, qual, nqual, index, out);
        if (prob <= 0.0) { *index = -1; return ORC_BC_NO_MATCH; }
        else if (prob >= opt.threshold) return ORC_BC_OK;
        else { *index = -1; return ORC_BC_LOW_CONF; }
    } else {
        *out = barcode;
        auto it = wl.barcode_counts.find(barcode);
        *index = it == wl.barcode_counts.end() ? -1 : (int64_t)it->second.index;
        return ORC_BC_OK;
    }
}

// ---------------------------------------------------------------------------------------
// align/fastq.rs : annotate / join / process_chunk ; qc.rs : QcFastq
// ---------------------------------------------------------------------------------------
struct Fq { std::string seq, qual; };                 // the parts of fastq::Record the path touches
struct Barcode { Fq raw; bool corrected_some = false; std::string corrected; };   // fastq.rs:528-545
struct Annotated {                                    // fastq.rs:549-556
    int bc_file = -1, r1_file = -1, r2_file = -1;     // the file whose record (name, description) each part carries
    bool has_barcode = false; Barcode barcode;
    bool has_umi = false; Fq umi;
    bool has_read1 = false; Fq read1;
    bool has_read2 = false; Fq read2;
};

void extend_fastq_record(Fq &a, const Fq &b) { a.seq += b.seq; a.qual += b.qual; }   // fastq.rs:606-610

void barcode_extend(Barcode &self, const Barcode &other) {                            // fastq.rs:535-544
    extend_fastq_record(self.raw, other.raw);
    if (other.corrected_some) {
        if (self.corrected_some) self.corrected += other.corrected;
    } else {
        self.corrected_some = false;
        self.corrected.clear();
    }
}

// AnnotatedFastq::join (fastq.rs:571-603); returns false on the reference's panics.
bool annotated_join(Annotated &self, const Annotated &other) {
    if (self.has_barcode) { if (other.has_barcode) barcode_extend(self.barcode, other.barcode); }
    else { self.has_barcode = other.has_barcode; self.barcode = other.barcode; self.bc_file = other.bc_file; }
    if (self.has_umi) { if (other.has_umi) extend_fastq_record(self.umi, other.umi); }
    else { self.has_umi = other.has_umi; self.umi = other.umi; }
    if (self.has_read1) { if (other.has_read1) return false; }
    else { self.has_read1 = other.has_read1; self.read1 = other.read1; self.r1_file = other.r1_file; }
    if (self.has_read2) { if (other.has_read2) return false; }
    else { self.has_read2 = other.has_read2; self.read2 = other.read2; self.r2_file = other.r2_file; }
    return true;
}

struct QcFastq {                                       // qc.rs:21-27, keyed by small ints instead of strings
    std::vector<uint64_t> num_reads, num_defect;       // per file (read_id)
    uint64_t cat_reads[4] = {0, 0, 0, 0};              // umi, barcode, read1, read2 (only read1/read2 are bumped)
    uint64_t cat_total[4] = {0, 0, 0, 0};
    uint64_t cat_q30[4] = {0, 0, 0, 0};
    explicit QcFastq(size_t n_files = 0) : num_reads(n_files, 0), num_defect(n_files, 0) {}

    static uint64_t q30(const std::string &q) {
        uint64_t c = 0;
        // `**s - 33 >= 30` on u8 (qc.rs:37): release builds wrap for s < 33, so such a byte counts as Q30.
        for (char s : q) c += (uint8_t)((uint8_t)s - 33) >= 30;
        return c;
    }
    void update(const Annotated &fq) {                 // qc.rs:30-69
        if (fq.has_umi) { cat_total[0] += fq.umi.seq.size(); cat_q30[0] += q30(fq.umi.qual); }
        if (fq.has_barcode) { cat_total[1] += fq.barcode.raw.seq.size(); cat_q30[1] += q30(fq.barcode.raw.qual); }
        if (fq.has_read1) { cat_reads[2] += 1; cat_total[2] += fq.read1.seq.size(); cat_q30[2] += q30(fq.read1.qual); }
        if (fq.has_read2) { cat_reads[3] += 1; cat_total[3] += fq.read2.seq.size(); cat_q30[3] += q30(fq.read2.qual); }
    }
    void extend(const QcFastq &o) {                    // qc.rs:72-89
        for (size_t i = 0; i < num_reads.size(); ++i) { num_reads[i] += o.num_reads[i]; num_defect[i] += o.num_defect[i]; }
        for (int i = 0; i < 4; ++i) { cat_reads[i] += o.cat_reads[i]; cat_total[i] += o.cat_total[i]; cat_q30[i] += o.cat_q30[i]; }
    }
};

struct FileInput {
    const uint8_t *seq = nullptr, *qual = nullptr;
    const uint16_t *len = nullptr;
    uint32_t stride = 0;
    uint64_t n = 0;
    bool set = false;
};

struct ReadDef {
    SegInfo info;
    uint32_t reader_max_len = 0;
    bool has_barcode = false;
    bool annotates = false;        // FastqAnnotator::new is Some (fastq.rs:460-471)
    std::vector<int> slot_of_elem; // emit slot k for barcode/UMI elements, else -1
    FileInput in;
};

}  // namespace

struct orc {
    std::vector<ReadDef> reads;
    std::vector<Whitelist> wl;     // indexed by region slot
    std::string err;
    int n_slots = 0;
    std::vector<int> bcseg_of_slot;  // emit slot -> barcode segment j or -1
    int n_bcsegs = 0;
    bool counted = false;
    bool keep_records = false;
    // file-level driver (orc_fastq_load): planes and definitions read from FASTQ files, owned here
    std::vector<std::vector<uint8_t>> own_seq, own_qual;
    std::vector<std::vector<uint16_t>> own_len;
    std::vector<std::vector<std::string>> defs;
    std::vector<Annotated> kept;     // per group, when keep_records
    std::vector<uint8_t> kept_flag;
    QcFastq qc;

    void index_slots() {
        n_slots = 0; n_bcsegs = 0; bcseg_of_slot.clear();
        for (ReadDef &r : reads) {
            r.slot_of_elem.assign(r.info.elems.size(), -1);
            for (size_t e = 0; e < r.info.elems.size(); ++e) {
                const Elem &el = r.info.elems[e];
                if (el.is_barcode() || el.is_umi()) {
                    r.slot_of_elem[e] = n_slots++;
                    bcseg_of_slot.push_back(el.is_barcode() ? n_bcsegs++ : -1);
                }
            }
        }
    }
};

namespace {

// FastqReader::read_record truncation (seqspec/src/read.rs:98-103): the sequence the path sees.
inline size_t record_len(const ReadDef &r, uint64_t i) {
    size_t l = r.in.len[i];
    if (r.reader_max_len > 0) l = std::min(l, (size_t)r.reader_max_len);
    return l;
}

struct RecordView {
    std::string seq, qual;   // owned copies, like noodles' Record (Vec<u8>)
};

inline void load_record(const ReadDef &r, uint64_t i, RecordView &rec) {
    size_t l = record_len(r, i);
    if (r.in.seq) rec.seq.assign((const char *)r.in.seq + (size_t)i * r.in.stride, l);
    else rec.seq.assign(l, 'N');
    if (r.in.qual) rec.qual.assign((const char *)r.in.qual + (size_t)i * r.in.stride, l);
    else rec.qual.assign(l, '!');
}

struct AnnotateOut {
    int status = ORC_SPLIT_OK;
    std::vector<Seg> segs;
    // per barcode segment in encounter order: (elem, err, index)
    struct BcRes { int elem; int err; int64_t index; };
    std::vector<BcRes> bc;
    bool panic_multi_target = false;
};

// FastqAnnotator::annotate (fastq.rs:474-525)
int annotate(const orc &o, const ReadDef &rd, const RecordView &rec, const CorrectOptions &opt, Annotated &out,
             AnnotateOut &aux) {
    const SegInfo &info = rd.info;
    out = Annotated();
    const int this_file = (int)(&rd - o.reads.data());
    aux.bc.clear();
    int st = split_with_tolerance(info, (const uint8_t *)rec.seq.data(), rec.seq.size(), 1.0, 0.2, aux.segs, nullptr);
    aux.status = st;
    if (st != ORC_SPLIT_OK) return st;
    for (const Seg &segment : aux.segs) {
        bool is_bc = seg_is_barcode(info, segment), is_umi = seg_is_umi(info, segment);
        if (is_bc || is_umi) {
            Fq fq;                                                         // segment.into_fq
            fq.seq = rec.seq.substr(segment.start, segment.len);
            fq.qual = rec.qual.substr(segment.start, segment.len);
            if (info.is_reverse) {                                         // rev_compl_fastq_record (utils.rs:132-136)
                std::reverse(fq.qual.begin(), fq.qual.end());
                fq.seq = rev_compl_precellar((const uint8_t *)fq.seq.data(), fq.seq.size());
            }
            if (is_bc) {
                const Elem &el = info.elems[segment.elem];
                Barcode b;
                int64_t index = -1;
                int err = correct_barcode(o.wl[el.region], opt, fq.seq, (const uint8_t *)fq.qual.data(), fq.qual.size(),
                                          &index, &b.corrected);
                b.corrected_some = (err == ORC_BC_OK);                     // .ok().map(to_vec)
                if (!b.corrected_some) b.corrected.clear();
                b.raw = fq;
                aux.bc.push_back({segment.elem, err, b.corrected_some ? index : -1});
                if (out.has_barcode) barcode_extend(out.barcode, b);
                else { out.has_barcode = true; out.barcode = b; out.bc_file = this_file; }
            } else {
                out.has_umi = true;                                        // `umi = Some(fq)`: a later UMI segment of the
                out.umi = fq;                                              // same file REPLACES the earlier one (fastq.rs:502)
            }
        } else if (seg_contains_target(info, segment)) {
            if (out.has_read1 || out.has_read2) { aux.panic_multi_target = true; return SPLIT_PANIC; }
            Fq fq;
            fq.seq = rec.seq.substr(segment.start, segment.len);
            fq.qual = rec.qual.substr(segment.start, segment.len);
            if (info.is_reverse) { out.has_read2 = true; out.read2 = fq; out.r2_file = this_file; }
            else { out.has_read1 = true; out.read1 = fq; out.r1_file = this_file; }
        }
    }
    return ORC_SPLIT_OK;
}

void set_err(orc *o, const std::string &s) { o->err = s; }

}  // namespace

// ---------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------
extern "C" {

orc_t *orc_create(void) { return new orc(); }
void orc_destroy(orc_t *o) { delete o; }
const char *orc_last_error(orc_t *o) { return o->err.c_str(); }

static Elem to_elem(const orc_elem &e) {
    Elem x;
    x.rtype = e.rtype; x.seq_fixed = e.seq_fixed != 0; x.min_len = e.min_len; x.max_len = e.max_len; x.region = e.region;
    if (e.sequence) x.sequence = e.sequence;
    return x;
}

int orc_add_read(orc_t *o, int is_reverse, uint32_t reader_max_len, const orc_elem *elems, int n_elems) {
    ReadDef r;
    r.info.is_reverse = is_reverse != 0;
    r.reader_max_len = reader_max_len;
    for (int i = 0; i < n_elems; ++i) {
        r.info.elems.push_back(to_elem(elems[i]));
        const Elem &el = r.info.elems.back();
        r.has_barcode |= el.is_barcode();
        r.annotates |= el.is_barcode() || el.is_umi() || el.is_target();
        if (el.is_barcode()) {
            if (el.region < 0) { set_err(o, "barcode element without region slot"); return -1; }
            if ((size_t)el.region >= o->wl.size()) o->wl.resize(el.region + 1);
        }
    }
    o->reads.push_back(r);
    o->index_slots();
    return (int)o->reads.size() - 1;
}

// WhitelistBuilder::new (barcode.rs:393-407) fed from Onlist::read (region.rs:459-469)
int orc_set_whitelist(orc_t *o, int region, const uint8_t *bytes, const uint64_t *offsets, uint64_t n) {
    if (region < 0) { set_err(o, "bad region"); return -1; }
    if ((size_t)region >= o->wl.size()) o->wl.resize(region + 1);
    Whitelist &w = o->wl[region];
    w = Whitelist();
    w.configured = true;
    w.barcode_counts.reserve(n * 2);
    for (uint64_t i = 0; i < n; ++i) {
        std::string s((const char *)bytes + offsets[i], offsets[i + 1] - offsets[i]);
        w.barcode_counts.emplace(s, CountEntry{0, (uint64_t)w.barcode_counts.size()});   // IndexSet: first occurrence wins
    }
    w.whitelist_exists = !w.barcode_counts.empty();
    w.by_index.assign(w.barcode_counts.size(), nullptr);
    for (auto &kv : w.barcode_counts) w.by_index[kv.second.index] = &kv.first;
    return 0;
}

int orc_set_input(orc_t *o, int file, const uint8_t *seq, const uint8_t *qual, const uint16_t *len, uint32_t stride,
                  uint64_t n) {
    if (file < 0 || (size_t)file >= o->reads.size()) { set_err(o, "bad file index"); return -1; }
    FileInput &in = o->reads[file].in;
    in.seq = seq; in.qual = qual; in.len = len; in.stride = stride; in.n = n; in.set = true;
    return 0;
}

void orc_set_keep_records(orc_t *o, int keep) { o->keep_records = keep != 0; }

// BarcodeAnalyzer::new (barcode.rs:59-137): serial pass over every barcode-bearing read.
int orc_count(orc_t *o) {
    for (Whitelist &w : o->wl) {
        for (auto &kv : w.barcode_counts) kv.second.count = 0;
        if (!w.whitelist_exists) { w.barcode_counts.clear(); w.by_index.clear(); }
        w.mismatch_count = 0; w.total_count = 0;
    }
    RecordView rec;
    std::vector<Seg> segs;
    for (ReadDef &rd : o->reads) {
        if (!rd.has_barcode) continue;                                  // barcode.rs:81-83
        if (!rd.in.set) { set_err(o, "Failed to open FASTQ file (no input set)"); return -1; }   // barcode.rs:88-90
        if (!rd.in.seq) { set_err(o, "barcode read without sequence bytes"); return -1; }
        for (uint64_t i = 0; i < rd.in.n; ++i) {
            load_record(rd, i, rec);
            int st = split_with_tolerance(rd.info, (const uint8_t *)rec.seq.data(), rec.seq.size(), 1.0, 0.2, segs, nullptr);
            if (st == SPLIT_PANIC) { set_err(o, "split would panic in the reference"); return -1; }
            if (st != ORC_SPLIT_OK) continue;                           // `if let Ok(segments)` (barcode.rs:102)
            for (const Seg &s : segs) {
                if (!seg_is_barcode(rd.info, s)) continue;
                Whitelist &w = o->wl[rd.info.elems[s.elem].region];
                if (rd.info.is_reverse)                                 // barcode.rs:110-114
                    w.add(rev_compl_precellar((const uint8_t *)rec.seq.data() + s.start, s.len));
                else
                    w.add(rec.seq.substr(s.start, s.len));
            }
        }
    }
    // assert all whitelists have the same total (barcode.rs:125-129)
    bool first = true; uint64_t t0 = 0;
    for (Whitelist &w : o->wl) {
        if (!w.configured && w.total_count == 0 && w.barcode_counts.empty()) continue;
        if (first) { t0 = w.total_count; first = false; }
        else if (w.total_count != t0) { set_err(o, "All whitelists should have the same total read count"); return -2; }
    }
    // no-whitelist regions: rebuild index order for reporting
    for (Whitelist &w : o->wl) {
        if (!w.whitelist_exists) {
            w.by_index.assign(w.barcode_counts.size(), nullptr);
            for (auto &kv : w.barcode_counts) w.by_index[kv.second.index] = &kv.first;
        }
    }
    o->counted = true;
    return 0;
}

int orc_get_counts(orc_t *o, int region, uint64_t *counts, uint64_t *mismatch, uint64_t *total) {
    if (region < 0 || (size_t)region >= o->wl.size()) { set_err(o, "bad region"); return -1; }
    Whitelist &w = o->wl[region];
    if (counts) for (auto &kv : w.barcode_counts) counts[kv.second.index] = kv.second.count;
    if (mismatch) *mismatch = w.mismatch_count;
    if (total) *total = w.total_count;
    return 0;
}

int orc_annotate(orc_t *o, int correct, double threshold, double max_expected_errors, int n_threads,
                 uint64_t chunk_bases, orc_out *out) {
    if (!o->counted) { set_err(o, "orc_count must run first"); return -1; }
    // readers: reads with a FastqAnnotator and files (fastq.rs:139-148)
    std::vector<int> files;
    for (size_t f = 0; f < o->reads.size(); ++f)
        if (o->reads[f].annotates && o->reads[f].in.set) files.push_back((int)f);
    if (files.empty()) { set_err(o, "no annotated reads"); return -1; }
    uint64_t n = o->reads[files[0]].in.n;
    for (int f : files)
        if (o->reads[f].in.n != n) { set_err(o, "Unequal number of reads in the chunk"); return -3; }   // fastq.rs:435-436

    CorrectOptions opt;                                   // fastq.rs:131-137
    opt.some = correct != 0; opt.threshold = threshold; opt.max_expected_errors = max_expected_errors;

    const size_t n_files_all = o->reads.size();
    if (out) {
        if (out->fstatus) memset(out->fstatus, 0, n_files_all * n);
        if (out->tcoord) memset(out->tcoord, 0xFF, n_files_all * n * 4);
        if (out->bucoord) memset(out->bucoord, 0xFF, (size_t)o->n_slots * n * 4);
        if (out->bc_index) for (size_t k = 0; k < (size_t)o->n_bcsegs * n; ++k) out->bc_index[k] = -1;
        if (out->bc_err) memset(out->bc_err, ORC_BC_ABSENT, (size_t)o->n_bcsegs * n);
        if (out->g_flags) memset(out->g_flags, 0, n);
    }
    o->qc = QcFastq(n_files_all);
    if (o->keep_records) { o->kept.assign(n, Annotated()); o->kept_flag.assign(n, 0); }

    // BatchedFqReader (fastq.rs:400-448): batches of >= chunk_bases accumulated sequence length.
    std::vector<std::pair<uint64_t, uint64_t>> batches;
    {
        uint64_t i = 0;
        while (i < n) {
            uint64_t b0 = i, acc = 0;
            while (i < n && acc < chunk_bases) {
                for (int f : files) acc += record_len(o->reads[f], i);
                ++i;
            }
            batches.push_back({b0, i});
        }
    }

    std::mutex qc_mutex;
    std::atomic<int> failed{0};
    std::string fail_msg;

    // process_chunk (fastq.rs:353-389) on records [lo, hi)
    auto process_chunk = [&](uint64_t lo, uint64_t hi) {
        QcFastq qc(n_files_all);
        RecordView rec;
        Annotated anno, joined;
        AnnotateOut aux;
        for (uint64_t i = lo; i < hi; ++i) {
            bool have = false;
            for (int f : files) {
                const ReadDef &rd = o->reads[f];
                qc.num_reads[f] += 1;
                load_record(rd, i, rec);
                int st = annotate(*o, rd, rec, opt, anno, aux);
                if (st == SPLIT_PANIC) {
                    std::lock_guard<std::mutex> g(qc_mutex);
                    failed = 1;
                    fail_msg = aux.panic_multi_target ? "Multiple target regions found in one fastq record!"
                                                      : "split would panic in the reference";
                    return;
                }
                if (out && out->fstatus) out->fstatus[(size_t)f * n + i] = (uint8_t)st;
                if (st != ORC_SPLIT_OK) { qc.num_defect[f] += 1; continue; }
                if (out) {
                    for (const Seg &s : aux.segs) {
                        uint32_t c = ((uint32_t)s.start << 16) | (uint32_t)s.len;
                        if (!s.joined && rd.slot_of_elem[s.elem] >= 0) {
                            if (out->bucoord) out->bucoord[(size_t)rd.slot_of_elem[s.elem] * n + i] = c;
                        } else if (seg_contains_target(rd.info, s)) {
                            if (out->tcoord) out->tcoord[(size_t)f * n + i] = c;
                        }
                    }
                    for (const AnnotateOut::BcRes &b : aux.bc) {
                        int j = o->bcseg_of_slot[rd.slot_of_elem[b.elem]];
                        if (out->bc_index) out->bc_index[(size_t)j * n + i] = b.index;
                        if (out->bc_err) out->bc_err[(size_t)j * n + i] = (uint8_t)b.err;
                    }
                }
                if (!have) { joined = anno; have = true; }
                else if (!annotated_join(joined, anno)) {
                    std::lock_guard<std::mutex> g(qc_mutex);
                    failed = 1; fail_msg = "Read1/Read2 already exists";
                    return;
                }
            }
            if (!have) continue;                       // reduce() of an empty iterator -> `?` (fastq.rs:376-379)
            qc.update(joined);                         // fastq.rs:380
            uint8_t flags = 0;
            if (joined.has_barcode) flags |= 1;
            if (joined.has_barcode && joined.barcode.corrected_some) flags |= 2;
            if (joined.has_umi) flags |= 4;
            if (joined.has_read1) flags |= 8;
            if (joined.has_read2) flags |= 16;
            if (out && out->g_flags) out->g_flags[i] = flags;
            if (o->keep_records && joined.has_barcode) { o->kept[i] = joined; o->kept_flag[i] = 1; }   // fastq.rs:381-385
        }
        std::lock_guard<std::mutex> g(qc_mutex);       // self.qc.lock().unwrap().extend(..) (fastq.rs:345)
        o->qc.extend(qc);
    };

    if (n_threads < 1) n_threads = 1;
    for (auto &b : batches) {
        uint64_t nb = b.second - b.first;
        // chunk.par_chunks(n / 128) (fastq.rs:341-348).  n/128 == 0 panics in rayon; the oracle
        // uses 1-record sub-chunks there instead (documented divergence, results are batching-invariant).
        uint64_t sub = std::max<uint64_t>(1, nb / 128);
        uint64_t n_sub = (nb + sub - 1) / sub;
        std::atomic<uint64_t> next{0};
        auto worker = [&]() {
            for (;;) {
                uint64_t s = next.fetch_add(1);
                if (s >= n_sub || failed) break;
                uint64_t lo = b.first + s * sub, hi = std::min(b.second, lo + sub);
                process_chunk(lo, hi);
            }
        };
        if (n_threads == 1) worker();
        else {
            std::vector<std::thread> th;
            for (int t = 0; t < n_threads; ++t) th.emplace_back(worker);
            for (auto &t : th) t.join();
        }
        if (failed) { set_err(o, fail_msg); return -4; }
    }
    return 0;
}

int orc_get_qc(orc_t *o, uint64_t *per_file, uint64_t *cat) {
    for (size_t f = 0; f < o->qc.num_reads.size(); ++f) {
        per_file[2 * f] = o->qc.num_reads[f];
        per_file[2 * f + 1] = o->qc.num_defect[f];
    }
    for (int c = 0; c < 4; ++c) {
        cat[3 * c] = o->qc.cat_reads[c]; cat[3 * c + 1] = o->qc.cat_total[c]; cat[3 * c + 2] = o->qc.cat_q30[c];
    }
    return 0;
}

int64_t orc_render(orc_t *o, uint64_t i, char *buf, uint64_t cap) {
    if (!o->keep_records || i >= o->kept.size()) return -1;
    std::string s;
    if (o->kept_flag[i]) {
        const Annotated &a = o->kept[i];
        s += a.barcode.raw.seq; s += '\t'; s += a.barcode.raw.qual; s += '\t';
        s += a.barcode.corrected_some ? a.barcode.corrected : std::string("*"); s += '\t';
        s += a.has_umi ? a.umi.seq : ""; s += '\t'; s += a.has_umi ? a.umi.qual : ""; s += '\t';
        s += a.has_read1 ? a.read1.seq : ""; s += '\t'; s += a.has_read1 ? a.read1.qual : ""; s += '\t';
        s += a.has_read2 ? a.read2.seq : ""; s += '\t'; s += a.has_read2 ? a.read2.qual : "";
    }
    if (s.size() + 1 > cap) return -2;
    memcpy(buf, s.c_str(), s.size() + 1);
    return (int64_t)s.size();
}

int orc_split_str(const orc_elem *elems, int n_elems, int is_reverse, const uint8_t *seq, const uint8_t *qual,
                  uint32_t len, double linker_tol, double anchor_tol, char *buf, uint32_t cap) {
    (void)qual;
    SegInfo info;
    info.is_reverse = is_reverse != 0;
    for (int i = 0; i < n_elems; ++i) info.elems.push_back(to_elem(elems[i]));
    std::vector<Seg> segs;
    int st = split_with_tolerance(info, seq, len, linker_tol, anchor_tol, segs, nullptr);
    if (st != ORC_SPLIT_OK) return st;
    std::string s;                                           // Display for Segment / SegmentType (segment.rs:16-49)
    for (size_t k = 0; k < segs.size(); ++k) {
        if (k) s += ',';
        s += '[';
        if (segs[k].joined) for (int m : segs[k].members) s += type_letter(info.elems[m].rtype);
        else s += type_letter(info.elems[segs[k].elem].rtype);
        s += ']';
        s.append((const char *)seq + segs[k].start, segs[k].len);
    }
    if (s.size() + 1 > cap) return -2;
    memcpy(buf, s.c_str(), s.size() + 1);
    return 0;
}

int orc_truncate_max(const orc_elem *elems, int n_elems, uint32_t len) {
    std::vector<Elem> v;
    for (int i = 0; i < n_elems; ++i) v.push_back(to_elem(elems[i]));
    return (int)truncate_max(v, len);
}

void orc_rev_compl_seqspec(const uint8_t *in, uint32_t n, uint8_t *out) {
    std::string s = rev_compl_seqspec(in, n);
    memcpy(out, s.data(), n);
}
void orc_rev_compl_precellar(const uint8_t *in, uint32_t n, uint8_t *out) {
    std::string s = rev_compl_precellar(in, n);
    memcpy(out, s.data(), n);
}
double orc_error_probability(uint8_t q) { return error_probability(q); }

int orc_correct_one(const uint8_t *wl_bytes, const uint64_t *wl_offsets, const uint64_t *wl_counts, uint64_t n_wl,
                    const uint8_t *bc, const uint8_t *qual, uint32_t len, double threshold, double max_expected_errors,
                    int64_t *index_out, double *prob_out) {
    Whitelist w;
    for (uint64_t i = 0; i < n_wl; ++i) {
        std::string s((const char *)wl_bytes + wl_offsets[i], wl_offsets[i + 1] - wl_offsets[i]);
        w.barcode_counts.emplace(s, CountEntry{wl_counts ? wl_counts[i] : 0, (uint64_t)w.barcode_counts.size()});
    }
    w.whitelist_exists = !w.barcode_counts.empty();
    std::string query((const char *)bc, len), key;
    if (prob_out) {
        int64_t idx;
        *prob_out = likelihood1(w.barcode_counts, query, qual, len, &idx, &key);
    }
    CorrectOptions opt;
    opt.some = true; opt.threshold = threshold; opt.max_expected_errors = max_expected_errors;
    return correct_barcode(w, opt, query, qual, len, index_out, &key);
}

// ---------------------------------------------------------------------------------------
// File-level driver: the reference's make_fastq end to end on the host (bench.py api_e2e, reference arm).
//   orc_fastq_load   Read::open + FastqReader::read_record (seqspec/src/read.rs:98-108,137-155; open_file,
//                    seqspec/src/utils.rs:44-56) + strip_fq_suffix (precellar/src/align/fastq.rs:612-621)
//   orc_make_fastq   the record loop of make_fastq (python/src/lib.rs:136-168) over the AnnotatedFastq kept by
//                    orc_annotate, written through zstd at `level` with `threads` workers (create_file,
//                    seqspec/src/utils.rs:17-42: level 9, multithread(8))
// FASTQ dialect: four lines per record, definition = name [blank description]; the un-vendored noodles reader's
// corner cases (blank lines, wrapped sequences) are not modelled ("parity unpinned").
// ---------------------------------------------------------------------------------------
struct LineReader {
    gzFile f = nullptr;
    std::vector<char> buf;
    size_t pos = 0, end = 0;
    bool open(const char *p) { f = gzopen(p, "rb"); if (f) { gzbuffer(f, 1 << 20); buf.resize(1 << 20); } return f != nullptr; }
    void close() { if (f) gzclose(f); f = nullptr; }
    bool getline(std::string &out) {
        out.clear();
        for (;;) {
            if (pos == end) {
                int n = gzread(f, buf.data(), (unsigned)buf.size());
                if (n <= 0) return !out.empty();
                pos = 0; end = (size_t)n;
            }
            const char *nl = (const char *)memchr(buf.data() + pos, '\n', end - pos);
            if (nl) {
                out.append(buf.data() + pos, nl - (buf.data() + pos));
                pos = (size_t)(nl - buf.data()) + 1;
                if (!out.empty() && out.back() == '\r') out.pop_back();
                return true;
            }
            out.append(buf.data() + pos, end - pos);
            pos = end;
        }
    }
};

int orc_fastq_load(orc_t *o, int file, const char *const *paths, int n_paths, uint32_t stride) {
    if (file < 0 || (size_t)file >= o->reads.size()) { set_err(o, "bad file index"); return -1; }
    const size_t nf = o->reads.size();
    if (o->own_seq.size() < nf) { o->own_seq.resize(nf); o->own_qual.resize(nf); o->own_len.resize(nf); o->defs.resize(nf); }
    std::vector<uint8_t> &S = o->own_seq[file], &Q = o->own_qual[file];
    std::vector<uint16_t> &Ln = o->own_len[file];
    std::vector<std::string> &D = o->defs[file];
    S.clear(); Q.clear(); Ln.clear(); D.clear();
    std::string l0, l1, l2, l3;
    for (int k = 0; k < n_paths; ++k) {                                  // the files of a read are concatenated (read.rs:148-149)
        LineReader rd;
        if (!rd.open(paths[k])) { set_err(o, std::string("cannot open file: ") + paths[k]); return -2; }
        while (rd.getline(l0)) {
            if (l0.empty()) continue;
            if (l0[0] != '@' || !rd.getline(l1) || !rd.getline(l2) || !rd.getline(l3) || l1.size() != l3.size()) {
                rd.close(); set_err(o, "malformed FASTQ record"); return -3;
            }
            size_t e = 1;                                                // strip_fq_suffix on the name
            while (e < l0.size() && l0[e] != ' ' && l0[e] != '\t') ++e;
            size_t name_len = e - 1;
            if (name_len > 2 && l0[e - 2] == '/' && (l0[e - 1] == '1' || l0[e - 1] == '2')) name_len -= 2;
            D.emplace_back(l0.substr(1, name_len) + l0.substr(e));
            const size_t L = l1.size(), c = std::min<size_t>(L, stride), at = S.size();
            S.resize(at + stride, 'N'); Q.resize(at + stride, '!');
            memcpy(S.data() + at, l1.data(), c);
            memcpy(Q.data() + at, l3.data(), c);
            Ln.push_back((uint16_t)std::min<size_t>(L, 65535));
        }
        rd.close();
    }
    FileInput &in = o->reads[file].in;
    in.seq = S.data(); in.qual = Q.data(); in.len = Ln.data(); in.stride = stride; in.n = Ln.size(); in.set = true;
    return 0;
}

uint64_t orc_n_records(orc_t *o, int file) {
    return (file >= 0 && (size_t)file < o->reads.size() && o->reads[file].in.set) ? o->reads[file].in.n : 0;
}

namespace {
struct ZBuf { void *ptr; size_t size, pos; };
struct ZstdApi {
    void *h = nullptr;
    void *(*createCCtx)() = nullptr;
    size_t (*freeCCtx)(void *) = nullptr;
    size_t (*setParameter)(void *, int, int) = nullptr;
    size_t (*compressStream2)(void *, ZBuf *, ZBuf *, int) = nullptr;
    unsigned (*isError)(size_t) = nullptr;
    bool ok = false;
};
ZstdApi &zstd_api() {
    static ZstdApi a;
    static bool tried = false;
    if (tried) return a;
    tried = true;
    a.h = dlopen("libzstd.so.1", RTLD_NOW);
    if (!a.h) return a;
    a.createCCtx = (void *(*)())dlsym(a.h, "ZSTD_createCCtx");
    a.freeCCtx = (size_t(*)(void *))dlsym(a.h, "ZSTD_freeCCtx");
    a.setParameter = (size_t(*)(void *, int, int))dlsym(a.h, "ZSTD_CCtx_setParameter");
    a.compressStream2 = (size_t(*)(void *, ZBuf *, ZBuf *, int))dlsym(a.h, "ZSTD_compressStream2");
    a.isError = (unsigned (*)(size_t))dlsym(a.h, "ZSTD_isError");
    a.ok = a.createCCtx && a.freeCCtx && a.setParameter && a.compressStream2 && a.isError;
    return a;
}
struct ZWriter {                                     // zstd::stream::Encoder::new(file, level) + multithread(n)
    FILE *f = nullptr;
    void *cctx = nullptr;
    std::vector<char> in, out;
    size_t fill = 0;
    int workers = 0;
    bool open(const char *path, int level, int threads) {
        ZstdApi &z = zstd_api();
        if (!z.ok) return false;
        f = fopen(path, "wb");
        if (!f) return false;
        cctx = z.createCCtx();
        z.setParameter(cctx, 100 /* ZSTD_c_compressionLevel */, level);
        workers = z.isError(z.setParameter(cctx, 400 /* ZSTD_c_nbWorkers */, threads)) ? 0 : threads;
        in.resize(4 << 20); out.resize(4 << 20);
        return true;
    }
    bool flush(int mode) {                           // 0 continue, 2 end
        ZstdApi &z = zstd_api();
        ZBuf ib{in.data(), fill, 0};
        for (;;) {
            ZBuf ob{out.data(), out.size(), 0};
            const size_t r = z.compressStream2(cctx, &ob, &ib, mode);
            if (z.isError(r)) return false;
            if (ob.pos && fwrite(out.data(), 1, ob.pos, f) != ob.pos) return false;
            if (mode == 2 ? r == 0 : ib.pos == ib.size) break;
        }
        fill = 0;
        return true;
    }
    bool write(const char *p, size_t n) {
        while (n) {
            const size_t c = std::min(n, in.size() - fill);
            memcpy(in.data() + fill, p, c);
            fill += c; p += c; n -= c;
            if (fill == in.size() && !flush(0)) return false;
        }
        return true;
    }
    bool close() {
        bool ok = flush(2);
        zstd_api().freeCCtx(cctx);
        ok = fclose(f) == 0 && ok;
        return ok;
    }
};
}  // namespace

// Returns the number of records written, or < 0.  *zstd_workers = worker threads the system libzstd accepted.
int64_t orc_make_fastq(orc_t *o, int correct_barcode, const char *r1_path, const char *r2_path, int level, int threads,
                       int *zstd_workers) {
    if (!o->keep_records || o->kept.empty()) { set_err(o, "orc_annotate with kept records must run first"); return -1; }
    ZWriter w1, w2;
    if (!w1.open(r1_path, level, threads)) { set_err(o, "cannot create R1 output (libzstd.so.1 needed)"); return -2; }
    if (r2_path && !w2.open(r2_path, level, threads)) { set_err(o, "cannot create R2 output"); return -2; }
    if (zstd_workers) *zstd_workers = w1.workers;
    int64_t written = 0;
    std::string rec;
    bool ok = true;
    for (size_t i = 0; i < o->kept.size() && ok; ++i) {                  // for record in fq_reader.flatten() (lib.rs:149)
        if (!o->kept_flag[i]) continue;                                  // groups without a barcode were dropped
        const Annotated &a = o->kept[i];
        if (correct_barcode && !a.barcode.corrected_some) continue;      // lib.rs:154
        if (!a.has_read1) { set_err(o, "called `Option::unwrap()` on a `None` value: read1"); ok = false; break; }
        rec.clear();
        rec += '@'; rec += o->defs[a.bc_file][i]; rec += '\n';
        rec += a.barcode.corrected_some ? a.barcode.corrected : a.barcode.raw.seq;
        if (a.has_umi) rec += a.umi.seq;
        rec += a.read1.seq; rec += "\n+\n";
        rec += a.barcode.raw.qual;
        if (a.has_umi) rec += a.umi.qual;
        rec += a.read1.qual; rec += '\n';
        ok = w1.write(rec.data(), rec.size());
        if (r2_path) {
            if (!a.has_read2) { set_err(o, "called `Option::unwrap()` on a `None` value: read2"); ok = false; break; }
            rec.clear();
            rec += '@'; rec += o->defs[a.r2_file][i]; rec += '\n';
            rec += a.read2.seq; rec += "\n+\n"; rec += a.read2.qual; rec += '\n';
            ok = ok && w2.write(rec.data(), rec.size());
        }
        ++written;
    }
    ok = w1.close() && ok;
    if (r2_path) ok = w2.close() && ok;
    if (!ok) { if (o->err.empty()) set_err(o, "write failed"); return -3; }
    return written;
}

}  // extern "C"
