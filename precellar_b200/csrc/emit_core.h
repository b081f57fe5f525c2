// emit_core.h -- the make_fastq records of one read group, assembled from the FASTQ text kept on the device and the
// result planes of the two passes, as __host__ __device__ code (kernels: emit.cuh; tests/emul runs it on the host).
//
// SURVEY.md §8(f) row 3.  Replaces, for runs whose text was parsed on the device, the record loop of make_fastq
// (python/src/lib.rs:148-168) together with what it reads from AnnotatedFastq:
//   R1 record   '@' + definition of the first read file holding a barcode (name stripped of /1 /2 by
//               BatchedFqReader::next, fastq.rs:421,612-621) + '\n' + corrected barcode(s) + UMI + read1 + "\n+\n"
//               + barcode qualities + UMI qualities + read1 qualities + '\n'     (Barcode::extend / join, fastq.rs:534-604)
//   R2 record   read2 under its own definition (paired-end layouts)
//   dropped     groups without a barcode (fastq.rs:381-385) and, with correct_barcode, groups with a barcode that was
//               not corrected (lib.rs:154)
// Segments of neg-strand reads are reverse-complemented, their qualities reversed (fastq.rs:487-499).
// The same function measures a record (EmitCount) and writes it (a byte sink), so sizes and bytes cannot disagree.
#ifndef PCL_EMIT_CORE_H
#define PCL_EMIT_CORE_H

#include "hd_core.h"

struct EmitRead {
    const uint8_t *text;                           // the read file's text block as parsed (pcl_text_keep)
    const uint32_t *nl;                            // offset of the '\n' ending each line, four per record
    const uint16_t *len;
    const uint16_t *fstatus;
    const uint32_t *tcoord;
    const void *bc_key[PCL_MAX_BC_PER_READ];
    const uint32_t *bcoord[PCL_MAX_BC_PER_READ];
    const uint32_t *ucoord[PCL_MAX_UMI_PER_READ];
    uint32_t max_len;
    uint8_t is_reverse, n_bc, n_umi, has_target, fixed_layout, pad[3];
    uint8_t bc_key_bytes[PCL_MAX_BC_PER_READ], bc_l16[PCL_MAX_BC_PER_READ];
    uint16_t bc_start[PCL_MAX_BC_PER_READ], bc_min[PCL_MAX_BC_PER_READ], bc_max[PCL_MAX_BC_PER_READ];
    uint16_t umi_start[PCL_MAX_UMI_PER_READ], umi_min[PCL_MAX_UMI_PER_READ], umi_max[PCL_MAX_UMI_PER_READ];
};

struct EmitParams {
    EmitRead rd[PCL_MAX_READS];
    int n_reads, correct, paired;
    uint64_t n;
};

// the layout-derived half of an EmitRead (host side: pcl_emit_fastq and the emulation harness)
inline void pcl_emit_describe(EmitRead &er, const pcl_read_desc &rd, const pcl_read_info &inf) {
    er.max_len = rd.max_len; er.is_reverse = rd.is_reverse; er.n_bc = inf.n_bc; er.n_umi = inf.n_umi;
    er.has_target = inf.has_target; er.fixed_layout = inf.fixed_layout;
    for (int j = 0; j < inf.n_bc; ++j) {
        const pcl_elem_desc &el = rd.elems[inf.bc_elem[j]];
        er.bc_key_bytes[j] = inf.bc_key_bytes[j];
        er.bc_l16[j] = (inf.bc_key_bytes[j] == 4 && el.min_len == 16 && el.max_len == 16) ? 1 : 0;
        er.bc_start[j] = inf.static_start[inf.bc_elem[j]]; er.bc_min[j] = el.min_len; er.bc_max[j] = el.max_len;
    }
    for (int j = 0; j < inf.n_umi; ++j) {
        const pcl_elem_desc &el = rd.elems[inf.umi_elem[j]];
        er.umi_start[j] = inf.static_start[inf.umi_elem[j]]; er.umi_min[j] = el.min_len; er.umi_max[j] = el.max_len;
    }
}

enum { PCL_EMIT_OK = 0, PCL_EMIT_MULTI_TARGET = 1, PCL_EMIT_BCOORD = 2, PCL_EMIT_KEY = 3, PCL_EMIT_READ1_TWICE = 4,
       PCL_EMIT_READ2_TWICE = 5, PCL_EMIT_NO_READ1 = 6, PCL_EMIT_NO_READ2 = 7 };

// where the four lines of record i lie in the text: [s, e) without the line end ("\r\n" tolerated)
struct EmitLines { uint32_t def_s, def_e, seq_s, qual_s; };

PCL_HD EmitLines pcl_emit_lines(const EmitRead &r, uint64_t i) {
    EmitLines o;
    const uint32_t e0 = r.nl[4 * i], e1 = r.nl[4 * i + 1], e2 = r.nl[4 * i + 2];
    o.def_s = i ? r.nl[4 * i - 1] + 1u : 0u;
    o.def_e = (e0 > o.def_s && r.text[e0 - 1u] == '\r') ? e0 - 1u : e0;
    o.seq_s = e0 + 1u;
    o.qual_s = e2 + 1u;
    (void)e1;
    return o;
}

PCL_HD uint32_t pcl_emit_static_coord(uint32_t s, uint32_t mn, uint32_t mx, uint32_t L) {
    if (s >= L || s + mn > L) return PCL_COORD_ABSENT;
    const uint32_t hi = s + mx < L ? s + mx : L;
    return (s << 16) | (hi - s);
}

PCL_HD uint8_t pcl_emit_complement(uint8_t x) {       // precellar::utils::rev_compl (utils.rs:138-149)
    switch (x) {
        case 'A': case 'a': return 'T';
        case 'T': case 't': return 'A';
        case 'C': case 'c': return 'G';
        case 'G': case 'g': return 'C';
        default: return x;
    }
}

// corrected key of barcode j -> number of bases, or -1 for a key that is no packed barcode
PCL_HD int pcl_emit_key(const EmitRead &r, int j, uint64_t i, uint64_t *key) {
    uint64_t k = r.bc_key_bytes[j] == 4 ? (uint64_t)((const uint32_t *)r.bc_key[j])[i] : ((const uint64_t *)r.bc_key[j])[i];
    if (r.bc_key_bytes[j] == 4 && r.bc_l16[j]) k |= 1ull << 32;
    if (k == 0 || k == PCL_KEY_ABSENT64) return -1;
    int top = 63;
    while (top >= 0 && !((k >> top) & 1ull)) --top;
    if (top & 1) return -1;
    *key = k;
    return top / 2;
}

// byte sink that only measures
struct EmitCount {
    uint32_t n = 0;
    PCL_HD void put(uint8_t) { n += 1; }
    PCL_HD void copy(const uint8_t *, uint32_t len) { n += len; }
    PCL_HD void rcopy(const uint8_t *, uint32_t len, bool) { n += len; }
    PCL_HD void bases(uint64_t, uint32_t L) { n += L; }
};

// byte sink of one thread writing front to back (host emulation)
struct EmitSerial {
    uint8_t *dst;
    PCL_HD void put(uint8_t c) { *dst++ = c; }
    PCL_HD void copy(const uint8_t *src, uint32_t len) { for (uint32_t k = 0; k < len; ++k) dst[k] = src[k]; dst += len; }
    PCL_HD void rcopy(const uint8_t *src, uint32_t len, bool compl_) {
        for (uint32_t k = 0; k < len; ++k) { const uint8_t x = src[len - 1u - k]; dst[k] = compl_ ? pcl_emit_complement(x) : x; }
        dst += len;
    }
    PCL_HD void bases(uint64_t key, uint32_t L) {
        for (uint32_t p = 0; p < L; ++p) dst[p] = (uint8_t)"ACGT"[(key >> (2u * (L - 1u - p))) & 3ull];
        dst += L;
    }
};

struct EmitDecision {
    uint32_t err;                 // PCL_EMIT_*; a group with an error makes the whole call fail (the reference panics)
    bool write;
    int bc_def, r1_file, r2_file;
    uint32_t r1c, r2c;
};

// The checks of the record loop, in the reference's order, without producing bytes.
PCL_HD EmitDecision pcl_emit_decide(const EmitParams &p, uint64_t i) {
    EmitDecision d;
    d.err = PCL_EMIT_OK; d.write = false; d.bc_def = -1; d.r1_file = -1; d.r2_file = -1; d.r1c = 0; d.r2c = 0;
    bool has_bc = false, all_ok = true;
    for (int r = 0; r < p.n_reads; ++r) {
        const EmitRead &rd = p.rd[r];
        const uint32_t fs = rd.fstatus[i], st = fs & 3u;
        if (st == PCL_SPLIT_MULTI_TARGET) { d.err = PCL_EMIT_MULTI_TARGET; return d; }
        if (st != PCL_SPLIT_OK) continue;
        uint32_t L = rd.len[i];
        if (rd.max_len && L > rd.max_len) L = rd.max_len;
        for (int j = 0; j < rd.n_bc; ++j) {
            const uint32_t code = PCL_BC_CODE(fs, j);
            if (code == PCL_BC_ABSENT) continue;
            const uint32_t c = rd.fixed_layout ? pcl_emit_static_coord(rd.bc_start[j], rd.bc_min[j], rd.bc_max[j], L) : rd.bcoord[j][i];
            if (c == PCL_COORD_ABSENT) { d.err = PCL_EMIT_BCOORD; return d; }
            if (d.bc_def < 0) d.bc_def = r;
            has_bc = true;
            if (!p.correct || code == PCL_BC_EXACT) continue;
            if (code == PCL_BC_CORRECTED) {
                uint64_t key;
                if (pcl_emit_key(rd, j, i, &key) < 0) { d.err = PCL_EMIT_KEY; return d; }
            } else all_ok = false;
        }
        if (rd.has_target && (fs & PCL_FSTATUS_TARGET)) {
            if (rd.is_reverse) {
                if (d.r2_file >= 0) { d.err = PCL_EMIT_READ2_TWICE; return d; }
                d.r2_file = r; d.r2c = rd.tcoord[i];
            } else {
                if (d.r1_file >= 0) { d.err = PCL_EMIT_READ1_TWICE; return d; }
                d.r1_file = r; d.r1c = rd.tcoord[i];
            }
        }
    }
    if (!has_bc) return d;
    if (p.correct && !all_ok) return d;
    if (d.r1_file < 0) { d.err = PCL_EMIT_NO_READ1; return d; }
    if (p.paired && d.r2_file < 0) { d.err = PCL_EMIT_NO_READ2; return d; }
    d.write = true;
    return d;
}

// '@' + definition with the name's /1 or /2 removed + '\n'
template <class Sink>
PCL_HD void pcl_emit_definition(const EmitRead &rd, uint64_t i, Sink &o) {
    const EmitLines ln = pcl_emit_lines(rd, i);
    const uint32_t b = ln.def_s + 1u;                   // behind the '@'
    uint32_t q = b;
    while (q < ln.def_e) { const uint8_t ch = rd.text[q]; if (ch == ' ' || ch == '\t') break; ++q; }
    uint32_t name_e = q;
    if (q - b > 2u && rd.text[q - 2u] == '/' && (rd.text[q - 1u] == '1' || rd.text[q - 1u] == '2')) name_e = q - 2u;
    o.put('@');
    o.copy(rd.text + b, name_e - b);
    o.copy(rd.text + q, ln.def_e - q);
    o.put('\n');
}

// one segment of a record's sequence (which = 0) or quality (which = 1) line, in the orientation the reference reports
template <class Sink>
PCL_HD void pcl_emit_segment(const EmitRead &rd, uint64_t i, uint32_t coord, int which, Sink &o) {
    const EmitLines ln = pcl_emit_lines(rd, i);
    const uint8_t *src = rd.text + (which ? ln.qual_s : ln.seq_s) + (coord >> 16);
    const uint32_t len = coord & 0xFFFFu;
    if (rd.is_reverse) o.rcopy(src, len, which == 0);
    else o.copy(src, len);
}

// barcode pieces (which = 0: corrected sequence, 1: qualities) of every read file, then the UMI of every read file
template <class Sink>
PCL_HD void pcl_emit_barcode_umi(const EmitParams &p, uint64_t i, int which, Sink &o) {
    for (int r = 0; r < p.n_reads; ++r) {
        const EmitRead &rd = p.rd[r];
        const uint32_t fs = rd.fstatus[i];
        if ((fs & 3u) != PCL_SPLIT_OK) continue;
        uint32_t L = rd.len[i];
        if (rd.max_len && L > rd.max_len) L = rd.max_len;
        for (int j = 0; j < rd.n_bc; ++j) {
            const uint32_t code = PCL_BC_CODE(fs, j);
            if (code == PCL_BC_ABSENT) continue;
            const uint32_t c = rd.fixed_layout ? pcl_emit_static_coord(rd.bc_start[j], rd.bc_min[j], rd.bc_max[j], L) : rd.bcoord[j][i];
            if (which == 0 && p.correct && code == PCL_BC_CORRECTED) {
                uint64_t key = 0;
                const int kl = pcl_emit_key(rd, j, i, &key);
                o.bases(key, (uint32_t)(kl > 0 ? kl : 0));
            } else pcl_emit_segment(rd, i, c, which, o);
        }
    }
    for (int r = 0; r < p.n_reads; ++r) {
        const EmitRead &rd = p.rd[r];
        if ((rd.fstatus[i] & 3u) != PCL_SPLIT_OK) continue;
        uint32_t L = rd.len[i];
        if (rd.max_len && L > rd.max_len) L = rd.max_len;
        uint32_t uc = PCL_COORD_ABSENT;                     // the last UMI segment of a file wins (fastq.rs:502)
        for (int j = 0; j < rd.n_umi; ++j) {
            const uint32_t c = rd.fixed_layout ? pcl_emit_static_coord(rd.umi_start[j], rd.umi_min[j], rd.umi_max[j], L) : rd.ucoord[j][i];
            if (c != PCL_COORD_ABSENT) uc = c;
        }
        if (uc != PCL_COORD_ABSENT) pcl_emit_segment(rd, i, uc, which, o);
    }
}

template <class Sink>
PCL_HD void pcl_emit_r1(const EmitParams &p, uint64_t i, const EmitDecision &d, Sink &o) {
    const EmitRead &h1 = p.rd[d.r1_file];
    const EmitLines ln = pcl_emit_lines(h1, i);
    pcl_emit_definition(p.rd[d.bc_def], i, o);
    pcl_emit_barcode_umi(p, i, 0, o);
    o.copy(h1.text + ln.seq_s + (d.r1c >> 16), d.r1c & 0xFFFFu);
    o.put('\n'); o.put('+'); o.put('\n');
    pcl_emit_barcode_umi(p, i, 1, o);
    o.copy(h1.text + ln.qual_s + (d.r1c >> 16), d.r1c & 0xFFFFu);
    o.put('\n');
}

template <class Sink>
PCL_HD void pcl_emit_r2(const EmitParams &p, uint64_t i, const EmitDecision &d, Sink &o) {
    const EmitRead &h2 = p.rd[d.r2_file];
    const EmitLines ln = pcl_emit_lines(h2, i);
    pcl_emit_definition(h2, i, o);
    o.copy(h2.text + ln.seq_s + (d.r2c >> 16), d.r2c & 0xFFFFu);
    o.put('\n'); o.put('+'); o.put('\n');
    o.copy(h2.text + ln.qual_s + (d.r2c >> 16), d.r2c & 0xFFFFu);
    o.put('\n');
}

#endif  // PCL_EMIT_CORE_H
