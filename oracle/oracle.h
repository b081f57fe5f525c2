/*
 * oracle.h -- C ABI of the CPU parity oracle.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a deliberately literal, byte-string based CPU
 * restatement of precellar's barcode hot path (segment split -> whitelist count ->
 * Phred-weighted 1-Hamming correction -> annotate/join -> FASTQ QC).  It is used by
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * as the checker and the reported CPU baseline.  Nothing under precellar_b200/ may
 * include, link or call it; the product path has no CPU fallback.
 *
 * Every function in oracle.cpp cites the reference file:line it follows
 * (paths relative to the precellar repository root).
 *
 * Parity pinning: split() is pinned by the reference's own 11 known-answer vectors
 * (seqspec/src/read/segment.rs:617-705) -- see tests/test_oracle_golden.py -- and the
 * extraction end to end (YAML -> region table -> split -> barcode join -> read1/read2) by the
 * reference's golden output files precellar/data/test{2,3,4}.out.gz (750 lines of three real
 * assays; inputs rebuilt from the outputs, tests/test_reference_outputs.py).
 * Counting / correction have no test in the reference: "parity unpinned" for
 * WhitelistBuilder, likelihood1 and correct_barcode (restatement reviewed line by line,
 * exercised with hand-computed cases in tests/test_oracle_correction.py, and required to
 * agree bit for bit with a second, independent restatement, tests/test_oracle_independent.py).
 */
#ifndef PRECELLAR_ORACLE_H
#define PRECELLAR_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* region_type classes that the path distinguishes (seqspec/src/region.rs:321-376). */
enum { ORC_OTHER = 0, ORC_BARCODE = 1, ORC_UMI = 2, ORC_TARGET = 3, ORC_PRIMER = 4 };

/* split() outcomes (seqspec/src/read/segment.rs:99-103). */
enum { ORC_SPLIT_OK = 0, ORC_SPLIT_ANCHOR_NOT_FOUND = 1, ORC_SPLIT_PATTERN_MISMATCH = 2 };

/* correct_barcode() outcomes (precellar/src/barcode.rs:43-48,160-184). */
enum {
    ORC_BC_OK = 0,          /* Ok(bytes)                 */
    ORC_BC_NO_MATCH = 1,    /* Err(NoMatch)              */
    ORC_BC_LOW_CONF = 2,    /* Err(LowConfidence)        */
    ORC_BC_EXCEED_EE = 3,   /* Err(ExceedExpectedError)  */
    ORC_BC_ABSENT = 255     /* segment not produced      */
};

typedef struct {
    uint8_t rtype;        /* ORC_* region class                                  */
    uint8_t seq_fixed;    /* sequence_type == Fixed                              */
    uint8_t pad[2];
    uint32_t min_len;
    uint32_t max_len;
    int32_t region;       /* whitelist region slot for barcode elements, else -1 */
    const char *sequence; /* NUL-terminated; only read when seq_fixed            */
} orc_elem;

typedef struct orc orc_t;

orc_t *orc_create(void);
void orc_destroy(orc_t *);
const char *orc_last_error(orc_t *);

/* One sequenced read (FASTQ file) of the modality, in sequence_spec order.
 * reader_max_len: Read.max_len applied by FastqReader::read_record (0 = no truncation).
 * Returns the file index or <0. */
int orc_add_read(orc_t *, int is_reverse, uint32_t reader_max_len, const orc_elem *elems, int n_elems);

/* Whitelist of one barcode region, lines in file order (duplicates allowed: the first
 * occurrence defines the index, like IndexSet).  n == 0 -> "no whitelist" mode. */
int orc_set_whitelist(orc_t *, int region, const uint8_t *bytes, const uint64_t *offsets, uint64_t n);

/* Borrow one file's records as an SoA batch: record i occupies seq[i*stride .. +len[i]).
 * seq/qual may be NULL for files whose layout has no barcode/UMI/anchor (lengths only). */
int orc_set_input(orc_t *, int file, const uint8_t *seq, const uint8_t *qual, const uint16_t *len,
                  uint32_t stride, uint64_t n);

/* Keep the joined AnnotatedFastq of every group for orc_render (small cases only). */
void orc_set_keep_records(orc_t *, int keep);

/* Pass 1 == BarcodeAnalyzer::new (serial). */
int orc_count(orc_t *);
int orc_get_counts(orc_t *, int region, uint64_t *counts /* n_whitelist */, uint64_t *mismatch, uint64_t *total);

/* Pass 2 == gen_barcoded_fastq(correct, chunk_size) drained through process_chunk.
 * n_threads workers over n/128-record sub-chunks of >= chunk_bases batches. */
typedef struct {
    /* per file f, record i : index f*n+i */
    uint8_t *fstatus;      /* ORC_SPLIT_*                                         */
    uint32_t *tcoord;      /* target segment (start<<16|len), 0xFFFFFFFF if none  */
    /* per emit slot k (barcode/UMI elements of all files, file-major, element order): k*n+i */
    uint32_t *bucoord;     /* (start<<16|len) in the file's record, 0xFFFFFFFF if absent */
    /* per barcode segment j (subset of emit slots, same order): j*n+i */
    int64_t *bc_index;     /* corrected whitelist index, -1 if not corrected / absent   */
    uint8_t *bc_err;       /* ORC_BC_*                                                  */
    /* per group i */
    uint8_t *g_flags;      /* bit0 has_barcode (kept), bit1 corrected is Some, bit2 umi, bit3 read1, bit4 read2 */
} orc_out;

int orc_annotate(orc_t *, int correct, double threshold, double max_expected_errors, int n_threads,
                 uint64_t chunk_bases, orc_out *out);

/* QcFastq counters after orc_annotate (precellar/src/qc.rs:21-89).
 * per_file: [num_reads, num_defect] * n_files ; cat: for umi, barcode, read1, read2:
 * [num_reads, total_bases, q30_bases]. */
int orc_get_qc(orc_t *, uint64_t *per_file, uint64_t *cat /* 12 */);

/* Render group i of the last orc_annotate as
 * "bc_seq\tbc_qual\tcorrected|*\tumi_seq\tumi_qual\tr1_seq\tr1_qual\tr2_seq\tr2_qual"
 * (empty string if the group was dropped).  Returns bytes written (excl. NUL) or <0. */
int64_t orc_render(orc_t *, uint64_t i, char *buf, uint64_t cap);

/* Standalone helpers used by the golden-vector tests. */
/* split one read; writes "[B]AAAA,[U]TT" style text (segment.rs Display impls) */
int orc_split_str(const orc_elem *elems, int n_elems, int is_reverse, const uint8_t *seq, const uint8_t *qual,
                  uint32_t len, double linker_tol, double anchor_tol, char *buf, uint32_t cap);
/* truncate_max: returns number of elements kept (segment.rs:144-164) */
int orc_truncate_max(const orc_elem *elems, int n_elems, uint32_t len);
void orc_rev_compl_seqspec(const uint8_t *in, uint32_t n, uint8_t *out);   /* seqspec/src/utils.rs:176-187 */
void orc_rev_compl_precellar(const uint8_t *in, uint32_t n, uint8_t *out); /* precellar/src/utils.rs:138-149 */
double orc_error_probability(uint8_t q);                                   /* precellar/src/barcode.rs:464-468 */
/* likelihood1 against an ad-hoc count table: returns ORC_BC_* of correct_barcode and the
 * chosen key index / posterior. */
int orc_correct_one(const uint8_t *wl_bytes, const uint64_t *wl_offsets, const uint64_t *wl_counts, uint64_t n_wl,
                    const uint8_t *bc, const uint8_t *qual, uint32_t len, double threshold,
                    double max_expected_errors, int64_t *index_out, double *prob_out);

/* File-level driver for the API end-to-end baseline (bench.py api_e2e, reference arm): read the concatenated FASTQ
 * files (plain or gzip) of one read into planes owned by the oracle, then -- after orc_count and orc_annotate with kept
 * records -- write R1.fq.zst / R2.fq.zst like make_fastq (python/src/lib.rs:136-168; zstd level 9, 8 workers). */
int orc_fastq_load(orc_t *, int file, const char *const *paths, int n_paths, uint32_t stride);
uint64_t orc_n_records(orc_t *, int file);
int64_t orc_make_fastq(orc_t *, int correct_barcode, const char *r1_path, const char *r2_path /* or NULL */, int level, int threads,
                       int *zstd_workers);

#ifdef __cplusplus
}
#endif
#endif
