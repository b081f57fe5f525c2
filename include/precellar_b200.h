/*
 * precellar_b200.h -- C ABI of the B200-native barcode hot path.
 *
 * This is the drop-in boundary for precellar's read-preprocessing hot path:
 *     seqspec-driven segment extraction  ->  whitelist frequency count  ->
 *     Phred-weighted 1-Hamming barcode correction
 * implemented as hand-written sm_100a CUDA kernels.  Plain pointers and sizes only; no C++,
 * CUDA or torch types cross this boundary.  There is NO CPU fallback behind these entry points:
 * every compute call fails with PCL_ERR_CUDA when no usable device is present.
 *
 * What each entry point replaces in the reference (paths relative to the precellar repo):
 *   pcl_layout_set          Assay::get_segments_by_modality -> Read::get_segments -> truncate_max
 *                           (seqspec/src/lib.rs:416-440, seqspec/src/read.rs:182-212,
 *                           seqspec/src/read/segment.rs:144-164): the per-read SegmentInfo tables,
 *                           already derived by the host (precellar_b200/layout.py).
 *   pcl_whitelist_load      Assay::get_whitelists + WhitelistBuilder::new
 *                           (seqspec/src/lib.rs:443-472, precellar/src/barcode.rs:393-407)
 *   pcl_chunk_* (inputs)    BatchedFqReader::next batches (precellar/src/align/fastq.rs:400-448)
 *   pcl_count               BarcodeAnalyzer::new, hot loop A: split + WhitelistBuilder::add
 *                           (precellar/src/barcode.rs:59-137,410-426; segment.rs:180-368)
 *   pcl_counts_allreduce    (new) merges per-GPU counts; the reference is single-process
 *   pcl_correct             FastqAnnotator::annotate + BarcodeAnalyzer::correct_barcode +
 *                           OligoFrequncy::likelihood1, hot loop B
 *                           (precellar/src/align/fastq.rs:474-525, precellar/src/barcode.rs:160-184,311-358)
 *   pcl_chunk_fetch         the AnnotatedFastq fields a consumer reads
 *                           (precellar/src/align/fastq.rs:528-556)
 *   pcl_counts_fetch        Whitelist{barcode_counts, mismatch_count, total_count}
 *                           (precellar/src/barcode.rs:362-366,431-437)
 *   pcl_qc_fetch            QcFastq::num_reads / num_defect (precellar/src/qc.rs:21-27)
 *   pcl_qc_accumulate /     QcFastq::update: bases and Q30 bases of barcode / umi / read1 / read2
 *   pcl_qc_fetch_categories (precellar/src/qc.rs:30-69)
 *   pcl_fastq_*             Read::open + FastqReader::read_record + the per-file half of BatchedFqReader::next
 *                           (seqspec/src/read.rs:98-108,137-155; precellar/src/align/fastq.rs:400-448,612-621)
 *   pcl_gzsrc_*,            open_file's MultiGzDecoder (seqspec/src/utils.rs:44-56): ordinary gzip on a host thread,
 *   pcl_bgzf_walk,          BGZF members on the device (one warp per member), the text indexed where it is
 *   pcl_inflater_*,
 *   pcl_text_scan_device
 *
 * Threading: one host thread drives one ctx (one GPU).  Calls on a ctx are not re-entrant.
 * Different ctxs are independent.  Work on a chunk is stream-ordered on the chunk's stream;
 * pcl_chunk_fetch / pcl_sync block.  A pcl_inflater has its own stream and staging and may be driven
 * from a thread of its own (one per read file) next to the thread that drives the ctx.
 *
 * Errors: every function returns 0 or a negative pcl_status; pcl_last_error() gives the message.
 * Nothing throws or aborts across this boundary.  Per-read structural failures are data
 * (status bits in the results), not errors.
 */
#ifndef PRECELLAR_B200_H
#define PRECELLAR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PCL_ABI_VERSION 1

#define PCL_MAX_READS 8          /* read files per read-group                     */
#define PCL_MAX_ELEMS 16         /* SegmentInfo elements per read                 */
#define PCL_MAX_ANCHOR 64        /* bytes of a Fixed (anchor) sequence            */
#define PCL_MAX_REGIONS 8        /* barcode regions (whitelists) per modality     */
#define PCL_MAX_BC_PER_READ 4    /* barcode elements per read                     */
#define PCL_MAX_UMI_PER_READ 4   /* UMI elements per read                         */
#define PCL_MAX_KEY_LEN 31       /* bases in one packed barcode / UMI key         */
#define PCL_MAX_STRIDE 4096      /* bytes per record slot                         */

typedef enum {
    PCL_OK = 0,
    PCL_ERR_ARG = -1,            /* bad argument / call order                              */
    PCL_ERR_CUDA = -2,           /* CUDA runtime failure, or no device                     */
    PCL_ERR_NCCL = -3,           /* NCCL failure / libnccl not loadable                    */
    PCL_ERR_UNSUPPORTED = -4,    /* layout / whitelist outside what the kernels implement  */
    PCL_ERR_INPUT = -5,          /* inconsistent input (the reference would panic/assert)  */
    PCL_ERR_NOMEM = -6
} pcl_status;

/* Region classes the path distinguishes (seqspec/src/region.rs:342-376). */
typedef enum { PCL_KIND_OTHER = 0, PCL_KIND_BARCODE = 1, PCL_KIND_UMI = 2, PCL_KIND_TARGET = 3 } pcl_kind;

/* One SegmentInfoElem (seqspec/src/read/segment.rs:371-378). */
typedef struct {
    uint8_t kind;                /* pcl_kind                                                   */
    uint8_t is_fixed_seq;        /* sequence_type == Fixed: acts as an anchor after a variable run */
    uint16_t min_len;
    uint16_t max_len;
    int16_t region_slot;         /* barcode elements: index of the region's whitelist; else -1 */
    uint8_t seq_len;             /* strlen(sequence) when is_fixed_seq                         */
    uint8_t reserved[3];
    char sequence[PCL_MAX_ANCHOR];
} pcl_elem_desc;

/* One sequenced read of the modality (seqspec/src/read.rs:56-66) with its derived SegmentInfo. */
typedef struct {
    uint16_t n_elems;
    uint8_t is_reverse;          /* strand: neg                                                */
    uint8_t reserved;
    uint32_t max_len;            /* Read.max_len: records are truncated to it when > 0 (read.rs:98-103) */
    pcl_elem_desc elems[PCL_MAX_ELEMS];
} pcl_read_desc;

/* ---- results ------------------------------------------------------------------------------
 * fstatus (uint16, one per read file and record):
 *   bits 1:0   split status: 0 ok, 1 AnchorNotFound, 2 PatternMismatch (record is a "defect",
 *              precellar/src/align/fastq.rs:369-374), 3 more than one target segment (reference panics)
 *   bit  2     target segment present (tcoord valid)
 *   bit  3     reserved
 *   bits 15:4  four 3-bit codes, one per barcode element of this read in element order:        */
#define PCL_SPLIT_OK 0
#define PCL_SPLIT_ANCHOR_NOT_FOUND 1
#define PCL_SPLIT_PATTERN_MISMATCH 2
#define PCL_SPLIT_MULTI_TARGET 3
#define PCL_FSTATUS_TARGET 0x4
#define PCL_BC_EXACT 0u          /* raw barcode is in the whitelist: corrected == raw (posterior 1.0)  */
#define PCL_BC_ABSENT 3u         /* segment not produced by split (short read / defect)                */
#define PCL_BC_NO_MATCH 4u       /* Err(NoMatch)          -- also the state between pcl_count and pcl_correct */
#define PCL_BC_CORRECTED 5u      /* Ok(key) by one substitution, posterior >= threshold                */
#define PCL_BC_LOW_CONF 6u       /* Err(LowConfidence)                                                 */
#define PCL_BC_EXCEED_EE 7u      /* Err(ExceedExpectedError)                                           */
#define PCL_BC_CODE(fstatus, j) (((fstatus) >> (4 + 3 * (j))) & 7u)
/*
 * Packed keys (barcodes and UMIs): 2 bits per base, A=0 C=1 G=2 T=3, first base most significant,
 * and a marker bit above the first base: key = (1 << 2L) | bases, in the orientation the reference
 * reports (reverse-complemented for neg-strand reads).  Keys are stored 4 bytes wide when the element
 * is at most 15 bases long, or exactly 16 (min_len == max_len == 16), and 8 bytes wide otherwise
 * (pcl_read_info.*_key_bytes); a 4-byte key of a 16-base element has its marker bit (bit 32) dropped.
 *   bc_key   valid only for PCL_BC_EXACT (the raw key) and PCL_BC_CORRECTED (the whitelist key).
 *   umi_key  0 = present but not packable (a byte outside ACGT, or longer than PCL_MAX_KEY_LEN;
 *            16-base UMIs are stored 8 bytes wide so that 0 stays free); all-ones = segment absent.
 *
 * Coordinates: (start << 16) | len inside the (max_len-truncated) record; 0xFFFFFFFF = absent.
 * For reads whose layout is all fixed-length (pcl_read_info.fixed_layout) barcode/UMI coordinates are
 * static (pcl_read_info.static_start, element lengths) and bcoord/ucoord are not produced.
 */
#define PCL_KEY_ABSENT32 0xFFFFFFFFu
#define PCL_KEY_ABSENT64 0xFFFFFFFFFFFFFFFFull
#define PCL_COORD_ABSENT 0xFFFFFFFFu

/* pcl_read_info.route: which pass-1 kernel a read with 4-byte keys takes.  WALK = generic element walk; FAST = fixed
 * layout within 32 bytes; FAST_SPEC = the same with a compile-time field geometry; PLAN = one variable element then an
 * anchor; PLAN_GEO = the same with a compile-time geometry instance (sci-RNA-seq3 R1).  8-byte-key tables and the
 * PCL_* environment switches may still send a read down the WALK kernel at launch time. */
enum { PCL_ROUTE_WALK = 0, PCL_ROUTE_FAST = 1, PCL_ROUTE_FAST_SPEC = 2, PCL_ROUTE_PLAN = 3, PCL_ROUTE_PLAN_GEO = 4, PCL_ROUTE_LENONLY = 5 };

typedef struct {
    uint8_t n_bc;                          /* barcode elements, in element order                      */
    uint8_t n_umi;                         /* UMI elements                                            */
    uint8_t has_target;
    uint8_t fixed_layout;                  /* all elements fixed-length except an optional last one   */
    uint8_t needs_payload;                 /* seq/qual bytes are read (else lengths only)             */
    uint8_t qual_only;                     /* QC mode, insert read: only the quality plane is used    */
    uint8_t route;                         /* pass-1 kernel family of this read (PCL_ROUTE_*, informational) */
    uint8_t reserved;
    uint8_t bc_elem[PCL_MAX_BC_PER_READ];  /* element index of each barcode element                   */
    uint8_t bc_key_bytes[PCL_MAX_BC_PER_READ];   /* 4 or 8                                            */
    uint8_t umi_elem[PCL_MAX_UMI_PER_READ];      /* element index of each UMI element                 */
    uint8_t umi_key_bytes[PCL_MAX_UMI_PER_READ]; /* 4 or 8                                            */
    uint16_t static_start[PCL_MAX_ELEMS];  /* fixed_layout: start of every element                    */
    uint16_t qual_offset;                  /* the quality plane holds record bytes [qual_offset,       */
    uint16_t qual_stride;                  /*   qual_offset + qual_stride) of every record             */
} pcl_read_info;

/* Host-side views handed to pcl_chunk_fetch; any pointer may be NULL to skip that array.
 * bc_key[j] / umi_key[j] point at n keys of the width given by pcl_read_info. */
typedef struct {
    uint16_t *fstatus;                     /* [n]                                              */
    uint32_t *tcoord;                      /* [n]          if has_target                       */
    void *bc_key[PCL_MAX_BC_PER_READ];     /* [n] each                                         */
    void *umi_key[PCL_MAX_UMI_PER_READ];   /* [n] each                                         */
    uint32_t *bcoord[PCL_MAX_BC_PER_READ]; /* [n] each     only when !fixed_layout             */
    uint32_t *ucoord[PCL_MAX_UMI_PER_READ];/* [n] each     only when !fixed_layout             */
} pcl_read_results;

typedef struct pcl_ctx pcl_ctx;
typedef struct pcl_chunk pcl_chunk;

/* ---- context ---------------------------------------------------------------------------- */
int pcl_abi_version(void);
/* Message of the last failure on this ctx (ctx == NULL: last failure of pcl_ctx_create on this thread). */
const char *pcl_last_error(const pcl_ctx *ctx);
int pcl_device_count(void);                                     /* >= 0, or PCL_ERR_CUDA            */
int pcl_ctx_create(pcl_ctx **out, int device);
int pcl_ctx_destroy(pcl_ctx *ctx);
int pcl_sync(pcl_ctx *ctx);                                     /* wait for all streams of the ctx  */

/* Multi-GPU (one ctx per rank; ranks may be threads or processes).  uid is ncclUniqueId (128 bytes). */
int pcl_comm_unique_id(void *uid128);
int pcl_comm_init(pcl_ctx *ctx, int nranks, int rank, const void *uid128);

/* ---- configuration ----------------------------------------------------------------------- */
int pcl_layout_set(pcl_ctx *ctx, const pcl_read_desc *reads, int n_reads);
/* FASTQ QC of inserts (QcFastq, precellar/src/qc.rs:30-69) needs every quality byte of the target segments.
 * Call before pcl_layout_set: reads that hold a target then carry a quality plane of ceil16(Read.max_len)
 * bytes per record (Read.max_len must be > 0), uploaded with pcl_chunk_upload like any other plane. */
int pcl_qc_enable(pcl_ctx *ctx, int on);
/* Assay::verify (seqspec/src/lib.rs:481-548) checks a sample of every read it is handed with
 * split_with_tolerance(read, 0.2, 0.2): fixed sequences that the walk does not search for are compared with the read
 * as well and more than 20 % mismatches make the record invalid (segment.rs:325-333).  Call before pcl_layout_set; the
 * ctx then runs pass 1 with that rule (interpreter kernel) -- for verification samples, not for production runs. */
int pcl_verify_enable(pcl_ctx *ctx, int on);
int pcl_read_info_get(const pcl_ctx *ctx, int read_slot, pcl_read_info *out);
/* Whitelist of one barcode region: n byte strings, string i = bytes[offsets[i] .. offsets[i+1]).
 * File order; duplicates are dropped keeping the first (IndexSet semantics).  Entries must be
 * upper-case ACGT of length 1..PCL_MAX_KEY_LEN, else PCL_ERR_UNSUPPORTED. */
int pcl_whitelist_load(pcl_ctx *ctx, int region_slot, const uint8_t *bytes, const uint64_t *offsets, uint64_t n);
uint64_t pcl_whitelist_size(const pcl_ctx *ctx, int region_slot);   /* after dedup */

/* ---- chunks: one batch of read-groups, all read files of a group in the same chunk -------- */
int pcl_chunk_create(pcl_ctx *ctx, uint64_t capacity_records, pcl_chunk **out);
int pcl_chunk_destroy(pcl_chunk *chunk);
/* Set the number of records of the next fill (<= capacity) and invalidate previous contents. */
int pcl_chunk_reset(pcl_chunk *chunk, uint64_t n_records);
/* Record slot size (multiple of 16) the ctx uses for this read's seq/qual planes. */
uint32_t pcl_read_stride(const pcl_ctx *ctx, int read_slot);
/* Asynchronous H2D of one read file's SoA planes: record i occupies [i*stride, i*stride+len[i]) of seq.
 * The quality plane only has to cover what the correction pass reads -- the barcode bases: row i holds the
 * quality bytes [qual_offset, qual_offset + qual_stride) of record i (pcl_read_info; for v3 that is 16 of the 28
 * bytes).  Variable layouts and QC mode use the whole record (qual_offset 0, qual_stride == stride).
 * seq/qual may be NULL when !needs_payload.  Host buffers should be pinned (pcl_host_alloc) and must
 * stay valid until the next blocking call on the chunk. */
int pcl_chunk_upload(pcl_chunk *chunk, int read_slot, const uint8_t *seq, const uint8_t *qual, const uint16_t *len);
/* The same with sequence rows of `row_bytes` instead of pcl_read_stride bytes: the host-to-device copy bounds the
 * end-to-end rate, so a caller may hand over only the bytes the layout can touch (pcl_read_payload_bytes, a multiple
 * of 4: 28 instead of 32 for 10x v3 R1); the rows are widened on the device.  row_bytes must be a multiple of 4 in
 * [pcl_read_payload_bytes, pcl_read_stride].  The quality and length planes are as for pcl_chunk_upload. */
int pcl_chunk_upload_compact(pcl_chunk *chunk, int read_slot, const uint8_t *seq, uint32_t row_bytes, const uint8_t *qual,
                             const uint16_t *len);
uint32_t pcl_read_payload_bytes(const pcl_ctx *ctx, int read_slot);
/* Every record of this read file has the same length (the usual case for a sequencing run): the length plane is filled
 * on the device and never crosses PCIe; later pcl_chunk_upload* calls on this slot may pass len == NULL (until the next
 * pcl_chunk_reset).  For a read that uploads lengths only (an insert read) this call alone makes the slot ready. */
int pcl_chunk_set_uniform_len(pcl_chunk *chunk, int read_slot, uint16_t len);
/* The result planes of an insert read are the same for every record when all lengths agree (fstatus, trimmed insert
 * coordinates).  This asks the device whether they are and brings back 16 bytes instead of 6 per record: stream-ordered,
 * `out` (pinned) is valid after pcl_sync / a blocking fetch.  uniform == 0: fetch the planes with pcl_chunk_fetch. */
typedef struct { uint32_t uniform, fstatus, tcoord, reserved; } pcl_uniform_result;
int pcl_chunk_uniform_results_async(pcl_chunk *chunk, int read_slot, pcl_uniform_result *out);
/* Zero-copy producers (a GPU-side decoder or generator) may fill the device planes directly. */
int pcl_chunk_device_ptrs(pcl_chunk *chunk, int read_slot, void **seq, void **qual, void **len);
void *pcl_chunk_stream(pcl_chunk *chunk);                        /* cudaStream_t of the chunk        */

/* ---- the three stages --------------------------------------------------------------------- */
int pcl_counts_reset(pcl_ctx *ctx);                             /* stream-ordered against every chunk (async) */
/* Pass 1 on this chunk (async).  For reads whose whitelists are 32-bit tables it also runs the stage of pass 2 that needs
 * no counts (the 1-Hamming neighbourhood filter of the missed barcodes), so that it overlaps the all-reduce. */
int pcl_count(pcl_ctx *ctx, pcl_chunk *chunk);
/* Sums the counts over the ranks after every chunk's pass 1; later chunk work waits for it (async).  Once per round:
 * with more than one rank a second call, or a pcl_count, before the next pcl_counts_reset is PCL_ERR_ARG (the counts
 * already hold the global sums; adding to them or summing them again would multiply them by the number of ranks). */
int pcl_counts_allreduce(pcl_ctx *ctx);
int pcl_correct(pcl_ctx *ctx, pcl_chunk *chunk, double threshold, double max_expected_errors); /* pass 2 (async) */
/* For callers that cannot keep every chunk resident between the passes (the reference streams its files twice in bounded
 * memory, precellar/src/align/fastq.rs:109-153): after pcl_counts_allreduce, with the counts frozen, pcl_count recomputes a
 * re-uploaded chunk's pass-1 outputs (split status, keys, coordinates, miss worklist) WITHOUT touching the whitelist
 * counts, totals or QC counters; pcl_correct then works as usual. */
int pcl_counts_freeze(pcl_ctx *ctx, int on);
/* Device bytes one chunk of `capacity_records` takes (pcl_chunk_create allocates them in one piece), and the free /
 * total memory of the ctx's device: what a host layer needs to decide how many chunks it can keep resident. */
uint64_t pcl_chunk_bytes(const pcl_ctx *ctx, uint64_t capacity_records);
int pcl_mem_info(pcl_ctx *ctx, uint64_t *free_bytes, uint64_t *total_bytes);

/* ---- outputs ------------------------------------------------------------------------------ */
int pcl_chunk_fetch(pcl_chunk *chunk, int read_slot, const pcl_read_results *out);   /* D2H + wait */
/* Same copies without the wait (stream-ordered on the chunk's stream; pcl_sync or a later pcl_chunk_fetch
 * completes them).  UMI keys, coordinates and the split status are final after pcl_count, so they can travel
 * back while later chunks are still uploading; fstatus barcode codes and bc_key are final after pcl_correct. */
int pcl_chunk_fetch_async(pcl_chunk *chunk, int read_slot, const pcl_read_results *out);
/* counts[i] for whitelist entry i (deduplicated file order); any pointer may be NULL. */
int pcl_counts_fetch(pcl_ctx *ctx, int region_slot, uint64_t *counts, uint64_t *mismatch, uint64_t *total);
/* per read file: num_reads, num_defect (2 * n_reads values) accumulated by pcl_count since the last reset */
int pcl_qc_fetch(pcl_ctx *ctx, uint64_t *per_read);
/* QcFastq::update over one chunk (after pcl_count; independent of pcl_correct): accumulates, for the categories
 * umi, barcode, read1, read2 (in that order), {num_reads, num_total_bases, num_q30_bases}.  Needs pcl_qc_enable
 * when a read holds a target.  pcl_counts_reset clears the accumulators. */
int pcl_qc_accumulate(pcl_ctx *ctx, pcl_chunk *chunk);
int pcl_qc_fetch_categories(pcl_ctx *ctx, uint64_t *cat12);
/* Unpack a key into ASCII (host utility): returns the length or -1. */
int pcl_key_decode(uint64_t key, char *out32);

/* ---- host ingest: FASTQ (plain or gzip; the files of one read are concatenated) -> SoA planes ---------- */
typedef struct pcl_fastq pcl_fastq;
int pcl_fastq_open(const char *const *paths, int n_paths, pcl_fastq **out);
void pcl_fastq_close(pcl_fastq *r);
const char *pcl_fastq_error(const pcl_fastq *r);
/* Reads up to max_records records into [max_records][stride] planes (seq/qual may be NULL), u16 lengths and
 * the hash of the read name (definition up to the first blank, trailing /1 or /2 stripped: the value
 * BatchedFqReader compares across files).  text/off (optional) keep definition, sequence and quality of every
 * record: off[3i] .. off[3i+3] delimit them inside text; *text_used is advanced.  Returns the number of records
 * (0 at end of input) or a negative pcl_status (pcl_fastq_error has the message). */
int64_t pcl_fastq_read(pcl_fastq *r, uint64_t max_records, uint32_t stride, uint32_t qual_offset, uint32_t qual_stride,
                       uint8_t *seq, uint8_t *qual, uint16_t *len, uint64_t *name_hash, uint8_t *text, uint64_t text_cap,
                       uint64_t *off, uint64_t *text_used);
/* 1 when that call stopped early because `text` was full.  Nothing is lost: call pcl_fastq_read again with a larger
 * arena (the old contents copied, *text_used carried over) and seq / qual / len / name_hash / off advanced by the
 * records already delivered (off by 3 per record), so that a batch keeps its size whatever the header lengths. */
int pcl_fastq_text_full(const pcl_fastq *r);

/* ---- host-side record assembly for writers ------------------------------------------------------
 * make_fastq's record loop (python/src/lib.rs:149-167) on top of FastqAnnotator::annotate / AnnotatedFastq::join /
 * Barcode::extend (precellar/src/align/fastq.rs:474-604), for one fetched chunk: R1 = (corrected or raw) barcode +
 * UMI + read1, R2 = read2; groups without a barcode are skipped and, with correct_barcode, so are groups whose
 * barcode is not corrected.  Text is FASTQ ("@definition\nseq\n+\nqual\n"), appended at *r1_used / *r2_used. */
typedef struct {
    const uint8_t *text;         /* arena filled by pcl_fastq_read (definition, sequence, quality per record) */
    const uint64_t *off;         /* 3 * n + 1 offsets into text                                               */
    const uint16_t *len;         /* record lengths as uploaded                                                */
    pcl_read_results res;        /* host copies of the result planes (pcl_chunk_fetch)                        */
} pcl_host_read;
int pcl_make_fastq_batch(pcl_ctx *ctx, const pcl_host_read *reads, uint64_t n, int correct_barcode, int paired,
                         uint8_t *r1_out, uint64_t r1_cap, uint64_t *r1_used, uint8_t *r2_out, uint64_t r2_cap,
                         uint64_t *r2_used, uint64_t *n_written);

/* ---- device ingest: raw FASTQ text -> SoA planes, parsed on the GPU ------------------------------------
 * Alternative to pcl_fastq_read + pcl_chunk_upload for pass 1: the host only moves
 * uncompressed text; no host core touches individual records.  Replaces FastqReader::read_record
 * (seqspec/src/read.rs:137-155) and the name check of BatchedFqReader::next (precellar/src/align/fastq.rs:417-431)
 * for that pass.  Dialect: exactly four '\n'-terminated lines per record, "\r\n" tolerated, no blank lines
 * between records (pcl_fastq_read skips those; here they are reported as a malformed record).
 *
 * Per chunk and read file:
 *   pcl_text_scan(chunk, slot, text, n)   copy a block of text that STARTS AT A RECORD BOUNDARY to the device and
 *                                         index its line ends (asynchronous; text must stay valid until
 *                                         pcl_text_lines returns; n < 4 GiB - 4 KiB)
 *   pcl_text_lines(chunk, slot, &lines)   number of complete lines indexed (blocks).  lines / 4 complete records are
 *                                         available; the caller picks n_records <= min over the read files and
 *                                         calls pcl_chunk_reset(chunk, n_records)
 *   pcl_text_fill(chunk, slot, &used)     fill the slot's seq / qual / len planes with the first n_records records
 *                                         (same bytes as pcl_fastq_read would produce, incl. the quality window and
 *                                         the 'N' / '!' padding); *used = bytes of text consumed, the remainder is
 *                                         the head of the caller's next block.  PCL_ERR_INPUT for a malformed record.
 *   pcl_text_names_check(chunk, &bad)     first group whose read names differ between the text-filled read files
 *                                         (names compared after stripping a trailing /1 or /2), or -1
 *   pcl_text_line_ends(chunk, slot, out, n) byte offset of the end of each of the first n lines (for writers that
 *                                         slice names / records out of the host copy of the text)              */
int pcl_text_scan(pcl_chunk *chunk, int read_slot, const uint8_t *text, uint64_t n_bytes);
int pcl_text_lines(pcl_chunk *chunk, int read_slot, uint64_t *n_lines);
int pcl_text_fill(pcl_chunk *chunk, int read_slot, uint64_t *consumed);
int pcl_text_names_check(pcl_chunk *chunk, int64_t *first_bad);
int pcl_text_line_ends(pcl_chunk *chunk, int read_slot, uint32_t *out, uint64_t n_lines);

/* ---- device-side make_fastq writer --------------------------------------------------------------------
 * Replaces, for chunks whose text was parsed on the device, the record loop of make_fastq and the zstd encoder behind
 * it (python/src/lib.rs:136-168, seqspec/src/utils.rs:33-36), i.e. what pcl_make_fastq_batch + a host zstd do for
 * the host reader: the R1 (and R2) records of a chunk are assembled in HBM from the kept text and the result planes
 * and, with `compress`, coded there into one standard zstd frame per file and call (RFC 8878: 128 KB blocks of
 * Huffman-coded literals plus sequences - matches of every line against the line four lines earlier, FSE-coded with
 * the predefined distributions; any zstd decoder reads them; files come out about 1.3-1.4 x the size of level 9).
 *   pcl_text_keep(chunk, slot)   after pcl_text_fill: the chunk keeps its own copy of the consumed text and line
 *                                index (device to device, asynchronous) - the staging buffers move on to the next chunk
 *   pcl_emit_fastq(chunk, correct_barcode, paired, compress, &res)
 *                                after pcl_count (and pcl_correct when correct_barcode is set); blocks.  res.r1 / res.r2
 *                                point at pinned host memory owned by the ctx holding the bytes to append to R1.fq(.zst)
 *                                / R2.fq(.zst); they stay valid until the SECOND next pcl_emit_fastq on the ctx (two
 *                                buffers alternate, so one call's bytes can be written to disk during the next call).
 *                                Errors mirror the reference's panics (PCL_ERR_INPUT: a group with two targets, no read1 ...). */
typedef struct {
    const uint8_t *r1;           /* NULL when nothing was written                               */
    uint64_t r1_bytes;           /* bytes at r1 (a whole zstd frame when compress != 0)         */
    uint64_t r1_raw_bytes;       /* bytes of FASTQ text they hold                               */
    const uint8_t *r2;
    uint64_t r2_bytes;
    uint64_t r2_raw_bytes;
    uint64_t n_written;          /* read groups written                                         */
} pcl_emit_result;
int pcl_text_keep(pcl_chunk *chunk, int read_slot);
int pcl_emit_fastq(pcl_chunk *chunk, int correct_barcode, int paired, int compress, pcl_emit_result *out);
/* The block coder alone: n < 4 GiB host bytes -> one zstd frame in dst (at most 14 + n + 3 * (n / 131072 + 1) bytes);
 * what zstd::stream::Encoder does for create_file (seqspec/src/utils.rs:20-42), on the device.  Blocks. */
int pcl_zstd_frame(pcl_ctx *ctx, const uint8_t *src, uint64_t n, uint8_t *dst, uint64_t cap, uint64_t *used);

/* ---- decompressed byte stream of one gzip file (feeds pcl_text_scan) ---------------------------------
 * Replaces open_file's MultiGzDecoder (seqspec/src/utils.rs:44-56) for the text path.  BGZF files (independent
 * <= 64 KB members with their sizes in the header) are inflated by n_threads threads straight into the caller's
 * buffer; any other gzip file (single or multi-member) is one sequential zlib stream.
 * pcl_gzsrc_read returns the bytes produced (0 at the end of the file) or a negative pcl_status. */
typedef struct pcl_gzsrc pcl_gzsrc;
int pcl_gzsrc_open(const char *path, int n_threads, pcl_gzsrc **out);
int64_t pcl_gzsrc_read(pcl_gzsrc *src, uint8_t *dst, uint64_t cap);
int pcl_gzsrc_is_bgzf(const pcl_gzsrc *src);
const char *pcl_gzsrc_error(const pcl_gzsrc *src);
void pcl_gzsrc_close(pcl_gzsrc *src);

/* ---- BGZF members inflated on the device (csrc/inflate.cuh, inflate_core.h) -------------------------------
 * Replaces open_file's MultiGzDecoder (seqspec/src/utils.rs:44-56; the reference inflates every file on one host thread,
 * twice per run: seqspec/src/read.rs:83-155) for BGZF input: the host walks the member headers (pcl_bgzf_walk), the
 * compressed bytes go to the device as they are and one warp inflates each member - CRC-32 and ISIZE of the trailer
 * checked - into a text block that pcl_text_scan_device indexes without a trip through host memory.
 *
 * pcl_bgzf_walk      whole members at the front of buf[0..n_bytes): stops once they inflate to >= want_out bytes, at
 *                    `cap` members, or at the first incomplete / non-BGZF member.  Members that inflate to nothing (the
 *                    end-of-file marker) are consumed, not listed.  src_off is relative to buf, dst_off counts from 0.
 *                    Returns PCL_ERR_INPUT when the bytes at buf[*consumed] are complete but no BGZF header.
 * pcl_inflater_*     one handle per producer thread (own stream and staging; calls on one handle are not re-entrant,
 *                    handles of one ctx may be driven from different threads).  A handle owns up to 4 text buffers
 *                    ("slots") in device memory; pcl_inflater_buffer (re)allocates one (bytes = 0 frees it) and returns
 *                    its address.  pcl_inflater_run copies comp[0..comp_bytes) (pageable or pinned host memory) to the
 *                    device and inflates the members to slot bytes [dst_off + member.dst_off, ...); it blocks, and a
 *                    damaged member is PCL_ERR_INPUT (pcl_inflater_error names the member and the defect).  move / fetch /
 *                    poke copy device-to-device (between or inside slots, ranges must not overlap), device-to-host and
 *                    host-to-device; all block.  pcl_inflater_stats: kernel time (CUDA events), inflated bytes, launches. */
typedef struct pcl_bgzf_member {
    uint32_t src_off;            /* first byte of the member's DEFLATE stream                    */
    uint32_t clen;               /* length of the stream                                         */
    uint32_t dst_off;            /* offset of the member's bytes in the inflated text            */
    uint32_t isize;              /* ISIZE of the trailer                                         */
    uint32_t crc;                /* CRC32 of the trailer                                         */
} pcl_bgzf_member;
int pcl_bgzf_walk(const uint8_t *buf, uint64_t n_bytes, uint64_t want_out, pcl_bgzf_member *members, uint32_t cap,
                  uint32_t *n_members, uint64_t *consumed, uint64_t *out_bytes);
typedef struct pcl_inflater pcl_inflater;
int pcl_inflater_create(pcl_ctx *ctx, pcl_inflater **out);
void pcl_inflater_destroy(pcl_inflater *inf);
const char *pcl_inflater_error(const pcl_inflater *inf);
int pcl_inflater_buffer(pcl_inflater *inf, int slot, uint64_t bytes, void **device_ptr);
int pcl_inflater_run(pcl_inflater *inf, const uint8_t *comp, uint64_t comp_bytes, const pcl_bgzf_member *members,
                     uint32_t n_members, int slot, uint64_t dst_off);
int pcl_inflater_move(pcl_inflater *inf, int src_slot, uint64_t src_off, int dst_slot, uint64_t dst_off, uint64_t n_bytes);
int pcl_inflater_fetch(pcl_inflater *inf, int slot, uint64_t off, uint64_t n_bytes, void *host_dst);
int pcl_inflater_poke(pcl_inflater *inf, int slot, uint64_t off, const void *host_src, uint64_t n_bytes);
int pcl_inflater_stats(const pcl_inflater *inf, double *kernel_ms, uint64_t *out_bytes, uint64_t *launches);
/* pcl_text_scan for a block that is already in device memory (a slot of a pcl_inflater of the same ctx). */
int pcl_text_scan_device(pcl_chunk *chunk, int read_slot, const void *device_text, uint64_t n_bytes);

/* ---- per-barcode grouping of a chunk ---------------------------------------------
 * The first data-parallel step behind the aligner, keyed on this path's output:
 *   sort_by_barcode + chunk_by(barcode)          precellar/src/fragment/mod.rs:359-363,434-455
 *   RemoveDuplicates::remove_duplicates          precellar/src/fragment/dedups.rs:320-340 (FingerPrint: dedups.rs:52-66)
 * pcl_group_build (after pcl_count / pcl_correct) sorts the read groups of the chunk whose barcode is assigned -- every
 * present barcode segment EXACT or CORRECTED, i.e. the read carries a CB tag made of whitelist entries -- stably on
 * (barcode, fingerprint, UMI) with a radix sort on the device.  (Raw barcodes that are not whitelist members may hold
 * non-ACGT bytes, which the packed keys do not keep: they are not grouped.)
 *   fingerprint  one 64-bit word per read group (host pointer, n values) with whatever the aligner contributes to
 *                dedups.rs's FingerPrint (reference id, 5' coordinates, orientation), or NULL: barcode + UMI only
 *   order        [n_assigned]       record indices in sorted order; ties keep input order
 *   frag_start   [n_fragments + 1]  positions in `order` where a run of equal (barcode, fingerprint, UMI) starts: a
 *                duplicate set.  order[frag_start[f]] is the record the reference keeps, frag_start[f+1] - frag_start[f]
 *                its Fragment.count
 *   bc_frag_start[n_barcodes + 1]   first fragment of every barcode, barcodes in the reference's (byte-string) order
 *   bc_key       [n_barcodes]       the barcode: 2 bits per base left-aligned in bits 63..6, its length in bits 4..0
 *                (pcl_group_key_decode); multi-segment barcodes are concatenated in join order
 * UMIs are compared as byte strings over A C G T N; a read group whose UMI holds any other byte is not grouped
 * (n_ungroupable).  Barcodes longer than 29 bases and UMIs longer than 19 are PCL_ERR_UNSUPPORTED / ungroupable. */
typedef struct pcl_group pcl_group;
typedef struct { uint64_t n_records, n_assigned, n_ungroupable, n_barcodes, n_fragments; } pcl_group_info;
int pcl_group_build(pcl_ctx *ctx, pcl_chunk *chunk, const uint64_t *fingerprint, pcl_group **out);
/* The same over several chunks of one ctx (a run that did not fit one chunk): read group i of chunks[k] is record
 * (n_0 + ... + n_{k-1}) + i of the grouping, `fingerprint` (or NULL) holds one word per record in that order, and
 * `order` refers to those record numbers.  Fewer than 2^32 - 1 read groups in total. */
int pcl_group_build_multi(pcl_ctx *ctx, pcl_chunk *const *chunks, int n_chunks, const uint64_t *fingerprint, pcl_group **out);
int pcl_group_info_get(const pcl_group *g, pcl_group_info *out);
int pcl_group_fetch(pcl_group *g, uint32_t *order, uint32_t *frag_start, uint32_t *bc_frag_start, uint64_t *bc_key);
int pcl_group_key_decode(uint64_t key, char *out32);
int pcl_group_destroy(pcl_group *g);

/* ---- pinned host memory ------------------------------------------------------------------- */
void *pcl_host_alloc(uint64_t bytes);
void pcl_host_free(void *p);

/* ---- instrumentation (used by bench.py; CUDA events on the launching stream) --------------- */
enum { PCL_K_COUNT = 0, PCL_K_CORRECT = 1, PCL_K_COUNT_LENONLY = 2, PCL_K_ENUM = 3, PCL_K_QC = 4, PCL_K_TEXT_INDEX = 5,
       PCL_K_TEXT_FILL = 6, PCL_K_FILTER = 7, PCL_K_EMIT = 8, PCL_K_ZSTD = 9, PCL_K_MAX = 16 };
int pcl_profile_enable(pcl_ctx *ctx, int on);                    /* bracket every kernel with events */
/* accumulated ms and launches of kernel k since the last pcl_profile_reset (syncs the ctx) */
int pcl_profile_get(pcl_ctx *ctx, int k, double *ms, uint64_t *launches);
int pcl_profile_reset(pcl_ctx *ctx);
uint64_t pcl_launch_count(const pcl_ctx *ctx);                   /* kernels launched by this ctx     */
/* cudaEvent-based stopwatch on a chunk's stream */
int pcl_timer_start(pcl_chunk *chunk);
int pcl_timer_stop_ms(pcl_chunk *chunk, double *ms);             /* blocks                           */

#ifdef __cplusplus
}
#endif
#endif /* PRECELLAR_B200_H */
