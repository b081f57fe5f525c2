// v3_demo.cpp -- the hot path driven from compiled code through the C ABI only (no Python, no torch):
//   FASTQ files -> pcl_fastq_read (pinned SoA planes) -> pcl_chunk_upload_compact -> pcl_count -> pcl_counts_allreduce
//   -> pcl_correct -> pcl_chunk_fetch, for a 10x-v3-like layout (R1 = 16 nt barcode + 12 nt UMI, R2 = insert).
// It is what the body of FastqProcessor::gen_barcoded_fastq (precellar/src/align/fastq.rs:109-153) turns into when the
// reference binds libprecellar_b200.so; tests/test_gpu_cabi_demo.py runs it against the Python host layer.
//
//   g++ -O2 -std=c++17 -I include examples/v3_demo.cpp -o examples/v3_demo -L precellar_b200 -lprecellar_b200 \
//       -Wl,-rpath,'$ORIGIN/../precellar_b200'
//   examples/v3_demo R1.fq[.gz] R2.fq[.gz] whitelist.txt [threshold] [chunk_records]
//
// Prints one JSON line: records, status counts, whitelist totals and an order-dependent checksum of the result planes.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>

#include "precellar_b200.h"

#define CHECK(ctx, call)                                                                      \
    do {                                                                                      \
        int rc_ = (call);                                                                     \
        if (rc_ < 0) { fprintf(stderr, "%s -> %d: %s\n", #call, rc_, pcl_last_error(ctx)); return 1; } \
    } while (0)

static uint64_t mix(uint64_t h, uint64_t v) { h ^= v + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2); return h; }

int main(int argc, char **argv) {
    if (argc < 4) { fprintf(stderr, "usage: %s R1.fq R2.fq whitelist.txt [threshold] [chunk_records]\n", argv[0]); return 2; }
    const double threshold = argc > 4 ? atof(argv[4]) : 0.975;
    const uint64_t chunk_records = argc > 5 ? strtoull(argv[5], nullptr, 10) : (1u << 20);

    // region table of the two reads (what Assay::get_segments_by_modality derives for 10x v3)
    pcl_read_desc reads[2];
    memset(reads, 0, sizeof reads);
    reads[0].n_elems = 2; reads[0].max_len = 28;
    reads[0].elems[0].kind = PCL_KIND_BARCODE; reads[0].elems[0].min_len = reads[0].elems[0].max_len = 16; reads[0].elems[0].region_slot = 0;
    reads[0].elems[1].kind = PCL_KIND_UMI; reads[0].elems[1].min_len = reads[0].elems[1].max_len = 12; reads[0].elems[1].region_slot = -1;
    reads[1].n_elems = 1; reads[1].is_reverse = 1; reads[1].max_len = 0;
    reads[1].elems[0].kind = PCL_KIND_TARGET; reads[1].elems[0].min_len = 1; reads[1].elems[0].max_len = 1000; reads[1].elems[0].region_slot = -1;

    pcl_ctx *ctx = nullptr;
    if (int rc = pcl_ctx_create(&ctx, 0); rc < 0) { fprintf(stderr, "pcl_ctx_create -> %d: %s\n", rc, pcl_last_error(nullptr)); return 1; }
    CHECK(ctx, pcl_layout_set(ctx, reads, 2));

    // whitelist: one barcode per line, file order (Onlist::read, seqspec/src/region.rs:459-469)
    std::vector<uint8_t> blob;
    std::vector<uint64_t> offs{0};
    {
        std::ifstream in(argv[3]);
        std::string line;
        while (std::getline(in, line)) {
            if (!line.empty() && line.back() == '\r') line.pop_back();
            if (line.empty()) continue;
            blob.insert(blob.end(), line.begin(), line.end());
            offs.push_back(blob.size());
        }
    }
    CHECK(ctx, pcl_whitelist_load(ctx, 0, blob.data(), offs.data(), offs.size() - 1));

    const uint32_t row1 = pcl_read_payload_bytes(ctx, 0);                 // 28: the bytes of R1 the layout uses
    pcl_read_info info1;
    CHECK(ctx, pcl_read_info_get(ctx, 0, &info1));
    const uint32_t qoff = info1.qual_offset, qstride = info1.qual_stride;  // the 16 barcode qualities

    pcl_fastq *f1 = nullptr, *f2 = nullptr;
    const char *p1[1] = {argv[1]}, *p2[1] = {argv[2]};
    if (pcl_fastq_open(p1, 1, &f1) < 0 || pcl_fastq_open(p2, 1, &f2) < 0) { fprintf(stderr, "cannot open the FASTQ files\n"); return 1; }

    struct Host { uint8_t *seq, *qual; uint16_t *len1, *len2; uint64_t *nh1, *nh2; uint64_t n; pcl_chunk *chunk; };
    std::vector<Host> batches;
    uint64_t total = 0;
    for (;;) {                                                            // pass 1: read, upload, count (BarcodeAnalyzer::new)
        Host h{};
        h.seq = (uint8_t *)pcl_host_alloc(chunk_records * row1);
        h.qual = (uint8_t *)pcl_host_alloc(chunk_records * qstride);
        h.len1 = (uint16_t *)pcl_host_alloc(chunk_records * 2);
        h.len2 = (uint16_t *)pcl_host_alloc(chunk_records * 2);
        h.nh1 = (uint64_t *)malloc(chunk_records * 8);
        h.nh2 = (uint64_t *)malloc(chunk_records * 8);
        const int64_t n1 = pcl_fastq_read(f1, chunk_records, row1, qoff, qstride, h.seq, h.qual, h.len1, h.nh1, nullptr, 0, nullptr, nullptr);
        const int64_t n2 = pcl_fastq_read(f2, chunk_records, 0, 0, 0, nullptr, nullptr, h.len2, h.nh2, nullptr, 0, nullptr, nullptr);
        if (n1 < 0 || n2 < 0) { fprintf(stderr, "FASTQ error: %s%s\n", pcl_fastq_error(f1), pcl_fastq_error(f2)); return 1; }
        if (n1 != n2) { fprintf(stderr, "Unequal number of reads in the chunk\n"); return 1; }       // fastq.rs:435-436
        if (n1 == 0) break;
        for (int64_t i = 0; i < n1; ++i)
            if (h.nh1[i] != h.nh2[i]) { fprintf(stderr, "read names mismatch (group %llu)\n", (unsigned long long)(total + i)); return 1; }
        h.n = (uint64_t)n1;
        CHECK(ctx, pcl_chunk_create(ctx, h.n, &h.chunk));
        CHECK(ctx, pcl_chunk_reset(h.chunk, h.n));
        CHECK(ctx, pcl_chunk_upload_compact(h.chunk, 0, h.seq, row1, h.qual, h.len1));
        CHECK(ctx, pcl_chunk_upload(h.chunk, 1, nullptr, nullptr, h.len2));
        CHECK(ctx, pcl_count(ctx, h.chunk));
        batches.push_back(h);
        total += h.n;
    }
    CHECK(ctx, pcl_counts_allreduce(ctx));                                // a no-op on one GPU: the event hand-over between the passes

    uint64_t n_code[8] = {0}, n_target = 0, checksum = 0;
    for (Host &h : batches) {                                             // pass 2: correct and fetch (AnnotatedFastqReader)
        CHECK(ctx, pcl_correct(ctx, h.chunk, threshold, 1.7976931348623157e308));
        std::vector<uint16_t> fs1(h.n), fs2(h.n);
        std::vector<uint32_t> bck(h.n), umi(h.n), tc(h.n);
        pcl_read_results r1, r2;
        memset(&r1, 0, sizeof r1); memset(&r2, 0, sizeof r2);
        r1.fstatus = fs1.data(); r1.bc_key[0] = bck.data(); r1.umi_key[0] = umi.data();
        r2.fstatus = fs2.data(); r2.tcoord = tc.data();
        CHECK(ctx, pcl_chunk_fetch(h.chunk, 0, &r1));
        CHECK(ctx, pcl_chunk_fetch(h.chunk, 1, &r2));
        for (uint64_t i = 0; i < h.n; ++i) {
            const uint32_t code = PCL_BC_CODE(fs1[i], 0);
            ++n_code[code];
            n_target += (fs2[i] & PCL_FSTATUS_TARGET) != 0;
            const bool assigned = code == PCL_BC_EXACT || code == PCL_BC_CORRECTED;
            checksum = mix(checksum, fs1[i]); checksum = mix(checksum, assigned ? bck[i] : 0u);
            checksum = mix(checksum, umi[i]); checksum = mix(checksum, fs2[i]); checksum = mix(checksum, tc[i]);
        }
        pcl_chunk_destroy(h.chunk);
        pcl_host_free(h.seq); pcl_host_free(h.qual); pcl_host_free(h.len1); pcl_host_free(h.len2); free(h.nh1); free(h.nh2);
    }
    std::vector<uint64_t> counts(offs.size() - 1);
    uint64_t mismatch = 0, tot = 0, csum = 0;
    CHECK(ctx, pcl_counts_fetch(ctx, 0, counts.data(), &mismatch, &tot));
    for (uint64_t c : counts) csum = mix(csum, c);
    printf("{\"records\": %llu, \"exact\": %llu, \"corrected\": %llu, \"no_match\": %llu, \"low_confidence\": %llu, \"absent\": %llu, "
           "\"targets\": %llu, \"whitelist_total\": %llu, \"whitelist_mismatch\": %llu, \"counts_checksum\": %llu, \"results_checksum\": %llu, "
           "\"kernels_launched\": %llu}\n",
           (unsigned long long)total, (unsigned long long)n_code[PCL_BC_EXACT], (unsigned long long)n_code[PCL_BC_CORRECTED],
           (unsigned long long)n_code[PCL_BC_NO_MATCH], (unsigned long long)n_code[PCL_BC_LOW_CONF], (unsigned long long)n_code[PCL_BC_ABSENT],
           (unsigned long long)n_target, (unsigned long long)tot, (unsigned long long)mismatch, (unsigned long long)csum,
           (unsigned long long)checksum, (unsigned long long)pcl_launch_count(ctx));
    pcl_fastq_close(f1); pcl_fastq_close(f2);
    pcl_ctx_destroy(ctx);
    return 0;
}
