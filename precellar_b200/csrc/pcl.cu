// pcl.cu -- implementation of the C ABI in include/precellar_b200.h (sm_100a only).
//
// Host side: context / layout / whitelist-table construction / chunk memory / stream ordering.
// Device side: kernels.cuh.  No CPU fallback: compute entry points need a CUDA device.

#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "host_build.h"
#include "kernels.cuh"
#include "textparse.cuh"
#include "group.cuh"
#include "tablebuild.cuh"
#include "emit.cuh"
#include "inflate.cuh"

// ------------------------------------------------------------------------------------------
// NCCL through dlopen: the library has no link-time dependency on libnccl, and inside a process
// that already loaded NCCL (e.g. through torch) the same copy is reused.
// ------------------------------------------------------------------------------------------
namespace {

struct NcclUid { char internal[128]; };
typedef void *NcclComm;
struct NcclApi {
    void *h = nullptr;
    int (*GetUniqueId)(NcclUid *) = nullptr;
    int (*CommInitRank)(NcclComm *, int, NcclUid, int) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*CommDestroy)(NcclComm) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    bool ok = false;
};
const int kNcclUint32 = 3, kNcclUint64 = 5, kNcclSum = 0;

NcclApi &nccl() {
    static NcclApi api;
    static bool tried = false;
    if (tried) return api;
    tried = true;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
        api.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (api.h) break;
    }
    if (!api.h) return api;
    api.GetUniqueId = (int (*)(NcclUid *))dlsym(api.h, "ncclGetUniqueId");
    api.CommInitRank = (int (*)(NcclComm *, int, NcclUid, int))dlsym(api.h, "ncclCommInitRank");
    api.AllReduce = (int (*)(const void *, void *, size_t, int, int, NcclComm, cudaStream_t))dlsym(api.h, "ncclAllReduce");
    api.GroupStart = (int (*)())dlsym(api.h, "ncclGroupStart");
    api.GroupEnd = (int (*)())dlsym(api.h, "ncclGroupEnd");
    api.CommDestroy = (int (*)(NcclComm))dlsym(api.h, "ncclCommDestroy");
    api.GetErrorString = (const char *(*)(int))dlsym(api.h, "ncclGetErrorString");
    api.ok = api.GetUniqueId && api.CommInitRank && api.AllReduce && api.GroupStart && api.GroupEnd && api.CommDestroy;
    return api;
}

thread_local std::string g_create_err;

struct Table {
    bool loaded = false;
    DTable d{};
    uint64_t n_entries = 0;
    uint64_t n_slots = 0;
    void *d_keys = nullptr;
    uint32_t *d_counts = nullptr;
    uint32_t *d_index_to_slot = nullptr;
    uint32_t *d_nkeys[4] = {nullptr, nullptr, nullptr, nullptr};
    uint32_t *d_nbits = nullptr;               // 4 x PCL_NBITS_WORDS: bitmaps of the partial keys (DTable::nbits)
    uint32_t *d_dense = nullptr;               // n_entries counts of the occupied slots: what the all-reduce sends
    uint32_t *d_order_slots = nullptr;         // the occupied slots, (nearly) ascending: the order of d_dense, the same on every rank
    uint32_t *d_scratch_counts = nullptr;      // pcl_counts_freeze: where a re-run of pass 1 puts its (discarded) increments
    unsigned long long *d_scratch_scalars = nullptr;
    unsigned long long *d_scalars = nullptr;   // [0] total, [1] mismatch
};

struct EvPair { cudaEvent_t a, b; int k; };

}  // namespace

struct pcl_chunk;

// Staging of one read file's text block (pcl_text_*): shared by the chunks of a ctx, one block per read file in
// flight (pcl_text_fill blocks, so the next pcl_text_scan of the same read file finds it free).
struct TextStage {
    uint8_t *d_text = nullptr;
    uint32_t *d_nl = nullptr, *d_tile_cnt = nullptr, *d_tile_off = nullptr;
    uint64_t text_cap = 0, nl_cap = 0, text_bytes = 0;
    pcl_chunk *owner = nullptr;
    // small result words and name hashes of the parse in flight: owned by the stage (one block per read file at a time),
    // so that creating a chunk costs no pinned or device allocation on this path
    uint64_t *d_hash = nullptr;
    uint64_t hash_cap = 0;
    unsigned long long *d_tmeta = nullptr;    // device: [0] newline total, [1] first bad record << 8 | code
    unsigned long long *h_tmeta = nullptr;    // pinned: [0] total, [1] err, [2] end of the last consumed line
};

// Scratch of pcl_emit_fastq, shared by the chunks of a ctx (the call blocks; one emission in flight per ctx).  The pinned
// result buffers alternate, so the bytes of one call stay valid during the next one (the caller writes a file meanwhile).
struct EmitStage {
    uint32_t *d_sz[2] = {nullptr, nullptr}, *d_off[2] = {nullptr, nullptr}, *d_scan_tmp = nullptr;
    uint64_t rec_cap = 0;
    unsigned long long *d_meta = nullptr;        // [0..1] raw bytes of R1 / R2, [2] first error, [3..4] stream bytes, [5] groups written
    unsigned long long *h_meta = nullptr;        // pinned copy
    uint8_t *d_raw[2] = {nullptr, nullptr};
    uint64_t raw_cap[2] = {0, 0};
    uint8_t *d_slots[2] = {nullptr, nullptr}, *d_stream[2] = {nullptr, nullptr};
    uint32_t *d_bsize[2] = {nullptr, nullptr};
    ZhScratch *d_scratch[2] = {nullptr, nullptr};
    uint64_t block_cap[2] = {0, 0};
    uint8_t *h_out[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};     // [flip][file]
    uint64_t h_cap[2][2] = {{0, 0}, {0, 0}};
    int flip = 0;
};

struct pcl_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_reduced = nullptr;
    std::string err;
    int n_reads = 0;
    pcl_read_desc descs[PCL_MAX_READS];
    DRead dreads[PCL_MAX_READS];
    pcl_read_info infos[PCL_MAX_READS];
    int bc_region[PCL_MAX_READS][PCL_MAX_BC_PER_READ];
    Table tables[PCL_MAX_REGIONS];
    double *d_lut = nullptr;                  // [0..255] capped, [256..511] raw
    unsigned long long *d_qc = nullptr;       // [PCL_MAX_READS] num_defect, then [12] QcFastq category counters
    bool qc_enabled = false;
    bool verify_mode = false;                 // pcl_verify_enable: Assay::verify's split_with_tolerance(read, 0.2, 0.2)
    uint64_t qc_reads[PCL_MAX_READS] = {0};
    std::vector<pcl_chunk *> chunks;
    TextStage text[PCL_MAX_READS];
    EmitStage emit;
    // pcl_text_keep: the chunks' kept text and line indices are slices of a few large slabs (a cudaMalloc per chunk and
    // read file costs more than the copies); the slabs are recycled when the last chunk holding a slice goes away
    struct KeepSlab { uint8_t *p; uint64_t cap, used; };
    std::vector<KeepSlab> keep_slabs;
    uint64_t keep_live = 0;
    // nccl
    NcclComm comm = nullptr;
    int nranks = 1, rank = 0;
    bool frozen = false;          // pcl_counts_freeze: pcl_count recomputes a chunk's outputs without touching counts / totals / QC
    unsigned long long *d_qc_scratch = nullptr;
    bool reduced = false;         // counts hold the sum over ranks: pcl_count / a second all-reduce need pcl_counts_reset first
    // instrumentation
    bool profile = false;
    std::vector<EvPair> evs;
    double prof_ms[PCL_K_MAX] = {0};
    uint64_t prof_n[PCL_K_MAX] = {0};
    uint64_t launches = 0;
    int sm_count = 148;
    int occ[16] = {1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1};   // resident CTAs per SM of each kernel (see OCC_*)
    int fast_bits = 14, fast_waves = 1;
    uint32_t hist_admit_mask = 7, hist_two_choice = 1;
    bool no_fast_spec = false;    // PCL_NO_FAST_SPEC=1: the mask-driven instance of k_count_fast for every FastLayout
    bool no_prefetch = false;     // PCL_NO_PREFETCH=1: k_count_spec (v3 forward instance) without the L2 prefetch of the next rows (experiments)
    bool old_fast = false;        // PCL_OLD_FAST=1: the software-pipelined k_count_fast instances instead of k_count_spec
    bool no_plan = false;         // PCL_NO_PLAN=1: AnchorPlan layouts go through the interpreter kernels (tests compare both)
    bool force_generic = false;   // PCL_FORCE_GENERIC=1: only the generic kernels (tests compare both paths)
    bool no_neighbour_tables = false;   // PCL_NO_NEIGHBOUR_TABLES=1: 3*L direct probes instead of the 4 neighbourhood tables
    uint64_t device_build_min = 100000; // whitelists from this many entries on are deduplicated and hashed on the device (PCL_DEVICE_BUILD_MIN)
    bool reduce_after_filter = false;   // PCL_REDUCE_AFTER_FILTER=1: the all-reduce waits for k_filter instead of overlapping it (experiments)
    // PCL_TIMELINE=1 (diagnostic): events at the hand-over points of a step; pcl_correct prints their offsets to stderr
    bool timeline = false;
    cudaEvent_t tl[12] = {};
    bool tl_set[12] = {};
    int filter_waves = 16;              // PCL_FILTER_WAVES: resident waves of k_filter (see launch_filter)
    bool reduce_slots = false;          // PCL_REDUCE_SLOTS=1: all-reduce the whole slot array instead of the occupied slots
    bool no_plan_geo = false;           // PCL_NO_PLAN_GEO=1: AnchorPlan layouts always run the run-time geometry (tests compare the two)
    bool no_nbits = false;              // PCL_NO_NBITS=1: no partial-key bitmaps in front of the neighbourhood tables
    bool zh_no_match = false;           // PCL_ZH_NO_MATCH=1: the device zstd coder emits literals only (no sequences; experiments)
};

struct ReadBuf {
    uint8_t *seq = nullptr, *qual = nullptr;
    uint16_t *len = nullptr;
    uint16_t *fstatus = nullptr;
    uint32_t *tcoord = nullptr;
    void *bc_key[PCL_MAX_BC_PER_READ] = {nullptr};
    void *umi_key[PCL_MAX_UMI_PER_READ] = {nullptr};
    uint32_t *bcoord[PCL_MAX_BC_PER_READ] = {nullptr};
    uint32_t *ucoord[PCL_MAX_UMI_PER_READ] = {nullptr};
    unsigned long long *worklist = nullptr;
    unsigned long long *wl_count = nullptr;
    uint32_t *wl_key = nullptr;               // reads whose tables are all 32-bit: packed key and candidate mask per entry
    unsigned long long *wl_cand = nullptr;
    bool filtered = false;                    // k_filter has run on the current worklist
    bool uniform_len = false;                 // pcl_chunk_set_uniform_len since the last reset: uploads may pass len == NULL
    uint32_t *d_uniform = nullptr;            // pcl_chunk_uniform_results_async: flag, fstatus, tcoord
    uint8_t *seq_stage = nullptr;             // pcl_chunk_upload_compact: the rows as uploaded, before k_repack_rows
    bool filled = false;
    // device-side text parse (pcl_text_*): name hashes and the small result words of this chunk's parse
    uint64_t *d_hash = nullptr;               // = the stage's, while this chunk owns it
    unsigned long long *d_tmeta = nullptr;
    unsigned long long *h_tmeta = nullptr;
    uint64_t text_consumed = 0;               // bytes of the text block the filled records took
    uint64_t text_lines = 0;
    bool text_scanned = false, text_filled = false;
    // pcl_text_keep: this chunk's own copy of the consumed text and its line index (the staging buffers move on)
    uint8_t *k_text = nullptr;                // slices of ctx->keep_slabs
    uint32_t *k_nl = nullptr;
    uint64_t k_text_cap = 0, k_nl_cap = 0, k_records = 0;
    bool text_kept = false;
};

struct pcl_chunk {
    pcl_ctx *ctx = nullptr;
    uint64_t cap = 0, n = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev = nullptr, t0 = nullptr, t1 = nullptr;
    cudaEvent_t ev_counted = nullptr;    // recorded by pcl_count behind the last kernel that touches whitelist counts
    bool counted = false;
    bool has_results = false;            // pcl_count has run on the records of the last reset
    ReadBuf rb[PCL_MAX_READS];
    std::vector<void *> allocs;
};

static void keep_release(pcl_ctx *ctx);

namespace {

void emit_stage_free(EmitStage &e) {
    for (int f = 0; f < 2; ++f) {
        cudaFree(e.d_sz[f]); cudaFree(e.d_off[f]); cudaFree(e.d_raw[f]); cudaFree(e.d_slots[f]); cudaFree(e.d_stream[f]); cudaFree(e.d_bsize[f]);
        cudaFree(e.d_scratch[f]);
        for (int k = 0; k < 2; ++k) if (e.h_out[k][f]) cudaFreeHost(e.h_out[k][f]);
    }
    cudaFree(e.d_scan_tmp); cudaFree(e.d_meta);
    if (e.h_meta) cudaFreeHost(e.h_meta);
    e = EmitStage();
}

int fail(pcl_ctx *ctx, int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf; else g_create_err = buf;
    return code;
}

#define CU(ctx, call)                                                                               \
    do {                                                                                            \
        cudaError_t e_ = (call);                                                                    \
        if (e_ != cudaSuccess) return fail(ctx, PCL_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); \
    } while (0)

int bind(pcl_ctx *ctx) {
    cudaError_t e = cudaSetDevice(ctx->device);
    if (e != cudaSuccess) return fail(ctx, PCL_ERR_CUDA, "cudaSetDevice(%d): %s", ctx->device, cudaGetErrorString(e));
    return PCL_OK;
}

// error_probability (precellar/src/barcode.rs:464-468) with the host libm, as the reference does.
double error_probability(int q) { return std::pow(10.0, -(((double)q - 33.0) / 10.0)); }

void prof_begin(pcl_ctx *ctx, cudaStream_t s, int k) {
    ctx->launches += 1;
    if (!ctx->profile) return;
    EvPair p;
    cudaEventCreate(&p.a);
    cudaEventCreate(&p.b);
    p.k = k;
    cudaEventRecord(p.a, s);
    ctx->evs.push_back(p);
}
void prof_end(pcl_ctx *ctx, cudaStream_t s) {
    if (!ctx->profile) return;
    cudaEventRecord(ctx->evs.back().b, s);
}

int device_scan(pcl_ctx *ctx, cudaStream_t s, const uint32_t *in, uint32_t *out, uint64_t n, uint32_t *tmp, unsigned long long *d_total);
int device_radix_sort(pcl_ctx *ctx, cudaStream_t s, unsigned long long **keys, uint32_t **vals, unsigned long long **keys_tmp,
                      uint32_t **vals_tmp, uint64_t n, uint32_t byte_mask, uint32_t *hist, uint32_t *scan_tmp, unsigned long long *d_scratch);
uint32_t bytes_used(unsigned long long or_of_keys);

enum { OCC_COUNT = 0, OCC_COUNT_FAST, OCC_COUNT_LENONLY, OCC_ENUM, OCC_FILTER, OCC_CORRECT, OCC_QC, OCC_GATHER, OCC_COUNT_CODES, OCC_COUNT_PLAN, OCC_RESOLVE, OCC_COUNT_PLAN2 };

// every barcode of read r goes through the 32-bit tables (k_filter / k_resolve)
bool read_all32(const pcl_ctx *ctx, int r) {
    const pcl_read_info &inf = ctx->infos[r];
    if (ctx->force_generic || inf.n_bc == 0) return false;
    for (int j = 0; j < inf.n_bc; ++j) {
        const Table &t = ctx->tables[ctx->bc_region[r][j]];
        if (!t.loaded || t.d.mode == PCL_TAB_K64 || inf.bc_key_bytes[j] != 4) return false;
    }
    return true;
}

int grid_for(const pcl_ctx *ctx, uint64_t n_threads_wanted, int blocks_per_sm) {
    uint64_t blocks = (n_threads_wanted + PCL_BLOCK - 1) / PCL_BLOCK;
    uint64_t cap = (uint64_t)ctx->sm_count * blocks_per_sm;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

}  // namespace

extern "C" {

int pcl_abi_version(void) { return PCL_ABI_VERSION; }

const char *pcl_last_error(const pcl_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

int pcl_device_count(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) return fail(nullptr, PCL_ERR_CUDA, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
    return n;
}

int pcl_ctx_create(pcl_ctx **out, int device) {
    if (!out) return fail(nullptr, PCL_ERR_ARG, "out is NULL");
    *out = nullptr;
    int n = pcl_device_count();
    if (n < 0) return n;
    if (n == 0) return fail(nullptr, PCL_ERR_CUDA, "no CUDA device: this library has no CPU fallback");
    if (device < 0 || device >= n) return fail(nullptr, PCL_ERR_ARG, "device %d out of range (0..%d)", device, n - 1);
    cudaDeviceProp prop;
    cudaError_t e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) return fail(nullptr, PCL_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    if (prop.major != 10)
        return fail(nullptr, PCL_ERR_CUDA, "device %d is sm_%d%d; this build contains sm_100a code only", device, prop.major, prop.minor);
    pcl_ctx *ctx = new pcl_ctx();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    { const char *g = getenv("PCL_FORCE_GENERIC"); ctx->force_generic = g && g[0] == '1'; }
    { const char *g = getenv("PCL_NO_PLAN"); ctx->no_plan = g && g[0] == '1'; }
    { const char *g = getenv("PCL_NO_FAST_SPEC"); ctx->no_fast_spec = g && g[0] == '1'; }
    { const char *g = getenv("PCL_NO_NEIGHBOUR_TABLES"); ctx->no_neighbour_tables = g && g[0] == '1'; }
    { const char *g = getenv("PCL_NO_NBITS"); ctx->no_nbits = g && g[0] == '1'; }
    { const char *g = getenv("PCL_ZH_NO_MATCH"); ctx->zh_no_match = g && g[0] == '1'; }
    { const char *g = getenv("PCL_NO_PLAN_GEO"); ctx->no_plan_geo = g && g[0] == '1'; }
    { const char *g = getenv("PCL_TIMELINE"); ctx->timeline = g && g[0] == '1'; }
    { const char *g = getenv("PCL_REDUCE_SLOTS"); ctx->reduce_slots = g && g[0] == '1'; }
    { const char *g = getenv("PCL_FILTER_WAVES"); if (g && atoi(g) > 0) ctx->filter_waves = atoi(g); }
    { const char *g = getenv("PCL_REDUCE_AFTER_FILTER"); ctx->reduce_after_filter = g && g[0] == '1'; }
    { const char *g = getenv("PCL_DEVICE_BUILD_MIN"); if (g) ctx->device_build_min = strtoull(g, nullptr, 10); }
    memset(ctx->descs, 0, sizeof ctx->descs);
    memset(ctx->dreads, 0, sizeof ctx->dreads);
    memset(ctx->infos, 0, sizeof ctx->infos);
    auto bail = [&](const char *what, cudaError_t ce) {
        fail(nullptr, PCL_ERR_CUDA, "%s: %s", what, cudaGetErrorString(ce));
        delete ctx;
        return PCL_ERR_CUDA;
    };
    if ((e = cudaSetDevice(device)) != cudaSuccess) return bail("cudaSetDevice", e);
    {
        // the ctx stream carries the all-reduce of the counts: highest priority, so that its CTAs take the first slots the
        // kernels of the chunk streams free up (k_filter runs several waves for that reason)
        int prio_lo = 0, prio_hi = 0;
        cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
        if ((e = cudaStreamCreateWithPriority(&ctx->stream, cudaStreamNonBlocking, prio_hi)) != cudaSuccess) return bail("cudaStreamCreate", e);
    }
    if ((e = cudaEventCreateWithFlags(&ctx->ev_reduced, cudaEventDisableTiming)) != cudaSuccess) return bail("cudaEventCreate", e);
    if ((e = cudaMalloc(&ctx->d_lut, 512 * sizeof(double))) != cudaSuccess) return bail("cudaMalloc", e);
    if ((e = cudaMalloc(&ctx->d_qc, (PCL_MAX_READS + 12) * sizeof(unsigned long long))) != cudaSuccess) return bail("cudaMalloc", e);
    double lut[512];
    for (int q = 0; q < 256; ++q) {
        lut[q] = error_probability(q < 66 ? q : 66);     // BC_MAX_QV cap (barcode.rs:18,320)
        lut[256 + q] = error_probability(q);
    }
    if ((e = cudaMemcpy(ctx->d_lut, lut, sizeof lut, cudaMemcpyHostToDevice)) != cudaSuccess) return bail("cudaMemcpy", e);
    if ((e = cudaMemset(ctx->d_qc, 0, (PCL_MAX_READS + 12) * sizeof(unsigned long long))) != cudaSuccess) return bail("cudaMemset", e);
    // persistent grids are exactly one resident wave: CTAs per SM come from the compiled kernels' real occupancy
    {
        auto occ = [&](const void *fn) {
            int nblk = 0;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nblk, fn, PCL_BLOCK, 0) != cudaSuccess || nblk < 1) { cudaGetLastError(); nblk = 1; }
            return nblk;
        };
        ctx->occ[OCC_COUNT] = occ((const void *)k_count<0>);
        ctx->occ[OCC_COUNT_CODES] = occ((const void *)k_count<1>);
        ctx->occ[OCC_COUNT_PLAN] = occ((const void *)k_count<2>);
        ctx->occ[OCC_COUNT_PLAN2] = occ((const void *)k_count_plan<false, PlanGeoRt>);
        {
            // k_count_fast: 1024-thread CTAs with a 16 K-entry histogram cache (128 KB), one resident wave, two
            // candidate lines per key, a new key admitted on 1 of 8 sightings.  Measured on 125 M v3 records:
            // 2.09 ms against 2.49 ms for 256 threads / 4 K entries / 3 waves / first-come lines (profiles/).
            // The environment overrides exist for experiments (tools/kbench.py).
            const char *fb = getenv("PCL_FAST_BITS"), *fw = getenv("PCL_WAVES"), *fa = getenv("PCL_ADMIT"), *ft = getenv("PCL_TWO");
            ctx->fast_bits = fb ? atoi(fb) : 14;
            if (ctx->fast_bits < 12 || ctx->fast_bits > 14) ctx->fast_bits = 14;
            ctx->fast_waves = fw ? std::max(1, atoi(fw)) : 1;
            ctx->hist_admit_mask = fa ? (uint32_t)atoi(fa) : 7u;
            ctx->hist_two_choice = ft ? (uint32_t)atoi(ft) : 1u;
            const void *fn = ctx->fast_bits == 12 ? (const void *)k_count_fast<12> : ctx->fast_bits == 13 ? (const void *)k_count_fast<13> : (const void *)k_count_fast<14>;
            const int smem = 8 << ctx->fast_bits, threads = PCL_BLOCK << (ctx->fast_bits - 12);
            cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            if (ctx->fast_bits == 14) {
                cudaFuncSetAttribute((const void *)k_count_fast<14, PCL_FAST_SPEC_B0_U16x12>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
                cudaFuncSetAttribute((const void *)k_count_fast<14, PCL_FAST_SPEC_B0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
                cudaFuncSetAttribute((const void *)k_count_fast<14, PCL_FAST_SPEC_B8>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            }
            int nb = 1;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, fn, threads, smem) != cudaSuccess || nb < 1) nb = 1;
            ctx->occ[OCC_COUNT_FAST] = nb;
            const void *spec_fns[6] = {(const void *)k_count_spec<PCL_FAST_SPEC_B0_U16x12, false>, (const void *)k_count_spec<PCL_FAST_SPEC_B0_U16x12, true>,
                                       (const void *)k_count_spec<PCL_FAST_SPEC_B0, false>, (const void *)k_count_spec<PCL_FAST_SPEC_B0, true>,
                                       (const void *)k_count_spec<PCL_FAST_SPEC_B8, false>, (const void *)k_count_spec<PCL_FAST_SPEC_B8, true>};
            for (const void *f : spec_fns) cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 << 14);
            cudaFuncSetAttribute((const void *)k_count_spec<PCL_FAST_SPEC_B0_U16x12, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 << 14);
            { const char *g = getenv("PCL_NO_PREFETCH"); ctx->no_prefetch = g && g[0] == '1'; }
            { const char *g = getenv("PCL_OLD_FAST"); ctx->old_fast = g && g[0] == '1'; }
        }
        ctx->occ[OCC_COUNT_LENONLY] = occ((const void *)k_count_lenonly);
        ctx->occ[OCC_ENUM] = occ((const void *)k_enum_all);
        ctx->occ[OCC_FILTER] = occ((const void *)k_filter);
        ctx->occ[OCC_RESOLVE] = occ((const void *)k_resolve);
        ctx->occ[OCC_CORRECT] = occ((const void *)k_correct);
        ctx->occ[OCC_QC] = occ((const void *)k_qc);
        ctx->occ[OCC_GATHER] = occ((const void *)k_gather_counts);
    }
    *out = ctx;
    return PCL_OK;
}

static void table_free(Table &t) {
    if (t.d_keys) cudaFree(t.d_keys);
    if (t.d_counts) cudaFree(t.d_counts);
    if (t.d_index_to_slot) cudaFree(t.d_index_to_slot);
    for (int q = 0; q < 4; ++q) if (t.d_nkeys[q]) cudaFree(t.d_nkeys[q]);
    if (t.d_nbits) cudaFree(t.d_nbits);
    if (t.d_dense) cudaFree(t.d_dense);
    if (t.d_order_slots) cudaFree(t.d_order_slots);
    if (t.d_scratch_counts) cudaFree(t.d_scratch_counts);
    if (t.d_scratch_scalars) cudaFree(t.d_scratch_scalars);
    if (t.d_scalars) cudaFree(t.d_scalars);
    t = Table();
}

int pcl_ctx_destroy(pcl_ctx *ctx) {
    if (!ctx) return PCL_OK;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    while (!ctx->chunks.empty()) pcl_chunk_destroy(ctx->chunks.back());
    for (Table &t : ctx->tables) table_free(t);
    for (TextStage &t : ctx->text) {
        cudaFree(t.d_text); cudaFree(t.d_nl); cudaFree(t.d_tile_cnt); cudaFree(t.d_tile_off); cudaFree(t.d_hash); cudaFree(t.d_tmeta);
        if (t.h_tmeta) cudaFreeHost(t.h_tmeta);
    }
    emit_stage_free(ctx->emit);
    for (pcl_ctx::KeepSlab &k : ctx->keep_slabs) cudaFree(k.p);
    for (EvPair &p : ctx->evs) { cudaEventDestroy(p.a); cudaEventDestroy(p.b); }
    if (ctx->comm && nccl().ok) nccl().CommDestroy(ctx->comm);
    if (ctx->d_lut) cudaFree(ctx->d_lut);
    if (ctx->d_qc) cudaFree(ctx->d_qc);
    if (ctx->d_qc_scratch) cudaFree(ctx->d_qc_scratch);
    if (ctx->ev_reduced) cudaEventDestroy(ctx->ev_reduced);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return PCL_OK;
}

int pcl_sync(pcl_ctx *ctx) {
    if (!ctx) return PCL_ERR_ARG;
    if (int r = bind(ctx)) return r;
    for (pcl_chunk *c : ctx->chunks) CU(ctx, cudaStreamSynchronize(c->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return PCL_OK;
}

int pcl_comm_unique_id(void *uid128) {
    if (!uid128) return fail(nullptr, PCL_ERR_ARG, "uid is NULL");
    if (!nccl().ok) return fail(nullptr, PCL_ERR_NCCL, "libnccl.so.2 could not be loaded");
    NcclUid id;
    int r = nccl().GetUniqueId(&id);
    if (r != 0) return fail(nullptr, PCL_ERR_NCCL, "ncclGetUniqueId: %d", r);
    memcpy(uid128, &id, sizeof id);
    return PCL_OK;
}

int pcl_comm_init(pcl_ctx *ctx, int nranks, int rank, const void *uid128) {
    if (!ctx) return PCL_ERR_ARG;
    if (nranks < 1 || rank < 0 || rank >= nranks) return fail(ctx, PCL_ERR_ARG, "bad nranks/rank");
    ctx->nranks = nranks;
    ctx->rank = rank;
    if (nranks == 1) return PCL_OK;
    if (!uid128) return fail(ctx, PCL_ERR_ARG, "uid is NULL");
    if (!nccl().ok) return fail(ctx, PCL_ERR_NCCL, "libnccl.so.2 could not be loaded");
    if (int r = bind(ctx)) return r;
    NcclUid id;
    memcpy(&id, uid128, sizeof id);
    int r = nccl().CommInitRank(&ctx->comm, nranks, id, rank);
    if (r != 0) return fail(ctx, PCL_ERR_NCCL, "ncclCommInitRank: %s", nccl().GetErrorString ? nccl().GetErrorString(r) : "?");
    return PCL_OK;
}

// ------------------------------------------------------------------------------------------
// layout
// ------------------------------------------------------------------------------------------
int pcl_layout_set(pcl_ctx *ctx, const pcl_read_desc *reads, int n_reads) {
    if (!ctx || !reads) return PCL_ERR_ARG;
    if (n_reads < 1 || n_reads > PCL_MAX_READS) return fail(ctx, PCL_ERR_UNSUPPORTED, "n_reads %d outside 1..%d", n_reads, PCL_MAX_READS);
    if (!ctx->chunks.empty()) return fail(ctx, PCL_ERR_ARG, "layout cannot change while chunks exist");
    for (int r = 0; r < n_reads; ++r) {
        DRead d;
        pcl_read_info info;
        int rc = pcl_build_read(r, reads[r], &d, &info, ctx->bc_region[r], &ctx->err, ctx->qc_enabled);
        if (rc != PCL_OK) return rc;
        if (ctx->verify_mode) d.check_linkers = 1;
        ctx->descs[r] = reads[r];
        ctx->dreads[r] = d;
        ctx->infos[r] = info;
    }
    ctx->n_reads = n_reads;
    return PCL_OK;
}

int pcl_verify_enable(pcl_ctx *ctx, int on) {
    if (!ctx) return PCL_ERR_ARG;
    if (ctx->n_reads) return fail(ctx, PCL_ERR_ARG, "pcl_verify_enable must be called before pcl_layout_set");
    ctx->verify_mode = on != 0;
    if (on) ctx->force_generic = true;        // the byte interpreter is the kernel that compares unsearched fixed sequences
    return PCL_OK;
}

int pcl_qc_enable(pcl_ctx *ctx, int on) {
    if (!ctx) return PCL_ERR_ARG;
    if (ctx->n_reads) return fail(ctx, PCL_ERR_ARG, "pcl_qc_enable must be called before pcl_layout_set");
    ctx->qc_enabled = on != 0;
    return PCL_OK;
}

int pcl_read_info_get(const pcl_ctx *ctx, int read_slot, pcl_read_info *out) {
    if (!ctx || !out || read_slot < 0 || read_slot >= ctx->n_reads) return PCL_ERR_ARG;
    *out = ctx->infos[read_slot];
    return PCL_OK;
}

uint32_t pcl_read_payload_bytes(const pcl_ctx *ctx, int read_slot) {
    if (!ctx || read_slot < 0 || read_slot >= ctx->n_reads) return 0;
    return ctx->dreads[read_slot].needs_payload ? ctx->dreads[read_slot].payload_bytes : 0;
}

uint32_t pcl_read_stride(const pcl_ctx *ctx, int read_slot) {
    if (!ctx || read_slot < 0 || read_slot >= ctx->n_reads) return 0;
    return ctx->dreads[read_slot].stride;
}

// ------------------------------------------------------------------------------------------
// whitelist table
// ------------------------------------------------------------------------------------------
}  // extern "C"

// Device build of a 32-bit whitelist table (tablebuild.cuh).  Returns PCL_OK, an error, or 1 when the list needs 64-bit keys.
static int whitelist_load_device(pcl_ctx *ctx, int region_slot, const uint8_t *bytes, const uint64_t *offsets, uint64_t n) {
    std::vector<uint64_t> hk;
    bool all16 = true, all_le15 = true;
    if (int rc = pcl_pack_whitelist(bytes, offsets, n, &hk, &all16, &all_le15, &ctx->err)) return rc;
    if (!all16 && !all_le15) return 1;
    if (n >= (1ull << 28)) return fail(ctx, PCL_ERR_UNSUPPORTED, "whitelist too large (%llu entries)", (unsigned long long)n);
    cudaStream_t s = ctx->stream;
    unsigned long long *kA = nullptr, *kB = nullptr, *d_tot = nullptr;
    uint32_t *iA = nullptr, *iB = nullptr, *flag = nullptr, *pos = nullptr, *hist = nullptr, *scan_tmp = nullptr, *dense = nullptr;
    const uint64_t n_ctas = (n + PCL_SORT_TILE - 1) / PCL_SORT_TILE;
    std::vector<void *> tmp;
    auto A = [&](void **p, size_t b) { cudaError_t e = cudaMalloc(p, b ? b : 16); if (e == cudaSuccess) tmp.push_back(*p); return e; };
    auto done = [&](int code) { for (void *p : tmp) cudaFree(p); return code; };
#define TCU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return done(fail(ctx, e_ == cudaErrorMemoryAllocation ? PCL_ERR_NOMEM : PCL_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_))); } while (0)
    TCU(A((void **)&kA, n * 8)); TCU(A((void **)&kB, n * 8)); TCU(A((void **)&iA, n * 4)); TCU(A((void **)&iB, n * 4));
    TCU(A((void **)&flag, n * 4)); TCU(A((void **)&pos, n * 4)); TCU(A((void **)&dense, n * 4)); TCU(A((void **)&d_tot, 16));
    TCU(A((void **)&hist, 256 * n_ctas * 4 + 16)); TCU(A((void **)&scan_tmp, (256 * n_ctas / PCL_SCAN_TILE + n / PCL_SCAN_TILE + 4) * 4));
    // keys in file order stay in `orig`; (key, file index) pairs are sorted to find the duplicates
    unsigned long long *orig = nullptr;
    TCU(A((void **)&orig, n * 8));
    TCU(cudaMemcpyAsync(orig, hk.data(), n * 8, cudaMemcpyHostToDevice, s));
    TCU(cudaMemcpyAsync(kA, orig, n * 8, cudaMemcpyDeviceToDevice, s));
    {   // iA = identity: an exclusive scan of ones
        k_fill_u32<<<grid_for(ctx, n, 8), 256, 0, s>>>(flag, n, 1u);
        if (int rc = device_scan(ctx, s, flag, iA, n, scan_tmp, d_tot)) return done(rc);
    }
    const uint32_t key_bytes_mask = all16 ? 0x1Fu : 0x0Fu;                 // marker-form keys: 33 bits for 16-mers, <= 31 bits otherwise
    if (int rc = device_radix_sort(ctx, s, &kA, &iA, &kB, &iB, n, key_bytes_mask, hist, scan_tmp, d_tot)) return done(rc);
    k_wl_mark_first<<<grid_for(ctx, n, 8), 256, 0, s>>>(kA, iA, n, flag);
    if (int rc = device_scan(ctx, s, flag, pos, n, scan_tmp, d_tot)) return done(rc);
    k_wl_compact<<<grid_for(ctx, n, 8), 256, 0, s>>>(orig, flag, pos, n, dense);
    unsigned long long m = 0;
    // the empty marker of a table of marker-less 16-mers: the largest 32-bit value that is not a key
    std::vector<unsigned long long> top(std::min<uint64_t>(n, 4096));
    TCU(cudaMemcpyAsync(&m, d_tot, 8, cudaMemcpyDeviceToHost, s));
    TCU(cudaMemcpyAsync(top.data(), kA + (n - top.size()), top.size() * 8, cudaMemcpyDeviceToHost, s));
    TCU(cudaStreamSynchronize(s));
    uint32_t empty = 0xFFFFFFFFu;
    if (all16)
        for (size_t k = top.size(); k-- > 0;) {                            // sorted ascending: walk down from the largest key
            const uint32_t v = (uint32_t)top[k];
            if (v == empty) --empty; else if (v < empty) break;
        }
    Table &t = ctx->tables[region_slot];
    table_free(t);
    const uint32_t spb = 8;
    uint64_t n_buckets = (100 * m / PCL_TABLE_LOAD_PCT + spb - 1) / spb;
    if (n_buckets < 2) n_buckets = 2;
    const uint64_t n_slots = n_buckets * spb;
    if (n_slots >= (1ull << 29)) return done(fail(ctx, PCL_ERR_UNSUPPORTED, "whitelist too large (%llu entries)", m));
    auto P = [&](void **p, size_t b) { return cudaMalloc(p, b ? b : 16); };     // persistent allocations: freed by table_free
    TCU(P(&t.d_keys, n_slots * 4)); TCU(P((void **)&t.d_counts, n_slots * 4)); TCU(P((void **)&t.d_index_to_slot, m * 4));
    TCU(P((void **)&t.d_scalars, 16)); TCU(P((void **)&t.d_dense, (m + 1) * 4)); TCU(P((void **)&t.d_order_slots, (m + 1) * 4));
    for (int q = 0; q < 4; ++q) TCU(P((void **)&t.d_nkeys[q], n_slots * 4));
    const int gfill = grid_for(ctx, n_slots, 8), gins = grid_for(ctx, m, 8);
    k_fill_u32<<<gfill, 256, 0, s>>>((uint32_t *)t.d_keys, n_slots, empty);
    {   // main table: placement by rank inside the home bucket (every rank builds the same slots; see tablebuild.cuh)
        k_wl_home<<<gins, 256, 0, s>>>(dense, m, (uint32_t)n_buckets, kA, iA);
        if (int rc = device_radix_sort(ctx, s, &kA, &iA, &kB, &iB, m, 0x0Fu, hist, scan_tmp, d_tot)) return done(rc);
        k_wl_place<<<gins, 256, 0, s>>>(kA, iA, dense, m, (uint32_t *)t.d_keys, t.d_index_to_slot, t.d_order_slots, flag);
        if (int rc = device_scan(ctx, s, flag, pos, m, scan_tmp, d_tot)) return done(rc);
        k_wl_overflow_list<<<gins, 256, 0, s>>>(flag, pos, m, iB);
        k_wl_overflow<<<1, 32, 0, s>>>(kA, iA, dense, iB, d_tot, (uint32_t *)t.d_keys, (uint32_t)n_buckets, empty, t.d_index_to_slot, t.d_order_slots);
        ctx->launches += 4;
    }
    for (uint32_t q = 0; q < 4; ++q) {
        k_fill_u32<<<gfill, 256, 0, s>>>(t.d_nkeys[q], n_slots, empty);
        k_wl_insert<false><<<gins, 256, 0, s>>>(dense, m, t.d_nkeys[q], (uint32_t)n_buckets, empty, q, nullptr);
    }
    ctx->launches += 11;
    TCU(cudaMemsetAsync(t.d_counts, 0, n_slots * 4, s));
    TCU(cudaMemsetAsync(t.d_scalars, 0, 16, s));
    for (int q = 0; q < 4; ++q) { t.d.nkeys[q] = t.d_nkeys[q]; t.d.nbits[q] = nullptr; }
    if (!ctx->no_nbits) {
        TCU(P((void **)&t.d_nbits, 4ull * PCL_NBITS_WORDS * 4));
        TCU(cudaMemsetAsync(t.d_nbits, 0, 4ull * PCL_NBITS_WORDS * 4, s));
        k_build_nbits<<<gfill, PCL_BLOCK, 0, s>>>((const uint32_t *)t.d_keys, n_slots, empty, t.d_nbits, t.d_nbits + PCL_NBITS_WORDS,
                                                  t.d_nbits + 2 * PCL_NBITS_WORDS, t.d_nbits + 3 * PCL_NBITS_WORDS);
        for (int q = 0; q < 4; ++q) t.d.nbits[q] = t.d_nbits + (uint64_t)q * PCL_NBITS_WORDS;
    }
    TCU(cudaGetLastError());
    TCU(cudaStreamSynchronize(s));
#undef TCU
    t.d.keys = t.d_keys; t.d.counts = t.d_counts; t.d.total = t.d_scalars; t.d.mismatch = t.d_scalars + 1;
    t.d.empty = empty; t.d.n_buckets = (uint32_t)n_buckets; t.d.mode = all16 ? PCL_TAB_K32_L16 : PCL_TAB_K32_MARKER;
    t.n_entries = m; t.n_slots = n_slots; t.loaded = true;
    return done(PCL_OK);
}

extern "C" {

int pcl_whitelist_load(pcl_ctx *ctx, int region_slot, const uint8_t *bytes, const uint64_t *offsets, uint64_t n) {
    if (!ctx || region_slot < 0 || region_slot >= PCL_MAX_REGIONS) return PCL_ERR_ARG;
    if (n && (!bytes || !offsets)) return fail(ctx, PCL_ERR_ARG, "NULL whitelist buffers");
    if (int r = bind(ctx)) return r;
    if (n == 0)
        return fail(ctx, PCL_ERR_UNSUPPORTED, "region %d has no whitelist: observed-barcode mode (precellar/src/barcode.rs:418-423) is not implemented on the device", region_slot);
    if (n >= ctx->device_build_min && !ctx->no_neighbour_tables) {
        const int rc = whitelist_load_device(ctx, region_slot, bytes, offsets, n);
        if (rc != 1) return rc;                     // 1: not a 32-bit table -> host build below
    }
    HostTable ht;
    if (int rc = pcl_build_table(bytes, offsets, n, &ht, &ctx->err)) return rc;
    Table &t = ctx->tables[region_slot];
    table_free(t);
    const uint64_t m = ht.n_entries, n_slots = ht.n_slots;
    const size_t key_bytes = ht.key_bytes();
    CU(ctx, cudaMalloc(&t.d_keys, n_slots * key_bytes));
    CU(ctx, cudaMalloc(&t.d_counts, n_slots * sizeof(uint32_t)));
    CU(ctx, cudaMalloc(&t.d_index_to_slot, m * sizeof(uint32_t)));
    CU(ctx, cudaMalloc(&t.d_scalars, 2 * sizeof(unsigned long long)));
    CU(ctx, cudaMemcpy(t.d_keys, ht.data(), n_slots * key_bytes, cudaMemcpyHostToDevice));
    CU(ctx, cudaMemcpy(t.d_index_to_slot, ht.index_to_slot.data(), m * sizeof(uint32_t), cudaMemcpyHostToDevice));
    CU(ctx, cudaMemset(t.d_counts, 0, n_slots * sizeof(uint32_t)));
    CU(ctx, cudaMemset(t.d_scalars, 0, 2 * sizeof(unsigned long long)));
    for (int q = 0; q < 4; ++q) {
        t.d.nkeys[q] = nullptr;
        t.d.nbits[q] = nullptr;
        if (ht.mode != PCL_TAB_K64 && !ctx->no_neighbour_tables) {
            CU(ctx, cudaMalloc(&t.d_nkeys[q], n_slots * sizeof(uint32_t)));
            CU(ctx, cudaMemcpy(t.d_nkeys[q], ht.ntab32[q].data(), n_slots * sizeof(uint32_t), cudaMemcpyHostToDevice));
            t.d.nkeys[q] = t.d_nkeys[q];
        }
    }
    if (ht.mode != PCL_TAB_K64 && !ctx->no_neighbour_tables && !ctx->no_nbits) {
        CU(ctx, cudaMalloc(&t.d_nbits, 4ull * PCL_NBITS_WORDS * sizeof(uint32_t)));
        CU(ctx, cudaMemsetAsync(t.d_nbits, 0, 4ull * PCL_NBITS_WORDS * sizeof(uint32_t), ctx->stream));
        k_build_nbits<<<grid_for(ctx, n_slots, 8), PCL_BLOCK, 0, ctx->stream>>>(
            reinterpret_cast<const uint32_t *>(t.d_keys), n_slots, (uint32_t)ht.empty, t.d_nbits, t.d_nbits + PCL_NBITS_WORDS,
            t.d_nbits + 2 * PCL_NBITS_WORDS, t.d_nbits + 3 * PCL_NBITS_WORDS);
        CU(ctx, cudaGetLastError());
        for (int q = 0; q < 4; ++q) t.d.nbits[q] = t.d_nbits + (uint64_t)q * PCL_NBITS_WORDS;
    }
    t.d.keys = t.d_keys;
    t.d.counts = t.d_counts;
    t.d.total = t.d_scalars;
    t.d.mismatch = t.d_scalars + 1;
    t.d.empty = ht.empty;
    t.d.n_buckets = (uint32_t)ht.n_buckets;
    t.d.mode = ht.mode;
    t.n_entries = m;
    t.n_slots = n_slots;
    CU(ctx, cudaMalloc(&t.d_dense, (m + 1) * sizeof(uint32_t)));
    {   // the occupied slots in ascending order: the order the all-reduce sends the counts in
        std::vector<uint32_t> os(ht.index_to_slot);
        std::sort(os.begin(), os.end());
        CU(ctx, cudaMalloc(&t.d_order_slots, (m + 1) * sizeof(uint32_t)));
        CU(ctx, cudaMemcpy(t.d_order_slots, os.data(), m * sizeof(uint32_t), cudaMemcpyHostToDevice));
    }
    t.loaded = true;
    // the uploads above come from pageable memory and the memsets may be asynchronous: nothing else orders them against
    // the (non-blocking) chunk streams
    CU(ctx, cudaDeviceSynchronize());
    return PCL_OK;
}

uint64_t pcl_whitelist_size(const pcl_ctx *ctx, int region_slot) {
    if (!ctx || region_slot < 0 || region_slot >= PCL_MAX_REGIONS) return 0;
    return ctx->tables[region_slot].n_entries;
}

// ------------------------------------------------------------------------------------------
// chunks
// ------------------------------------------------------------------------------------------
static void stage_free(TextStage &t) {              // the text buffers only: the hash / meta words stay
    cudaFree(t.d_text); cudaFree(t.d_nl); cudaFree(t.d_tile_cnt); cudaFree(t.d_tile_off);
    t.d_text = nullptr; t.d_nl = nullptr; t.d_tile_cnt = nullptr; t.d_tile_off = nullptr;
    t.text_cap = t.nl_cap = t.text_bytes = 0;
    t.owner = nullptr;
}

static int chunk_alloc(pcl_chunk *c, void **p, size_t bytes) {
    if (bytes == 0) bytes = 16;
    cudaError_t e = cudaMalloc(p, bytes);
    if (e != cudaSuccess) return fail(c->ctx, e == cudaErrorMemoryAllocation ? PCL_ERR_NOMEM : PCL_ERR_CUDA, "cudaMalloc(%zu): %s", bytes, cudaGetErrorString(e));
    c->allocs.push_back(*p);
    return PCL_OK;
}

int pcl_chunk_create(pcl_ctx *ctx, uint64_t capacity_records, pcl_chunk **out) {
    if (!ctx || !out) return PCL_ERR_ARG;
    *out = nullptr;
    if (ctx->n_reads == 0) return fail(ctx, PCL_ERR_ARG, "pcl_layout_set must be called first");
    if (capacity_records == 0 || capacity_records >= (1ull << 32)) return fail(ctx, PCL_ERR_ARG, "capacity must be in 1..2^32-1");
    if (int r = bind(ctx)) return r;
    pcl_chunk *c = new pcl_chunk();
    c->ctx = ctx;
    c->cap = capacity_records;
    const uint64_t cap = capacity_records;
    int rc = PCL_OK;
    // every plane of the chunk comes out of ONE device allocation (cudaMalloc / cudaFree are slow and synchronise: an API run
    // that creates a chunk per batch used to spend more time in them than in the kernels)
    std::vector<std::pair<void **, size_t>> wants;
    size_t slab_bytes = 0;
    auto A = [&](void **p, size_t bytes) {
        if (bytes == 0) bytes = 16;
        bytes = (bytes + 255) & ~(size_t)255;
        wants.push_back({p, slab_bytes});
        slab_bytes += bytes;
    };
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_counted, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreate(&c->t0) != cudaSuccess || cudaEventCreate(&c->t1) != cudaSuccess)
        rc = fail(ctx, PCL_ERR_CUDA, "stream/event creation failed: %s", cudaGetErrorString(cudaGetLastError()));
    for (int r = 0; r < ctx->n_reads && rc == PCL_OK; ++r) {
        const DRead &d = ctx->dreads[r];
        const pcl_read_info &inf = ctx->infos[r];
        ReadBuf &b = c->rb[r];
        if (d.needs_payload) A((void **)&b.seq, cap * d.stride);
        if ((d.needs_payload || d.qual_only) && d.qstride) A((void **)&b.qual, cap * d.qstride);
        A((void **)&b.len, cap * 2 + 16);
        A((void **)&b.fstatus, cap * 2 + 16);
        if (inf.has_target) A((void **)&b.tcoord, cap * 4);
        for (int j = 0; j < inf.n_bc; ++j) {
            A(&b.bc_key[j], cap * inf.bc_key_bytes[j]);
            if (!inf.fixed_layout) A((void **)&b.bcoord[j], cap * 4);
        }
        for (int j = 0; j < inf.n_umi; ++j) {
            A(&b.umi_key[j], cap * inf.umi_key_bytes[j]);
            if (!inf.fixed_layout) A((void **)&b.ucoord[j], cap * 4);
        }
        // misses + the unused tails of the per-warp reservations (largest grid: 3 waves of <= 8 CTAs per SM, 8 warps each)
        if (inf.n_bc) {
            uint64_t warps = (uint64_t)ctx->sm_count * 24 * (PCL_BLOCK / 32), by_size = (cap + 31) / 32 + 32;
            if (warps > by_size) warps = by_size;             // a grid never has more warps than the chunk has records / 32
            const uint64_t wl_cap = (cap + warps * PCL_WL_BATCH) * inf.n_bc;
            A((void **)&b.worklist, wl_cap * 8);
            bool k32 = true;                                  // whitelists may still be loaded later: go by the key widths
            for (int j = 0; j < inf.n_bc; ++j) k32 = k32 && inf.bc_key_bytes[j] == 4;
            if (k32) { A((void **)&b.wl_key, wl_cap * 4); A((void **)&b.wl_cand, wl_cap * 8); }
        }
        A((void **)&b.wl_count, 16);
    }
    if (rc == PCL_OK) {
        void *slab = nullptr;
        rc = chunk_alloc(c, &slab, slab_bytes);
        if (rc == PCL_OK) for (auto &w : wants) *w.first = (uint8_t *)slab + w.second;
    }
    if (rc != PCL_OK) {
        for (void *p : c->allocs) cudaFree(p);
        if (c->stream) cudaStreamDestroy(c->stream);
        delete c;
        return rc;
    }
    ctx->chunks.push_back(c);
    *out = c;
    return PCL_OK;
}

int pcl_chunk_destroy(pcl_chunk *c) {
    if (!c) return PCL_OK;
    pcl_ctx *ctx = c->ctx;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(c->stream);
    for (void *p : c->allocs) cudaFree(p);
    for (int r = 0; r < PCL_MAX_READS; ++r) {
        ReadBuf &b = c->rb[r];
        if (b.k_text || b.k_nl) keep_release(ctx);
        if (ctx->text[r].owner == c) ctx->text[r].owner = nullptr;
    }
    cudaEventDestroy(c->ev); cudaEventDestroy(c->ev_counted); cudaEventDestroy(c->t0); cudaEventDestroy(c->t1);
    cudaStreamDestroy(c->stream);
    for (size_t i = 0; i < ctx->chunks.size(); ++i)
        if (ctx->chunks[i] == c) { ctx->chunks.erase(ctx->chunks.begin() + i); break; }
    delete c;
    return PCL_OK;
}

int pcl_chunk_reset(pcl_chunk *c, uint64_t n_records) {
    if (!c) return PCL_ERR_ARG;
    if (n_records > c->cap) return fail(c->ctx, PCL_ERR_ARG, "n_records %llu exceeds chunk capacity %llu", (unsigned long long)n_records, (unsigned long long)c->cap);
    c->n = n_records;
    c->has_results = false;
    for (int r = 0; r < c->ctx->n_reads; ++r) c->rb[r].filled = c->rb[r].text_filled = c->rb[r].uniform_len = c->rb[r].text_kept = false;
    return PCL_OK;
}

int pcl_chunk_set_uniform_len(pcl_chunk *c, int read_slot, uint16_t len) {
    if (!c) return PCL_ERR_ARG;
    pcl_ctx *ctx = c->ctx;
    if (read_slot < 0 || read_slot >= ctx->n_reads) return fail(ctx, PCL_ERR_ARG, "bad read_slot");
    if (int r = bind(ctx)) return r;
    ReadBuf &b = c->rb[read_slot];
    if (c->n) {
        k_fill_len<<<grid_for(ctx, c->n / 2 + 1, 8), PCL_BLOCK, 0, c->stream>>>(b.len, c->n, len);
        ctx->launches += 1;
        CU(ctx, cudaGetLastError());
    }
    b.uniform_len = true;
    if (!ctx->dreads[read_slot].needs_payload && !ctx->dreads[read_slot].qual_only) b.filled = true;     // lengths are all such a read uploads
    return PCL_OK;
}

int pcl_chunk_uniform_results_async(pcl_chunk *c, int read_slot, pcl_uniform_result *out) {
    if (!c || !out) return PCL_ERR_ARG;
    pcl_ctx *ctx = c->ctx;
    if (read_slot < 0 || read_slot >= ctx->n_reads) return fail(ctx, PCL_ERR_ARG, "bad read_slot");
    const pcl_read_info &inf = ctx->infos[read_slot];
    if (inf.n_bc || inf.n_umi) return fail(ctx, PCL_ERR_ARG, "read %d holds barcodes / UMIs: its results are per record", read_slot);
    if (int r = bind(ctx)) return r;
    ReadBuf &b = c->rb[read_slot];
    if (!b.d_uniform) if (int r = chunk_alloc(c, (void **)&b.d_uniform, 16)) return r;
    if (c->n == 0) { out->uniform = 1; out->fstatus = 0; out->tcoord = PCL_COORD_ABSENT; return PCL_OK; }
    static const uint32_t one[4] = {1u, 0u, 0u, 0u};
    CU(ctx, cudaMemcpyAsync(b.d_uniform, one, 16, cudaMemcpyHostToDevice, c->stream));
    k_uniform_check<<<grid_for(ctx, c->n, 8), PCL_BLOCK, 0, c->stream>>>(b.fstatus, inf.has_target ? b.tcoord : nullptr, c->n, b.d_uniform);
    ctx->launches += 1;
    CU(ctx, cudaGetLastError());
    CU(ctx, cudaMemcpyAsync(out, b.d_uniform, 16, cudaMemcpyDeviceToHost, c->stream));
    return PCL_OK;
}

int pcl_chunk_upload(pcl_chunk *c, int read_slot, const uint8_t *seq, const uint8_t *qual, const uint16_t *len) {
    if (!c) return PCL_ERR_ARG;
    pcl_ctx *ctx = c->ctx;
    if (read_slot < 0 || read_slot >= ctx->n_reads) return fail(ctx, PCL_ERR_ARG, "bad read_slot");
    if (!len && !c->rb[read_slot].uniform_len) return fail(ctx, PCL_ERR_ARG, "len is NULL (and no pcl_chunk_set_uniform_len since the last reset)");
    if (int r = bind(ctx)) return r;
    const DRead &d = ctx->dreads[read_slot];
    ReadBuf &b = c->rb[read_slot];
    if (d.needs_payload) {
        if (!seq || (!qual && d.qstride)) return fail(ctx, PCL_ERR_ARG, "read %d needs seq and qual planes", read_slot);
        CU(ctx, cudaMemcpyAsync(b.seq, seq, c->n * d.stride, cudaMemcpyHostToDevice, c->stream));
        if (d.qstride) CU(ctx, cudaMemcpyAsync(b.qual, qual, c->n * d.qstride, cudaMemcpyHostToDevice, c->stream));
    }
    if (d.qual_only) {
        if (!qual) return fail(ctx, PCL_ERR_ARG, "read %d needs its quality plane (FASTQ QC is enabled)", read_slot);
        CU(ctx, cudaMemcpyAsync(b.qual, qual, c->n * d.qstride, cudaMemcpyHostToDevice, c->stream));
    }
    if (len) CU(ctx, cudaMemcpyAsync(b.len, len, c->n * 2, cudaMemcpyHostToDevice, c->stream));
    b.filled = true;
    return PCL_OK;
}

int pcl_chunk_upload_compact(pcl_chunk *c, int read_slot, const uint8_t *seq, uint32_t row_bytes, const uint8_t *qual,
                             const uint16_t *len) {
    if (!c) return PCL_ERR_ARG;
    pcl_ctx *ctx = c->ctx;
    if (read_slot < 0 || read_slot >= ctx->n_reads) return fail(ctx, PCL_ERR_ARG, "bad read_slot");
    const DRead &d = ctx->dreads[read_slot];
    if (!d.needs_payload || row_bytes == d.stride) return pcl_chunk_upload(c, read_slot, seq, qual, len);
    if ((row_bytes & 3u) || row_bytes < d.payload_bytes || row_bytes > d.stride)
        return fail(ctx, PCL_ERR_ARG, "read %d: compact rows must be a multiple of 4 bytes in [%u, %u]", read_slot,
                    (unsigned)d.payload_bytes, d.stride);
    if (!seq || (!len && !c->rb[read_slot].uniform_len) || (!qual && d.qstride))
        return fail(ctx, PCL_ERR_ARG, "read %d needs seq, qual and len planes", read_slot);
    if (int r = bind(ctx)) return r;
    ReadBuf &b = c->rb[read_slot];
    if (!b.seq_stage) if (int r = chunk_alloc(c, (void **)&b.seq_stage, c->cap * (uint64_t)d.stride)) return r;
    CU(ctx, cudaMemcpyAsync(b.seq_stage, seq, c->n * (uint64_t)row_bytes, cudaMemcpyHostToDevice, c->stream));
    if (c->n) {
        k_repack_rows<<<grid_for(ctx, c->n * (d.stride / 4), 8), PCL_BLOCK, 0, c->stream>>>(
            reinterpret_cast<const uint32_t *>(b.seq_stage), row_bytes / 4, reinterpret_cast<uint32_t *>(b.seq), d.stride / 4, c->n);
        ctx->launches += 1;
        CU(ctx, cudaGetLastError());
    }
    if (d.qstride) CU(ctx, cudaMemcpyAsync(b.qual, qual, c->n * (uint64_t)d.qstride, cudaMemcpyHostToDevice, c->stream));
    if (len) CU(ctx, cudaMemcpyAsync(b.len, len, c->n * 2, cudaMemcpyHostToDevice, c->stream));
    b.filled = true;
    return PCL_OK;
}

int pcl_chunk_device_ptrs(pcl_chunk *c, int read_slot, void **seq, void **qual, void **len) {
    if (!c) return PCL_ERR_ARG;
    if (read_slot < 0 || read_slot >= c->ctx->n_reads) return fail(c->ctx, PCL_ERR_ARG, "bad read_slot");
    ReadBuf &b = c->rb[read_slot];
    if (seq) *seq = b.seq;
    if (qual) *qual = b.qual;
    if (len) *len = b.len;
    b.filled = true;
    return PCL_OK;
}

void *pcl_chunk_stream(pcl_chunk *c) { return c ? (void *)c->stream : nullptr; }

// ------------------------------------------------------------------------------------------
// device-side FASTQ text parse (textparse.cuh)

static int text_reserve(pcl_chunk *c, ReadBuf &b, TextStage &t, uint64_t n_bytes) {
    pcl_ctx *ctx = c->ctx;
    if (!t.d_tmeta) {
        CU(ctx, cudaMalloc((void **)&t.d_tmeta, 16));
        CU(ctx, cudaMallocHost((void **)&t.h_tmeta, 32));
    }
    if (c->cap > t.hash_cap) {
        if (t.owner) CU(ctx, cudaStreamSynchronize(t.owner->stream));
        cudaFree(t.d_hash);
        t.d_hash = nullptr; t.hash_cap = 0;
        CU(ctx, cudaMalloc((void **)&t.d_hash, c->cap * 8 + 16));
        t.hash_cap = c->cap;
    }
    b.d_tmeta = t.d_tmeta; b.h_tmeta = t.h_tmeta; b.d_hash = t.d_hash;
    if (n_bytes + 16 <= t.text_cap) return PCL_OK;
    if (t.owner) CU(ctx, cudaStreamSynchronize(t.owner->stream));
    stage_free(t);
    const uint64_t cap = ((n_bytes + n_bytes / 8 + 16 + 4095) / 4096) * 4096;      // some head room: blocks vary a little
    const uint64_t tiles = cap / PCL_TEXT_TILE + 1;
    // one index entry per 8 bytes of text: a record needs >= 8 bytes for its four lines, real data ~30 per line;
    // denser input is consumed in several calls (the index is simply cut off at nl_cap)
    const uint64_t nl_cap = cap / 8 + 16;
    cudaError_t e = cudaMalloc((void **)&t.d_text, cap);
    if (e == cudaSuccess) e = cudaMalloc((void **)&t.d_nl, nl_cap * 4);
    if (e == cudaSuccess) e = cudaMalloc((void **)&t.d_tile_cnt, tiles * 4);
    if (e == cudaSuccess) e = cudaMalloc((void **)&t.d_tile_off, tiles * 4);
    if (e != cudaSuccess) {
        stage_free(t);
        return fail(ctx, e == cudaErrorMemoryAllocation ? PCL_ERR_NOMEM : PCL_ERR_CUDA, "text staging (%llu bytes): %s",
                    (unsigned long long)cap, cudaGetErrorString(e));
    }
    t.text_cap = cap;
    t.nl_cap = nl_cap;
    return PCL_OK;
}

static int text_scan_from(pcl_chunk *c, int read_slot, const uint8_t *text, uint64_t n_bytes, cudaMemcpyKind kind);

int pcl_text_scan(pcl_chunk *c, int read_slot, const uint8_t *text, uint64_t n_bytes) {
    return text_scan_from(c, read_slot, text, n_bytes, cudaMemcpyHostToDevice);
}

int pcl_text_scan_device(pcl_chunk *c, int read_slot, const void *device_text, uint64_t n_bytes) {
    return text_scan_from(c, read_slot, (const uint8_t *)device_text, n_bytes, cudaMemcpyDeviceToDevice);
}

static int text_scan_from(pcl_chunk *c, int read_slot, const uint8_t *text, uint64_t n_bytes, cudaMemcpyKind kind) {
    if (!c) return PCL_ERR_ARG;
    pcl_ctx *ctx = c->ctx;
    if (read_slot < 0 || read_slot >= ctx->n_reads) return fail(ctx, PCL_ERR_ARG, "bad read_slot");
    if (!text && n_bytes) return fail(ctx, PCL_ERR_ARG, "text is NULL");
    if (n_bytes >= (1ull << 32) - 4096) return fail(ctx, PCL_ERR_ARG, "a text block must be smaller than 4 GiB");
    if (int r = bind(ctx)) return r;
    ReadBuf &b = c->rb[read_slot];
    TextStage &t = ctx->text[read_slot];
    if (int r = text_reserve(c, b, t, n_bytes)) return r;
    cudaStream_t s = c->stream;
    // the previous user of the staging buffers has finished its fill (pcl_text_fill blocks) unless it never filled
    if (t.owner && t.owner != c) { CU(ctx, cudaStreamSynchronize(t.owner->stream)); t.owner->rb[read_slot].text_scanned = false; }
    t.owner = c;
    if (n_bytes) CU(ctx, cudaMemcpyAsync(t.d_text, text, n_bytes, kind, s));
    CU(ctx, cudaMemsetAsync(t.d_text + n_bytes, 0, 16, s));
    const uint64_t n_tiles = (n_bytes + PCL_TEXT_TILE - 1) / PCL_TEXT_TILE;
    const int grid = (int)std::min<uint64_t>(std::max<uint64_t>(n_tiles, 1), (uint64_t)ctx->sm_count * 8);
    prof_begin(ctx, s, PCL_K_TEXT_INDEX);
    k_text_count<<<grid, PCL_TEXT_BLOCK, 0, s>>>(t.d_text, n_bytes, n_tiles, t.d_tile_cnt);
    k_text_scan<<<1, 1024, 0, s>>>(t.d_tile_cnt, n_tiles, t.d_tile_off, b.d_tmeta);
    k_text_index<<<grid, PCL_TEXT_BLOCK, 0, s>>>(t.d_text, n_bytes, n_tiles, t.d_tile_off, t.d_nl, t.nl_cap);
    prof_end(ctx, s);
    ctx->launches += 2;
    CU(ctx, cudaGetLastError());
    CU(ctx, cudaMemcpyAsync(b.h_tmeta, b.d_tmeta, 8, cudaMemcpyDeviceToHost, s));
    t.text_bytes = n_bytes;
    b.text_scanned = true;
    b.text_lines = ~0ull;
    return PCL_OK;
}

int pcl_text_lines(pcl_chunk *c, int read_slot, uint64_t *n_lines) {
    if (!c) return PCL_ERR_ARG;
    pcl_ctx *ctx = c->ctx;
    if (read_slot < 0 || read_slot >= ctx->n_reads || !n_lines) return fail(ctx, PCL_ERR_ARG, "bad argument");
    ReadBuf &b = c->rb[read_slot];
    const TextStage &t = ctx->text[read_slot];
    if (!b.text_scanned || t.owner != c) return fail(ctx, PCL_ERR_ARG, "pcl_text_scan must be called first (one text block per read file in flight)");
    if (int r = bind(ctx)) return r;
    if (b.text_lines == ~0ull) {
        CU(ctx, cudaStreamSynchronize(c->stream));
        b.text_lines = b.h_tmeta[0] < t.nl_cap ? b.h_tmeta[0] : t.nl_cap;
    }
    *n_lines = b.text_lines;
    return PCL_OK;
}

int pcl_text_fill(pcl_chunk *c, int read_slot, uint64_t *consumed) {
    if (!c) return PCL_ERR_ARG;
    pcl_ctx *ctx = c->ctx;
    if (read_slot < 0 || read_slot >= ctx->n_reads) return fail(ctx, PCL_ERR_ARG, "bad read_slot");
    ReadBuf &b = c->rb[read_slot];
    uint64_t lines = 0;
    if (int r = pcl_text_lines(c, read_slot, &lines)) return r;
    if (4 * c->n > lines) return fail(ctx, PCL_ERR_ARG, "chunk holds %llu records but the text block has only %llu complete lines",
                                      (unsigned long long)c->n, (unsigned long long)lines);
    const DRead &d = ctx->dreads[read_slot];
    const TextStage &t = ctx->text[read_slot];
    cudaStream_t s = c->stream;
    b.h_tmeta[2] = 0;
    if (c->n) {
        TextFillParams p;
        p.text = t.d_text; p.nl = t.d_nl; p.n = c->n;
        p.stride = d.stride; p.qoff = d.qoff; p.qstride = d.qstride;
        p.seq = d.needs_payload ? b.seq : nullptr;
        p.qual = ((d.needs_payload || d.qual_only) && d.qstride) ? b.qual : nullptr;
        p.len = b.len; p.name_hash = b.d_hash; p.err = b.d_tmeta + 1;
        CU(ctx, cudaMemsetAsync(b.d_tmeta + 1, 0xFF, 8, s));
        prof_begin(ctx, s, PCL_K_TEXT_FILL);
        k_text_fill<<<grid_for(ctx, c->n, 8), PCL_TEXT_BLOCK, 0, s>>>(p);
        prof_end(ctx, s);
        CU(ctx, cudaGetLastError());
        CU(ctx, cudaMemcpyAsync(b.h_tmeta + 1, b.d_tmeta + 1, 8, cudaMemcpyDeviceToHost, s));
        CU(ctx, cudaMemcpyAsync(b.h_tmeta + 2, t.d_nl + (4 * c->n - 1), 4, cudaMemcpyDeviceToHost, s));
        CU(ctx, cudaStreamSynchronize(s));
        if (b.h_tmeta[1] != ~0ull) {
            static const char *const what[] = {"", "does not start with '@'", "missing '+' line", "sequence and quality lengths differ"};
            const unsigned code = (unsigned)(b.h_tmeta[1] & 0xFF);
            return fail(ctx, PCL_ERR_INPUT, "FASTQ record %llu of the text block (read %d): %s", (unsigned long long)(b.h_tmeta[1] >> 8),
                        read_slot, what[code < 4 ? code : 0]);
        }
    }
    b.text_consumed = c->n ? (b.h_tmeta[2] & 0xFFFFFFFFull) + 1 : 0;
    if (consumed) *consumed = b.text_consumed;
    b.filled = true;
    b.text_filled = true;
    return PCL_OK;
}

int pcl_text_names_check(pcl_chunk *c, int64_t *first_bad) {
    if (!c || !first_bad) return PCL_ERR_ARG;
    pcl_ctx *ctx = c->ctx;
    if (int r = bind(ctx)) return r;
    NamesCheckParams p;
    p.n_reads = 0;
    p.n = c->n;
    ReadBuf *first = nullptr;
    for (int r = 0; r < ctx->n_reads; ++r)
        if (c->rb[r].text_filled) { if (!first) first = &c->rb[r]; p.hash[p.n_reads++] = c->rb[r].d_hash; }
    *first_bad = -1;
    if (p.n_reads < 2 || c->n == 0) return PCL_OK;
    p.bad = first->d_tmeta + 1;
    cudaStream_t s = c->stream;
    CU(ctx, cudaMemsetAsync(p.bad, 0xFF, 8, s));
    prof_begin(ctx, s, PCL_K_TEXT_FILL);
    k_names_check<<<grid_for(ctx, c->n, 8), PCL_TEXT_BLOCK, 0, s>>>(p);
    prof_end(ctx, s);
    CU(ctx, cudaGetLastError());
    CU(ctx, cudaMemcpyAsync(first->h_tmeta + 1, p.bad, 8, cudaMemcpyDeviceToHost, s));
    CU(ctx, cudaStreamSynchronize(s));
    if (first->h_tmeta[1] != ~0ull) *first_bad = (int64_t)first->h_tmeta[1];
    return PCL_OK;
}

int pcl_text_line_ends(pcl_chunk *c, int read_slot, uint32_t *out, uint64_t n_lines) {
    if (!c) return PCL_ERR_ARG;
    pcl_ctx *ctx = c->ctx;
    if (read_slot < 0 || read_slot >= ctx->n_reads || (!out && n_lines)) return fail(ctx, PCL_ERR_ARG, "bad argument");
    uint64_t lines = 0;
    if (int r = pcl_text_lines(c, read_slot, &lines)) return r;
    if (n_lines > lines) return fail(ctx, PCL_ERR_ARG, "only %llu lines are indexed", (unsigned long long)lines);
    if (n_lines) {
        CU(ctx, cudaMemcpyAsync(out, ctx->text[read_slot].d_nl, n_lines * 4, cudaMemcpyDeviceToHost, c->stream));
        CU(ctx, cudaStreamSynchronize(c->stream));
    }
    return PCL_OK;
}


// ------------------------------------------------------------------------------------------
// device-side make_fastq writer (emit.cuh)
// ------------------------------------------------------------------------------------------
}  // extern "C"

namespace {
const uint64_t kKeepSlabBytes = 1ull << 30;

// `bytes` of device memory out of the keep slabs (256-byte aligned); NULL with the error set when the device is full
uint8_t *keep_alloc(pcl_ctx *ctx, uint64_t bytes) {
    bytes = (bytes + 255) & ~255ull;
    if (!ctx->keep_slabs.empty()) {
        pcl_ctx::KeepSlab &k = ctx->keep_slabs.back();
        if (k.used + bytes <= k.cap) { uint8_t *p = k.p + k.used; k.used += bytes; return p; }
    }
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    uint64_t cap = std::max<uint64_t>(bytes, std::min<uint64_t>(kKeepSlabBytes, free_b / 4));
    cap = (cap + 255) & ~255ull;
    uint8_t *p = nullptr;
    cudaError_t e = cudaMalloc((void **)&p, cap);
    if (e != cudaSuccess) {
        cudaGetLastError();
        fail(ctx, e == cudaErrorMemoryAllocation ? PCL_ERR_NOMEM : PCL_ERR_CUDA, "kept text (%llu bytes): %s", (unsigned long long)cap, cudaGetErrorString(e));
        return nullptr;
    }
    ctx->keep_slabs.push_back({p, cap, bytes});
    return p;
}
}  // namespace

// one read file of a chunk gave its slices back; with the last one the slabs start over (the first is kept for the next run)
static void keep_release(pcl_ctx *ctx) {
    if (ctx->keep_live) --ctx->keep_live;
    if (ctx->keep_live) return;
    while (ctx->keep_slabs.size() > 1) { cudaFree(ctx->keep_slabs.back().p); ctx->keep_slabs.pop_back(); }
    if (!ctx->keep_slabs.empty()) ctx->keep_slabs[0].used = 0;
}

extern "C" {

int pcl_text_keep(pcl_chunk *c, int read_slot) {
    if (!c) return PCL_ERR_ARG;
    pcl_ctx *ctx = c->ctx;
    if (read_slot < 0 || read_slot >= ctx->n_reads) return fail(ctx, PCL_ERR_ARG, "bad read_slot");
    ReadBuf &b = c->rb[read_slot];
    const TextStage &t = ctx->text[read_slot];
    if (!b.text_filled || t.owner != c) return fail(ctx, PCL_ERR_ARG, "pcl_text_keep follows pcl_text_fill of the same chunk and read file");
    if (int r = bind(ctx)) return r;
    const uint64_t bytes = b.text_consumed, lines = 4 * c->n;
    if (bytes + 16 > b.k_text_cap || lines + 4 > b.k_nl_cap) {     // a chunk that is reused with larger blocks takes new slices
        if (!b.k_text) ++ctx->keep_live;
        const uint64_t tcap = bytes + bytes / 16 + 16, lcap = std::max<uint64_t>(lines, 4 * c->cap) + 4;
        uint8_t *p = keep_alloc(ctx, ((tcap + 255) & ~255ull) + lcap * 4);
        if (!p) { if (!b.k_text) --ctx->keep_live; return PCL_ERR_NOMEM; }
        b.k_text = p; b.k_text_cap = tcap;
        b.k_nl = (uint32_t *)(p + ((tcap + 255) & ~255ull)); b.k_nl_cap = lcap;
    }
    if (bytes) CU(ctx, cudaMemcpyAsync(b.k_text, t.d_text, bytes, cudaMemcpyDeviceToDevice, c->stream));
    if (lines) CU(ctx, cudaMemcpyAsync(b.k_nl, t.d_nl, lines * 4, cudaMemcpyDeviceToDevice, c->stream));
    b.k_records = c->n;
    b.text_kept = true;
    return PCL_OK;
}

// k_zh_blocks takes more than the 48 KB of shared memory a kernel gets without asking
static int zh_kernel_ready(pcl_ctx *ctx) {
    static bool done[64] = {false};
    if (ctx->device < 64 && done[ctx->device]) return PCL_OK;
    CU(ctx, cudaFuncSetAttribute(k_zh_blocks, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ZhShared)));
    if (ctx->device < 64) done[ctx->device] = true;
    return PCL_OK;
}

// slots, block stream, sizes and per-block scratch of file f for nb blocks
static int zh_reserve_blocks(pcl_ctx *ctx, EmitStage &e, int f, uint64_t nb) {
    cudaFree(e.d_slots[f]); cudaFree(e.d_stream[f]); cudaFree(e.d_bsize[f]); cudaFree(e.d_scratch[f]);
    e.d_slots[f] = e.d_stream[f] = nullptr; e.d_bsize[f] = nullptr; e.d_scratch[f] = nullptr; e.block_cap[f] = 0;
    const uint64_t cap = nb + nb / 8 + 8;
    CU(ctx, cudaMalloc((void **)&e.d_slots[f], cap * ZH_SLOT_BYTES));
    CU(ctx, cudaMalloc((void **)&e.d_stream[f], cap * ZH_SLOT_BYTES));
    CU(ctx, cudaMalloc((void **)&e.d_bsize[f], cap * 4));
    CU(ctx, cudaMalloc((void **)&e.d_scratch[f], cap * sizeof(ZhScratch)));
    e.block_cap[f] = cap;
    return PCL_OK;
}

static int emit_reserve(pcl_ctx *ctx, void **p, uint64_t *cap, uint64_t need, bool pinned) {
    if (need <= *cap && *p) return PCL_OK;
    if (*p) { if (pinned) cudaFreeHost(*p); else cudaFree(*p); }
    *p = nullptr; *cap = 0;
    const uint64_t want = ((need + need / 8 + 4095) / 4096) * 4096;
    cudaError_t e = pinned ? cudaHostAlloc(p, want, cudaHostAllocDefault) : cudaMalloc(p, want);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(ctx, e == cudaErrorMemoryAllocation ? PCL_ERR_NOMEM : PCL_ERR_CUDA, "writer scratch (%llu bytes): %s", (unsigned long long)want, cudaGetErrorString(e)); }
    *cap = want;
    return PCL_OK;
}

int pcl_emit_fastq(pcl_chunk *c, int correct_barcode, int paired, int compress, pcl_emit_result *out) {
    if (!c || !out) return PCL_ERR_ARG;
    pcl_ctx *ctx = c->ctx;
    memset(out, 0, sizeof *out);
    if (!c->has_results) return fail(ctx, PCL_ERR_ARG, "pcl_emit_fastq needs a counted chunk (pcl_count, and pcl_correct when correct_barcode is set)");
    if (int r = bind(ctx)) return r;
    EmitStage &e = ctx->emit;
    cudaStream_t s = c->stream;
    const uint64_t n = c->n;
    const int nf = paired ? 2 : 1;
    EmitKernelParams k;
    memset(&k, 0, sizeof k);
    k.p.n_reads = ctx->n_reads; k.p.correct = correct_barcode ? 1 : 0; k.p.paired = paired ? 1 : 0; k.p.n = n;
    uint64_t text_bytes = 0;
    for (int r = 0; r < ctx->n_reads; ++r) {
        const ReadBuf &b = c->rb[r];
        const pcl_read_desc &rd = ctx->descs[r];
        const pcl_read_info &inf = ctx->infos[r];
        if (!b.text_kept || b.k_records != n) return fail(ctx, PCL_ERR_ARG, "read file %d: the chunk holds no kept text (pcl_text_keep)", r);
        EmitRead &er = k.p.rd[r];
        er.text = b.k_text; er.nl = b.k_nl; er.len = b.len; er.fstatus = b.fstatus; er.tcoord = b.tcoord;
        pcl_emit_describe(er, rd, inf);
        for (int j = 0; j < inf.n_bc; ++j) { er.bc_key[j] = b.bc_key[j]; er.bcoord[j] = b.bcoord[j]; }
        for (int j = 0; j < inf.n_umi; ++j) er.ucoord[j] = b.ucoord[j];
        text_bytes += b.text_consumed;
    }
    e.flip ^= 1;
    if (!e.d_meta) {
        CU(ctx, cudaMalloc((void **)&e.d_meta, 64));
        CU(ctx, cudaHostAlloc((void **)&e.h_meta, 64, cudaHostAllocDefault));
    }
    if (n == 0) return PCL_OK;
    if (n > e.rec_cap) {
        for (int f = 0; f < 2; ++f) { cudaFree(e.d_sz[f]); cudaFree(e.d_off[f]); e.d_sz[f] = e.d_off[f] = nullptr; }
        cudaFree(e.d_scan_tmp); e.d_scan_tmp = nullptr;
        e.rec_cap = 0;
        const uint64_t cap = n + n / 8 + 1024;
        for (int f = 0; f < 2; ++f) {
            CU(ctx, cudaMalloc((void **)&e.d_sz[f], cap * 4));
            CU(ctx, cudaMalloc((void **)&e.d_off[f], cap * 4));
        }
        CU(ctx, cudaMalloc((void **)&e.d_scan_tmp, (cap / PCL_SCAN_TILE + 2) * 4));
        e.rec_cap = cap;
    }
    // every output byte comes from the text of the group (plus '@', "\n+\n" and line ends): the text bounds the output
    const uint64_t bound = text_bytes + 8 * n + 64;
    if (bound >= (1ull << 32)) return fail(ctx, PCL_ERR_UNSUPPORTED, "a chunk's output must stay below 4 GiB (use smaller chunks)");
    for (int f = 0; f < nf; ++f)
        if (int r = emit_reserve(ctx, (void **)&e.d_raw[f], &e.raw_cap[f], bound, false)) return r;
    k.sz[0] = e.d_sz[0]; k.sz[1] = e.d_sz[1]; k.off[0] = e.d_off[0]; k.off[1] = e.d_off[1];
    k.out[0] = e.d_raw[0]; k.out[1] = e.d_raw[1]; k.err = e.d_meta + 2; k.n_written = e.d_meta + 5;
    CU(ctx, cudaMemsetAsync(e.d_meta, 0, 64, s));
    CU(ctx, cudaMemsetAsync(e.d_meta + 2, 0xFF, 8, s));
    prof_begin(ctx, s, PCL_K_EMIT);
    k_emit_plan<<<grid_for(ctx, n, 8), PCL_EMIT_BLOCK, 0, s>>>(k);
    prof_end(ctx, s);
    CU(ctx, cudaGetLastError());
    for (int f = 0; f < nf; ++f) {
        if (int r = device_scan(ctx, s, e.d_sz[f], e.d_off[f], n, e.d_scan_tmp, e.d_meta + f)) return r;
    }
    CU(ctx, cudaMemcpyAsync(e.h_meta, e.d_meta, 48, cudaMemcpyDeviceToHost, s));
    CU(ctx, cudaStreamSynchronize(s));
    if (e.h_meta[2] != ~0ull) {
        static const char *const what[] = {"", "Multiple target regions found in one fastq record!", "inconsistent barcode coordinates",
                                           "corrupt corrected key", "Read1 already exists", "Read2 already exists",
                                           "called `Option::unwrap()` on a `None` value: read1", "called `Option::unwrap()` on a `None` value: read2"};
        const unsigned code = (unsigned)(e.h_meta[2] & 0xFF);
        return fail(ctx, PCL_ERR_INPUT, "%s", what[code < 8 ? code : 0]);
    }
    const uint64_t raw[2] = {e.h_meta[0], paired ? e.h_meta[1] : 0};
    if (raw[0] > bound || raw[1] > bound) return fail(ctx, PCL_ERR_CUDA, "writer: record sizes exceed the text they come from");
    if (raw[0]) {
        prof_begin(ctx, s, PCL_K_EMIT);
        k_emit_fill<<<grid_for(ctx, n * 32, 8), PCL_EMIT_BLOCK, 0, s>>>(k);
        prof_end(ctx, s);
        CU(ctx, cudaGetLastError());
    }
    uint64_t out_bytes[2] = {raw[0], raw[1]};
    const uint8_t *d_src[2] = {e.d_raw[0], e.d_raw[1]};
    uint64_t head[2] = {0, 0};
    if (compress) {
        ZhParams zp;
        memset(&zp, 0, sizeof zp);
        uint32_t total_blocks = 0;
        for (int f = 0; f < nf; ++f) {
            const uint64_t nb = (raw[f] + ZH_BLOCK_BYTES - 1) / ZH_BLOCK_BYTES;
            if (nb > e.block_cap[f]) {
                if (int r = zh_reserve_blocks(ctx, e, f, nb)) return r;
            }
            ZhJob &j = zp.job[f];
            j.src = e.d_raw[f]; j.n = raw[f]; j.slots = e.d_slots[f]; j.sizes = e.d_bsize[f]; j.stream = e.d_stream[f];
            j.scratch = ctx->zh_no_match ? nullptr : e.d_scratch[f];
            j.total = e.d_meta + 3 + f; j.n_blocks = (uint32_t)nb;
            total_blocks += (uint32_t)nb;
        }
        if (total_blocks) {
            prof_begin(ctx, s, PCL_K_ZSTD);
            if (int r = zh_kernel_ready(ctx)) return r;
            k_zh_blocks<<<total_blocks, ZH_THREADS, sizeof(ZhShared), s>>>(zp);
            k_zh_compact<<<total_blocks, ZH_THREADS, 0, s>>>(zp);
            prof_end(ctx, s);
            ctx->launches += 1;
            CU(ctx, cudaGetLastError());
            CU(ctx, cudaMemcpyAsync(e.h_meta + 3, e.d_meta + 3, 16, cudaMemcpyDeviceToHost, s));
            CU(ctx, cudaStreamSynchronize(s));
        }
        for (int f = 0; f < nf; ++f) {
            out_bytes[f] = raw[f] ? e.h_meta[3 + f] : 0;
            d_src[f] = e.d_stream[f];
            head[f] = raw[f] ? ZH_FRAME_HEADER_BYTES : 0;
        }
    }
    for (int f = 0; f < nf; ++f) {
        if (!raw[f]) continue;
        if (int r = emit_reserve(ctx, (void **)&e.h_out[e.flip][f], &e.h_cap[e.flip][f], head[f] + out_bytes[f], true)) return r;
        if (compress) zh_put_frame_header(e.h_out[e.flip][f], raw[f]);
        CU(ctx, cudaMemcpyAsync(e.h_out[e.flip][f] + head[f], d_src[f], out_bytes[f], cudaMemcpyDeviceToHost, s));
    }
    CU(ctx, cudaStreamSynchronize(s));
    out->n_written = e.h_meta[5];
    out->r1 = raw[0] ? e.h_out[e.flip][0] : nullptr; out->r1_bytes = head[0] + out_bytes[0]; out->r1_raw_bytes = raw[0];
    out->r2 = raw[1] ? e.h_out[e.flip][1] : nullptr; out->r2_bytes = raw[1] ? head[1] + out_bytes[1] : 0; out->r2_raw_bytes = raw[1];
    return PCL_OK;
}


// host bytes -> one zstd frame built by the block coder of the device writer (k_zh_blocks / k_zh_compact); blocks
int pcl_zstd_frame(pcl_ctx *ctx, const uint8_t *src, uint64_t n, uint8_t *dst, uint64_t cap, uint64_t *used) {
    if (!ctx || (!src && n) || !dst || !used) return PCL_ERR_ARG;
    if (n >= (1ull << 32)) return fail(ctx, PCL_ERR_UNSUPPORTED, "pcl_zstd_frame takes less than 4 GiB per call");
    if (int r = bind(ctx)) return r;
    EmitStage &e = ctx->emit;
    cudaStream_t s = ctx->stream;
    const uint64_t nb = (n + ZH_BLOCK_BYTES - 1) / ZH_BLOCK_BYTES;
    if (cap < ZH_FRAME_HEADER_BYTES + 3) return fail(ctx, PCL_ERR_NOMEM, "output buffer too small");
    zh_put_frame_header(dst, n);
    if (n == 0) {                                 // an empty frame: one empty raw block marked last
        dst[ZH_FRAME_HEADER_BYTES] = 1; dst[ZH_FRAME_HEADER_BYTES + 1] = 0; dst[ZH_FRAME_HEADER_BYTES + 2] = 0;
        *used = ZH_FRAME_HEADER_BYTES + 3;
        return PCL_OK;
    }
    if (!e.d_meta) {
        CU(ctx, cudaMalloc((void **)&e.d_meta, 64));
        CU(ctx, cudaHostAlloc((void **)&e.h_meta, 64, cudaHostAllocDefault));
    }
    if (int r = emit_reserve(ctx, (void **)&e.d_raw[0], &e.raw_cap[0], n + 16, false)) return r;
    if (nb > e.block_cap[0]) {
        if (int r = zh_reserve_blocks(ctx, e, 0, nb)) return r;
    }
    CU(ctx, cudaMemcpyAsync(e.d_raw[0], src, n, cudaMemcpyHostToDevice, s));
    ZhParams zp;
    memset(&zp, 0, sizeof zp);
    ZhJob &j = zp.job[0];
    j.src = e.d_raw[0]; j.n = n; j.slots = e.d_slots[0]; j.sizes = e.d_bsize[0]; j.stream = e.d_stream[0];
    j.scratch = ctx->zh_no_match ? nullptr : e.d_scratch[0];
    j.total = e.d_meta + 3; j.n_blocks = (uint32_t)nb;
    prof_begin(ctx, s, PCL_K_ZSTD);
    if (int r = zh_kernel_ready(ctx)) return r;
    k_zh_blocks<<<(unsigned)nb, ZH_THREADS, sizeof(ZhShared), s>>>(zp);
    k_zh_compact<<<(unsigned)nb, ZH_THREADS, 0, s>>>(zp);
    prof_end(ctx, s);
    ctx->launches += 1;
    CU(ctx, cudaGetLastError());
    CU(ctx, cudaMemcpyAsync(e.h_meta + 3, e.d_meta + 3, 8, cudaMemcpyDeviceToHost, s));
    CU(ctx, cudaStreamSynchronize(s));
    const uint64_t bytes = e.h_meta[3];
    if (ZH_FRAME_HEADER_BYTES + bytes > cap) return fail(ctx, PCL_ERR_NOMEM, "output buffer too small (%llu bytes needed)", (unsigned long long)(ZH_FRAME_HEADER_BYTES + bytes));
    CU(ctx, cudaMemcpy(dst + ZH_FRAME_HEADER_BYTES, e.d_stream[0], bytes, cudaMemcpyDeviceToHost));
    *used = ZH_FRAME_HEADER_BYTES + bytes;
    return PCL_OK;
}

// ------------------------------------------------------------------------------------------
// stages
// ------------------------------------------------------------------------------------------
int pcl_counts_reset(pcl_ctx *ctx) {
    if (!ctx) return PCL_ERR_ARG;
    if (int r = bind(ctx)) return r;
    // stream-ordered: after everything already queued on the chunks, before anything queued later
    for (pcl_chunk *c : ctx->chunks) {
        CU(ctx, cudaEventRecord(c->ev, c->stream));
        CU(ctx, cudaStreamWaitEvent(ctx->stream, c->ev, 0));
    }
    for (Table &t : ctx->tables) {
        if (!t.loaded) continue;
        CU(ctx, cudaMemsetAsync(t.d_counts, 0, t.n_slots * sizeof(uint32_t), ctx->stream));
        CU(ctx, cudaMemsetAsync(t.d_scalars, 0, 2 * sizeof(unsigned long long), ctx->stream));
    }
    CU(ctx, cudaMemsetAsync(ctx->d_qc, 0, (PCL_MAX_READS + 12) * sizeof(unsigned long long), ctx->stream));
    CU(ctx, cudaEventRecord(ctx->ev_reduced, ctx->stream));
    for (pcl_chunk *c : ctx->chunks) CU(ctx, cudaStreamWaitEvent(c->stream, ctx->ev_reduced, 0));
    memset(ctx->qc_reads, 0, sizeof ctx->qc_reads);
    ctx->reduced = false;
    return PCL_OK;
}

// CorrectParams of read r of a chunk (pass 2 and the filter stage of pass 1)
static void correct_params(pcl_ctx *ctx, pcl_chunk *c, int r, CorrectParams &p) {
    const pcl_read_info &inf = ctx->infos[r];
    ReadBuf &b = c->rb[r];
    memset(&p, 0, sizeof p);
    p.rd = ctx->dreads[r];
    for (int j = 0; j < inf.n_bc; ++j) {
        p.tab[j] = ctx->tables[ctx->bc_region[r][j]].d;
        p.bc_key[j] = b.bc_key[j]; p.bcoord[j] = b.bcoord[j];
        p.bc_key_bytes[j] = inf.bc_key_bytes[j]; p.bc_elem[j] = inf.bc_elem[j];
    }
    memcpy(p.static_start, inf.static_start, sizeof p.static_start);
    p.seq = b.seq; p.qual = b.qual; p.len = b.len; p.n = c->n;
    p.fstatus = b.fstatus;
    p.worklist = b.worklist; p.wl_key = b.wl_key; p.wl_cand = b.wl_cand; p.wl_count = b.wl_count;
    p.lut_cap = ctx->d_lut; p.lut_raw = ctx->d_lut + 256;
}

// worst-case grid of a kernel that walks the worklist (its size lives on the device; idle CTAs exit at once)
static int worklist_grid(const pcl_ctx *ctx, const pcl_chunk *c, int n_bc, bool all_entries, int occ) {
    return grid_for(ctx, all_entries ? c->n * n_bc : (c->n * n_bc + 3) / 4 + 4096, occ);
}

enum { TL_COUNT_BEGIN = 0, TL_COUNTED, TL_FILTER_END, TL_COUNT_END, TL_REDUCE_BEGIN, TL_GATHERED, TL_ALLREDUCED, TL_REDUCED, TL_RESOLVE_BEGIN,
       TL_RESOLVE_END, TL_N };
static void tl_mark(pcl_ctx *ctx, int k, cudaStream_t s) {
    if (!ctx->timeline) return;
    if (!ctx->tl[k]) cudaEventCreate(&ctx->tl[k]);
    cudaEventRecord(ctx->tl[k], s);
    ctx->tl_set[k] = true;
}
static void tl_print(pcl_ctx *ctx) {
    if (!ctx->timeline || !ctx->tl_set[TL_COUNT_BEGIN]) return;
    static const char *names[TL_N] = {"count_begin", "counted", "filter_end", "count_end", "reduce_begin", "gathered", "allreduced", "reduced",
                                      "resolve_begin", "resolve_end"};
    cudaDeviceSynchronize();
    fprintf(stderr, "[pcl timeline rank %d]", ctx->rank);
    for (int k = 0; k < TL_N; ++k) {
        float ms = 0;
        if (ctx->tl_set[k] && cudaEventElapsedTime(&ms, ctx->tl[TL_COUNT_BEGIN], ctx->tl[k]) == cudaSuccess) fprintf(stderr, " %s=%.3f", names[k], ms);
        ctx->tl_set[k] = false;
    }
    fprintf(stderr, "\n");
}

static int launch_filter(pcl_ctx *ctx, pcl_chunk *c, int r) {
    CorrectParams p;
    correct_params(ctx, c, r, p);
    prof_begin(ctx, c->stream, PCL_K_FILTER);
    // Sixteen short waves instead of one resident wave: CTA slots free up every ~70 us, so the gather / all-reduce /
    // scatter kernels waiting on the (high-priority) ctx stream start at once instead of a quarter of the kernel later.
    // Measured, v3, 125 M pairs per GPU: step at N=2 3.59 ms (4 waves) / 3.54 (8) / 3.49 (16) / 3.50 (32) / 3.52 (64);
    // at N=1 3.47 (4) / 3.44 (16).
    k_filter<<<worklist_grid(ctx, c, ctx->infos[r].n_bc, false, ctx->filter_waves * ctx->occ[OCC_FILTER]), PCL_BLOCK, 0, c->stream>>>(p);
    prof_end(ctx, c->stream);
    CU(ctx, cudaGetLastError());
    c->rb[r].filtered = true;
    return PCL_OK;
}

static int check_ready(pcl_ctx *ctx, pcl_chunk *c) {
    if (!ctx || !c || c->ctx != ctx) return PCL_ERR_ARG;
    for (int r = 0; r < ctx->n_reads; ++r) {
        if (!c->rb[r].filled) return fail(ctx, PCL_ERR_INPUT, "read %d of the chunk has no data (Unequal number of reads in the chunk)", r);
        for (int j = 0; j < ctx->infos[r].n_bc; ++j)
            if (!ctx->tables[ctx->bc_region[r][j]].loaded)
                return fail(ctx, PCL_ERR_ARG, "whitelist of region %d not loaded", ctx->bc_region[r][j]);
    }
    return PCL_OK;
}

int pcl_count(pcl_ctx *ctx, pcl_chunk *c) {
    if (int r = check_ready(ctx, c)) return r;
    if (int r = bind(ctx)) return r;
    c->has_results = true;
    if (ctx->reduced && ctx->nranks > 1 && !ctx->frozen)
        return fail(ctx, PCL_ERR_ARG, "the counts already hold the sum over ranks: call pcl_counts_reset before counting again");
    if (ctx->frozen) {                         // discarded increments go to scratch arrays (never read, never cleared)
        for (int r = 0; r < ctx->n_reads; ++r)
            for (int j = 0; j < ctx->infos[r].n_bc; ++j) {
                Table &t = ctx->tables[ctx->bc_region[r][j]];
                if (!t.d_scratch_counts) {
                    CU(ctx, cudaMalloc(&t.d_scratch_counts, t.n_slots * sizeof(uint32_t)));
                    CU(ctx, cudaMalloc(&t.d_scratch_scalars, 2 * sizeof(unsigned long long)));
                }
            }
        if (!ctx->d_qc_scratch) CU(ctx, cudaMalloc(&ctx->d_qc_scratch, (PCL_MAX_READS + 12) * sizeof(unsigned long long)));
    }
    if (c->n == 0) return PCL_OK;
    // Reads that hold barcodes first: the all-reduce of the counts only has to wait for their kernels (ev_counted), so it
    // runs while the length-only kernels of the insert reads are still going.
    int order[PCL_MAX_READS], n_order = 0, n_with_bc = 0;
    for (int r = 0; r < ctx->n_reads; ++r) if (ctx->infos[r].n_bc) order[n_order++] = r;
    n_with_bc = n_order;
    for (int r = 0; r < ctx->n_reads; ++r) if (!ctx->infos[r].n_bc) order[n_order++] = r;
    if (n_with_bc == 0) { CU(ctx, cudaEventRecord(c->ev_counted, c->stream)); c->counted = true; }
    tl_mark(ctx, TL_COUNT_BEGIN, c->stream);
    for (int oi = 0; oi < n_order; ++oi) {
        const int r = order[oi];
        const pcl_read_info &inf = ctx->infos[r];
        ReadBuf &b = c->rb[r];
        CountParams p;
        memset(&p, 0, sizeof p);
        p.rd = ctx->dreads[r];
        for (int j = 0; j < inf.n_bc; ++j) {
            const Table &t = ctx->tables[ctx->bc_region[r][j]];
            p.tab[j] = t.d;
            if (ctx->frozen) { p.tab[j].counts = t.d_scratch_counts; p.tab[j].total = t.d_scratch_scalars; p.tab[j].mismatch = t.d_scratch_scalars + 1; }
            p.bc_key[j] = b.bc_key[j]; p.bcoord[j] = b.bcoord[j];
            p.bc_key_bytes[j] = inf.bc_key_bytes[j]; p.bc_elem[j] = inf.bc_elem[j];
        }
        for (int j = 0; j < inf.n_umi; ++j) {
            p.umi_key[j] = b.umi_key[j]; p.ucoord[j] = b.ucoord[j];
            p.umi_key_bytes[j] = inf.umi_key_bytes[j]; p.umi_elem[j] = inf.umi_elem[j];
        }
        p.seq = b.seq; p.len = b.len; p.n = c->n;
        p.fstatus = b.fstatus; p.tcoord = b.tcoord;
        p.worklist = b.worklist; p.wl_count = b.wl_count;
        p.wl_key = read_all32(ctx, r) ? b.wl_key : nullptr;
        b.filtered = false;
        p.num_defect = (ctx->frozen ? ctx->d_qc_scratch : ctx->d_qc) + r;
        CU(ctx, cudaMemsetAsync(b.wl_count, 0, sizeof(unsigned long long), c->stream));
        const DRead &d = ctx->dreads[r];
        const bool fast = d.fast.enabled && ctx->tables[ctx->bc_region[r][0]].d.mode != PCL_TAB_K64 && !ctx->force_generic &&
                          c->n < 0xFF000000ull;      // 32-bit record indices inside k_count_fast
        const bool lenonly = !d.needs_payload && d.fixed_layout && !ctx->force_generic;
        bool plan = d.plan.enabled && !ctx->force_generic && !ctx->no_plan;
        for (int j = 0; j < inf.n_bc; ++j) plan = plan && ctx->tables[ctx->bc_region[r][j]].d.mode != PCL_TAB_K64;
        prof_begin(ctx, c->stream, d.needs_payload ? PCL_K_COUNT : PCL_K_COUNT_LENONLY);
        // The generic counting kernels run THREE resident waves rather than one: with first-come cache lines a CTA
        // that lives a third as long lets fewer one-off barcodes occupy lines (measured on the former fast kernel:
        // 2.79 ms at one wave, 2.49 ms at three).  k_count_fast has an admission filter instead and runs one wave.
        const bool spec = fast && d.fast.spec != PCL_FAST_SPEC_NONE && !d.fast.has_tail_target && !ctx->no_fast_spec && !ctx->old_fast &&
                          ctx->fast_bits == 14 && p.wl_key;
        if (spec) {
            // one 1024-thread CTA per SM (128 KB histogram cache), two records per thread and iteration
            p.hist_admit_mask = ctx->hist_admit_mask;
            p.hist_two_choice = ctx->hist_two_choice;
            const int threads = 1024, smem = 8 << 14;
            const uint64_t blocks = (c->n + 2 * threads - 1) / (2 * threads);
            const int grid = (int)std::max<uint64_t>(1, std::min<uint64_t>(blocks, (uint64_t)ctx->sm_count * ctx->fast_waves));
            const bool rv = d.is_reverse != 0;
            switch (d.fast.spec) {
                case PCL_FAST_SPEC_B0_U16x12:
                    if (rv) k_count_spec<PCL_FAST_SPEC_B0_U16x12, true><<<grid, threads, smem, c->stream>>>(p);
                    else if (ctx->no_prefetch) k_count_spec<PCL_FAST_SPEC_B0_U16x12, false, false><<<grid, threads, smem, c->stream>>>(p);
                    else k_count_spec<PCL_FAST_SPEC_B0_U16x12, false><<<grid, threads, smem, c->stream>>>(p);
                    break;
                case PCL_FAST_SPEC_B0:
                    if (rv) k_count_spec<PCL_FAST_SPEC_B0, true><<<grid, threads, smem, c->stream>>>(p);
                    else k_count_spec<PCL_FAST_SPEC_B0, false><<<grid, threads, smem, c->stream>>>(p);
                    break;
                default:
                    if (rv) k_count_spec<PCL_FAST_SPEC_B8, true><<<grid, threads, smem, c->stream>>>(p);
                    else k_count_spec<PCL_FAST_SPEC_B8, false><<<grid, threads, smem, c->stream>>>(p);
                    break;
            }
        }
        else if (fast) {
            p.hist_admit_mask = ctx->hist_admit_mask;
            p.hist_two_choice = ctx->hist_two_choice;
            const int threads = PCL_BLOCK << (ctx->fast_bits - 12), smem = 8 << ctx->fast_bits;
            uint64_t blocks = (c->n + threads - 1) / threads, cap = (uint64_t)ctx->sm_count * ctx->fast_waves * ctx->occ[OCC_COUNT_FAST];
            cap = std::min<uint64_t>(cap, (uint64_t)ctx->sm_count * 24 * (PCL_BLOCK / 32) / (threads / 32));   // worklist slack (pcl_chunk_create)
            const int grid = (int)std::max<uint64_t>(1, std::min(blocks, cap));
            if (ctx->fast_bits == 12) k_count_fast<12><<<grid, threads, smem, c->stream>>>(p);
            else if (ctx->fast_bits == 13) k_count_fast<13><<<grid, threads, smem, c->stream>>>(p);
            else if (d.fast.spec == PCL_FAST_SPEC_B0_U16x12 && !ctx->no_fast_spec) k_count_fast<14, PCL_FAST_SPEC_B0_U16x12><<<grid, threads, smem, c->stream>>>(p);
            else if (d.fast.spec == PCL_FAST_SPEC_B0 && !ctx->no_fast_spec) k_count_fast<14, PCL_FAST_SPEC_B0><<<grid, threads, smem, c->stream>>>(p);
            else if (d.fast.spec == PCL_FAST_SPEC_B8 && !ctx->no_fast_spec) k_count_fast<14, PCL_FAST_SPEC_B8><<<grid, threads, smem, c->stream>>>(p);
            else k_count_fast<14><<<grid, threads, smem, c->stream>>>(p);
        }
        else if (lenonly) k_count_lenonly<<<grid_for(ctx, c->n / 8 + 8, ctx->occ[OCC_COUNT_LENONLY]), PCL_BLOCK, 0, c->stream>>>(p);
        else if (plan && !ctx->old_fast && p.wl_key) {
            const int g = grid_for(ctx, c->n, 3 * ctx->occ[OCC_COUNT_PLAN2]);
            // compile-time geometry for the layouts that have an instance (pcl_plan_geo_matches), else the run-time plan
            if (!d.is_reverse && !ctx->no_plan_geo && pcl_plan_geo_matches<PlanGeoSci3>(d)) k_count_plan<false, PlanGeoSci3><<<g, PCL_BLOCK, 0, c->stream>>>(p);
            else if (d.is_reverse) k_count_plan<true, PlanGeoRt><<<g, PCL_BLOCK, 0, c->stream>>>(p);
            else k_count_plan<false, PlanGeoRt><<<g, PCL_BLOCK, 0, c->stream>>>(p);
        }
        else if (plan) k_count<2><<<grid_for(ctx, c->n, 3 * ctx->occ[OCC_COUNT_PLAN]), PCL_BLOCK, 0, c->stream>>>(p);
        else if (d.codes_ok && !ctx->force_generic) k_count<1><<<grid_for(ctx, c->n, 3 * ctx->occ[OCC_COUNT_CODES]), PCL_BLOCK, 0, c->stream>>>(p);
        else k_count<0><<<grid_for(ctx, c->n, 3 * ctx->occ[OCC_COUNT]), PCL_BLOCK, 0, c->stream>>>(p);
        prof_end(ctx, c->stream);
        CU(ctx, cudaGetLastError());
        if (!ctx->frozen) ctx->qc_reads[r] += c->n;
        if (oi + 1 == n_with_bc) {
            if (!ctx->reduce_after_filter) { CU(ctx, cudaEventRecord(c->ev_counted, c->stream)); c->counted = true; }
            tl_mark(ctx, TL_COUNTED, c->stream);
            // The neighbourhood filter of pass 2 needs no counts: it runs here, behind the event the all-reduce waits for,
            // so on several GPUs the exchange of the counts is hidden under it.
            for (int k = 0; k < n_with_bc; ++k)
                if (read_all32(ctx, order[k]) && c->rb[order[k]].wl_key)
                    if (int rc = launch_filter(ctx, c, order[k])) return rc;
            if (ctx->reduce_after_filter) { CU(ctx, cudaEventRecord(c->ev_counted, c->stream)); c->counted = true; }
            tl_mark(ctx, TL_FILTER_END, c->stream);
        }
    }
    tl_mark(ctx, TL_COUNT_END, c->stream);
    return PCL_OK;
}

int pcl_counts_freeze(pcl_ctx *ctx, int on) {
    if (!ctx) return PCL_ERR_ARG;
    ctx->frozen = on != 0;
    return PCL_OK;
}

uint64_t pcl_chunk_bytes(const pcl_ctx *ctx, uint64_t cap) {
    if (!ctx || ctx->n_reads == 0) return 0;
    uint64_t total = 0;
    auto A = [&](uint64_t b) { total += ((b ? b : 16) + 255) & ~255ull; };
    for (int r = 0; r < ctx->n_reads; ++r) {                       // mirrors pcl_chunk_create
        const DRead &d = ctx->dreads[r];
        const pcl_read_info &inf = ctx->infos[r];
        if (d.needs_payload) A(cap * d.stride);
        if ((d.needs_payload || d.qual_only) && d.qstride) A(cap * d.qstride);
        A(cap * 2 + 16); A(cap * 2 + 16);
        if (inf.has_target) A(cap * 4);
        for (int j = 0; j < inf.n_bc; ++j) { A(cap * inf.bc_key_bytes[j]); if (!inf.fixed_layout) A(cap * 4); }
        for (int j = 0; j < inf.n_umi; ++j) { A(cap * inf.umi_key_bytes[j]); if (!inf.fixed_layout) A(cap * 4); }
        if (inf.n_bc) {
            uint64_t warps = (uint64_t)ctx->sm_count * 24 * (PCL_BLOCK / 32), by_size = (cap + 31) / 32 + 32;
            if (warps > by_size) warps = by_size;
            const uint64_t wl_cap = (cap + warps * PCL_WL_BATCH) * inf.n_bc;
            A(wl_cap * 8);
            bool k32 = true;
            for (int j = 0; j < inf.n_bc; ++j) k32 = k32 && inf.bc_key_bytes[j] == 4;
            if (k32) { A(wl_cap * 4); A(wl_cap * 8); }
        }
        A(16);
    }
    return total;
}

int pcl_mem_info(pcl_ctx *ctx, uint64_t *free_bytes, uint64_t *total_bytes) {
    if (!ctx) return PCL_ERR_ARG;
    if (int r = bind(ctx)) return r;
    size_t f = 0, t = 0;
    CU(ctx, cudaMemGetInfo(&f, &t));
    if (free_bytes) *free_bytes = f;
    if (total_bytes) *total_bytes = t;
    return PCL_OK;
}

int pcl_counts_allreduce(pcl_ctx *ctx) {
    if (!ctx) return PCL_ERR_ARG;
    if (int r = bind(ctx)) return r;
    // ctx stream waits for every chunk's count pass (for the kernels that touch the counts, when pcl_count marked them) ...
    for (pcl_chunk *c : ctx->chunks) {
        if (c->counted) {
            CU(ctx, cudaStreamWaitEvent(ctx->stream, c->ev_counted, 0));
            c->counted = false;
        } else {
            CU(ctx, cudaEventRecord(c->ev, c->stream));
            CU(ctx, cudaStreamWaitEvent(ctx->stream, c->ev, 0));
        }
    }
    tl_mark(ctx, TL_REDUCE_BEGIN, ctx->stream);
    if (ctx->nranks > 1) {
        if (!ctx->comm) return fail(ctx, PCL_ERR_NCCL, "communicator not initialised");
        if (ctx->reduced) return fail(ctx, PCL_ERR_ARG, "the counts were already summed over the ranks (pcl_counts_reset starts a new round)");
        NcclApi &n = nccl();
        // Every rank builds the same table, slot for slot (host build: sequential; device build: placement by rank,
        // tablebuild.cuh), so the counts may travel in any order the ranks agree on.  Default: the n_entries occupied
        // slots in ascending slot order (a third of the bytes of the slot array at the table's load factor; gather and
        // scatter stream through the counts).  PCL_REDUCE_SLOTS=1: the whole slot array, no kernels around the collective.
        const bool by_slot = ctx->reduce_slots;
        if (!by_slot)
            for (Table &t : ctx->tables) {
                if (!t.loaded || !t.n_entries) continue;
                k_counts_gather32<<<grid_for(ctx, t.n_entries, 4), 256, 0, ctx->stream>>>(t.d_counts, t.d_order_slots, t.d_dense, t.n_entries);
                ctx->launches += 1;
            }
        tl_mark(ctx, TL_GATHERED, ctx->stream);
        n.GroupStart();
        for (Table &t : ctx->tables) {
            if (!t.loaded) continue;
            int r1 = 0;
            if (t.n_entries) r1 = by_slot ? n.AllReduce(t.d_counts, t.d_counts, t.n_slots, kNcclUint32, kNcclSum, ctx->comm, ctx->stream)
                                          : n.AllReduce(t.d_dense, t.d_dense, t.n_entries, kNcclUint32, kNcclSum, ctx->comm, ctx->stream);
            int r2 = n.AllReduce(t.d_scalars, t.d_scalars, 2, kNcclUint64, kNcclSum, ctx->comm, ctx->stream);
            if (r1 || r2) { n.GroupEnd(); return fail(ctx, PCL_ERR_NCCL, "ncclAllReduce failed (%d,%d)", r1, r2); }
        }
        int r = n.GroupEnd();
        if (r) return fail(ctx, PCL_ERR_NCCL, "ncclGroupEnd: %d", r);
        tl_mark(ctx, TL_ALLREDUCED, ctx->stream);
        if (!by_slot)
            for (Table &t : ctx->tables) {
                if (!t.loaded || !t.n_entries) continue;
                k_counts_scatter32<<<grid_for(ctx, t.n_entries, 4), 256, 0, ctx->stream>>>(t.d_counts, t.d_order_slots, t.d_dense, t.n_entries);
                ctx->launches += 1;
            }
        CU(ctx, cudaGetLastError());
    }
    ctx->reduced = true;
    tl_mark(ctx, TL_REDUCED, ctx->stream);
    // ... and every chunk's later work waits for the reduced counts
    CU(ctx, cudaEventRecord(ctx->ev_reduced, ctx->stream));
    for (pcl_chunk *c : ctx->chunks) CU(ctx, cudaStreamWaitEvent(c->stream, ctx->ev_reduced, 0));
    return PCL_OK;
}

int pcl_correct(pcl_ctx *ctx, pcl_chunk *c, double threshold, double max_expected_errors) {
    if (int r = check_ready(ctx, c)) return r;
    if (int r = bind(ctx)) return r;
    if (c->n == 0) return PCL_OK;
    const bool check_ee = max_expected_errors < DBL_MAX;
    for (int r = 0; r < ctx->n_reads; ++r) {
        const pcl_read_info &inf = ctx->infos[r];
        if (inf.n_bc == 0) continue;
        ReadBuf &b = c->rb[r];
        CorrectParams p;
        correct_params(ctx, c, r, p);
        p.threshold = threshold; p.max_expected_errors = max_expected_errors; p.check_ee = check_ee;
        if (check_ee) {
            // the expected-error test precedes the whitelist lookup: every present barcode needs its qualities
            CU(ctx, cudaMemsetAsync(b.wl_count, 0, sizeof(unsigned long long), c->stream));
            b.filtered = false;
            prof_begin(ctx, c->stream, PCL_K_ENUM);
            k_enum_all<<<grid_for(ctx, c->n, ctx->occ[OCC_ENUM]), PCL_BLOCK, 0, c->stream>>>(p);
            prof_end(ctx, c->stream);
            CU(ctx, cudaGetLastError());
        }
        const bool all32 = !check_ee && read_all32(ctx, r) && b.wl_key;
        if (all32) {
            if (!b.filtered) if (int rc = launch_filter(ctx, c, r)) return rc;
            tl_mark(ctx, TL_RESOLVE_BEGIN, c->stream);
            prof_begin(ctx, c->stream, PCL_K_CORRECT);
            k_resolve<<<worklist_grid(ctx, c, inf.n_bc, false, ctx->occ[OCC_RESOLVE]), PCL_BLOCK, 0, c->stream>>>(p);
        } else {
            prof_begin(ctx, c->stream, PCL_K_CORRECT);
            k_correct<<<worklist_grid(ctx, c, inf.n_bc, check_ee, ctx->occ[OCC_CORRECT]), PCL_BLOCK, 0, c->stream>>>(p);
        }
        prof_end(ctx, c->stream);
        CU(ctx, cudaGetLastError());
    }
    tl_mark(ctx, TL_RESOLVE_END, c->stream);
    tl_print(ctx);
    return PCL_OK;
}

// ------------------------------------------------------------------------------------------
// outputs
// ------------------------------------------------------------------------------------------
static int chunk_fetch(pcl_chunk *c, int read_slot, const pcl_read_results *out, bool wait) {
    if (!c || !out) return PCL_ERR_ARG;
    pcl_ctx *ctx = c->ctx;
    if (read_slot < 0 || read_slot >= ctx->n_reads) return fail(ctx, PCL_ERR_ARG, "bad read_slot");
    if (int r = bind(ctx)) return r;
    const pcl_read_info &inf = ctx->infos[read_slot];
    ReadBuf &b = c->rb[read_slot];
    const uint64_t n = c->n;
    auto D2H = [&](void *dst, const void *src, size_t bytes) -> cudaError_t {
        if (!dst || !src || !bytes) return cudaSuccess;
        return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c->stream);
    };
    CU(ctx, D2H(out->fstatus, b.fstatus, n * 2));
    if (inf.has_target) CU(ctx, D2H(out->tcoord, b.tcoord, n * 4));
    for (int j = 0; j < inf.n_bc; ++j) {
        CU(ctx, D2H(out->bc_key[j], b.bc_key[j], n * inf.bc_key_bytes[j]));
        if (!inf.fixed_layout) CU(ctx, D2H(out->bcoord[j], b.bcoord[j], n * 4));
    }
    for (int j = 0; j < inf.n_umi; ++j) {
        CU(ctx, D2H(out->umi_key[j], b.umi_key[j], n * inf.umi_key_bytes[j]));
        if (!inf.fixed_layout) CU(ctx, D2H(out->ucoord[j], b.ucoord[j], n * 4));
    }
    if (wait) CU(ctx, cudaStreamSynchronize(c->stream));
    return PCL_OK;
}

int pcl_chunk_fetch(pcl_chunk *c, int read_slot, const pcl_read_results *out) { return chunk_fetch(c, read_slot, out, true); }
int pcl_chunk_fetch_async(pcl_chunk *c, int read_slot, const pcl_read_results *out) { return chunk_fetch(c, read_slot, out, false); }

int pcl_counts_fetch(pcl_ctx *ctx, int region_slot, uint64_t *counts, uint64_t *mismatch, uint64_t *total) {
    if (!ctx || region_slot < 0 || region_slot >= PCL_MAX_REGIONS) return PCL_ERR_ARG;
    Table &t = ctx->tables[region_slot];
    if (!t.loaded) return fail(ctx, PCL_ERR_ARG, "whitelist of region %d not loaded", region_slot);
    if (int r = pcl_sync(ctx)) return r;
    if (counts && t.n_entries) {
        unsigned long long *d_out = nullptr;
        CU(ctx, cudaMalloc(&d_out, t.n_entries * sizeof(unsigned long long)));
        ctx->launches += 1;
        k_gather_counts<<<grid_for(ctx, t.n_entries, ctx->occ[OCC_GATHER]), PCL_BLOCK, 0, ctx->stream>>>(t.d_counts, t.d_index_to_slot, d_out, t.n_entries);
        cudaError_t e = cudaMemcpyAsync(counts, d_out, t.n_entries * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        cudaFree(d_out);
        if (e != cudaSuccess) return fail(ctx, PCL_ERR_CUDA, "counts fetch: %s", cudaGetErrorString(e));
    }
    unsigned long long sc[2];
    CU(ctx, cudaMemcpy(sc, t.d_scalars, sizeof sc, cudaMemcpyDeviceToHost));
    if (total) *total = sc[0];
    if (mismatch) *mismatch = sc[1];
    return PCL_OK;
}

int pcl_qc_fetch(pcl_ctx *ctx, uint64_t *per_read) {
    if (!ctx || !per_read) return PCL_ERR_ARG;
    if (int r = pcl_sync(ctx)) return r;
    unsigned long long d[PCL_MAX_READS];
    CU(ctx, cudaMemcpy(d, ctx->d_qc, sizeof d, cudaMemcpyDeviceToHost));
    for (int r = 0; r < ctx->n_reads; ++r) { per_read[2 * r] = ctx->qc_reads[r]; per_read[2 * r + 1] = d[r]; }
    return PCL_OK;
}

int pcl_qc_accumulate(pcl_ctx *ctx, pcl_chunk *c) {
    if (int r = check_ready(ctx, c)) return r;
    if (int r = bind(ctx)) return r;
    if (c->n == 0) return PCL_OK;
    for (int r = 0; r < ctx->n_reads; ++r) {
        const pcl_read_info &inf = ctx->infos[r];
        const DRead &d = ctx->dreads[r];
        if (!inf.n_bc && !inf.n_umi && !inf.has_target) continue;
        if (inf.has_target && !ctx->qc_enabled)
            return fail(ctx, PCL_ERR_ARG, "read %d holds a target: call pcl_qc_enable before pcl_layout_set", r);
        ReadBuf &b = c->rb[r];
        QcParams p;
        memset(&p, 0, sizeof p);
        p.rd = d;
        p.qual = b.qual; p.len = b.len; p.fstatus = b.fstatus; p.tcoord = b.tcoord; p.n = c->n;
        for (int j = 0; j < inf.n_bc; ++j) { p.bcoord[j] = b.bcoord[j]; p.bc_elem[j] = inf.bc_elem[j]; }
        for (int j = 0; j < inf.n_umi; ++j) { p.ucoord[j] = b.ucoord[j]; p.umi_elem[j] = inf.umi_elem[j]; }
        memcpy(p.static_start, inf.static_start, sizeof p.static_start);
        p.cat = ctx->d_qc + PCL_MAX_READS;
        prof_begin(ctx, c->stream, PCL_K_QC);
        k_qc<<<grid_for(ctx, c->n, ctx->occ[OCC_QC]), PCL_BLOCK, 0, c->stream>>>(p);
        prof_end(ctx, c->stream);
        CU(ctx, cudaGetLastError());
    }
    return PCL_OK;
}

int pcl_qc_fetch_categories(pcl_ctx *ctx, uint64_t *cat12) {
    if (!ctx || !cat12) return PCL_ERR_ARG;
    if (int r = pcl_sync(ctx)) return r;
    unsigned long long d[12];
    CU(ctx, cudaMemcpy(d, ctx->d_qc + PCL_MAX_READS, sizeof d, cudaMemcpyDeviceToHost));
    for (int k = 0; k < 12; ++k) cat12[k] = d[k];
    return PCL_OK;
}

int pcl_key_decode(uint64_t key, char *out32) {
    if (!out32 || key == 0 || key == PCL_KEY_ABSENT64) return -1;
    int top = 63;
    while (top >= 0 && !((key >> top) & 1ull)) --top;
    if (top & 1) return -1;                      // marker must sit on an even bit
    int L = top / 2;
    static const char B[4] = {'A', 'C', 'G', 'T'};
    for (int p = 0; p < L; ++p) out32[p] = B[(key >> (2 * (L - 1 - p))) & 3ull];
    out32[L] = 0;
    return L;
}

// ------------------------------------------------------------------------------------------
// host-side assembly of make_fastq records from the result planes and the record text
// ------------------------------------------------------------------------------------------
static inline void rc_append(std::string &dst, const uint8_t *p, uint32_t n) {       // precellar::utils::rev_compl
    for (uint32_t i = 0; i < n; ++i) {
        uint8_t x = p[n - 1 - i];
        switch (x) {
            case 'A': case 'a': x = 'T'; break;
            case 'T': case 't': x = 'A'; break;
            case 'C': case 'c': x = 'G'; break;
            case 'G': case 'g': x = 'C'; break;
            default: break;
        }
        dst.push_back((char)x);
    }
}
static inline void rev_append(std::string &dst, const uint8_t *p, uint32_t n) {
    for (uint32_t i = 0; i < n; ++i) dst.push_back((char)p[n - 1 - i]);
}
static inline uint32_t static_coord(const pcl_read_desc &rd, const pcl_read_info &inf, uint32_t e, uint32_t L) {
    const uint32_t s = inf.static_start[e], mn = rd.elems[e].min_len, mx = rd.elems[e].max_len;
    if (s >= L || s + mn > L) return PCL_COORD_ABSENT;
    const uint32_t hi = s + mx < L ? s + mx : L;
    return (s << 16) | (hi - s);
}

// One contiguous range of read groups -> FASTQ text appended to o1 / o2.  Returns 0 or a negative status with *msg set.
static int make_fastq_range(const pcl_ctx *ctx, const pcl_host_read *reads, uint64_t i0, uint64_t i1, int correct_barcode, int paired,
                            std::string &o1, std::string &o2, uint64_t *n_written, std::string *msg) {
    const int nr = ctx->n_reads;
    uint64_t written = 0;
    std::string bseq, bqual, corr, useq, uqual, piece;
    auto bad = [&](int code, const char *m) { *msg = m; return code; };
    for (uint64_t i = i0; i < i1; ++i) {
        bseq.clear(); bqual.clear(); corr.clear(); useq.clear(); uqual.clear();
        bool has_bc = false, all_ok = true;
        int bc_def = -1, r1_file = -1, r2_file = -1;
        uint32_t r1c = 0, r2c = 0;
        for (int r = 0; r < nr; ++r) {
            const pcl_read_desc &rd = ctx->descs[r];
            const pcl_read_info &inf = ctx->infos[r];
            const pcl_host_read &h = reads[r];
            const uint32_t fs = h.res.fstatus[i];
            const uint32_t st = fs & 3u;
            if (st == PCL_SPLIT_MULTI_TARGET) return bad(PCL_ERR_INPUT, "Multiple target regions found in one fastq record!");
            if (st != PCL_SPLIT_OK) continue;                                    // a defect file contributes nothing
            uint32_t L = h.len[i];
            if (rd.max_len && L > rd.max_len) L = rd.max_len;
            const uint8_t *seq = h.text + h.off[3 * i + 1], *qual = h.text + h.off[3 * i + 2];
            for (int j = 0; j < inf.n_bc; ++j) {
                const uint32_t code = PCL_BC_CODE(fs, j);
                if (code == PCL_BC_ABSENT) continue;
                const uint32_t e = inf.bc_elem[j];
                const uint32_t c = inf.fixed_layout ? static_coord(rd, inf, e, L) : h.res.bcoord[j][i];
                if (c == PCL_COORD_ABSENT) return bad(PCL_ERR_INPUT, "inconsistent barcode coordinates");
                const uint32_t s0 = c >> 16, ln = c & 0xFFFFu;
                piece.clear();
                if (rd.is_reverse) { rc_append(piece, seq + s0, ln); rev_append(bqual, qual + s0, ln); }
                else { piece.assign((const char *)seq + s0, ln); bqual.append((const char *)qual + s0, ln); }
                bseq += piece;
                if (bc_def < 0) bc_def = r;
                has_bc = true;
                if (!correct_barcode || code == PCL_BC_EXACT) corr += piece;
                else if (code == PCL_BC_CORRECTED) {
                    uint64_t key = inf.bc_key_bytes[j] == 4 ? ((const uint32_t *)h.res.bc_key[j])[i] : ((const uint64_t *)h.res.bc_key[j])[i];
                    if (inf.bc_key_bytes[j] == 4 && rd.elems[e].min_len == 16 && rd.elems[e].max_len == 16) key |= 1ull << 32;
                    char buf[40];
                    const int kl = pcl_key_decode(key, buf);
                    if (kl < 0) return bad(PCL_ERR_INPUT, "corrupt corrected key");
                    corr.append(buf, kl);
                } else all_ok = false;
            }
            uint32_t uc = PCL_COORD_ABSENT;                                      // the last UMI segment of a file wins (fastq.rs:502)
            for (int j = 0; j < inf.n_umi; ++j) {
                const uint32_t c = inf.fixed_layout ? static_coord(rd, inf, inf.umi_elem[j], L) : h.res.ucoord[j][i];
                if (c != PCL_COORD_ABSENT) uc = c;
            }
            if (uc != PCL_COORD_ABSENT) {
                const uint32_t s0 = uc >> 16, ln = uc & 0xFFFFu;
                if (rd.is_reverse) { rc_append(useq, seq + s0, ln); rev_append(uqual, qual + s0, ln); }
                else { useq.append((const char *)seq + s0, ln); uqual.append((const char *)qual + s0, ln); }
            }
            if (inf.has_target && (fs & PCL_FSTATUS_TARGET)) {
                if (rd.is_reverse) {
                    if (r2_file >= 0) return bad(PCL_ERR_INPUT, "Read2 already exists");
                    r2_file = r; r2c = h.res.tcoord[i];
                } else {
                    if (r1_file >= 0) return bad(PCL_ERR_INPUT, "Read1 already exists");
                    r1_file = r; r1c = h.res.tcoord[i];
                }
            }
        }
        if (!has_bc) continue;                                                   // dropped by process_chunk (fastq.rs:381-385)
        if (correct_barcode && !all_ok) continue;                                // lib.rs:154
        if (r1_file < 0) return bad(PCL_ERR_INPUT, "called `Option::unwrap()` on a `None` value: read1");
        if (paired && r2_file < 0) return bad(PCL_ERR_INPUT, "called `Option::unwrap()` on a `None` value: read2");
        {
            const pcl_host_read &hb = reads[bc_def], &h1 = reads[r1_file];
            const uint32_t s0 = r1c >> 16, ln = r1c & 0xFFFFu;
            o1 += '@'; o1.append((const char *)hb.text + hb.off[3 * i], hb.off[3 * i + 1] - hb.off[3 * i]); o1 += '\n';
            o1 += corr; o1 += useq; o1.append((const char *)h1.text + h1.off[3 * i + 1] + s0, ln);
            o1 += "\n+\n";
            o1 += bqual; o1 += uqual; o1.append((const char *)h1.text + h1.off[3 * i + 2] + s0, ln);
            o1 += '\n';
        }
        if (paired) {
            const pcl_host_read &h2 = reads[r2_file];
            const uint32_t s0 = r2c >> 16, ln = r2c & 0xFFFFu;
            o2 += '@'; o2.append((const char *)h2.text + h2.off[3 * i], h2.off[3 * i + 1] - h2.off[3 * i]); o2 += '\n';
            o2.append((const char *)h2.text + h2.off[3 * i + 1] + s0, ln);
            o2 += "\n+\n";
            o2.append((const char *)h2.text + h2.off[3 * i + 2] + s0, ln);
            o2 += '\n';
        }
        ++written;
    }
    *n_written = written;
    return PCL_OK;
}

int pcl_make_fastq_batch(pcl_ctx *ctx, const pcl_host_read *reads, uint64_t n, int correct_barcode, int paired,
                         uint8_t *r1_out, uint64_t r1_cap, uint64_t *r1_used, uint8_t *r2_out, uint64_t r2_cap,
                         uint64_t *r2_used, uint64_t *n_written) {
    if (!ctx || !reads || !r1_out || !r1_used) return PCL_ERR_ARG;
    if (paired && (!r2_out || !r2_used)) return PCL_ERR_ARG;
    // contiguous ranges, one host thread each (the reference joins its records under rayon, fastq.rs:341-348); the pieces
    // are concatenated in range order, so the output does not depend on the number of threads
    unsigned T = std::thread::hardware_concurrency();
    T = std::max(1u, std::min(T ? T : 1u, 16u));
    if (n < 32768) T = 1;
    std::vector<std::string> o1(T), o2(T), msgs(T);
    std::vector<uint64_t> wr(T, 0);
    std::vector<int> rcs(T, PCL_OK);
    auto work = [&](unsigned t) {
        const uint64_t i0 = n * t / T, i1 = n * (t + 1) / T;
        size_t est = 8 * (size_t)(i1 - i0) + 64;                  // the text of the range's records bounds what it can emit
        for (int r = 0; r < ctx->n_reads; ++r) est += (size_t)(reads[r].off[3 * i1] - reads[r].off[3 * i0]);
        o1[t].reserve(est);
        if (paired) o2[t].reserve(est);
        rcs[t] = make_fastq_range(ctx, reads, i0, i1, correct_barcode, paired, o1[t], o2[t], &wr[t], &msgs[t]);
    };
    if (T == 1) work(0);
    else {
        std::vector<std::thread> th;
        for (unsigned t = 0; t < T; ++t) th.emplace_back(work, t);
        for (auto &x : th) x.join();
    }
    uint64_t u1 = *r1_used, u2 = r2_used ? *r2_used : 0, written = 0;
    std::vector<uint64_t> at1(T), at2(T);
    for (unsigned t = 0; t < T; ++t) {
        if (rcs[t] != PCL_OK) return fail(ctx, rcs[t], "%s", msgs[t].c_str());      // the first failing range in input order
        at1[t] = u1; at2[t] = u2;
        u1 += o1[t].size();
        if (paired) u2 += o2[t].size();
        written += wr[t];
    }
    if (u1 > r1_cap) return fail(ctx, PCL_ERR_NOMEM, "R1 output buffer too small");
    if (paired && u2 > r2_cap) return fail(ctx, PCL_ERR_NOMEM, "R2 output buffer too small");
    auto place = [&](unsigned t) {
        memcpy(r1_out + at1[t], o1[t].data(), o1[t].size());
        if (paired) memcpy(r2_out + at2[t], o2[t].data(), o2[t].size());
    };
    if (T == 1) place(0);
    else {
        std::vector<std::thread> th;
        for (unsigned t = 0; t < T; ++t) th.emplace_back(place, t);
        for (auto &x : th) x.join();
    }
    *r1_used = u1;
    if (r2_used) *r2_used = u2;
    if (n_written) *n_written = written;
    return PCL_OK;
}

// ------------------------------------------------------------------------------------------
// per-barcode grouping (group.cuh)
// ------------------------------------------------------------------------------------------
}  // extern "C"

struct pcl_group {
    pcl_ctx *ctx = nullptr;
    pcl_group_info info{};
    uint32_t *d_order = nullptr, *d_frag_start = nullptr, *d_bc_frag_start = nullptr;
    unsigned long long *d_bc_key = nullptr;
    std::vector<void *> allocs;
};

namespace {

// exclusive scan of n uint32 (out may alias in); *d_total receives the sum.  tmp: (n / PCL_SCAN_TILE + 1) uint32.
int device_scan(pcl_ctx *ctx, cudaStream_t s, const uint32_t *in, uint32_t *out, uint64_t n, uint32_t *tmp, unsigned long long *d_total) {
    const uint64_t tiles = (n + PCL_SCAN_TILE - 1) / PCL_SCAN_TILE;
    if (tiles == 0) { CU(ctx, cudaMemsetAsync(d_total, 0, 8, s)); return PCL_OK; }
    k_scan_tile_sums<<<(unsigned)tiles, 256, 0, s>>>(in, n, tmp);
    k_scan_sums<<<1, 256, 0, s>>>(tmp, tiles, d_total);
    k_scan_apply<<<(unsigned)tiles, 256, 0, s>>>(in, out, n, tmp);
    ctx->launches += 3;
    CU(ctx, cudaGetLastError());
    return PCL_OK;
}

// stable sort of (keys, vals) on the bytes of the keys named by byte_mask; the result is in *keys / *vals (the buffers swap)
int device_radix_sort(pcl_ctx *ctx, cudaStream_t s, unsigned long long **keys, uint32_t **vals, unsigned long long **keys_tmp,
                      uint32_t **vals_tmp, uint64_t n, uint32_t byte_mask, uint32_t *hist, uint32_t *scan_tmp, unsigned long long *d_scratch) {
    const uint32_t n_ctas = (uint32_t)((n + PCL_SORT_TILE - 1) / PCL_SORT_TILE);
    if (n_ctas == 0) return PCL_OK;
    for (uint32_t b = 0; b < 8; ++b) {
        if (!((byte_mask >> b) & 1u)) continue;
        k_radix_hist<<<n_ctas, PCL_SORT_BLOCK, 0, s>>>(*keys, n, 8u * b, hist, n_ctas);
        ctx->launches += 1;
        if (int rc = device_scan(ctx, s, hist, hist, 256ull * n_ctas, scan_tmp, d_scratch)) return rc;
        k_radix_scatter<<<n_ctas, PCL_SORT_BLOCK, 0, s>>>(*keys, *vals, *keys_tmp, *vals_tmp, n, 8u * b, hist, n_ctas);
        ctx->launches += 1;
        CU(ctx, cudaGetLastError());
        std::swap(*keys, *keys_tmp);
        std::swap(*vals, *vals_tmp);
    }
    return PCL_OK;
}

uint32_t bytes_used(unsigned long long or_of_keys) {
    uint32_t m = 0;
    for (uint32_t b = 0; b < 8; ++b) if ((or_of_keys >> (8 * b)) & 0xFFull) m |= 1u << b;
    return m;
}

}  // namespace

extern "C" {

// Key parameters of one chunk (which planes hold the barcode keys / statuses / UMI bytes); outputs are set by the caller.
static int group_key_params(pcl_ctx *ctx, pcl_chunk *c, GroupKeyParams &p) {
    memset(&p, 0, sizeof p);
    int umi_read = -1;
    for (int r = 0; r < ctx->n_reads; ++r) {
        const pcl_read_info &inf = ctx->infos[r];
        const pcl_read_desc &rd = ctx->descs[r];
        for (int j = 0; j < inf.n_bc; ++j) {
            if (p.n_seg >= PCL_GROUP_MAX_SEG) return fail(ctx, PCL_ERR_UNSUPPORTED, "more than %d barcode segments", PCL_GROUP_MAX_SEG);
            const pcl_elem_desc &el = rd.elems[inf.bc_elem[j]];
            GroupSeg &g = p.seg[p.n_seg++];
            g.key = c->rb[r].bc_key[j];
            g.fstatus = c->rb[r].fstatus;
            g.key_bytes = inf.bc_key_bytes[j];
            g.j = (uint8_t)j;
            g.fixed_len = (inf.bc_key_bytes[j] == 4 && el.min_len == 16 && el.max_len == 16) ? 16 : 0;
        }
        if (inf.n_umi) {
            if (umi_read >= 0) return fail(ctx, PCL_ERR_UNSUPPORTED, "grouping supports UMIs in one read file");
            umi_read = r;
        }
    }
    if (p.n_seg == 0) return fail(ctx, PCL_ERR_ARG, "the layout has no barcode");
    if (umi_read >= 0) {
        const pcl_read_info &inf = ctx->infos[umi_read];
        const pcl_read_desc &rd = ctx->descs[umi_read];
        const DRead &d = ctx->dreads[umi_read];
        const ReadBuf &b = c->rb[umi_read];
        const int ju = inf.n_umi - 1;                                            // the last UMI segment of a file wins (fastq.rs:502)
        const pcl_elem_desc &el = rd.elems[inf.umi_elem[ju]];
        p.useq = b.seq; p.ulen = b.len; p.ufstatus = b.fstatus;
        p.ucoord = inf.fixed_layout ? nullptr : b.ucoord[ju];
        p.ustride = d.stride; p.umax_len = d.max_len;
        p.ustart = inf.static_start[inf.umi_elem[ju]]; p.ubases = el.max_len; p.umin = el.min_len;
        if (p.ustart + p.ubases > d.stride && inf.fixed_layout) return fail(ctx, PCL_ERR_UNSUPPORTED, "UMI outside the uploaded record bytes");
    }
    p.n = c->n;
    return PCL_OK;
}

int pcl_group_build(pcl_ctx *ctx, pcl_chunk *c, const uint64_t *fingerprint, pcl_group **out) {
    return pcl_group_build_multi(ctx, &c, 1, fingerprint, out);
}

int pcl_group_build_multi(pcl_ctx *ctx, pcl_chunk *const *chunks, int n_chunks, const uint64_t *fingerprint, pcl_group **out) {
    if (!out) return PCL_ERR_ARG;
    *out = nullptr;
    if (!ctx || !chunks || n_chunks < 1) return PCL_ERR_ARG;
    uint64_t n = 0;
    std::vector<GroupKeyParams> kp((size_t)n_chunks);
    for (int k = 0; k < n_chunks; ++k) {
        if (int r = check_ready(ctx, chunks[k])) return r;
        for (int q = 0; q < k; ++q) if (chunks[q] == chunks[k]) return fail(ctx, PCL_ERR_ARG, "chunk %d is listed twice", k);
        if (int r = group_key_params(ctx, chunks[k], kp[(size_t)k])) return r;
        n += chunks[k]->n;
    }
    if (int r = bind(ctx)) return r;
    if (n >= (1ull << 32) - 1) return fail(ctx, PCL_ERR_ARG, "too many read groups for 32-bit record indices");
    const bool has_umi = kp[0].useq != nullptr;

    pcl_group *g = new pcl_group();
    g->ctx = ctx;
    g->info.n_records = n;
    cudaStream_t s = chunks[0]->stream;
    int rc = PCL_OK;
    auto A = [&](void **ptr, size_t bytes) {
        if (rc != PCL_OK) return;
        cudaError_t e = cudaMalloc(ptr, bytes ? bytes : 16);
        if (e != cudaSuccess) { rc = fail(ctx, e == cudaErrorMemoryAllocation ? PCL_ERR_NOMEM : PCL_ERR_CUDA, "cudaMalloc(%zu): %s", bytes, cudaGetErrorString(e)); return; }
        g->allocs.push_back(*ptr);
    };
    unsigned long long *d_bc = nullptr, *d_umi = nullptr, *d_fp = nullptr, *kA = nullptr, *kB = nullptr, *d_stats = nullptr;
    uint32_t *pA = nullptr, *pB = nullptr, *hist = nullptr, *scan_tmp = nullptr;
    const uint64_t n_ctas = (n + PCL_SORT_TILE - 1) / PCL_SORT_TILE;
    A((void **)&d_bc, n * 8); A((void **)&d_umi, n * 8); A((void **)&kA, n * 8); A((void **)&kB, n * 8);
    A((void **)&pA, n * 4); A((void **)&pB, n * 4);
    A((void **)&hist, 256 * n_ctas * 4 + 16); A((void **)&scan_tmp, (256 * n_ctas / PCL_SCAN_TILE + n / PCL_SCAN_TILE + 4) * 4);
    A((void **)&d_stats, 8 * 8);
    if (fingerprint) A((void **)&d_fp, n * 8);
    auto cleanup = [&](int code) { for (void *q : g->allocs) cudaFree(q); delete g; return code; };
    if (rc != PCL_OK) return cleanup(rc);
#define GCU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return cleanup(fail(ctx, PCL_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_))); } while (0)
    GCU(cudaMemsetAsync(d_stats, 0, 64, s));
    if (fingerprint) GCU(cudaMemcpyAsync(d_fp, fingerprint, n * 8, cudaMemcpyHostToDevice, s));
    unsigned long long st[8] = {0};
    if (n) {
        // the chunks' planes are final once their own streams have drained up to here; the grouping runs on chunk 0's
        uint64_t base = 0;
        for (int k = 0; k < n_chunks; ++k) {
            pcl_chunk *ck = chunks[k];
            if (k > 0) { GCU(cudaEventRecord(ck->ev, ck->stream)); GCU(cudaStreamWaitEvent(s, ck->ev, 0)); }
            GroupKeyParams &p = kp[(size_t)k];
            p.bc_key = d_bc + base; p.umi_key = d_umi + base; p.perm = pA + base; p.perm_base = (uint32_t)base; p.stats = d_stats;
            if (p.n) {
                k_group_keys<<<grid_for(ctx, p.n, 8), 256, 0, s>>>(p);
                ctx->launches += 1;
            }
            base += p.n;
        }
        GCU(cudaGetLastError());
        GCU(cudaMemcpyAsync(st, d_stats, 64, cudaMemcpyDeviceToHost, s));
        GCU(cudaStreamSynchronize(s));
    }
    g->info.n_assigned = st[0];
    g->info.n_ungroupable = st[1];
    // three stable sorts, least significant key first: UMI, fingerprint, barcode
    if (n) {
        if (has_umi && st[3]) {
            k_group_gather<<<grid_for(ctx, n, 8), 256, 0, s>>>(d_umi, pA, kA, n, nullptr);
            ctx->launches += 1;
            if (int r2 = device_radix_sort(ctx, s, &kA, &pA, &kB, &pB, n, bytes_used(st[3]), hist, scan_tmp, d_stats + 4)) return cleanup(r2);
        }
        if (fingerprint) {
            GCU(cudaMemsetAsync(d_stats + 5, 0, 8, s));
            k_group_gather<<<grid_for(ctx, n, 8), 256, 0, s>>>(d_fp, pA, kA, n, d_stats + 5);
            ctx->launches += 1;
            unsigned long long orfp = 0;
            GCU(cudaMemcpyAsync(&orfp, d_stats + 5, 8, cudaMemcpyDeviceToHost, s));
            GCU(cudaStreamSynchronize(s));
            if (int r2 = device_radix_sort(ctx, s, &kA, &pA, &kB, &pB, n, bytes_used(orfp), hist, scan_tmp, d_stats + 4)) return cleanup(r2);
        }
        k_group_gather<<<grid_for(ctx, n, 8), 256, 0, s>>>(d_bc, pA, kA, n, nullptr);
        ctx->launches += 1;
        // unassigned read groups carry the all-ones key: byte 0 tells them from every real key, so it is always sorted
        if (int r2 = device_radix_sort(ctx, s, &kA, &pA, &kB, &pB, n, bytes_used(st[2]) | 1u, hist, scan_tmp, d_stats + 4)) return cleanup(r2);
    }
    const uint64_t na = g->info.n_assigned;
    uint32_t *head_bc = nullptr, *head_frag = nullptr, *bc_id = nullptr, *frag_id = nullptr;
    A((void **)&head_bc, na * 4); A((void **)&head_frag, na * 4); A((void **)&bc_id, na * 4); A((void **)&frag_id, na * 4);
    if (rc != PCL_OK) return cleanup(rc);
    unsigned long long tot[2] = {0, 0};
    if (na) {
        k_group_flags<<<grid_for(ctx, na, 8), 256, 0, s>>>(kA, pA, d_umi, d_fp, na, head_bc, head_frag);
        ctx->launches += 1;
        if (int r2 = device_scan(ctx, s, head_bc, bc_id, na, scan_tmp, d_stats + 6)) return cleanup(r2);
        if (int r2 = device_scan(ctx, s, head_frag, frag_id, na, scan_tmp, d_stats + 7)) return cleanup(r2);
        GCU(cudaMemcpyAsync(tot, d_stats + 6, 16, cudaMemcpyDeviceToHost, s));
        GCU(cudaStreamSynchronize(s));
    }
    g->info.n_barcodes = tot[0];
    g->info.n_fragments = tot[1];
    A((void **)&g->d_frag_start, (tot[1] + 1) * 4); A((void **)&g->d_bc_frag_start, (tot[0] + 1) * 4); A((void **)&g->d_bc_key, (tot[0] + 1) * 8);
    if (rc != PCL_OK) return cleanup(rc);
    if (na) {
        k_group_emit<<<grid_for(ctx, na, 8), 256, 0, s>>>(kA, head_bc, head_frag, bc_id, frag_id, na, g->d_frag_start, g->d_bc_frag_start, g->d_bc_key);
        ctx->launches += 1;
        GCU(cudaGetLastError());
    }
    const uint32_t end_frag = (uint32_t)na, end_bc = (uint32_t)tot[1];
    GCU(cudaMemcpyAsync(g->d_frag_start + tot[1], &end_frag, 4, cudaMemcpyHostToDevice, s));
    GCU(cudaMemcpyAsync(g->d_bc_frag_start + tot[0], &end_bc, 4, cudaMemcpyHostToDevice, s));
    GCU(cudaStreamSynchronize(s));
#undef GCU
    g->d_order = pA;
    // the scratch buffers are no longer needed: keep only what pcl_group_fetch hands out
    for (void *q : g->allocs)
        if (q != (void *)g->d_order && q != (void *)g->d_frag_start && q != (void *)g->d_bc_frag_start && q != (void *)g->d_bc_key) cudaFree(q);
    g->allocs = {(void *)g->d_order, (void *)g->d_frag_start, (void *)g->d_bc_frag_start, (void *)g->d_bc_key};
    *out = g;
    return PCL_OK;
}

int pcl_group_info_get(const pcl_group *g, pcl_group_info *out) {
    if (!g || !out) return PCL_ERR_ARG;
    *out = g->info;
    return PCL_OK;
}

int pcl_group_fetch(pcl_group *g, uint32_t *order, uint32_t *frag_start, uint32_t *bc_frag_start, uint64_t *bc_key) {
    if (!g) return PCL_ERR_ARG;
    pcl_ctx *ctx = g->ctx;
    if (int r = bind(ctx)) return r;
    if (order && g->info.n_assigned) CU(ctx, cudaMemcpy(order, g->d_order, g->info.n_assigned * 4, cudaMemcpyDeviceToHost));
    if (frag_start) CU(ctx, cudaMemcpy(frag_start, g->d_frag_start, (g->info.n_fragments + 1) * 4, cudaMemcpyDeviceToHost));
    if (bc_frag_start) CU(ctx, cudaMemcpy(bc_frag_start, g->d_bc_frag_start, (g->info.n_barcodes + 1) * 4, cudaMemcpyDeviceToHost));
    if (bc_key && g->info.n_barcodes) CU(ctx, cudaMemcpy(bc_key, g->d_bc_key, g->info.n_barcodes * 8, cudaMemcpyDeviceToHost));
    return PCL_OK;
}

int pcl_group_key_decode(uint64_t key, char *out32) {
    if (!out32 || key == PCL_GROUP_NOKEY) return -1;
    const int L = (int)(key & 31ull);
    if (L > 29) return -1;
    static const char B[4] = {'A', 'C', 'G', 'T'};
    for (int k = 0; k < L; ++k) out32[k] = B[(key >> (62 - 2 * k)) & 3ull];
    out32[L] = 0;
    return L;
}

int pcl_group_destroy(pcl_group *g) {
    if (!g) return PCL_OK;
    cudaSetDevice(g->ctx->device);
    for (void *q : g->allocs) cudaFree(q);
    delete g;
    return PCL_OK;
}

// ------------------------------------------------------------------------------------------
// BGZF members inflated on the device (inflate.cuh)
// ------------------------------------------------------------------------------------------
#define INF_SLOTS 4
}  // extern "C"

struct pcl_inflater {
    int device = 0, sm_count = 1;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    uint8_t *d_comp = nullptr;
    uint64_t comp_cap = 0;
    InfMember *d_members = nullptr;
    uint64_t members_cap = 0;
    unsigned long long *d_err = nullptr, *h_err = nullptr;
    uint8_t *slot[INF_SLOTS] = {nullptr, nullptr, nullptr, nullptr};
    uint64_t slot_cap[INF_SLOTS] = {0, 0, 0, 0};
    double kernel_ms = 0;
    uint64_t out_bytes = 0, launches = 0;
    std::string err;
};

namespace {
int inf_fail(pcl_inflater *h, int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    h->err = buf;
    return code;
}
#define CUI(h, call)                                                                                \
    do {                                                                                            \
        cudaError_t e_ = (call);                                                                    \
        if (e_ != cudaSuccess) {                                                                    \
            cudaGetLastError();      /* the caller may fall back to the host inflate: leave no stale error for its thread */ \
            return inf_fail(h, e_ == cudaErrorMemoryAllocation ? PCL_ERR_NOMEM : PCL_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); \
        }                                                                                           \
    } while (0)

const char *inf_error_text(unsigned code) {
    static const char *const what[] = {"", "reserved block type", "bad stored block", "bad dynamic block header", "bad code lengths",
                                       "invalid code", "match before the start of the member", "more bytes than ISIZE",
                                       "compressed data ends early", "fewer bytes than ISIZE", "CRC-32 mismatch"};
    return code < sizeof what / sizeof what[0] ? what[code] : "error";
}

bool inf_range_ok(const pcl_inflater *h, int slot, uint64_t off, uint64_t n) {
    return slot >= 0 && slot < INF_SLOTS && h->slot[slot] && off <= h->slot_cap[slot] && n <= h->slot_cap[slot] - off;
}
}  // namespace

extern "C" {

int pcl_inflater_create(pcl_ctx *ctx, pcl_inflater **out) {
    if (!ctx || !out) return PCL_ERR_ARG;
    *out = nullptr;
    pcl_inflater *h = new pcl_inflater();
    h->device = ctx->device;
    h->sm_count = ctx->sm_count > 0 ? ctx->sm_count : 1;
    cudaError_t e = cudaSetDevice(h->device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreate(&h->ev0);
    if (e == cudaSuccess) e = cudaEventCreate(&h->ev1);
    if (e == cudaSuccess) e = cudaMalloc((void **)&h->d_err, 8);
    if (e == cudaSuccess) e = cudaMallocHost((void **)&h->h_err, 8);
    if (e != cudaSuccess) {
        const int rc = fail(ctx, PCL_ERR_CUDA, "pcl_inflater_create: %s", cudaGetErrorString(e));
        pcl_inflater_destroy(h);
        return rc;
    }
    *out = h;
    return PCL_OK;
}

void pcl_inflater_destroy(pcl_inflater *h) {
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    for (int k = 0; k < INF_SLOTS; ++k) cudaFree(h->slot[k]);
    cudaFree(h->d_comp); cudaFree(h->d_members); cudaFree(h->d_err);
    if (h->h_err) cudaFreeHost(h->h_err);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->stream) cudaStreamDestroy(h->stream);
    cudaGetLastError();
    delete h;
}

const char *pcl_inflater_error(const pcl_inflater *h) { return h ? h->err.c_str() : ""; }

int pcl_inflater_buffer(pcl_inflater *h, int slot, uint64_t bytes, void **device_ptr) {
    if (!h || slot < 0 || slot >= INF_SLOTS) return PCL_ERR_ARG;
    if (device_ptr) *device_ptr = nullptr;
    CUI(h, cudaSetDevice(h->device));
    if (bytes == 0 || bytes > h->slot_cap[slot]) {
        CUI(h, cudaStreamSynchronize(h->stream));
        cudaFree(h->slot[slot]);
        h->slot[slot] = nullptr;
        h->slot_cap[slot] = 0;
        if (bytes) {
            if (bytes >= (1ull << 32) - 4096) return inf_fail(h, PCL_ERR_ARG, "a text buffer must be smaller than 4 GiB");
            CUI(h, cudaMalloc((void **)&h->slot[slot], bytes + 16));
            h->slot_cap[slot] = bytes;
        }
    }
    if (device_ptr) *device_ptr = h->slot[slot];
    return PCL_OK;
}

int pcl_inflater_run(pcl_inflater *h, const uint8_t *comp, uint64_t comp_bytes, const pcl_bgzf_member *members, uint32_t n_members,
                     int slot, uint64_t dst_off) {
    if (!h || (!comp && comp_bytes) || (!members && n_members)) return PCL_ERR_ARG;
    if (n_members == 0) return PCL_OK;
    if (comp_bytes >= (1ull << 32) - 4096) return inf_fail(h, PCL_ERR_ARG, "a batch of compressed bytes must be smaller than 4 GiB");
    uint64_t out_end = 0;
    for (uint32_t i = 0; i < n_members; ++i) {
        const pcl_bgzf_member &m = members[i];
        if ((uint64_t)m.src_off + m.clen > comp_bytes) return inf_fail(h, PCL_ERR_ARG, "member %u leaves the compressed bytes", i);
        out_end = std::max<uint64_t>(out_end, (uint64_t)m.dst_off + m.isize);
    }
    if (!inf_range_ok(h, slot, dst_off, out_end)) return inf_fail(h, PCL_ERR_ARG, "the members do not fit text buffer %d", slot);
    CUI(h, cudaSetDevice(h->device));
    if (comp_bytes + INF_PAD_BYTES > h->comp_cap) {
        CUI(h, cudaStreamSynchronize(h->stream));
        cudaFree(h->d_comp);
        h->d_comp = nullptr; h->comp_cap = 0;
        const uint64_t cap = ((comp_bytes + comp_bytes / 4 + INF_PAD_BYTES + 4095) / 4096) * 4096;
        CUI(h, cudaMalloc((void **)&h->d_comp, cap));
        h->comp_cap = cap;
    }
    if (n_members > h->members_cap) {
        CUI(h, cudaStreamSynchronize(h->stream));
        cudaFree(h->d_members);
        h->d_members = nullptr; h->members_cap = 0;
        const uint64_t cap = (uint64_t)n_members + n_members / 4 + 64;
        CUI(h, cudaMalloc((void **)&h->d_members, cap * sizeof(InfMember)));
        h->members_cap = cap;
    }
    cudaStream_t s = h->stream;
    CUI(h, cudaMemcpyAsync(h->d_comp, comp, comp_bytes, cudaMemcpyHostToDevice, s));
    CUI(h, cudaMemsetAsync(h->d_comp + comp_bytes, 0, INF_PAD_BYTES, s));
    CUI(h, cudaMemcpyAsync(h->d_members, members, (uint64_t)n_members * sizeof(InfMember), cudaMemcpyHostToDevice, s));
    CUI(h, cudaMemsetAsync(h->d_err, 0xFF, 8, s));
    const uint32_t want = (n_members + INF_WARPS_PER_CTA - 1) / INF_WARPS_PER_CTA;
    const uint32_t grid = std::min<uint32_t>(want, (uint32_t)h->sm_count * 8u);
    CUI(h, cudaEventRecord(h->ev0, s));
    k_bgzf_inflate<<<grid, INF_WARPS_PER_CTA * 32u, 0, s>>>(h->d_comp, h->d_members, n_members, h->slot[slot] + dst_off, h->d_err);
    CUI(h, cudaGetLastError());
    CUI(h, cudaEventRecord(h->ev1, s));
    CUI(h, cudaMemcpyAsync(h->h_err, h->d_err, 8, cudaMemcpyDeviceToHost, s));
    CUI(h, cudaStreamSynchronize(s));
    float ms = 0;
    if (cudaEventElapsedTime(&ms, h->ev0, h->ev1) == cudaSuccess) h->kernel_ms += ms;
    h->launches += 1;
    h->out_bytes += out_end;
    if (*h->h_err != ~0ull) {
        const unsigned long long v = *h->h_err;
        return inf_fail(h, PCL_ERR_INPUT, "corrupt BGZF block (member %llu of the batch): %s", v >> 8, inf_error_text((unsigned)(v & 0xFF)));
    }
    return PCL_OK;
}

int pcl_inflater_move(pcl_inflater *h, int src_slot, uint64_t src_off, int dst_slot, uint64_t dst_off, uint64_t n_bytes) {
    if (!h) return PCL_ERR_ARG;
    if (!inf_range_ok(h, src_slot, src_off, n_bytes) || !inf_range_ok(h, dst_slot, dst_off, n_bytes)) return inf_fail(h, PCL_ERR_ARG, "pcl_inflater_move: bad range");
    if (src_slot == dst_slot && src_off < dst_off + n_bytes && dst_off < src_off + n_bytes && n_bytes)
        return inf_fail(h, PCL_ERR_ARG, "pcl_inflater_move: the ranges overlap");
    if (n_bytes == 0) return PCL_OK;
    CUI(h, cudaSetDevice(h->device));
    CUI(h, cudaMemcpyAsync(h->slot[dst_slot] + dst_off, h->slot[src_slot] + src_off, n_bytes, cudaMemcpyDeviceToDevice, h->stream));
    CUI(h, cudaStreamSynchronize(h->stream));
    return PCL_OK;
}

int pcl_inflater_fetch(pcl_inflater *h, int slot, uint64_t off, uint64_t n_bytes, void *host_dst) {
    if (!h || (!host_dst && n_bytes)) return PCL_ERR_ARG;
    if (!inf_range_ok(h, slot, off, n_bytes)) return inf_fail(h, PCL_ERR_ARG, "pcl_inflater_fetch: bad range");
    if (n_bytes == 0) return PCL_OK;
    CUI(h, cudaSetDevice(h->device));
    CUI(h, cudaMemcpyAsync(host_dst, h->slot[slot] + off, n_bytes, cudaMemcpyDeviceToHost, h->stream));
    CUI(h, cudaStreamSynchronize(h->stream));
    return PCL_OK;
}

int pcl_inflater_poke(pcl_inflater *h, int slot, uint64_t off, const void *host_src, uint64_t n_bytes) {
    if (!h || (!host_src && n_bytes)) return PCL_ERR_ARG;
    if (!inf_range_ok(h, slot, off, n_bytes)) return inf_fail(h, PCL_ERR_ARG, "pcl_inflater_poke: bad range");
    if (n_bytes == 0) return PCL_OK;
    CUI(h, cudaSetDevice(h->device));
    CUI(h, cudaMemcpyAsync(h->slot[slot] + off, host_src, n_bytes, cudaMemcpyHostToDevice, h->stream));
    CUI(h, cudaStreamSynchronize(h->stream));
    return PCL_OK;
}

int pcl_inflater_stats(const pcl_inflater *h, double *kernel_ms, uint64_t *out_bytes, uint64_t *launches) {
    if (!h) return PCL_ERR_ARG;
    if (kernel_ms) *kernel_ms = h->kernel_ms;
    if (out_bytes) *out_bytes = h->out_bytes;
    if (launches) *launches = h->launches;
    return PCL_OK;
}

void *pcl_host_alloc(uint64_t bytes) {
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 16, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
void pcl_host_free(void *p) { if (p) cudaFreeHost(p); }

// ------------------------------------------------------------------------------------------
// instrumentation
// ------------------------------------------------------------------------------------------
int pcl_profile_enable(pcl_ctx *ctx, int on) {
    if (!ctx) return PCL_ERR_ARG;
    ctx->profile = on != 0;
    return PCL_OK;
}

static int prof_drain(pcl_ctx *ctx) {
    if (int r = pcl_sync(ctx)) return r;
    for (EvPair &p : ctx->evs) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) { ctx->prof_ms[p.k] += ms; ctx->prof_n[p.k] += 1; }
        cudaEventDestroy(p.a);
        cudaEventDestroy(p.b);
    }
    ctx->evs.clear();
    return PCL_OK;
}

int pcl_profile_get(pcl_ctx *ctx, int k, double *ms, uint64_t *launches) {
    if (!ctx || k < 0 || k >= PCL_K_MAX) return PCL_ERR_ARG;
    if (int r = prof_drain(ctx)) return r;
    if (ms) *ms = ctx->prof_ms[k];
    if (launches) *launches = ctx->prof_n[k];
    return PCL_OK;
}

int pcl_profile_reset(pcl_ctx *ctx) {
    if (!ctx) return PCL_ERR_ARG;
    if (int r = prof_drain(ctx)) return r;
    for (int k = 0; k < PCL_K_MAX; ++k) { ctx->prof_ms[k] = 0; ctx->prof_n[k] = 0; }
    return PCL_OK;
}

uint64_t pcl_launch_count(const pcl_ctx *ctx) { return ctx ? ctx->launches : 0; }

int pcl_timer_start(pcl_chunk *c) {
    if (!c) return PCL_ERR_ARG;
    if (int r = bind(c->ctx)) return r;
    CU(c->ctx, cudaEventRecord(c->t0, c->stream));
    return PCL_OK;
}

int pcl_timer_stop_ms(pcl_chunk *c, double *ms) {
    if (!c || !ms) return PCL_ERR_ARG;
    if (int r = bind(c->ctx)) return r;
    CU(c->ctx, cudaEventRecord(c->t1, c->stream));
    CU(c->ctx, cudaEventSynchronize(c->t1));
    float f = 0;
    CU(c->ctx, cudaEventElapsedTime(&f, c->t0, c->t1));
    *ms = f;
    return PCL_OK;
}

}  // extern "C"
