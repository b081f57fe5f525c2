// emit.cuh -- device side of the make_fastq writer (included once, by pcl.cu).
//
// the records of R1.fq.zst / R2.fq.zst are assembled in HBM from the FASTQ text the device
// parsed (pcl_text_keep) and the result planes, and optionally entropy-coded into zstd blocks there, so that the host
// only appends finished bytes to the output files.  Per-group logic: emit_core.h; block coder: zhuff_core.h.
//
//   k_emit_plan     thread per read group: drop / error decision, bytes of its R1 and R2 records
//   (device_scan)   byte offset of every record
//   k_emit_fill     warp per read group: the record bytes, lanes striding over each copied piece
//   k_zh_blocks     CTA per 128 KB of output: matches of every line against the same line of the previous record ->
//                   literal bytes -> histogram -> length-limited Huffman code -> four bit streams; the matches as
//                   FSE-coded sequences -> one zstd block in the block's slot
//   k_zh_compact    CTA per block: slots -> one contiguous stream of blocks behind the room for the frame header
#ifndef PCL_EMIT_CUH
#define PCL_EMIT_CUH

#include <cuda_runtime.h>

#include "emit_core.h"
#include "zhuff_core.h"

#define PCL_EMIT_BLOCK 256

struct EmitKernelParams {
    EmitParams p;
    uint32_t *sz[2];                 // bytes of each group's R1 / R2 record (0: not written)
    const uint32_t *off[2];          // exclusive scan of sz
    uint8_t *out[2];
    unsigned long long *err;         // min over failing groups of (group << 8 | PCL_EMIT_*), ~0 if none
    unsigned long long *n_written;
};

__global__ void __launch_bounds__(PCL_EMIT_BLOCK) k_emit_plan(const __grid_constant__ EmitKernelParams k) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t first = (uint64_t)blockIdx.x * blockDim.x + (threadIdx.x - lane), stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t base = first; base < k.p.n; base += stride) {          // whole warps iterate together (the ballot below)
        const uint64_t i = base + lane;
        uint32_t s1 = 0, s2 = 0;
        if (i < k.p.n) {
            const EmitDecision d = pcl_emit_decide(k.p, i);
            if (d.err) atomicMin(k.err, ((unsigned long long)i << 8) | d.err);
            else if (d.write) {
                EmitCount c1;
                pcl_emit_r1(k.p, i, d, c1);
                s1 = c1.n;
                if (k.p.paired) {
                    EmitCount c2;
                    pcl_emit_r2(k.p, i, d, c2);
                    s2 = c2.n;
                }
            }
            k.sz[0][i] = s1;
            if (k.p.paired) k.sz[1][i] = s2;
        }
        const uint32_t wrote = __ballot_sync(0xFFFFFFFFu, s1 != 0);
        if (lane == 0 && wrote) atomicAdd(k.n_written, (unsigned long long)__popc(wrote));
    }
}

// byte sink of a warp: every lane sees the same pieces (the decision is warp-uniform) and copies its share of each
struct EmitWarp {
    uint8_t *dst;
    uint32_t lane;
    __device__ __forceinline__ void put(uint8_t c) { if (lane == 0) *dst = c; ++dst; }
    __device__ __forceinline__ void copy(const uint8_t *src, uint32_t len) {
        for (uint32_t o = lane; o < len; o += 32u) dst[o] = src[o];
        dst += len;
    }
    __device__ __forceinline__ void rcopy(const uint8_t *src, uint32_t len, bool compl_) {
        for (uint32_t o = lane; o < len; o += 32u) { const uint8_t x = src[len - 1u - o]; dst[o] = compl_ ? pcl_emit_complement(x) : x; }
        dst += len;
    }
    __device__ __forceinline__ void bases(uint64_t key, uint32_t L) {
        if (lane < L) dst[lane] = (uint8_t)"ACGT"[(key >> (2u * (L - 1u - lane))) & 3ull];
        dst += L;
    }
};

__global__ void __launch_bounds__(PCL_EMIT_BLOCK) k_emit_fill(const __grid_constant__ EmitKernelParams k) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t warp0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t i = warp0; i < k.p.n; i += n_warps) {
        if (k.sz[0][i] == 0) continue;                        // dropped group (a written record has at least six bytes)
        const EmitDecision d = pcl_emit_decide(k.p, i);
        EmitWarp w1{k.out[0] + k.off[0][i], lane};
        pcl_emit_r1(k.p, i, d, w1);
        if (k.p.paired) {
            EmitWarp w2{k.out[1] + k.off[1][i], lane};
            pcl_emit_r2(k.p, i, d, w2);
        }
    }
}

struct ZhJob {
    const uint8_t *src;
    uint64_t n;                      // bytes
    uint8_t *slots;                  // n_blocks * ZH_SLOT_BYTES
    uint32_t *sizes;                 // bytes used in each slot
    uint8_t *stream;                 // contiguous blocks (k_zh_compact)
    unsigned long long *total;       // bytes of the stream
    ZhScratch *scratch;              // one per block (match finder, literal buffer, sequence bits); NULL: literals only
    uint32_t n_blocks;
};
struct ZhParams { ZhJob job[2]; };

struct ZhDevExec {
    template <class F> __device__ __forceinline__ void each(F f) { f(threadIdx.x); __syncthreads(); }
};

// dynamic shared memory: sizeof(ZhShared) (about 72 KB: bit buffer, per-warp histograms, code table, per-sequence codes)
__global__ void __launch_bounds__(ZH_THREADS) k_zh_blocks(const __grid_constant__ ZhParams p) {
    extern __shared__ __align__(16) unsigned char zh_smem[];
    ZhShared &sh = *reinterpret_cast<ZhShared *>(zh_smem);
    uint32_t b = blockIdx.x;
    const int f = b >= p.job[0].n_blocks ? 1 : 0;
    if (f) b -= p.job[0].n_blocks;
    const ZhJob &j = p.job[f];
    const uint64_t lo = (uint64_t)b * ZH_BLOCK_BYTES;
    const uint32_t len = (uint32_t)(j.n - lo < ZH_BLOCK_BYTES ? j.n - lo : ZH_BLOCK_BYTES);
    ZhDevExec ex;
    const uint32_t used = zh_compress_block(ex, sh, j.scratch ? j.scratch + b : nullptr, j.src + lo, len, b + 1u == j.n_blocks ? 1u : 0u,
                                            j.slots + (uint64_t)b * ZH_SLOT_BYTES);
    if (threadIdx.x == 0) j.sizes[b] = used;
}

__global__ void __launch_bounds__(ZH_THREADS) k_zh_compact(const __grid_constant__ ZhParams p) {
    __shared__ unsigned long long part[ZH_THREADS / 32];
    __shared__ unsigned long long s_off;
    uint32_t b = blockIdx.x;
    const int f = b >= p.job[0].n_blocks ? 1 : 0;
    if (f) b -= p.job[0].n_blocks;
    const ZhJob &j = p.job[f];
    unsigned long long sum = 0;
    for (uint32_t k = threadIdx.x; k < b; k += ZH_THREADS) sum += j.sizes[k];
    for (int d = 16; d; d >>= 1) sum += __shfl_xor_sync(0xFFFFFFFFu, sum, d);
    if ((threadIdx.x & 31u) == 0) part[threadIdx.x >> 5] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (uint32_t k = 0; k < ZH_THREADS / 32; ++k) t += part[k];
        s_off = t;
        if (b + 1u == j.n_blocks) *j.total = t + j.sizes[b];
    }
    __syncthreads();
    const uint8_t *src = j.slots + (uint64_t)b * ZH_SLOT_BYTES;
    uint8_t *dst = j.stream + s_off;
    const uint32_t n = j.sizes[b];
    for (uint32_t k = threadIdx.x; k < n; k += ZH_THREADS) dst[k] = src[k];
}

#endif  // PCL_EMIT_CUH
