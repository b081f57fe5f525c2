// textparse.cuh -- device-side FASTQ record parse (included once, by pcl.cu).
//
// the host hands over raw, uncompressed FASTQ text; the device finds the line ends and
// fills the SoA planes of a chunk directly in HBM, so no host core touches individual records in pass 1.
// Replaces, for this path (paths relative to the precellar repository):
//   FastqReader::read_record            seqspec/src/read.rs:137-155 (four lines per record: '@' definition,
//                                        sequence, '+' line, quality; noodles-fastq dialect, \r\n tolerated)
//   BatchedFqReader::next name check    precellar/src/align/fastq.rs:417-431 (names agree after strip_fq_suffix,
//                                        fastq.rs:612-621)
//
//   k_text_count   newline count of every 16 KB tile (64 contiguous bytes per thread)
//   k_text_scan    exclusive scan of the tile counts (one CTA)
//   k_text_index   byte offset of every newline, in order -> line index
//   k_text_fill    line index -> seq / qual / len planes + name hash (+ record validation)
//   k_names_check  first group whose read names differ between files
#ifndef PCL_TEXTPARSE_CUH
#define PCL_TEXTPARSE_CUH

#include <cuda_runtime.h>

#include <cstdint>

#define PCL_TEXT_BLOCK 256
#define PCL_TEXT_PER_THREAD 64                    // contiguous text bytes per thread: four 16-byte loads
#define PCL_TEXT_TILE (PCL_TEXT_BLOCK * PCL_TEXT_PER_THREAD)       // bytes per CTA iteration (16 KB)

// 16 bytes -> 16-bit mask of the positions holding '\n' (bit k = byte k)
__device__ __forceinline__ uint32_t pcl_nl_mask16(const uint4 v) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    uint32_t m = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const uint32_t x = w[k] ^ 0x0A0A0A0Au;
        const uint32_t z = ~(((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x) & 0x80808080u;     // 0x80 where the byte was '\n'
        m |= ((((z >> 7) * 0x01020408u) >> 24) & 0xFu) << (4 * k);
    }
    return m;
}

// newline masks of the thread's 64 bytes (bit k of m[q] = byte 16q + k); the thread's bytes are contiguous, so the
// newlines of a tile are ordered by (thread, q, bit) and one block scan per tile orders them all
__device__ __forceinline__ void pcl_tile_masks(const uint8_t *text, uint64_t n_bytes, uint64_t tile, uint32_t tid, uint32_t m[4]) {
    const uint64_t pos = tile * PCL_TEXT_TILE + (uint64_t)tid * PCL_TEXT_PER_THREAD;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const uint64_t p = pos + 16u * q;
        m[q] = 0u;
        if (p < n_bytes) {
            const uint4 v = __ldg(reinterpret_cast<const uint4 *>(text + p));          // the buffer is padded to 16 bytes with zeros
            m[q] = pcl_nl_mask16(v);
            const uint64_t left = n_bytes - p;
            if (left < 16u) m[q] &= (1u << left) - 1u;
        }
    }
}

__global__ void __launch_bounds__(PCL_TEXT_BLOCK) k_text_count(const uint8_t *__restrict__ text, uint64_t n_bytes,
                                                               uint64_t n_tiles, uint32_t *__restrict__ tile_cnt) {
    __shared__ uint32_t wsum[PCL_TEXT_BLOCK / 32];
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        uint32_t m[4];
        pcl_tile_masks(text, n_bytes, tile, tid, m);
        uint32_t c = __popc(m[0]) + __popc(m[1]) + __popc(m[2]) + __popc(m[3]);
        c = __reduce_add_sync(0xFFFFFFFFu, c);
        if (lane == 0) wsum[warp] = c;
        __syncthreads();
        if (tid == 0) {
            uint32_t s = 0;
#pragma unroll
            for (int k = 0; k < PCL_TEXT_BLOCK / 32; ++k) s += wsum[k];
            tile_cnt[tile] = s;
        }
        __syncthreads();
    }
}

// one CTA of 1024 threads; tile_off[t] = sum of tile_cnt[0..t), *total = number of newlines
__global__ void __launch_bounds__(1024) k_text_scan(const uint32_t *__restrict__ tile_cnt, uint64_t n_tiles,
                                                    uint32_t *__restrict__ tile_off, unsigned long long *__restrict__ total) {
    __shared__ uint32_t part[1024];
    const uint32_t tid = threadIdx.x;
    const uint64_t per = (n_tiles + 1023u) / 1024u;
    const uint64_t lo = (uint64_t)tid * per, hi = lo + per < n_tiles ? lo + per : n_tiles;
    uint32_t s = 0;
    for (uint64_t t = lo; t < hi; ++t) s += tile_cnt[t];
    part[tid] = s;
    __syncthreads();
    for (uint32_t d = 1; d < 1024u; d <<= 1) {                // Hillis-Steele inclusive scan of the 1024 partial sums
        const uint32_t v = tid >= d ? part[tid - d] : 0u;
        __syncthreads();
        part[tid] += v;
        __syncthreads();
    }
    uint32_t run = tid ? part[tid - 1] : 0u;
    for (uint64_t t = lo; t < hi; ++t) { const uint32_t c = tile_cnt[t]; tile_off[t] = run; run += c; }
    if (tid == 1023u) *total = part[1023];
}

__global__ void __launch_bounds__(PCL_TEXT_BLOCK) k_text_index(const uint8_t *__restrict__ text, uint64_t n_bytes,
                                                               uint64_t n_tiles, const uint32_t *__restrict__ tile_off,
                                                               uint32_t *__restrict__ nl, uint64_t nl_cap) {
    __shared__ uint32_t wsum[PCL_TEXT_BLOCK / 32];
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        uint32_t m[4];
        pcl_tile_masks(text, n_bytes, tile, tid, m);
        const uint32_t c = __popc(m[0]) + __popc(m[1]) + __popc(m[2]) + __popc(m[3]);
        uint32_t incl = c;                                   // inclusive warp scan
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if ((int)lane >= d) incl += v;
        }
        if (lane == 31u) wsum[warp] = incl;
        __syncthreads();
        uint32_t base = tile_off[tile];
        for (uint32_t k = 0; k < warp; ++k) base += wsum[k];
        uint64_t idx = (uint64_t)base + incl - c;
        const uint32_t pos = (uint32_t)(tile * PCL_TEXT_TILE + (uint64_t)tid * PCL_TEXT_PER_THREAD);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            uint32_t mq = m[q];
            while (mq) {
                const uint32_t j = (uint32_t)__ffs((int)mq) - 1u;
                mq &= mq - 1u;
                if (idx < nl_cap) nl[idx] = pos + 16u * q + j;
                ++idx;
            }
        }
        __syncthreads();
    }
}

struct TextFillParams {
    const uint8_t *text;
    const uint32_t *nl;          // byte offset of the end of every line
    uint64_t n;                  // records to extract: lines [0, 4n)
    uint32_t stride, qoff, qstride;
    uint8_t *seq, *qual;         // may be NULL (lengths only / quality only)
    uint16_t *len;
    uint64_t *name_hash;
    unsigned long long *err;     // min over bad records of (record << 8 | code), ~0 if none
};

#define PCL_TEXT_ERR_AT 1        // definition line does not start with '@'
#define PCL_TEXT_ERR_PLUS 2      // third line does not start with '+'
#define PCL_TEXT_ERR_LEN 3       // sequence and quality lengths differ

__global__ void __launch_bounds__(PCL_TEXT_BLOCK) k_text_fill(const __grid_constant__ TextFillParams p) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t n_warp_iters = (p.n + 31u) / 32u;
    const uint64_t warp0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t wi = warp0; wi < n_warp_iters; wi += n_warps) {
        const uint64_t i = wi * 32u + lane;
        uint32_t s1 = 0, L = 0, s3 = 0, Lq = 0;
        if (i < p.n) {
            const uint4 e = __ldg(reinterpret_cast<const uint4 *>(p.nl) + i);     // ends of the record's four lines
            const uint32_t s0 = i ? __ldg(p.nl + 4u * i - 1u) + 1u : 0u;
            uint32_t e0 = e.x, e1 = e.y, e3 = e.w;
            s1 = e.x + 1u;
            const uint32_t s2 = e.y + 1u;
            s3 = e.z + 1u;
            if (e0 > s0 && p.text[e0 - 1u] == '\r') --e0;
            if (e1 > s1 && p.text[e1 - 1u] == '\r') --e1;
            if (e3 > s3 && p.text[e3 - 1u] == '\r') --e3;
            L = e1 - s1;
            Lq = e3 - s3;
            uint32_t code = 0;
            if (e0 == s0 || p.text[s0] != '@') code = PCL_TEXT_ERR_AT;
            else if (e.z == s2 || p.text[s2] != '+') code = PCL_TEXT_ERR_PLUS;
            else if (L != Lq) code = PCL_TEXT_ERR_LEN;
            if (code) atomicMin(p.err, ((unsigned long long)i << 8) | code);
            p.len[i] = (uint16_t)(L > 65535u ? 65535u : L);
            if (p.name_hash) {
                // FNV-1a of the name: the definition up to the first blank, without a trailing /1 or /2 (strip_fq_suffix)
                uint32_t b = s0 + 1u, q = b;
                while (q < e0) { const uint8_t ch = p.text[q]; if (ch == ' ' || ch == '\t') break; ++q; }
                if (q - b > 2u && p.text[q - 2u] == '/' && (p.text[q - 1u] == '1' || p.text[q - 1u] == '2')) q -= 2u;
                uint64_t h = 1469598103934665603ull;
                for (uint32_t k = b; k < q; ++k) { h ^= p.text[k]; h *= 1099511628211ull; }
                p.name_hash[i] = h;
            }
        }
        if (!p.seq && !p.qual) continue;
        // the warp copies its 32 records one after the other: coalesced loads from the text, full-sector stores
        const uint64_t base = wi * 32u;
        const uint32_t cnt = p.n - base < 32u ? (uint32_t)(p.n - base) : 32u;
        for (uint32_t k = 0; k < cnt; ++k) {
            const uint32_t ks1 = __shfl_sync(0xFFFFFFFFu, s1, k), kL = __shfl_sync(0xFFFFFFFFu, L, k);
            const uint32_t ks3 = __shfl_sync(0xFFFFFFFFu, s3, k), kLq = __shfl_sync(0xFFFFFFFFu, Lq, k);
            if (p.seq) {
                uint8_t *dst = p.seq + (base + k) * (uint64_t)p.stride;
                for (uint32_t o = lane; o < p.stride; o += 32u) dst[o] = o < kL ? p.text[ks1 + o] : (uint8_t)'N';
            }
            if (p.qual) {
                uint8_t *dst = p.qual + (base + k) * (uint64_t)p.qstride;
                const uint32_t avail = kLq > p.qoff ? kLq - p.qoff : 0u;
                for (uint32_t o = lane; o < p.qstride; o += 32u) dst[o] = o < avail ? p.text[ks3 + p.qoff + o] : (uint8_t)'!';
            }
        }
    }
}

struct NamesCheckParams {
    const uint64_t *hash[PCL_MAX_READS];
    int n_reads;
    uint64_t n;
    unsigned long long *bad;     // min index of a group whose names differ, ~0 if none
};

__global__ void __launch_bounds__(PCL_TEXT_BLOCK) k_names_check(const __grid_constant__ NamesCheckParams p) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t h0 = p.hash[0][i];
        bool bad = false;
        for (int r = 1; r < p.n_reads; ++r) bad |= p.hash[r][i] != h0;
        if (bad) atomicMin(p.bad, (unsigned long long)i);
    }
}

#endif  // PCL_TEXTPARSE_CUH
