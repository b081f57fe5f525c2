// group.cuh -- per-barcode grouping on the device (included once, by pcl.cu).
//
// The first data-parallel step on the far side of the aligner is keyed on this path's output:
//     sort_by_barcode (stable sort on the corrected barcode string)        precellar/src/fragment/mod.rs:434-455
//     chunk_by(barcode)                                                    precellar/src/fragment/mod.rs:359-363
//     RemoveDuplicates::remove_duplicates: one Fragment per FingerPrint
//         (alignment coordinates + orientation + UMI), count = duplicates  precellar/src/fragment/dedups.rs:52-66,320-340
// Here: a stable LSD radix sort of the chunk's read groups on (barcode, fingerprint, UMI), the barcode boundaries, and
// the runs of equal (barcode, fingerprint, UMI) -- the duplicate sets; the first record of a run (input order) is the
// one the reference keeps (`or_insert(first)`), the run length is Fragment.count.  The alignment part of the fingerprint
// (reference id, 5' coordinates, orientation) comes from the aligner, which is out of scope: the caller hands it over as
// one 64-bit word per read group (or nothing: barcode + UMI only).
//
// Keys.  Barcode: the corrected keys of all barcode segments in join order (segments in read order, files in layout
// order -- AnnotatedFastq::join, align/fastq.rs:571-603), 2 bits per base, left-aligned in bits 63..6 with the number of
// bases in bits 5..0: numeric order == the reference's byte-string order (A < C < G < T, a proper prefix first).
// UMI: the raw bytes of the UMI segment, 3 bits per base (A C G T N -> 0..4; any other byte makes the read group
// "ungroupable": it is sorted behind the assigned ones and reported, never merged), left-aligned the same way.
#ifndef PCL_GROUP_CUH
#define PCL_GROUP_CUH

#include <cuda_runtime.h>
#include <stdint.h>

#define PCL_SORT_BLOCK 256
#define PCL_SORT_ROUNDS 8
#define PCL_SORT_TILE (PCL_SORT_BLOCK * PCL_SORT_ROUNDS)      // items per CTA; every warp owns 32 * ROUNDS consecutive ones
#define PCL_GROUP_NOKEY 0xFFFFFFFFFFFFFFFFull                // barcode key of a read group that is not assigned

// ------------------------------------------------------------------------------------------------
// exclusive scan of n uint32 values (in place allowed): tile sums -> one-CTA scan of the sums -> apply
// ------------------------------------------------------------------------------------------------
#define PCL_SCAN_TILE 4096        // 256 threads x 16 values

__device__ __forceinline__ uint32_t pcl_warp_incl_scan(uint32_t v, uint32_t lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, v, o);
        if (lane >= (uint32_t)o) v += t;
    }
    return v;
}

// exclusive scan across the 256 threads of a CTA; *total = sum of all (valid after the call in every thread)
__device__ __forceinline__ uint32_t pcl_block_excl_scan(uint32_t v, uint32_t *total) {
    __shared__ uint32_t wsum[8];
    const uint32_t lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
    const uint32_t inc = pcl_warp_incl_scan(v, lane);
    if (lane == 31u) wsum[w] = inc;
    __syncthreads();
    uint32_t base = 0, tot = 0;
#pragma unroll
    for (uint32_t k = 0; k < 8u; ++k) { const uint32_t s = wsum[k]; if (k < w) base += s; tot += s; }
    __syncthreads();
    *total = tot;
    return base + inc - v;
}

__global__ void __launch_bounds__(256) k_scan_tile_sums(const uint32_t *__restrict__ in, uint64_t n, uint32_t *__restrict__ sums) {
    const uint64_t base = (uint64_t)blockIdx.x * PCL_SCAN_TILE;
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const uint64_t i = base + (uint64_t)k * 256 + threadIdx.x;
        if (i < n) s += in[i];
    }
    uint32_t tot;
    (void)pcl_block_excl_scan(s, &tot);
    if (threadIdx.x == 0) sums[blockIdx.x] = tot;
}

// one CTA: exclusive scan of the tile sums, in place; *grand = total
__global__ void __launch_bounds__(256) k_scan_sums(uint32_t *sums, uint64_t n_tiles, unsigned long long *grand) {
    unsigned long long carry = 0;
    for (uint64_t b = 0; b < n_tiles; b += 256) {
        const uint64_t i = b + threadIdx.x;
        const uint32_t v = i < n_tiles ? sums[i] : 0u;
        uint32_t tot;
        const uint32_t ex = pcl_block_excl_scan(v, &tot);
        if (i < n_tiles) sums[i] = (uint32_t)(carry + ex);
        carry += tot;
    }
    if (threadIdx.x == 0 && grand) *grand = carry;
}

__global__ void __launch_bounds__(256) k_scan_apply(const uint32_t *__restrict__ in, uint32_t *__restrict__ out, uint64_t n,
                                                   const uint32_t *__restrict__ sums) {
    const uint64_t base = (uint64_t)blockIdx.x * PCL_SCAN_TILE + (uint64_t)threadIdx.x * 16;     // 16 consecutive values per thread
    uint32_t v[16], s = 0;
#pragma unroll
    for (int k = 0; k < 16; ++k) { v[k] = base + k < n ? in[base + k] : 0u; s += v[k]; }
    uint32_t tot;
    uint32_t run = sums[blockIdx.x] + pcl_block_excl_scan(s, &tot);
#pragma unroll
    for (int k = 0; k < 16; ++k) { if (base + k < n) out[base + k] = run; run += v[k]; }
}

// ------------------------------------------------------------------------------------------------
// stable LSD radix sort of (u64 key, u32 value) pairs, one 8-bit digit per pass
// ------------------------------------------------------------------------------------------------
// hist[digit * n_ctas + cta]: after the exclusive scan, the global position of the CTA's first item with that digit
__global__ void __launch_bounds__(PCL_SORT_BLOCK) k_radix_hist(const unsigned long long *__restrict__ keys, uint64_t n, uint32_t shift,
                                                               uint32_t *__restrict__ hist, uint32_t n_ctas) {
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const uint64_t base = (uint64_t)blockIdx.x * PCL_SORT_TILE;
#pragma unroll
    for (int r = 0; r < PCL_SORT_ROUNDS; ++r) {
        const uint64_t i = base + (uint64_t)r * PCL_SORT_BLOCK + threadIdx.x;
        if (i < n) atomicAdd(&h[(uint32_t)(keys[i] >> shift) & 0xFFu], 1u);
    }
    __syncthreads();
    hist[(uint64_t)threadIdx.x * n_ctas + blockIdx.x] = h[threadIdx.x];
}

__global__ void __launch_bounds__(PCL_SORT_BLOCK) k_radix_scatter(const unsigned long long *__restrict__ keys_in, const uint32_t *__restrict__ vals_in,
                                                                  unsigned long long *__restrict__ keys_out, uint32_t *__restrict__ vals_out, uint64_t n,
                                                                  uint32_t shift, const uint32_t *__restrict__ offs, uint32_t n_ctas) {
    __shared__ uint32_t wcount[PCL_SORT_BLOCK / 32][256];      // items of each digit in each warp's sub-tile, then their global base
    const uint32_t lane = threadIdx.x & 31u, w = threadIdx.x >> 5, lt = (1u << lane) - 1u;
    for (uint32_t k = threadIdx.x; k < (PCL_SORT_BLOCK / 32) * 256; k += PCL_SORT_BLOCK) (&wcount[0][0])[k] = 0;
    __syncthreads();
    // warp w owns items [base + w * 32 * ROUNDS, + 32 * ROUNDS) in order: round r, lane l -> item r * 32 + l
    const uint64_t wbase = (uint64_t)blockIdx.x * PCL_SORT_TILE + (uint64_t)w * 32 * PCL_SORT_ROUNDS;
    unsigned long long key[PCL_SORT_ROUNDS];
    uint32_t val[PCL_SORT_ROUNDS], rank[PCL_SORT_ROUNDS];
#pragma unroll
    for (int r = 0; r < PCL_SORT_ROUNDS; ++r) {
        const uint64_t i = wbase + (uint64_t)r * 32 + lane;
        const bool ok = i < n;
        key[r] = ok ? keys_in[i] : 0ull;
        val[r] = ok ? vals_in[i] : 0u;
        const uint32_t act = __ballot_sync(0xFFFFFFFFu, ok);
        rank[r] = 0;
        if (ok) {
            const uint32_t d = (uint32_t)(key[r] >> shift) & 0xFFu;
            const uint32_t peers = __match_any_sync(act, d);            // lanes of this round with the same digit
            rank[r] = wcount[w][d] + __popc(peers & lt);                // earlier rounds of this warp + earlier lanes
            __syncwarp(act);
            if ((peers & lt) == 0u) wcount[w][d] += __popc(peers);      // the first lane of every digit group
        }
        __syncwarp();
    }
    __syncthreads();
    {   // digit d = threadIdx.x: global base of every warp's run of that digit
        uint32_t run = offs[(uint64_t)threadIdx.x * n_ctas + blockIdx.x];
#pragma unroll
        for (uint32_t k = 0; k < PCL_SORT_BLOCK / 32; ++k) { const uint32_t c = wcount[k][threadIdx.x]; wcount[k][threadIdx.x] = run; run += c; }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < PCL_SORT_ROUNDS; ++r) {
        const uint64_t i = wbase + (uint64_t)r * 32 + lane;
        if (i < n) {
            const uint32_t d = (uint32_t)(key[r] >> shift) & 0xFFu;
            const uint32_t dst = wcount[w][d] + rank[r];
            keys_out[dst] = key[r];
            vals_out[dst] = val[r];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// grouping kernels
// ------------------------------------------------------------------------------------------------
#define PCL_GROUP_MAX_SEG 16

struct GroupSeg {
    const void *key;              // bc_key plane of the segment
    const uint16_t *fstatus;      // fstatus plane of its read file
    uint8_t key_bytes, j, fixed_len, pad;     // fixed_len: bases of a fixed-length element, 0 = read the marker
};

struct GroupKeyParams {
    GroupSeg seg[PCL_GROUP_MAX_SEG];
    uint32_t n_seg;
    uint32_t perm_base;           // record i of this chunk is read group perm_base + i of the grouping (several chunks: pcl_group_build_multi)
    // UMI: the raw bytes of one segment of one read file
    const uint8_t *useq;          // NULL: no UMI
    const uint16_t *ulen;
    const uint16_t *ufstatus;
    const uint32_t *ucoord;       // variable layouts; NULL: static
    uint32_t ustride, umax_len, ustart, ubases, umin;    // umin < ubases: a trailing variable element, clipped at the record end
    uint64_t n;
    unsigned long long *bc_key;   // out [n]
    unsigned long long *umi_key;  // out [n]
    uint32_t *perm;               // out [n]: identity
    unsigned long long *stats;    // [0] assigned, [1] ungroupable (UMI byte outside ACGTN), [2] OR of bc keys ^ first, [3] OR of umi keys
};

// left-aligned packing: `bits` of payload, `bases` in the low 6 bits
__device__ __forceinline__ unsigned long long pcl_left_align(unsigned long long payload, uint32_t bits, uint32_t bases) {
    return (bits ? payload << (64u - bits) : 0ull) | (unsigned long long)bases;
}

__global__ void __launch_bounds__(256) k_group_keys(const __grid_constant__ GroupKeyParams p) {
    unsigned long long n_assigned = 0, n_bad = 0, or_bc = 0, or_umi = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += (uint64_t)gridDim.x * blockDim.x) {
        unsigned long long bits = 0;
        uint32_t bases = 0, n_present = 0;
        bool ok = true;
        for (uint32_t s = 0; s < p.n_seg; ++s) {
            const GroupSeg &g = p.seg[s];
            const uint32_t fs = g.fstatus[i];
            if ((fs & 3u) != PCL_SPLIT_OK) continue;                                    // a defect file contributes nothing
            const uint32_t code = PCL_BC_CODE(fs, g.j);
            if (code == PCL_BC_ABSENT) continue;
            ++n_present;
            if (code != PCL_BC_EXACT && code != PCL_BC_CORRECTED) { ok = false; continue; }      // no CB made of whitelist entries
            unsigned long long k = g.key_bytes == 4 ? (unsigned long long)reinterpret_cast<const uint32_t *>(g.key)[i]
                                                    : reinterpret_cast<const unsigned long long *>(g.key)[i];
            uint32_t L = g.fixed_len;
            if (g.key_bytes == 4 && L == 16u) { /* marker dropped */ }
            else {
                if (k == 0ull) { ok = false; continue; }
                L = (63u - (uint32_t)__clzll((long long)k)) >> 1;                       // marker form: (1 << 2L) | bases
                k &= (1ull << (2u * L)) - 1ull;
            }
            if (bases + L > 29u) { ok = false; continue; }                               // 58 payload bits
            bits = (bits << (2u * L)) | k;
            bases += L;
        }
        ok = ok && n_present > 0;
        unsigned long long bk = ok ? pcl_left_align(bits, 2u * bases, bases) : PCL_GROUP_NOKEY;
        unsigned long long uk = 0;
        if (ok && p.useq) {
            uint32_t start = p.ustart, len = p.ubases;
            bool present;
            if (p.ucoord) {
                const uint32_t c = p.ucoord[i];
                present = c != PCL_COORD_ABSENT && (p.ufstatus[i] & 3u) == PCL_SPLIT_OK;
                start = c >> 16; len = c & 0xFFFFu;
            } else {
                uint32_t L = p.ulen[i];
                if (p.umax_len && L > p.umax_len) L = p.umax_len;
                present = start < L && start + p.umin <= L;                              // pcl_static_coord
                len = (start + len < L ? start + len : L) - start;
            }
            if (present) {
                if (len > 19u) { ok = false; }                                          // 58 payload bits at 3 per base
                else {
                    const uint8_t *u = p.useq + i * (uint64_t)p.ustride + start;
                    unsigned long long acc = 0;
                    for (uint32_t t = 0; t < len; ++t) {
                        const uint8_t b = u[t];
                        const uint32_t c = b == 'A' ? 0u : b == 'C' ? 1u : b == 'G' ? 2u : b == 'T' ? 3u : b == 'N' ? 4u : 7u;
                        if (c == 7u) ok = false;
                        acc = (acc << 3) | c;
                    }
                    uk = pcl_left_align(acc, 3u * len, len) | (1ull << 5);               // bit 5: a UMI is present
                }
                if (!ok) { bk = PCL_GROUP_NOKEY; ++n_bad; }
            }
        }
        p.bc_key[i] = bk;
        p.umi_key[i] = ok ? uk : 0ull;
        p.perm[i] = p.perm_base + (uint32_t)i;
        if (ok) { ++n_assigned; or_bc |= bk; or_umi |= uk; }
    }
    for (int o = 16; o > 0; o >>= 1) {
        n_assigned += __shfl_xor_sync(0xFFFFFFFFu, n_assigned, o);
        n_bad += __shfl_xor_sync(0xFFFFFFFFu, n_bad, o);
        or_bc |= __shfl_xor_sync(0xFFFFFFFFu, or_bc, o);
        or_umi |= __shfl_xor_sync(0xFFFFFFFFu, or_umi, o);
    }
    if ((threadIdx.x & 31u) == 0) {
        if (n_assigned) atomicAdd(p.stats + 0, n_assigned);
        if (n_bad) atomicAdd(p.stats + 1, n_bad);
        if (or_bc) atomicOr(p.stats + 2, or_bc);
        if (or_umi) atomicOr(p.stats + 3, or_umi);
    }
}

// keys[k] = src[perm[k]]: the next sort key in the order the previous (less significant) sorts left
__global__ void __launch_bounds__(256) k_group_gather(const unsigned long long *__restrict__ src, const uint32_t *__restrict__ perm,
                                                     unsigned long long *__restrict__ keys, uint64_t n, unsigned long long *or_out) {
    unsigned long long acc = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const unsigned long long k = src[perm[i]];
        keys[i] = k;
        acc |= k;
    }
    if (or_out) {
        for (int o = 16; o > 0; o >>= 1) acc |= __shfl_xor_sync(0xFFFFFFFFu, acc, o);
        if ((threadIdx.x & 31u) == 0 && acc) atomicOr(or_out, acc);
    }
}

// position k of the sorted order (k < n_assigned): does a new barcode / a new duplicate set start here?
__global__ void __launch_bounds__(256) k_group_flags(const unsigned long long *__restrict__ bc_sorted, const uint32_t *__restrict__ perm,
                                                    const unsigned long long *__restrict__ umi_key, const unsigned long long *__restrict__ fp,
                                                    uint64_t n_assigned, uint32_t *__restrict__ head_bc, uint32_t *__restrict__ head_frag) {
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n_assigned; k += (uint64_t)gridDim.x * blockDim.x) {
        bool nb = true, nf = true;
        if (k > 0) {
            nb = bc_sorted[k] != bc_sorted[k - 1];
            const uint32_t a = perm[k], b = perm[k - 1];
            nf = nb || umi_key[a] != umi_key[b] || (fp && fp[a] != fp[b]);
        }
        head_bc[k] = nb;
        head_frag[k] = nf;
    }
}

// frag_id / bc_id: exclusive scans of the head flags
__global__ void __launch_bounds__(256) k_group_emit(const unsigned long long *__restrict__ bc_sorted, const uint32_t *__restrict__ head_bc,
                                                   const uint32_t *__restrict__ head_frag, const uint32_t *__restrict__ bc_id,
                                                   const uint32_t *__restrict__ frag_id, uint64_t n_assigned, uint32_t *__restrict__ frag_start,
                                                   uint32_t *__restrict__ bc_frag_start, unsigned long long *__restrict__ bc_key_out) {
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n_assigned; k += (uint64_t)gridDim.x * blockDim.x) {
        if (head_frag[k]) frag_start[frag_id[k]] = (uint32_t)k;
        if (head_bc[k]) { bc_frag_start[bc_id[k]] = frag_id[k]; bc_key_out[bc_id[k]] = bc_sorted[k]; }
    }
}

#endif  // PCL_GROUP_CUH
