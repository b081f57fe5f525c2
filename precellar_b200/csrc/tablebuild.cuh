// tablebuild.cuh -- whitelist tables built on the device (included once, by pcl.cu).
//
// Large whitelists (the 6.8 M-entry 10x v3 list) used to be deduplicated and hashed into the main table and the four
// neighbourhood tables by host threads and uploaded (~470 MB, ~1.5-3 s per rank).  Here the host only packs the entries
// to 2-bit keys (27 MB up); the device
//   1. removes duplicates keeping the first occurrence (IndexSet semantics, seqspec/src/region.rs:459-469,
//      precellar/src/barcode.rs:393-407): stable radix sort of (key, file index), first-of-run flags scattered back to
//      file order, exclusive scan, compaction -- so whitelist index i keeps meaning "i-th distinct line of the file";
//   2. inserts every key into the five tables with compare-and-swap (slots of a bucket fill in order, a full bucket
//      sends the key to the next one: the invariants pcl_probe32 and pcl_filter_bucket rely on).
// The slot a key lands in depends on thread timing, so ranks no longer agree on slots: pcl_counts_allreduce sums the
// counts in whitelist (dense) order -- gather, all-reduce of n_entries words, scatter -- which also sends a third of the
// bytes (27 MB instead of 78 MB for v3).
#ifndef PCL_TABLEBUILD_CUH
#define PCL_TABLEBUILD_CUH

#include "hd_core.h"

// sorted (key, file index): flag[index] = 1 for the first entry of every run of equal keys
__global__ void __launch_bounds__(256) k_wl_mark_first(const unsigned long long *__restrict__ keys_sorted, const uint32_t *__restrict__ idx_sorted,
                                                      uint64_t n, uint32_t *__restrict__ flag) {
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (uint64_t)gridDim.x * blockDim.x)
        flag[idx_sorted[k]] = (k == 0 || keys_sorted[k] != keys_sorted[k - 1]) ? 1u : 0u;
}

// dense[pos[i]] = (uint32_t)key[i] for the kept entries (file order)
__global__ void __launch_bounds__(256) k_wl_compact(const unsigned long long *__restrict__ keys, const uint32_t *__restrict__ flag,
                                                   const uint32_t *__restrict__ pos, uint64_t n, uint32_t *__restrict__ dense) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        if (flag[i]) dense[pos[i]] = (uint32_t)keys[i];
}

// kMain: the table is bucketed by the whole key and the slot of entry i is recorded; else by the key without byte q
template <bool kMain>
__global__ void __launch_bounds__(256) k_wl_insert(const uint32_t *__restrict__ dense, uint64_t m, uint32_t *__restrict__ tab, uint32_t n_buckets,
                                                  uint32_t empty, uint32_t q, uint32_t *__restrict__ index_to_slot) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t key = dense[i];
        uint32_t b = pcl_hash32_bucket(kMain ? key : (key & ~(0xFFu << (8u * q))), n_buckets);
        for (;;) {
            uint32_t s = 0;
            for (; s < 8u; ++s) {
                uint32_t *slot = tab + (uint64_t)b * 8 + s;
                if (*(volatile uint32_t *)slot != empty) continue;                  // taken (slots never become free again)
                if (atomicCAS(slot, empty, key) == empty) break;
            }
            if (s < 8u) { if (kMain) index_to_slot[i] = b * 8u + s; break; }
            b = (b + 1u == n_buckets) ? 0u : b + 1u;
        }
    }
}

__global__ void __launch_bounds__(256) k_fill_u32(uint32_t *p, uint64_t n, uint32_t v) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) p[i] = v;
}

// counts in whitelist order and back (pcl_counts_allreduce)
__global__ void __launch_bounds__(256) k_counts_gather32(const uint32_t *__restrict__ counts, const uint32_t *__restrict__ index_to_slot,
                                                        uint32_t *__restrict__ dense, uint64_t m) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (uint64_t)gridDim.x * blockDim.x) dense[i] = counts[index_to_slot[i]];
}
__global__ void __launch_bounds__(256) k_counts_scatter32(uint32_t *__restrict__ counts, const uint32_t *__restrict__ index_to_slot,
                                                         const uint32_t *__restrict__ dense, uint64_t m) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (uint64_t)gridDim.x * blockDim.x) counts[index_to_slot[i]] = dense[i];
}

#endif  // PCL_TABLEBUILD_CUH
