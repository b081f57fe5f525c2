// tablebuild.cuh -- whitelist tables built on the device (included once, by pcl.cu).
//
// Large whitelists (the 6.8 M-entry 10x v3 list) used to be deduplicated and hashed into the main table and the four
// neighbourhood tables by host threads and uploaded (~470 MB, ~1.5-3 s per rank).  Here the host only packs the entries
// to 2-bit keys (27 MB up); the device
//   1. removes duplicates keeping the first occurrence (IndexSet semantics, seqspec/src/region.rs:459-469,
//      precellar/src/barcode.rs:393-407): stable radix sort of (key, file index), first-of-run flags scattered back to
//      file order, exclusive scan, compaction -- so whitelist index i keeps meaning "i-th distinct line of the file";
//   2. places every key in the main table by RANK, not by timing: the distinct keys are sorted by home bucket (stable, so
//      file order inside a bucket), the first eight of a bucket take its slots in that order and the few keys of
//      over-full buckets are walked to the next bucket with room by one thread, in sorted order.  Every rank therefore
//      builds the same table, slot for slot (the invariants pcl_probe32 relies on hold: a key sits in its home bucket
//      unless that was full of home keys, else in the first later bucket that had room);
//   3. inserts the keys into the four neighbourhood tables with compare-and-swap (their slots carry no counts, so their
//      order does not matter).
// Because the ranks agree on slots, pcl_counts_allreduce may sum the counts either in slot order or -- the default -- as
// the n_entries occupied slots in ascending slot order (a third of the bytes; the gather and scatter around the
// all-reduce then stream through the count array instead of hopping through it in whitelist order).
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

// sort key = home bucket of every distinct key, value = its whitelist index
__global__ void __launch_bounds__(256) k_wl_home(const uint32_t *__restrict__ dense, uint64_t m, uint32_t n_buckets,
                                                unsigned long long *__restrict__ home, uint32_t *__restrict__ idx) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (uint64_t)gridDim.x * blockDim.x) {
        home[i] = pcl_hash32_bucket(dense[i], n_buckets);
        idx[i] = (uint32_t)i;
    }
}

// (home, idx) sorted by home bucket: entry k has rank = number of equal homes in front of it; ranks 0..7 take the slots of
// the home bucket, the rest are flagged for k_wl_overflow.  order_slots[k] = slot of sorted entry k (ascending but for
// the overflow entries).
__global__ void __launch_bounds__(256) k_wl_place(const unsigned long long *__restrict__ home, const uint32_t *__restrict__ idx,
                                                 const uint32_t *__restrict__ dense, uint64_t m, uint32_t *__restrict__ tab,
                                                 uint32_t *__restrict__ index_to_slot, uint32_t *__restrict__ order_slots,
                                                 uint32_t *__restrict__ over) {
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < m; k += (uint64_t)gridDim.x * blockDim.x) {
        const unsigned long long b = home[k];
        uint32_t rank = 0;
        while (rank < 8u && k > rank && home[k - rank - 1u] == b) ++rank;
        if (rank < 8u) {
            const uint32_t i = idx[k], slot = (uint32_t)b * 8u + rank;
            tab[slot] = dense[i];
            index_to_slot[i] = slot;
            order_slots[k] = slot;
            over[k] = 0u;
        } else {
            over[k] = 1u;
        }
    }
}

// list[pos[k]] = k for the flagged entries (sorted order kept)
__global__ void __launch_bounds__(256) k_wl_overflow_list(const uint32_t *__restrict__ over, const uint32_t *__restrict__ pos, uint64_t m,
                                                         uint32_t *__restrict__ list) {
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < m; k += (uint64_t)gridDim.x * blockDim.x)
        if (over[k]) list[pos[k]] = (uint32_t)k;
}

// One thread walks the keys of over-full buckets, in sorted order, to the first later bucket with a free slot (a few
// thousand of 6.8 M keys at the table's load factor): the order of these insertions decides slots, so it is serial.
__global__ void k_wl_overflow(const unsigned long long *__restrict__ home, const uint32_t *__restrict__ idx, const uint32_t *__restrict__ dense,
                              const uint32_t *__restrict__ list, const unsigned long long *__restrict__ n_list, uint32_t *__restrict__ tab,
                              uint32_t n_buckets, uint32_t empty, uint32_t *__restrict__ index_to_slot, uint32_t *__restrict__ order_slots) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    const unsigned long long n = *n_list;
    for (unsigned long long p = 0; p < n; ++p) {
        const uint32_t k = list[p], i = idx[k], key = dense[i];
        uint32_t b = (uint32_t)home[k];
        for (;;) {
            b = (b + 1u == n_buckets) ? 0u : b + 1u;
            uint32_t s = 0;
            while (s < 8u && tab[(uint64_t)b * 8 + s] != empty) ++s;
            if (s < 8u) {
                const uint32_t slot = b * 8u + s;
                tab[slot] = key;
                index_to_slot[i] = slot;
                order_slots[k] = slot;
                break;
            }
        }
    }
}

__global__ void __launch_bounds__(256) k_fill_u32(uint32_t *p, uint64_t n, uint32_t v) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) p[i] = v;
}

// the counts of the occupied slots, in the order of `index_to_slot` (pcl_counts_allreduce passes the ascending slot order), and back
__global__ void __launch_bounds__(256) k_counts_gather32(const uint32_t *__restrict__ counts, const uint32_t *__restrict__ index_to_slot,
                                                        uint32_t *__restrict__ dense, uint64_t m) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (uint64_t)gridDim.x * blockDim.x) dense[i] = counts[index_to_slot[i]];
}
__global__ void __launch_bounds__(256) k_counts_scatter32(uint32_t *__restrict__ counts, const uint32_t *__restrict__ index_to_slot,
                                                         const uint32_t *__restrict__ dense, uint64_t m) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (uint64_t)gridDim.x * blockDim.x) counts[index_to_slot[i]] = dense[i];
}

#endif  // PCL_TABLEBUILD_CUH
