// kernels.cuh -- sm_100a kernels of the barcode hot path (included once, by pcl.cu).
//
//   k_count     pass 1: split -> pack -> whitelist probe -> histogram (+ miss worklist, UMI keys, coords)
//               replaces hot loop A, BarcodeAnalyzer::new  (precellar/src/barcode.rs:85-119)
//               and the split/rev-comp part of FastqAnnotator::annotate (align/fastq.rs:484-516)
//   k_enum_all  builds a worklist of every present barcode (only when max_expected_errors is finite)
//   k_correct   pass 2: likelihood1 posterior for every worklist entry (barcode.rs:160-184,311-358)
#ifndef PCL_KERNELS_CUH
#define PCL_KERNELS_CUH

#include <cuda_runtime.h>
#include <type_traits>

#include "hd_core.h"

#define PCL_BLOCK 256
#define PCL_HCACHE_BITS 12
#define PCL_HCACHE (1u << PCL_HCACHE_BITS)
#define PCL_HEMPTY 0xFFFFFFFFu

struct CountParams {
    DRead rd;
    DTable tab[PCL_MAX_BC_PER_READ];
    const uint8_t *seq;
    const uint16_t *len;
    uint64_t n;
    uint16_t *fstatus;
    uint32_t *tcoord;
    void *bc_key[PCL_MAX_BC_PER_READ];
    void *umi_key[PCL_MAX_UMI_PER_READ];
    uint32_t *bcoord[PCL_MAX_BC_PER_READ];
    uint32_t *ucoord[PCL_MAX_UMI_PER_READ];
    uint8_t bc_key_bytes[PCL_MAX_BC_PER_READ];
    uint8_t umi_key_bytes[PCL_MAX_UMI_PER_READ];
    uint8_t bc_elem[PCL_MAX_BC_PER_READ];
    uint8_t umi_elem[PCL_MAX_UMI_PER_READ];
    unsigned long long *worklist;        // entries: pcl_wl_pack (hd_core.h)
    uint32_t *wl_key;                    // the 4-byte device key of every entry (reads whose tables are all 32-bit), else NULL
    unsigned long long *wl_count;
    unsigned long long *num_defect;
    uint32_t hist_admit_mask;            // k_count_fast: admit a new key to the histogram cache on 1 of (mask + 1) sightings
    uint32_t hist_two_choice;            // k_count_fast: two candidate cache lines per key
};

struct CorrectParams {
    DRead rd;
    DTable tab[PCL_MAX_BC_PER_READ];
    const uint8_t *seq;
    const uint8_t *qual;
    const uint16_t *len;
    uint64_t n;
    uint16_t *fstatus;
    void *bc_key[PCL_MAX_BC_PER_READ];
    uint32_t *bcoord[PCL_MAX_BC_PER_READ];
    uint8_t bc_key_bytes[PCL_MAX_BC_PER_READ];
    uint8_t bc_elem[PCL_MAX_BC_PER_READ];
    uint16_t static_start[PCL_MAX_ELEMS];
    unsigned long long *worklist;
    uint32_t *wl_key;
    unsigned long long *wl_cand;         // k_filter -> k_resolve: candidate mask of every entry, bit (pos << 2 | base)
    unsigned long long *wl_count;
    const double *lut_cap;
    const double *lut_raw;
    double threshold;
    double max_expected_errors;
    int check_ee;
};


// Worklist allocation.  A single global cursor bumped once per warp iteration serialises on one L2 address
// (millions of same-address returning atomics per launch).  Each warp instead reserves PCL_WL_BATCH entries
// at a time and hands them out locally; only the unused tail of a warp's LAST reservation is left over and is
// filled with PCL_WL_HOLE, which k_correct skips.  *wl_count is the number of RESERVED entries
// (misses + at most PCL_WL_BATCH holes per warp).
#define PCL_WL_BATCH 256u
#define PCL_WL_HOLE 0xFFFFFFFFFFFFFFFFull

struct WlCursor { unsigned long long pos, end; };

template <typename P>
__device__ __forceinline__ void pcl_wl_push(const P &p, WlCursor &cur, bool m, unsigned long long entry, uint32_t key, uint32_t lane) {
    const uint32_t ballot = __ballot_sync(0xFFFFFFFFu, m);
    if (!ballot) return;
    const unsigned long long need = __popc(ballot), rank = __popc(ballot & ((1u << lane) - 1u));
    const unsigned long long rem = cur.end - cur.pos;   // warp-uniform
    unsigned long long at = cur.pos + rank;
    if (need > rem) {
        // finish the old reservation exactly (no holes in the middle of the list), continue in a new one
        unsigned long long b = 0;
        if (lane == 0) b = atomicAdd(p.wl_count, (unsigned long long)PCL_WL_BATCH);
        b = __shfl_sync(0xFFFFFFFFu, b, 0);
        if (rank >= rem) at = b + rank - rem;
        cur.pos = b + need - rem;
        cur.end = b + PCL_WL_BATCH;
    } else {
        cur.pos += need;
    }
    if (m) {
        p.worklist[at] = entry;
        if (p.wl_key) p.wl_key[at] = key;
    }
}

// two entries per lane in one reservation step (two records of a lane, or the two barcodes of a record)
template <typename P>
__device__ __forceinline__ void pcl_wl_push2(const P &p, WlCursor &cur, bool m0, uint32_t hi0, uint32_t rec0, uint32_t key0, bool m1,
                                             uint32_t hi1, uint32_t rec1, uint32_t key1, uint32_t lane) {
    const uint32_t b0 = __ballot_sync(0xFFFFFFFFu, m0), b1 = __ballot_sync(0xFFFFFFFFu, m1);
    if (!(b0 | b1)) return;
    const uint32_t lt = (1u << lane) - 1u;
    const uint32_t n0 = __popc(b0), need = n0 + __popc(b1);
    const uint32_t r0 = __popc(b0 & lt), r1 = n0 + __popc(b1 & lt);
    const unsigned long long rem = cur.end - cur.pos;   // warp-uniform
    unsigned long long at0 = cur.pos + r0, at1 = cur.pos + r1;
    if (need > rem) {
        unsigned long long b = 0;
        if (lane == 0) b = atomicAdd(p.wl_count, (unsigned long long)PCL_WL_BATCH);
        b = __shfl_sync(0xFFFFFFFFu, b, 0);
        if (r0 >= rem) at0 = b + r0 - rem;
        if (r1 >= rem) at1 = b + r1 - rem;
        cur.pos = b + need - rem;
        cur.end = b + PCL_WL_BATCH;
    } else {
        cur.pos += need;
    }
    if (m0) { p.worklist[at0] = ((unsigned long long)hi0 << 32) | rec0; p.wl_key[at0] = key0; }
    if (m1) { p.worklist[at1] = ((unsigned long long)hi1 << 32) | rec1; p.wl_key[at1] = key1; }
}

template <typename P>
__device__ __forceinline__ void pcl_wl_finish(const P &p, const WlCursor &cur, uint32_t lane) {
    for (unsigned long long q = cur.pos + lane; q < cur.end; q += 32) p.worklist[q] = PCL_WL_HOLE;
}

// Per-CTA histogram cache: a direct-mapped table in shared memory keyed on (j, slot).  Zipf-distributed
// barcodes send a large share of the reads to a few counters; the cache turns those into shared-memory
// atomics and one global atomic per cached key at the end of the CTA's life.
template <int kBits = PCL_HCACHE_BITS>
__device__ __forceinline__ void pcl_hist_add(uint32_t *ck, uint32_t *cc, uint32_t j, uint32_t slot, uint32_t *gcounts) {
    const uint32_t id = (j << 29) | slot;
    const uint32_t h = (id * 2654435761u) >> (32 - kBits);
    uint32_t cur = ((volatile uint32_t *)ck)[h];
    if (cur == PCL_HEMPTY) {
        uint32_t old = atomicCAS(&ck[h], PCL_HEMPTY, id);
        cur = (old == PCL_HEMPTY) ? id : old;
    }
    if (cur == id) atomicAdd(&cc[h], 1u);
    else atomicAdd(&gcounts[slot], 1u);
}

// Variant for the fast kernel: two candidate lines per key and an admission filter.  A line never changes its
// owner, so every increment lands in exactly one counter that belongs to its key whatever the interleaving;
// the policy only decides how many increments stay in shared memory.  `salt` varies per record: a key is
// admitted to a free line on 1 of (admit_mask + 1) sightings, which keeps one-off (ambient) barcodes from
// filling the cache before the frequent ones have claimed their lines.
template <int kBits>
__device__ __forceinline__ void pcl_hist_add2(uint32_t *ck, uint32_t *cc, uint32_t slot, uint32_t *gcounts, uint32_t salt,
                                              uint32_t admit_mask, bool two_choice) {
    const uint32_t id = slot;
    const uint32_t h1 = (id * 2654435761u) >> (32 - kBits);
    uint32_t c1 = ((volatile uint32_t *)ck)[h1];
    if (c1 == id) { atomicAdd(&cc[h1], 1u); return; }
    const uint32_t h2 = (id * 0x85EBCA6Bu + 0x6A09E667u) >> (32 - kBits);
    uint32_t c2 = PCL_HEMPTY - 1u;
    if (two_choice) {
        c2 = ((volatile uint32_t *)ck)[h2];
        if (c2 == id) { atomicAdd(&cc[h2], 1u); return; }
    }
    if ((((salt * 0x9E3779B1u) >> 20) & admit_mask) == 0u) {
        if (c1 == PCL_HEMPTY) {
            c1 = atomicCAS(&ck[h1], PCL_HEMPTY, id);
            if (c1 == PCL_HEMPTY || c1 == id) { atomicAdd(&cc[h1], 1u); return; }
        }
        if (two_choice && c2 == PCL_HEMPTY) {
            c2 = atomicCAS(&ck[h2], PCL_HEMPTY, id);
            if (c2 == PCL_HEMPTY || c2 == id) { atomicAdd(&cc[h2], 1u); return; }
        }
    }
    atomicAdd(&gcounts[slot], 1u);
}

__device__ __forceinline__ void pcl_st_stream(uint32_t *p, uint32_t v) { __stcs(p, v); }
__device__ __forceinline__ void pcl_st_stream(uint16_t *p, uint16_t v) { __stcs(p, v); }
__device__ __forceinline__ void pcl_st_stream(uint4 *p, uint4 v) { __stcs(p, v); }

__device__ __forceinline__ uint4 pcl_ld_stream(const uint4 *p) {
    uint4 r;
    asm volatile("ld.global.cs.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

// kMode 1: the record (<= 64 bytes) is converted once to 2-bit codes in registers; anchors are matched and fields
// are cut in code space (pcl_split_codes / pcl_pack_codes).  kMode 2: the same for AnchorPlan layouts with 32-bit
// strings, the collapsed split and warp-uniform field offsets (hd_core.h).  kMode 0: the record bytes are read on demand.
template <int kMode>
__global__ void __launch_bounds__(PCL_BLOCK, kMode == 2 ? 4 : 1) k_count(const __grid_constant__ CountParams p) {
    constexpr bool kCodes = kMode == 1;
    constexpr bool kPlan = kMode == 2;
    __shared__ uint32_t ck[PCL_HCACHE];
    __shared__ uint32_t cc[PCL_HCACHE];
    const DRead &rd = p.rd;
    const bool use_hist = rd.n_bc > 0;
    if (use_hist) {
        for (uint32_t h = threadIdx.x; h < PCL_HCACHE; h += blockDim.x) { ck[h] = PCL_HEMPTY; cc[h] = 0; }
        __syncthreads();
    }

    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t step = (uint64_t)gridDim.x * blockDim.x;
    unsigned long long n_total[PCL_MAX_BC_PER_READ] = {0, 0, 0, 0};
    unsigned long long n_mis[PCL_MAX_BC_PER_READ] = {0, 0, 0, 0};
    unsigned long long n_defect = 0;
    const bool rev = rd.is_reverse != 0;
    WlCursor wl = {0, 0};

    // every lane of a warp runs the same number of iterations (full-mask ballots below)
    for (uint64_t base = (uint64_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31u); base < p.n; base += step) {
        const uint64_t i = base + lane;
        const bool live = i < p.n;
        uint32_t miss_mask_j = 0;                 // bit j: barcode j of this record goes to the worklist
        unsigned long long miss_ent[PCL_MAX_BC_PER_READ];
        uint32_t miss_key[PCL_MAX_BC_PER_READ];
        if (live) {
            uint32_t L = p.len[i];
            if (rd.max_len && L > rd.max_len) L = rd.max_len;          // FastqReader::read_record (read.rs:98-103)
            const uint8_t *rec = p.seq + i * (uint64_t)rd.stride;
            uint32_t fs;
            if (kPlan) {
                uint32_t w[16];
                const uint4 *r4 = reinterpret_cast<const uint4 *>(rec);
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                    uint4 t = make_uint4(0u, 0u, 0u, 0u);
                    if ((uint32_t)v * 16u < rd.stride) t = pcl_ld_stream(r4 + v);
                    w[4 * v] = t.x; w[4 * v + 1] = t.y; w[4 * v + 2] = t.z; w[4 * v + 3] = t.w;
                }
                const uint32_t nbytes = L < rd.stride ? L : rd.stride;
                Codes32 k;
                if (rev) pcl_codes32_t<true>(w, (int)((nbytes + 3u) >> 2), k); else pcl_codes32_t<false>(w, (int)((nbytes + 3u) >> 2), k);
                AnchorSplit sp;
                pcl_anchor_split(rd, k, L, sp);
                uint32_t c2[4], bd2[4];                                 // the strings behind the anchor, at uniform offsets
                pcl_shr_bases(k.c, sp.pos, c2);
                pcl_shr_bases(k.bd, sp.pos, bd2);
                fs = sp.status;
                if (sp.status != PCL_SPLIT_OK) ++n_defect;
                if (rd.has_target) {
                    const uint32_t tc = rd.plan.t_elem != 0xFFu ? pcl_anchor_coord(rd, sp, rd.plan.t_elem) : PCL_COORD_ABSENT;
                    if (tc != PCL_COORD_ABSENT) fs |= PCL_FSTATUS_TARGET;
                    p.tcoord[i] = tc;
                }
#pragma unroll
                for (uint32_t j = 0; j < PCL_MAX_BC_PER_READ; ++j) {
                    if (j >= rd.n_bc) { fs |= PCL_BC_ABSENT << (4 + 3 * j); continue; }
                    const uint32_t e = p.bc_elem[j];
                    const uint32_t c = pcl_anchor_coord(rd, sp, e);
                    p.bcoord[j][i] = c;
                    uint32_t code = PCL_BC_ABSENT, key = 0;
                    if (c != PCL_COORD_ABSENT) {
                        const uint32_t blen = c & 0xFFFFu;
                        uint32_t ninv, ipos;
                        key = pcl_anchor_key(rd, k, c2, bd2, e, c, &ninv, &ipos);
                        const DTable &t = p.tab[j];
                        uint32_t slot = 0xFFFFFFFFu;
                        if (ninv == 0u && (t.mode == PCL_TAB_K32_L16 ? blen == 16u : blen <= 15u)) slot = pcl_probe32(t, key);
                        ++n_total[j];                                  // WhitelistBuilder::add (barcode.rs:410-426)
                        if (slot != 0xFFFFFFFFu) {
                            code = PCL_BC_EXACT;
                            pcl_hist_add(ck, cc, j, slot, t.counts);
                        } else {
                            code = PCL_BC_NO_MATCH;
                            ++n_mis[j];
                            if (ninv <= 1u) {                          // two or more non-ACGT bases: NoMatch is final
                                miss_mask_j |= 1u << j;
                                miss_ent[j] = pcl_wl_pack(i, c >> 16, blen, j, ninv, ipos);
                                miss_key[j] = key;
                            }
                        }
                    }
                    reinterpret_cast<uint32_t *>(p.bc_key[j])[i] = key;
                    fs |= code << (4 + 3 * j);
                }
#pragma unroll
                for (uint32_t j = 0; j < PCL_MAX_UMI_PER_READ; ++j) {
                    if (j >= rd.n_umi) continue;
                    const uint32_t e = p.umi_elem[j];
                    const uint32_t c = pcl_anchor_coord(rd, sp, e);
                    p.ucoord[j][i] = c;
                    uint32_t key = PCL_KEY_ABSENT32;
                    if (c != PCL_COORD_ABSENT) {
                        uint32_t ninv;
                        key = pcl_anchor_key(rd, k, c2, bd2, e, c, &ninv);
                        if (ninv) key = 0;
                    }
                    reinterpret_cast<uint32_t *>(p.umi_key[j])[i] = key;
                }
            } else {
            SplitOut so;
            RecCodes rc;
            if (kCodes) {
                uint32_t w[16];
                const uint4 *r4 = reinterpret_cast<const uint4 *>(rec);
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                    uint4 t = make_uint4(0u, 0u, 0u, 0u);
                    if ((uint32_t)v * 16u < rd.stride) t = pcl_ld_stream(r4 + v);
                    w[4 * v] = t.x; w[4 * v + 1] = t.y; w[4 * v + 2] = t.z; w[4 * v + 3] = t.w;
                }
                // only the words that hold record bytes (bytes past the record end are never addressed by split)
                const uint32_t nbytes = L < rd.stride ? L : rd.stride;
                pcl_rec_codes(w, (int)((nbytes + 3u) >> 2), rev, rc);
                pcl_split_codes(rd, rc, L, so);
            } else {
                pcl_split(rd, rec, L, so);
            }
            fs = so.status;
            const bool ok = so.status == PCL_SPLIT_OK;
            if (so.status == PCL_SPLIT_ANCHOR_NOT_FOUND || so.status == PCL_SPLIT_PATTERN_MISMATCH) ++n_defect;
            if (ok && so.tcoord != PCL_COORD_ABSENT) fs |= PCL_FSTATUS_TARGET;
            if (rd.has_target) p.tcoord[i] = ok ? so.tcoord : PCL_COORD_ABSENT;

#pragma unroll
            for (uint32_t j = 0; j < PCL_MAX_BC_PER_READ; ++j) {
                if (j >= rd.n_bc) { fs |= PCL_BC_ABSENT << (4 + 3 * j); continue; }
                const uint32_t c = ok ? so.seg[p.bc_elem[j]] : PCL_COORD_ABSENT;
                if (!rd.fixed_layout) p.bcoord[j][i] = c;
                uint32_t code = PCL_BC_ABSENT;
                uint64_t key = 0;
                if (c != PCL_COORD_ABSENT) {
                    const uint32_t start = c >> 16, blen = c & 0xFFFFu;
                    uint32_t ninv, ipos;
                    key = kCodes ? pcl_pack_codes(rc, start, blen, rev, &ninv, &ipos) : pcl_pack_key(rec, start, blen, rev, &ninv, &ipos);
                    uint32_t slot = 0xFFFFFFFFu;
                    if (ninv == 0u) slot = pcl_probe(p.tab[j], key, blen);
                    ++n_total[j];                                      // WhitelistBuilder::add (barcode.rs:410-426)
                    if (slot != 0xFFFFFFFFu) {
                        code = PCL_BC_EXACT;
                        pcl_hist_add(ck, cc, j, slot, p.tab[j].counts);
                    } else {
                        code = PCL_BC_NO_MATCH;
                        ++n_mis[j];
                        if (ninv <= 1u) {                              // two or more non-ACGT bases (or an unpackable length): NoMatch is final
                            miss_mask_j |= 1u << j;
                            miss_ent[j] = pcl_wl_pack(i, start, blen, j, ninv, ipos);
                            miss_key[j] = (uint32_t)key;
                        }
                    }
                }
                pcl_store_key(p.bc_key[j], p.bc_key_bytes[j], i, key);
                fs |= code << (4 + 3 * j);
            }
#pragma unroll
            for (uint32_t j = 0; j < PCL_MAX_UMI_PER_READ; ++j) {
                if (j >= rd.n_umi) continue;
                const uint32_t c = ok ? so.seg[p.umi_elem[j]] : PCL_COORD_ABSENT;
                if (!rd.fixed_layout) p.ucoord[j][i] = c;
                uint64_t key = PCL_KEY_ABSENT64;
                if (c != PCL_COORD_ABSENT) {
                    uint32_t ninv, ipos;
                    key = kCodes ? pcl_pack_codes(rc, c >> 16, c & 0xFFFFu, rev, &ninv)
                                 : pcl_pack_key(rec, c >> 16, c & 0xFFFFu, rev, &ninv, &ipos);
                    if (ninv) key = 0;
                }
                pcl_store_key(p.umi_key[j], p.umi_key_bytes[j], i, key);
            }
            }
            p.fstatus[i] = (uint16_t)fs;
        }
        // warp-batched append of the misses to the worklist
#pragma unroll
        for (uint32_t j = 0; j < PCL_MAX_BC_PER_READ; ++j) {
            if (j >= rd.n_bc) break;
            const bool m = (miss_mask_j >> j) & 1u;
            pcl_wl_push(p, wl, m, m ? miss_ent[j] : 0ull, m ? miss_key[j] : 0u, lane);
        }
    }
    pcl_wl_finish(p, wl, lane);

    // totals: warp reduce, one atomic per warp
#pragma unroll
    for (uint32_t j = 0; j < PCL_MAX_BC_PER_READ; ++j) {
        if (j >= rd.n_bc) break;
        unsigned long long t = n_total[j], m = n_mis[j];
        for (int o = 16; o > 0; o >>= 1) {
            t += __shfl_xor_sync(0xFFFFFFFFu, t, o);
            m += __shfl_xor_sync(0xFFFFFFFFu, m, o);
        }
        if (lane == 0) {
            if (t) atomicAdd(p.tab[j].total, t);
            if (m) atomicAdd(p.tab[j].mismatch, m);
        }
    }
    for (int o = 16; o > 0; o >>= 1) n_defect += __shfl_xor_sync(0xFFFFFFFFu, n_defect, o);
    if (lane == 0 && n_defect) atomicAdd(p.num_defect, n_defect);

    if (use_hist) {
        __syncthreads();
        for (uint32_t h = threadIdx.x; h < PCL_HCACHE; h += blockDim.x) {
            const uint32_t id = ck[h], c = cc[h];
            if (id != PCL_HEMPTY && c) atomicAdd(&p.tab[id >> 29].counts[id & 0x1FFFFFFFu], c);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// k_count_plan: pass 1 for AnchorPlan layouts (sci-RNA-seq3 R1, dscATAC R1 ...: one variable element, then an anchor).
// Thread per record; records long enough for every element go through the straight-line pcl_plan_record_fast (one window
// search, fields cut at warp-uniform offsets), anything shorter through the element-by-element walk (out of line).
// ------------------------------------------------------------------------------------------------
template <bool kRev, class G>
__global__ void __launch_bounds__(PCL_BLOCK, 4) k_count_plan(const __grid_constant__ CountParams p) {
    __shared__ uint32_t ck[PCL_HCACHE];
    __shared__ uint32_t cc[PCL_HCACHE];
    const DRead &rd = p.rd;
    const G geo(rd);
    for (uint32_t h = threadIdx.x; h < PCL_HCACHE; h += blockDim.x) { ck[h] = PCL_HEMPTY; cc[h] = 0; }
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t n = (uint32_t)p.n;                                   // chunks hold fewer than 2^32 records (pcl_chunk_create)
    const uint32_t step = gridDim.x * blockDim.x;
    const uint32_t stride = G::kConst ? geo.stride() : rd.stride;
    const uint32_t n_vec = stride >> 4;                                 // 16-byte vectors per record row (<= 4)
    const uint32_t lcap = rd.max_len ? rd.max_len : 0xFFFFFFFFu;
    const uint32_t n_bc = geo.n_bc(), n_umi = geo.n_umi();
    const bool want_t = G::kConst ? geo.has_target() : rd.has_target != 0;
    // a table of 16-mers only holds 16-base barcodes, the marker-form tables only shorter ones
    bool l16[PCL_MAX_BC_PER_READ];
#pragma unroll
    for (uint32_t j = 0; j < PCL_MAX_BC_PER_READ; ++j) l16[j] = j < n_bc && p.tab[j].mode == PCL_TAB_K32_L16;
    uint32_t n_total[PCL_MAX_BC_PER_READ] = {0, 0, 0, 0}, n_mis[PCL_MAX_BC_PER_READ] = {0, 0, 0, 0}, n_defect = 0;
    WlCursor wl = {0, 0};

    // One warp iteration = 32 consecutive records.  `full` (every lane holds a record: all but the last slice of the
    // chunk) is a compile-time property of the body, so that the common case carries no per-lane bounds predicates.
    auto body = [&](const uint32_t base, auto full_tag) {
        constexpr bool kFull = decltype(full_tag)::value;
        const uint32_t i = base + lane;
        const bool live = kFull || i < n;
        const uint32_t ld = live ? i : n - 1u;                          // out-of-range lanes re-read the last record
        if (G::kConst && n - base > step) {
            // the rows of the next iteration travel from DRAM to L2 while this one is processed (every 128-byte line of
            // the warp's slice holds at least one row start: strides are <= 64)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(p.seq + (uint64_t)(i + step) * stride));
            if (lane == 0u) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.len + (i + step)));
        }
        uint32_t w[16];
        const uint4 *r4 = reinterpret_cast<const uint4 *>(p.seq + (uint64_t)ld * stride);
#pragma unroll
        for (uint32_t v = 0; v < 4u; ++v) {
            uint4 t = make_uint4(0u, 0u, 0u, 0u);
            if (v < n_vec) t = pcl_ld_stream(r4 + v);
            w[4 * v] = t.x; w[4 * v + 1] = t.y; w[4 * v + 2] = t.z; w[4 * v + 3] = t.w;
        }
        const uint32_t L = min((uint32_t)__ldcs(p.len + ld), lcap);     // FastqReader::read_record (read.rs:98-103)
        PlanOut o;
        // warp-uniform choice (a warp with one short record walks all of its records: same results, and rare)
        if (__all_sync(0xFFFFFFFFu, L >= geo.full_len())) (void)pcl_plan_fast_g<kRev>(geo, w, L, o);
        else pcl_plan_record<kRev>(rd, w, L, o);
        uint32_t fs = o.status;
        n_defect += live & (o.status != PCL_SPLIT_OK);
        if (want_t) {
            if (o.tcoord != PCL_COORD_ABSENT) fs |= PCL_FSTATUS_TARGET;
            if (live) pcl_st_stream(p.tcoord + i, o.tcoord);
        }
        bool miss[PCL_MAX_BC_PER_READ];
#pragma unroll
        for (uint32_t j = 0; j < PCL_MAX_BC_PER_READ; ++j) {
            miss[j] = false;
            if (j >= n_bc) { fs |= PCL_BC_ABSENT << (4 + 3 * j); continue; }
            const uint32_t c = o.bcoord[j];
            uint32_t code = PCL_BC_ABSENT;
            if (c != PCL_COORD_ABSENT) {
                const uint32_t blen = c & 0xFFFFu;
                const DTable &t = p.tab[j];
                uint32_t slot = 0xFFFFFFFFu;
                if (o.bninv[j] == 0u && (l16[j] ? blen == 16u : blen <= 15u)) slot = pcl_probe32(t, o.bkey[j]);
                n_total[j] += live;                                     // WhitelistBuilder::add (barcode.rs:410-426)
                if (slot != 0xFFFFFFFFu) {
                    code = PCL_BC_EXACT;
                    if (live) pcl_hist_add(ck, cc, j, slot, t.counts);
                } else {
                    code = PCL_BC_NO_MATCH;
                    n_mis[j] += live;
                    miss[j] = live & (o.bninv[j] <= 1u);                // two or more non-ACGT bases: NoMatch is final
                }
            }
            if (live) {
                pcl_st_stream(p.bcoord[j] + i, c);
                pcl_st_stream(reinterpret_cast<uint32_t *>(p.bc_key[j]) + i, o.bkey[j]);
            }
            fs |= code << (4 + 3 * j);
        }
#pragma unroll
        for (uint32_t j = 0; j < PCL_MAX_UMI_PER_READ; ++j) {
            if (j >= n_umi) break;
            if (live) {
                pcl_st_stream(p.ucoord[j] + i, o.ucoord[j]);
                pcl_st_stream(reinterpret_cast<uint32_t *>(p.umi_key[j]) + i, o.ukey[j]);
            }
        }
        if (live) pcl_st_stream(p.fstatus + i, (uint16_t)fs);
        // warp-batched append of the misses to the worklist: one reservation for the first two barcodes of the record
        {
            const uint32_t c0 = o.bcoord[0], c1 = n_bc > 1u ? o.bcoord[1] : 0u;
            const uint32_t h0 = (c0 >> 16) | ((c0 & 0xFFu) << 16) | ((o.bninv[0] == 1u ? 1u : 0u) << 26) | ((o.bninv[0] == 1u ? o.bipos[0] : 0u) << 27);
            const uint32_t h1 = (c1 >> 16) | ((c1 & 0xFFu) << 16) | (1u << 24) | ((o.bninv[1] == 1u ? 1u : 0u) << 26) | ((o.bninv[1] == 1u ? o.bipos[1] : 0u) << 27);
            pcl_wl_push2(p, wl, miss[0], h0, i, o.bkey[0], n_bc > 1u && miss[1], h1, i, o.bkey[1], lane);
        }
#pragma unroll
        for (uint32_t j = 2; j < PCL_MAX_BC_PER_READ; ++j) {
            if (j >= n_bc) break;
            const uint32_t c = o.bcoord[j];
            pcl_wl_push(p, wl, miss[j], pcl_wl_pack(i, c >> 16, c & 0xFFFFu, j, o.bninv[j], o.bipos[j]), o.bkey[j], lane);
        }
    };

    // every lane of a warp runs the same iterations (full-mask ballots in the body); `n - base > step` instead of
    // `base + step < n`: the index may not wrap
    for (uint32_t base = blockIdx.x * blockDim.x + (threadIdx.x & ~31u); base < n;) {
        if (n - base >= 32u) body(base, std::true_type{});
        else body(base, std::false_type{});
        if (n - base <= step) break;
        base += step;
    }
    pcl_wl_finish(p, wl, lane);
#pragma unroll
    for (uint32_t j = 0; j < PCL_MAX_BC_PER_READ; ++j) {
        if (j >= n_bc) break;
        uint32_t t = n_total[j], m = n_mis[j];
        for (int o = 16; o > 0; o >>= 1) { t += __shfl_xor_sync(0xFFFFFFFFu, t, o); m += __shfl_xor_sync(0xFFFFFFFFu, m, o); }
        if (lane == 0) {
            if (t) atomicAdd(p.tab[j].total, (unsigned long long)t);
            if (m) atomicAdd(p.tab[j].mismatch, (unsigned long long)m);
        }
    }
    for (int o = 16; o > 0; o >>= 1) n_defect += __shfl_xor_sync(0xFFFFFFFFu, n_defect, o);
    if (lane == 0 && n_defect) atomicAdd(p.num_defect, (unsigned long long)n_defect);
    __syncthreads();
    for (uint32_t h = threadIdx.x; h < PCL_HCACHE; h += blockDim.x) {
        const uint32_t id = ck[h], c = cc[h];
        if (id != PCL_HEMPTY && c) atomicAdd(&p.tab[id >> 29].counts[id & 0x1FFFFFFFu], c);
    }
}

// ------------------------------------------------------------------------------------------------
// k_count_fast: pass 1 for reads with a FastLayout (10x v2/v3 R1, ATAC / multiome I2, ...).
// One thread per record; the 32 record bytes arrive as two 128-bit streaming loads and stay in registers;
// SWAR conversion to 2-bit codes; one bucket probe of the 32-bit key table.
// ------------------------------------------------------------------------------------------------

// kBits: log2 of the histogram cache entries; the CTA has 256 << (kBits - 12) threads (same cache per thread).
// kSpec: PCL_FAST_SPEC_* instance of the field extraction (0: driven by the layout's byte masks).
template <int kBits, int kSpec = 0>
__global__ void __launch_bounds__(PCL_BLOCK << (kBits - PCL_HCACHE_BITS)) k_count_fast(const __grid_constant__ CountParams p) {
    extern __shared__ uint32_t pcl_hist_smem[];
    uint32_t *ck = pcl_hist_smem, *cc = pcl_hist_smem + (1u << kBits);
    for (uint32_t h = threadIdx.x; h < (1u << kBits); h += blockDim.x) { ck[h] = PCL_HEMPTY; cc[h] = 0; }
    __syncthreads();

    const DRead &rd = p.rd;
    const FastLayout &f = rd.fast;
    const DTable &tab = p.tab[0];
    const bool rev = rd.is_reverse != 0;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t step = gridDim.x * blockDim.x;
    const uint32_t n = (uint32_t)p.n;                                   // chunks hold < 2^32 records
    const uint32_t bc_end = (uint32_t)f.bc_start + f.bc_len, umi_end = (uint32_t)f.umi_start + f.umi_len;
    const bool len_ok = tab.mode == PCL_TAB_K32_L16 ? f.bc_len == 16 : (tab.mode == PCL_TAB_K32_MARKER && f.bc_len <= 15);
    const uint32_t absent_rest = (PCL_BC_ABSENT << 7) | (PCL_BC_ABSENT << 10) | (PCL_BC_ABSENT << 13);
    const bool wide = rd.stride == 32u;
    const uint32_t *keys = reinterpret_cast<const uint32_t *>(tab.keys);
    const uint32_t empty = (uint32_t)tab.empty, nb = tab.n_buckets;
    uint32_t *bc_out = reinterpret_cast<uint32_t *>(p.bc_key[0]);
    uint32_t *umi_out = reinterpret_cast<uint32_t *>(p.umi_key[0]);
    unsigned long long n_total = 0, n_mis = 0;
    WlCursor wl = {0, 0};

    // Three-stage software pipeline over the records of a lane (k = iteration):
    //   load    record k+1 : two streaming 128-bit loads + the length              (in flight during k)
    //   probe   record k   : SWAR extraction, hash, issue the bucket loads          (in flight during k+1's extraction)
    //   retire  record k-1 : compare the bucket, histogram, result stores, worklist
    uint32_t base = blockIdx.x * blockDim.x + (threadIdx.x & ~31u);
    uint4 a = make_uint4(0u, 0u, 0u, 0u), b = make_uint4(0u, 0u, 0u, 0u);
    uint32_t Lraw = 0;
    if (base < n && base + lane < n) {
        const uint4 *r = reinterpret_cast<const uint4 *>(p.seq + (uint64_t)(base + lane) * rd.stride);
        a = pcl_ld_stream(r);
        if (wide) b = pcl_ld_stream(r + 1);
        Lraw = __ldcs(p.len + base + lane);
    }
    // pending (probed, not yet retired) record
    uint4 pv0 = make_uint4(0u, 0u, 0u, 0u), pv1 = pv0;
    uint32_t p_i = 0, p_bk = 0, p_uk = 0, p_fs = 0, p_bucket = 0, p_bad = 0;
    uint32_t p_flags = 0;                       // bit0 live, bit1 barcode present, bit2 probe issued, bit3 UMI present
    bool have_pending = false;

    for (;; base += step) {
        const bool more = base < n;             // warp-uniform
        // ---- probe stage for record k ----
        uint32_t c_i = base + lane, c_bk = 0, c_uk = 0, c_fs = 0, c_flags = 0, c_bucket = 0, c_bad = 0;
        uint4 cv0 = make_uint4(0u, 0u, 0u, 0u), cv1 = cv0;
        if (more) {
            const uint4 ca = a, cb = b;
            uint32_t L = Lraw;
            const uint32_t ni = c_i + step;
            if (ni < n && ni >= c_i) {          // ---- load stage for record k+1 ----
                const uint4 *r = reinterpret_cast<const uint4 *>(p.seq + (uint64_t)ni * rd.stride);
                a = pcl_ld_stream(r);
                if (wide) b = pcl_ld_stream(r + 1);
                Lraw = __ldcs(p.len + ni);
            }
            if (c_i < n) {
                if (rd.max_len && L > rd.max_len) L = rd.max_len;
                const uint32_t w[8] = {ca.x, ca.y, ca.z, ca.w, cb.x, cb.y, cb.z, cb.w};
                bool ubad;
                if (kSpec == PCL_FAST_SPEC_B0_U16x12) pcl_fast_extract_aligned<0, 4, 3>(w, rev, &c_bk, &c_bad, &c_uk, &ubad, false);
                else if (kSpec == PCL_FAST_SPEC_B0) pcl_fast_extract_aligned<0, 0, 0>(w, rev, &c_bk, &c_bad, &c_uk, &ubad, false);
                else if (kSpec == PCL_FAST_SPEC_B8) pcl_fast_extract_aligned<2, 0, 0>(w, rev, &c_bk, &c_bad, &c_uk, &ubad, false);
                else pcl_fast_extract(f, w, rev, &c_bk, &c_bad, &c_uk, &ubad, false);
                c_fs = PCL_SPLIT_OK | absent_rest;
                c_flags = 1u;
                if (f.has_tail_target) {
                    uint32_t tc = PCL_COORD_ABSENT;
                    if (f.t_start < L && (uint32_t)f.t_start + f.t_min <= L) {
                        const uint32_t hi = (uint32_t)f.t_start + f.t_max < L ? (uint32_t)f.t_start + f.t_max : L;
                        tc = ((uint32_t)f.t_start << 16) | (hi - f.t_start);
                        c_fs |= PCL_FSTATUS_TARGET;
                    }
                    pcl_st_stream(p.tcoord + c_i, tc);
                }
                if (umi_end <= L) c_flags |= 8u;
                if (bc_end <= L) {
                    c_flags |= 2u;
                    if (!c_bad && len_ok && c_bk != empty) {
                        c_flags |= 4u;
                        c_bucket = pcl_hash32_bucket(c_bk, nb);
                        cv0 = __ldg(reinterpret_cast<const uint4 *>(keys + (uint64_t)c_bucket * 8));
                        cv1 = __ldg(reinterpret_cast<const uint4 *>(keys + (uint64_t)c_bucket * 8 + 4));
                    }
                } else {
                    c_bk = 0;
                    c_bad = 0;
                }
            }
        }
        // ---- retire stage for record k-1 ----
        if (have_pending) {
            bool miss = false;
            unsigned long long ent = 0;
            if (p_flags & 1u) {
                uint32_t code = PCL_BC_ABSENT;
                if (p_flags & 2u) {
                    ++n_total;
                    uint32_t slot = 0xFFFFFFFFu;
                    if (p_flags & 4u) {
                        uint32_t j = 8;
                        j = (pv1.w == p_bk) ? 7u : j;
                        j = (pv1.z == p_bk) ? 6u : j;
                        j = (pv1.y == p_bk) ? 5u : j;
                        j = (pv1.x == p_bk) ? 4u : j;
                        j = (pv0.w == p_bk) ? 3u : j;
                        j = (pv0.z == p_bk) ? 2u : j;
                        j = (pv0.y == p_bk) ? 1u : j;
                        j = (pv0.x == p_bk) ? 0u : j;
                        if (j < 8u) slot = p_bucket * 8u + j;
                        else if (pv1.w != empty) {                      // full bucket: the chain continues (rare)
                            uint32_t bb = p_bucket;
                            bool full = true;
                            while (slot == 0xFFFFFFFFu && full) {
                                bb = (bb + 1u == nb) ? 0u : bb + 1u;
                                slot = pcl_bucket32(keys, bb, p_bk, empty, &full);
                            }
                        }
                    }
                    if (slot != 0xFFFFFFFFu) {
                        code = PCL_BC_EXACT;
                        pcl_hist_add2<kBits>(ck, cc, slot, tab.counts, p_i, p.hist_admit_mask, p.hist_two_choice != 0);
                    } else {
                        code = PCL_BC_NO_MATCH;
                        ++n_mis;
                        miss = true;                                    // a barcode with non-ACGT bytes: k_filter looks at the record
                        ent = pcl_wl_pack(p_i, f.bc_start, f.bc_len, 0u, p_bad ? 1u : 0u, PCL_WL_REPACK_POS);
                    }
                }
                pcl_st_stream(bc_out + p_i, p_bk);
                if (f.umi_len) pcl_st_stream(umi_out + p_i, (p_flags & 8u) ? p_uk : PCL_KEY_ABSENT32);
                pcl_st_stream(p.fstatus + p_i, (uint16_t)(p_fs | (code << 4)));
            }
            pcl_wl_push(p, wl, miss, ent, p_bk, lane);
        }
        if (!more) break;
        pv0 = cv0; pv1 = cv1; p_i = c_i; p_bk = c_bk; p_uk = c_uk; p_fs = c_fs; p_flags = c_flags; p_bucket = c_bucket; p_bad = c_bad;
        have_pending = true;
    }
    pcl_wl_finish(p, wl, lane);
    for (int o = 16; o > 0; o >>= 1) {
        n_total += __shfl_xor_sync(0xFFFFFFFFu, n_total, o);
        n_mis += __shfl_xor_sync(0xFFFFFFFFu, n_mis, o);
    }
    if (lane == 0) {
        if (n_total) atomicAdd(tab.total, n_total);
        if (n_mis) atomicAdd(tab.mismatch, n_mis);
    }
    __syncthreads();
    for (uint32_t h = threadIdx.x; h < (1u << kBits); h += blockDim.x) {
        const uint32_t id = ck[h], c = cc[h];
        if (id != PCL_HEMPTY && c) atomicAdd(&tab.counts[id & 0x1FFFFFFFu], c);
    }
}

// ------------------------------------------------------------------------------------------------
// k_count_spec: pass 1 for the word-aligned geometries of the BASELINE configurations (PCL_FAST_SPEC_*: a 16-base
// barcode at byte 0 or 8, optionally a 12-base UMI at byte 16; no target in the read), strand known at compile time.
//
// The kernel is bound by instruction issue, not by HBM, so it is written for few instructions per record: straight-line
// code over TWO records per thread and iteration (lanes take records base + lane and base + 32 + lane: every load and
// store of a warp is one contiguous run), no software pipeline -- the 32 resident warps of the 1024-thread CTA hide the
// two load latencies of an iteration (records, then buckets) --, loads of out-of-range lanes clamped to the last record
// instead of branched around, 32-bit counters, one worklist reservation for both records.
// Same results as k_count_fast on the same layouts (tests run both: PCL_OLD_FAST=1).
// ------------------------------------------------------------------------------------------------
template <bool kRev>
__device__ __forceinline__ void pcl_spec_codes4(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3, uint32_t *codes, uint32_t *bad) {
    uint32_t d0, d1, d2, d3;
    const uint32_t fold = kRev ? 0xDFDFDFDFu : 0xFFFFFFFFu;          // precellar::utils::rev_compl accepts lower case
    const uint32_t c0 = pcl_word_codes(w0 & fold, &d0), c1 = pcl_word_codes(w1 & fold, &d1);
    const uint32_t c2 = pcl_word_codes(w2 & fold, &d2), c3 = pcl_word_codes(w3 & fold, &d3);
    *codes = c0 | (c1 << 8) | (c2 << 16) | (c3 << 24);
    *bad = d0 | d1 | d2 | d3;
}

template <int kSpec, bool kRev, bool kPrefetch = true>
__global__ void __launch_bounds__(1024) k_count_spec(const __grid_constant__ CountParams p) {
    constexpr int kBits = 14;
    constexpr bool kWide = kSpec != PCL_FAST_SPEC_B0;                    // 32-byte record rows (16 for the bare index read)
    constexpr bool kUmi = kSpec == PCL_FAST_SPEC_B0_U16x12;
    constexpr uint32_t kBcStart = kSpec == PCL_FAST_SPEC_B8 ? 8u : 0u, kBcEnd = kBcStart + 16u, kUmiEnd = 28u;
    constexpr uint32_t kAhead = 1u;                                      // prefetch distance in iterations
    extern __shared__ uint32_t pcl_hist_smem[];
    uint32_t *ck = pcl_hist_smem, *cc = pcl_hist_smem + (1u << kBits);
    for (uint32_t h = threadIdx.x; h < (1u << kBits); h += blockDim.x) { ck[h] = PCL_HEMPTY; cc[h] = 0; }
    __syncthreads();

    const DTable &tab = p.tab[0];
    const uint32_t *keys = reinterpret_cast<const uint32_t *>(tab.keys);
    const uint32_t empty = (uint32_t)tab.empty, nb = tab.n_buckets;
    const bool len_ok = tab.mode == PCL_TAB_K32_L16;                     // only a table of 16-mers can hold a 16-base barcode
    const uint32_t n = (uint32_t)p.n, last = n - 1u;                     // n >= 1
    const uint32_t lcap = p.rd.max_len ? p.rd.max_len : 0xFFFFFFFFu;     // FastqReader::read_record truncation (read.rs:98-103)
    const uint32_t lane = threadIdx.x & 31u;
    // Grid-stride over 64-record slices: all warps together stream one region of the planes.  (A blocked partition -- every
    // warp its own contiguous range, so that the worklist comes out in record order -- was measured: pass 2 did not get
    // faster and this kernel lost 3 %.)
    const uint32_t first = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 64u, end = n;
    const uint32_t step = gridDim.x * blockDim.x * 2u;
    const uint32_t absent_rest = (PCL_BC_ABSENT << 7) | (PCL_BC_ABSENT << 10) | (PCL_BC_ABSENT << 13);
    // high words of the two kinds of worklist entry: clean barcode / barcode with non-ACGT bytes (k_filter re-derives)
    const uint32_t ent_hi = (uint32_t)(pcl_wl_pack(0, kBcStart, 16u, 0u, 0u, 0u) >> 32);
    const uint32_t ent_hi_bad = (uint32_t)(pcl_wl_pack(0, kBcStart, 16u, 0u, 1u, PCL_WL_REPACK_POS) >> 32);
    uint32_t *bc_out = reinterpret_cast<uint32_t *>(p.bc_key[0]);
    uint32_t *umi_out = reinterpret_cast<uint32_t *>(p.umi_key[0]);
    uint32_t n_total = 0, n_mis = 0;
    WlCursor wl = {0, 0};

    for (uint32_t base = first; base < end; base += step) {
        uint32_t idx[2], L[2], bk[2], uk[2], bad[2], bucket[2];
        uint4 v0[2], v1[2];
        bool live[2], act[2];
        // ---- records ----
        uint4 a[2], b[2];
        if (kPrefetch && base + kAhead * step < end) {
            // rows a few iterations ahead travel from DRAM to L2 while this one is processed (no registers held)
            const uint8_t *nx = p.seq + (uint64_t)(base + kAhead * step + lane) * (kWide ? 32u : 16u);
            asm volatile("prefetch.global.L2 [%0];" ::"l"(nx));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(nx + 32u * (kWide ? 32u : 16u)));
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            idx[r] = base + 32u * r + lane;
            live[r] = idx[r] < n;
            const uint32_t ld = live[r] ? idx[r] : last;                 // out-of-range lanes re-read the last record
            const uint4 *rp = reinterpret_cast<const uint4 *>(p.seq + (uint64_t)ld * (kWide ? 32u : 16u));
            a[r] = pcl_ld_stream(rp);
            if (kWide) b[r] = pcl_ld_stream(rp + 1);
            L[r] = __ldcs(p.len + ld);
        }
        // ---- extraction, probe issue ----
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            uint32_t codes;
            if (kSpec == PCL_FAST_SPEC_B8) pcl_spec_codes4<kRev>(a[r].z, a[r].w, b[r].x, b[r].y, &codes, &bad[r]);
            else pcl_spec_codes4<kRev>(a[r].x, a[r].y, a[r].z, a[r].w, &codes, &bad[r]);
            bk[r] = pcl_field32_to_key(codes, 16u, kRev, 0u);
            if (kUmi) {
                uint32_t d4, d5, d6;
                const uint32_t fold = kRev ? 0xDFDFDFDFu : 0xFFFFFFFFu;
                const uint32_t uc = pcl_word_codes(b[r].x & fold, &d4) | (pcl_word_codes(b[r].y & fold, &d5) << 8) |
                                    (pcl_word_codes(b[r].z & fold, &d6) << 16);
                uk[r] = (d4 | d5 | d6) ? 0u : (pcl_field32_to_key(uc, 12u, kRev, 0u) | (1u << 24));
            }
            L[r] = min(L[r], lcap);
            act[r] = (L[r] >= kBcEnd) & (bad[r] == 0u) & len_ok & (bk[r] != empty);
            bucket[r] = pcl_hash32_bucket(bk[r], nb);
            if (act[r]) {
                v0[r] = __ldg(reinterpret_cast<const uint4 *>(keys + (uint64_t)bucket[r] * 8));
                v1[r] = __ldg(reinterpret_cast<const uint4 *>(keys + (uint64_t)bucket[r] * 8 + 4));
            }
        }
        // ---- retire ----
        bool miss[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const bool present = live[r] & (L[r] >= kBcEnd);
            uint32_t slot = 0xFFFFFFFFu;
            if (act[r]) {
                const uint32_t k = bk[r];
                uint32_t j = 8;
                j = (v1[r].w == k) ? 7u : j;
                j = (v1[r].z == k) ? 6u : j;
                j = (v1[r].y == k) ? 5u : j;
                j = (v1[r].x == k) ? 4u : j;
                j = (v0[r].w == k) ? 3u : j;
                j = (v0[r].z == k) ? 2u : j;
                j = (v0[r].y == k) ? 1u : j;
                j = (v0[r].x == k) ? 0u : j;
                if (j < 8u) slot = bucket[r] * 8u + j;
                else if (v1[r].w != empty) {                              // full bucket: the chain continues (rare)
                    uint32_t bb = bucket[r];
                    bool full = true;
                    while (slot == 0xFFFFFFFFu && full) {
                        bb = (bb + 1u == nb) ? 0u : bb + 1u;
                        slot = pcl_bucket32(keys, bb, k, empty, &full);
                    }
                }
            }
            const bool hit = present & (slot != 0xFFFFFFFFu);
            if (hit) pcl_hist_add2<kBits>(ck, cc, slot, tab.counts, idx[r], p.hist_admit_mask, p.hist_two_choice != 0);
            miss[r] = present & !hit;
            n_total += present;                                          // WhitelistBuilder::add (barcode.rs:410-426)
            n_mis += miss[r];
            if (live[r]) {
                const uint32_t code = !present ? PCL_BC_ABSENT : (hit ? PCL_BC_EXACT : PCL_BC_NO_MATCH);
                pcl_st_stream(bc_out + idx[r], present ? bk[r] : 0u);
                if (kUmi) pcl_st_stream(umi_out + idx[r], L[r] >= kUmiEnd ? uk[r] : PCL_KEY_ABSENT32);
                pcl_st_stream(p.fstatus + idx[r], (uint16_t)(PCL_SPLIT_OK | absent_rest | (code << 4)));
            }
        }
        pcl_wl_push2(p, wl, miss[0], bad[0] ? ent_hi_bad : ent_hi, idx[0], bk[0], miss[1], bad[1] ? ent_hi_bad : ent_hi, idx[1], bk[1], lane);
    }
    pcl_wl_finish(p, wl, lane);
    for (int o = 16; o > 0; o >>= 1) {
        n_total += __shfl_xor_sync(0xFFFFFFFFu, n_total, o);
        n_mis += __shfl_xor_sync(0xFFFFFFFFu, n_mis, o);
    }
    if (lane == 0) {
        if (n_total) atomicAdd(tab.total, (unsigned long long)n_total);
        if (n_mis) atomicAdd(tab.mismatch, (unsigned long long)n_mis);
    }
    __syncthreads();
    for (uint32_t h = threadIdx.x; h < (1u << kBits); h += blockDim.x) {
        const uint32_t id = ck[h], c = cc[h];
        if (id != PCL_HEMPTY && c) atomicAdd(&tab.counts[id & 0x1FFFFFFFu], c);
    }
}

// ------------------------------------------------------------------------------------------------
// k_count_lenonly: pass 1 for fixed-layout reads without barcode / UMI / anchor (insert reads): only the
// record length is read; emits fstatus and the insert coordinates.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void pcl_lenonly_one(const DRead &rd, int ne, uint32_t absent_all, uint32_t L, uint32_t *fs_out, uint32_t *tc_out) {
    if (rd.max_len && L > rd.max_len) L = rd.max_len;
    uint32_t off = 0, tc = PCL_COORD_ABSENT, n_t = 0;
    for (int e = 0; e < ne; ++e) {
        const uint32_t mn = rd.elems[e].min_len, mx = rd.elems[e].max_len;
        if (off >= L || off + mn > L) break;                       // segment.rs:245-248
        const uint32_t hi = off + mx < L ? off + mx : L;           // a fixed element fits entirely; the trailing variable one is clipped
        if (rd.elems[e].kind == PCL_KIND_TARGET) { tc = (off << 16) | (hi - off); ++n_t; }
        off += mx;
    }
    uint32_t fs = absent_all | (n_t > 1u ? PCL_SPLIT_MULTI_TARGET : PCL_SPLIT_OK);
    if (n_t == 1u) fs |= PCL_FSTATUS_TARGET;
    *fs_out = fs;
    *tc_out = n_t == 1u ? tc : PCL_COORD_ABSENT;
}

// Eight records per thread: one 128-bit load of the lengths, one 128-bit store of fstatus, two of tcoord.
__global__ void __launch_bounds__(PCL_BLOCK) k_count_lenonly(const __grid_constant__ CountParams p) {
    const DRead &rd = p.rd;
    const uint32_t absent_all = (PCL_BC_ABSENT << 4) | (PCL_BC_ABSENT << 7) | (PCL_BC_ABSENT << 10) | (PCL_BC_ABSENT << 13);
    const int ne = rd.n_elems;
    const uint64_t n8 = p.n >> 3;
    const uint4 *len8 = reinterpret_cast<const uint4 *>(p.len);
    uint4 *fs8 = reinterpret_cast<uint4 *>(p.fstatus);
    uint4 *tc4 = reinterpret_cast<uint4 *>(p.tcoord);
    for (uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; g < n8; g += (uint64_t)gridDim.x * blockDim.x) {
        const uint4 lv = pcl_ld_stream(len8 + g);
        const uint32_t lw[4] = {lv.x, lv.y, lv.z, lv.w};
        uint32_t fs[8], tc[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) pcl_lenonly_one(rd, ne, absent_all, (lw[k >> 1] >> (16 * (k & 1))) & 0xFFFFu, &fs[k], &tc[k]);
        pcl_st_stream(fs8 + g, make_uint4(fs[0] | (fs[1] << 16), fs[2] | (fs[3] << 16), fs[4] | (fs[5] << 16), fs[6] | (fs[7] << 16)));
        if (rd.has_target) {
            pcl_st_stream(tc4 + 2 * g, make_uint4(tc[0], tc[1], tc[2], tc[3]));
            pcl_st_stream(tc4 + 2 * g + 1, make_uint4(tc[4], tc[5], tc[6], tc[7]));
        }
    }
    // the last n % 8 records
    const uint64_t i = (n8 << 3) + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < p.n) {
        uint32_t fs, tc;
        pcl_lenonly_one(rd, ne, absent_all, p.len[i], &fs, &tc);
        if (rd.has_target) p.tcoord[i] = tc;
        p.fstatus[i] = (uint16_t)fs;
    }
}

// Worklist of every present barcode (exact and not): the expected-error test of correct_barcode
// precedes the whitelist lookup, so with a finite max_expected_errors every barcode needs its qualities.
__global__ void __launch_bounds__(PCL_BLOCK) k_enum_all(const __grid_constant__ CorrectParams p) {
    const DRead &rd = p.rd;
    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t step = (uint64_t)gridDim.x * blockDim.x;
    WlCursor wl = {0, 0};
    for (uint64_t base = (uint64_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31u); base < p.n; base += step) {
        const uint64_t i = base + lane;
        const bool live = i < p.n;
        const uint32_t fs = live ? p.fstatus[i] : 0xFFFFu;
        for (uint32_t j = 0; j < rd.n_bc; ++j) {
            const uint32_t code = PCL_BC_CODE(fs, j);
            const bool m = live && (code == PCL_BC_EXACT || code == PCL_BC_NO_MATCH);
            unsigned long long entry = 0;
            if (m) {
                uint32_t c;
                if (rd.fixed_layout) {
                    const uint32_t e = p.bc_elem[j];
                    uint32_t L = p.len[i];
                    if (rd.max_len && L > rd.max_len) L = rd.max_len;
                    uint32_t s = p.static_start[e], l = rd.elems[e].max_len;
                    if (rd.elems[e].min_len != rd.elems[e].max_len) l = (s + l < L ? s + l : L) - s;   // trailing variable barcode
                    c = (s << 16) | l;
                } else {
                    c = p.bcoord[j][i];
                }
                entry = pcl_wl_pack(i, c >> 16, c & 0xFFFFu, j, 0u, 0u);      // k_correct re-derives the key from the record
            }
            pcl_wl_push(p, wl, m, entry, 0u, lane);
        }
    }
    pcl_wl_finish(p, wl, lane);
}

__device__ __forceinline__ void pcl_fstatus_or(uint16_t *fstatus, uint64_t i, uint32_t bits) {
    uintptr_t a = reinterpret_cast<uintptr_t>(fstatus + i);
    uint32_t *w = reinterpret_cast<uint32_t *>(a & ~(uintptr_t)3);
    atomicOr(w, (a & 2) ? bits << 16 : bits);
}

// Generic pass 2 (64-bit tables, or a finite max_expected_errors): thread per worklist entry, the barcode is re-packed
// from the record bytes, 3 * L (4 at an N) neighbour probes in (pos, ACGT) order.
__global__ void __launch_bounds__(PCL_BLOCK) k_correct(const __grid_constant__ CorrectParams p) {
    const DRead &rd = p.rd;
    const unsigned long long n_work = *p.wl_count;
    const uint64_t step = (uint64_t)gridDim.x * blockDim.x;
    const bool rev = rd.is_reverse != 0;
    for (uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; w < n_work; w += step) {
        const unsigned long long ent = p.worklist[w];
        if (ent == PCL_WL_HOLE) continue;
        const uint64_t rec_i = PCL_WL_REC(ent);
        const uint32_t start = PCL_WL_START(ent), blen = PCL_WL_LEN(ent), j = PCL_WL_J(ent);
        const uint8_t *rec = p.seq + rec_i * (uint64_t)rd.stride;
        const uint8_t *ql = p.qual + rec_i * (uint64_t)rd.qstride - rd.qoff;     // record-relative indexing inside the window
        uint32_t ninv, ipos;
        const uint64_t key = pcl_pack_key(rec, start, blen, rev, &ninv, &ipos);
        const bool is_exact = PCL_BC_CODE((uint32_t)p.fstatus[rec_i], j) == PCL_BC_EXACT;
        uint64_t out_key = 0;
        const uint32_t code = pcl_correct_one(p.tab[j], key, blen, ninv, ipos, ql, start, rev, is_exact, p.threshold,
                                              p.max_expected_errors, p.check_ee != 0, p.lut_cap, p.lut_raw, &out_key);
        if (code == PCL_BC_CORRECTED) pcl_store_key(p.bc_key[j], p.bc_key_bytes[j], rec_i, out_key);
        if (code != PCL_BC_EXACT && code != PCL_BC_NO_MATCH) pcl_fstatus_or(p.fstatus, rec_i, code << (4 + 3 * j));
    }
}

// ------------------------------------------------------------------------------------------------
// Pass 2 for reads whose tables are all 32-bit, in two kernels.
//   k_filter  (launched by pcl_count right behind the count kernel: it needs no counts, so on several GPUs it runs
//             while the all-reduce of the counts is in flight): worklist entry (record, packed key, N position) ->
//             64-bit mask of the 1-Hamming neighbours that are whitelist members.  Touches neither the records nor
//             the qualities: 12 bytes in, 8 bytes out per miss, plus the table lookups (bitmap words from L2, the
//             neighbourhood buckets of the partial keys that exist).
//   k_resolve (pcl_correct, after the all-reduce): candidates -> main-table slot -> global count -> fp64 posterior
//             (likelihood1 / update_best_option / correct_barcode, barcode.rs:160-184,311-358).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long pcl_ld_stream64(const unsigned long long *p) { return __ldcs(p); }

// k_filter works in two phases per warp, because under divergence every lane of a warp pays for what any lane does and
// only ~1.6 of the 4 neighbourhood tables of an entry are worth a look:
//   A. lane per entry (32 entries per iteration): decode (re-derive a barcode with non-ACGT bytes from its record), four
//      bitmap words; the (entry, table) pairs that survive go into a ring queue in shared memory and the entry's mask
//      starts as 0 (or the direct result) in wl_cand;
//   B. lane per queued pair, ALL 32 lanes busy: rounds run only while the queue holds 32 pairs -- what is left over waits
//      for the pairs of the next iteration (and is drained after the last one).  One bucket of the pair's table per round;
//      a chain that continues in the next bucket goes back to the tail of the queue.  Candidates are OR-ed into the
//      entry's mask in wl_cand (a pair may be served one iteration after its entry was read).
// Before the carry-over, an iteration ran 2.85 rounds at 18 active lanes (32 + 19 pairs, then the few continued chains);
// now it runs 1.65 full ones (764 M -> 651 M warp instructions per 21.5 M entries).  The kernel is bound by memory latency
// (ncu: long-scoreboard stalls, issue slots 55 % busy), hence the loads of the next iteration's entries ahead of time and
// the L2 prefetch of the records its REPACK entries will need: 1.05 -> 0.99 ms.  Five CTAs per SM (48 registers); at six
// (40 registers) the spills cost more than the occupancy gives (1.59 ms).
#define PCL_FQ_CAP 256u      // ring: < 32 left over + <= 128 new pairs + <= 32 continued chains of a round
struct FilterQueue {
    uint32_t key[PCL_FQ_CAP], meta[PCL_FQ_CAP], bucket[PCL_FQ_CAP], widx[PCL_FQ_CAP];
    uint32_t tail;
};
// meta: q << 5 | len << 7 | has_n << 12 | n_p2 << 13 | j << 16 | continued << 18;  widx: low 32 bits of the worklist index
__global__ void __launch_bounds__(PCL_BLOCK, 5) k_filter(const __grid_constant__ CorrectParams p) {
    __shared__ FilterQueue fq[PCL_BLOCK / 32];
    FilterQueue &Q = fq[threadIdx.x >> 5];
    const unsigned long long n_work = *p.wl_count;
    const DRead &rd = p.rd;
    const bool rev = rd.is_reverse != 0;
    const uint32_t lane = threadIdx.x & 31u, lt = (1u << lane) - 1u;
    const uint64_t step = (uint64_t)gridDim.x * blockDim.x;
    uint32_t head = 0, total = 0;                                   // ring positions (monotonic; masked on access)
    if (lane == 0) Q.tail = 0;
    __syncwarp();
    uint64_t wb_last = 0;

    // one round: the pairs at ring positions [head, stop), one per lane
    auto round = [&](const uint32_t stop, const uint64_t wb_now) {
        const uint32_t pos = head + lane;
        if (pos < stop) {
            const uint32_t at = pos & (PCL_FQ_CAP - 1u);
            const uint32_t k = Q.key[at], m = Q.meta[at], w32 = Q.widx[at];
            const uint32_t q = (m >> 5) & 3u, len = (m >> 7) & 31u;
            const DTable &t = p.tab[(m >> 16) & 3u];
            const uint32_t b = (m >> 18) & 1u ? Q.bucket[at] : pcl_hash32_bucket(k & ~(0xFFu << (8u * q)), t.n_buckets);
            bool full;
            const uint32_t sub = pcl_filter_bucket(t.nkeys[q], b, k, q, (uint32_t)t.empty, (m >> 12) & 1u, (m >> 13) & 7u, &full);
            if (sub) {
                uint64_t w = (wb_now & 0xFFFFFFFF00000000ull) | w32;      // the entry is at most a few iterations behind wb_now
                if (w > wb_now + 31u) w -= 1ull << 32;
                atomicOr(p.wl_cand + w, pcl_filter_place(sub, len, q));
            }
            if (full) {                                                  // rare: the chain continues in the next bucket
                const uint32_t nxt = atomicAdd(&Q.tail, 1u) & (PCL_FQ_CAP - 1u);
                Q.key[nxt] = k; Q.meta[nxt] = m | (1u << 18); Q.widx[nxt] = w32;
                Q.bucket[nxt] = (b + 1u == t.n_buckets) ? 0u : b + 1u;
            }
        }
        __syncwarp();
        head = stop;
        total = Q.tail;
    };

    // wb is warp-uniform: every lane runs the same number of iterations (ballots and __syncwarp below)
    // The entries of the NEXT iteration are loaded while this one is processed (their DRAM latency would otherwise head
    // every iteration), and the records of its REPACK entries are pulled into L2 before the rounds start.
    const uint64_t wb0 = (uint64_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31u);
    unsigned long long ent_n = PCL_WL_HOLE;
    uint32_t key_n = 0;
    if (wb0 + lane < n_work) { ent_n = pcl_ld_stream64(p.worklist + wb0 + lane); key_n = __ldcs(p.wl_key + wb0 + lane); }
    for (uint64_t wb = wb0; wb < n_work; wb += step) {
        wb_last = wb;
        const uint64_t w = wb + lane;
        const bool in = w < n_work;
        const unsigned long long ent = ent_n;
        uint32_t key32 = key_n;
        ent_n = PCL_WL_HOLE; key_n = 0;
        if (w + step < n_work) { ent_n = pcl_ld_stream64(p.worklist + w + step); key_n = __ldcs(p.wl_key + w + step); }
        unsigned long long pend = 0;
        uint32_t live = 0, meta = 0;
        if (ent != PCL_WL_HOLE) {
            const uint32_t len = PCL_WL_LEN(ent), j = PCL_WL_J(ent);
            uint32_t ninv = PCL_WL_NINV(ent), ipos = PCL_WL_IPOS(ent);
            if (PCL_WL_IS_REPACK(ent)) {
                // a FastLayout barcode with non-ACGT bytes: which bases, and the key with those bases read as 0
                const uint64_t rec_i = PCL_WL_REC(ent);
                const uint8_t *rec = p.seq + rec_i * (uint64_t)rd.stride;
                uint32_t k2;
                if (rd.fast.spec != PCL_FAST_SPEC_NONE) {                  // 16 bases in four aligned words (at byte 0 or 8)
                    const uint2 lo = __ldg(reinterpret_cast<const uint2 *>(rec + rd.fast.bc_start));
                    const uint2 hi = __ldg(reinterpret_cast<const uint2 *>(rec + rd.fast.bc_start) + 1);
                    k2 = pcl_repack16(lo.x, lo.y, hi.x, hi.y, rev, &ninv, &ipos);
                } else {
                    const uint4 *r4 = reinterpret_cast<const uint4 *>(rec);
                    const uint4 a = __ldg(r4), b = rd.stride == 32u ? __ldg(r4 + 1) : make_uint4(0u, 0u, 0u, 0u);
                    const uint32_t wds[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
                    k2 = pcl_fast_repack(rd.fast, wds, rev, &ninv, &ipos);
                }
                if (k2 != key32) {                                  // an invalid byte other than 'N' left a non-zero code behind
                    key32 = k2;
                    p.wl_key[w] = k2;
                    reinterpret_cast<uint32_t *>(p.bc_key[0])[rec_i] = k2;
                }
            }
            const DTable &t = p.tab[j];
            if (!pcl_filter_hopeless(t, len, ninv)) {
                if (!t.nkeys[0]) pend = pcl_filter_direct(t, key32, len, ninv, ipos);
                else {
#pragma unroll
                    for (uint32_t q = 0; q < 4u; ++q) live |= (uint32_t)pcl_filter_live(t, key32, len, ninv, ipos, q) << q;   // four independent loads
                    const uint32_t n_p2 = (2u * (len - 1u - (ninv == 1u ? ipos : 0u))) & 7u;
                    meta = (len << 7) | ((ninv == 1u ? 1u : 0u) << 12) | (n_p2 << 13) | (j << 16);
                }
            }
        }
        if (in) p.wl_cand[w] = pend;                                 // the rounds OR the candidates in (same warp: ordered by __syncwarp)
        // ---- queue the live (entry, table) pairs, table-major ----
        uint32_t added = 0;
#pragma unroll
        for (uint32_t q = 0; q < 4u; ++q) {
            const uint32_t bal = __ballot_sync(0xFFFFFFFFu, (live >> q) & 1u);
            if ((live >> q) & 1u) {
                const uint32_t at = (total + added + __popc(bal & lt)) & (PCL_FQ_CAP - 1u);
                Q.key[at] = key32; Q.meta[at] = meta | (q << 5); Q.widx[at] = (uint32_t)w;
            }
            added += __popc(bal);
        }
        total += added;
        if (lane == 0) Q.tail = total;
        __syncwarp();
        if (ent_n != PCL_WL_HOLE && PCL_WL_IS_REPACK(ent_n))
            asm volatile("prefetch.global.L2 [%0];" ::"l"(p.seq + PCL_WL_REC(ent_n) * (uint64_t)rd.stride));
        // ---- full rounds only; the rest waits for the next iteration ----
        while (total - head >= 32u) round(head + 32u, wb);
    }
    while (total != head) round(total - head >= 32u ? head + 32u : total, wb_last);       // drain
}

__global__ void __launch_bounds__(PCL_BLOCK) k_resolve(const __grid_constant__ CorrectParams p) {
    const DRead &rd = p.rd;
    const unsigned long long n_work = *p.wl_count;
    const uint64_t step = (uint64_t)gridDim.x * blockDim.x;
    const bool rev = rd.is_reverse != 0;
    for (uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; w < n_work; w += step) {
        const unsigned long long cand = pcl_ld_stream64(p.wl_cand + w);
        if (!cand) continue;                                                      // holes and barcodes without a neighbour: NoMatch stays
        const unsigned long long ent = pcl_ld_stream64(p.worklist + w);
        const uint32_t key32 = __ldcs(p.wl_key + w);
        const uint64_t rec_i = PCL_WL_REC(ent);
        const uint32_t start = PCL_WL_START(ent), blen = PCL_WL_LEN(ent), j = PCL_WL_J(ent);
        const uint8_t *ql = p.qual + rec_i * (uint64_t)rd.qstride - rd.qoff;     // record-relative indexing inside the window
        uint32_t out32 = 0;
        const uint32_t code = pcl_resolve32(p.tab[j], key32, blen, cand, ql, start, rev, p.threshold, p.lut_cap, &out32);
        if (code == PCL_BC_CORRECTED) reinterpret_cast<uint32_t *>(p.bc_key[j])[rec_i] = out32;
        if (code != PCL_BC_NO_MATCH) pcl_fstatus_or(p.fstatus, rec_i, code << (4 + 3 * j));
    }
}

// pcl_whitelist_load: the bitmaps of the partial keys (DTable::nbits) from the main key table
__global__ void __launch_bounds__(PCL_BLOCK) k_build_nbits(const uint32_t *__restrict__ keys, uint64_t n_slots, uint32_t empty,
                                                          uint32_t *b0, uint32_t *b1, uint32_t *b2, uint32_t *b3) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_slots; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t k = keys[i];
        if (k == empty) continue;
        uint32_t *bm[4] = {b0, b1, b2, b3};
#pragma unroll
        for (uint32_t q = 0; q < 4u; ++q) {
            const uint32_t idx = pcl_partial_index(k, q);
            atomicOr(bm[q] + (idx >> 5), 1u << (idx & 31u));
        }
    }
}

// ------------------------------------------------------------------------------------------------
// k_qc: QcFastq::update (precellar/src/qc.rs:30-69) for one read file of a chunk: bases and Q30 bases of the
// barcode / UMI / target segments.  Categories: 0 umi, 1 barcode, 2 read1, 3 read2; per category
// {num_reads, num_total_bases, num_q30_bases}.  `**s - 33 >= 30` is evaluated on u8 and wraps in release
// builds of the reference, so bytes below 33 count as Q30 as well.
// ------------------------------------------------------------------------------------------------
struct QcParams {
    DRead rd;
    const uint8_t *qual;
    const uint16_t *len;
    const uint16_t *fstatus;
    const uint32_t *tcoord;
    const uint32_t *bcoord[PCL_MAX_BC_PER_READ];
    const uint32_t *ucoord[PCL_MAX_UMI_PER_READ];
    uint8_t bc_elem[PCL_MAX_BC_PER_READ];
    uint8_t umi_elem[PCL_MAX_UMI_PER_READ];
    uint16_t static_start[PCL_MAX_ELEMS];
    uint64_t n;
    unsigned long long *cat;      // [12]
};

__device__ __forceinline__ uint32_t pcl_q30_range(const uint8_t *row, uint32_t start, uint32_t len) {
    // the row is 16-byte aligned; walk the 32-bit words that overlap [start, start+len)
    const uint32_t *w = reinterpret_cast<const uint32_t *>(row);
    const uint32_t hi = start + len;
    uint32_t cnt = 0;
    for (uint32_t k = start >> 2; (k << 2) < hi; ++k) {
        const uint32_t b = k << 2;
        const uint32_t c0 = start > b ? start - b : 0u, c1 = b + 4u > hi ? b + 4u - hi : 0u;
        uint32_t mask = 0xFFFFFFFFu;
        if (c0) mask &= 0xFFFFFFFFu << (8u * c0);
        if (c1) mask &= 0xFFFFFFFFu >> (8u * c1);
        const uint32_t v = __ldg(w + k);
        const uint32_t q30 = __vcmpgeu4(v, 0x3F3F3F3Fu) | __vcmpltu4(v, 0x21212121u);   // q >= 63 or (wrapping) q < 33
        cnt += __popc(q30 & mask) >> 3;
    }
    return cnt;
}

// coordinates of element e of a fixed layout for a record of (truncated) length L, or PCL_COORD_ABSENT
__device__ __forceinline__ uint32_t pcl_static_coord(const DRead &rd, const uint16_t *static_start, uint32_t e, uint32_t L) {
    const uint32_t s = static_start[e], mn = rd.elems[e].min_len, mx = rd.elems[e].max_len;
    if (s >= L || s + mn > L) return PCL_COORD_ABSENT;
    const uint32_t hi = s + mx < L ? s + mx : L;
    return (s << 16) | (hi - s);
}

__global__ void __launch_bounds__(PCL_BLOCK) k_qc(const __grid_constant__ QcParams p) {
    const DRead &rd = p.rd;
    unsigned long long acc[6] = {0, 0, 0, 0, 0, 0};      // umi bases,q30 ; barcode bases,q30 ; target bases,q30
    unsigned long long n_target = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t fs = p.fstatus[i];
        if ((fs & 3u) != PCL_SPLIT_OK) continue;          // defect records contribute no segment (fastq.rs:369-374)
        uint32_t L = p.len[i];
        if (rd.max_len && L > rd.max_len) L = rd.max_len;
        const uint8_t *row = p.qual + i * (uint64_t)rd.qstride;               // QC mode: qoff == 0, whole record
        for (uint32_t j = 0; j < rd.n_bc; ++j) {
            if (PCL_BC_CODE(fs, j) == PCL_BC_ABSENT) continue;
            const uint32_t c = rd.fixed_layout ? pcl_static_coord(rd, p.static_start, p.bc_elem[j], L) : p.bcoord[j][i];
            acc[2] += c & 0xFFFFu;
            acc[3] += pcl_q30_range(row, c >> 16, c & 0xFFFFu);
        }
        uint32_t uc = PCL_COORD_ABSENT;                   // the last UMI segment of a file wins (fastq.rs:502)
        for (uint32_t j = 0; j < rd.n_umi; ++j) {
            const uint32_t c = rd.fixed_layout ? pcl_static_coord(rd, p.static_start, p.umi_elem[j], L) : p.ucoord[j][i];
            if (c != PCL_COORD_ABSENT) uc = c;
        }
        if (uc != PCL_COORD_ABSENT) {
            acc[0] += uc & 0xFFFFu;
            acc[1] += pcl_q30_range(row, uc >> 16, uc & 0xFFFFu);
        }
        if (fs & PCL_FSTATUS_TARGET) {
            const uint32_t c = p.tcoord[i];
            ++n_target;
            acc[4] += c & 0xFFFFu;
            acc[5] += pcl_q30_range(row, c >> 16, c & 0xFFFFu);
        }
    }
    const uint32_t lane = threadIdx.x & 31u;
#pragma unroll
    for (int k = 0; k < 6; ++k)
        for (int o = 16; o > 0; o >>= 1) acc[k] += __shfl_xor_sync(0xFFFFFFFFu, acc[k], o);
    for (int o = 16; o > 0; o >>= 1) n_target += __shfl_xor_sync(0xFFFFFFFFu, n_target, o);
    if (lane == 0) {
        const int tcat = rd.is_reverse ? 3 : 2;           // read2 for neg-strand reads, read1 otherwise (fastq.rs:510-514)
        if (acc[0]) atomicAdd(&p.cat[0 * 3 + 1], acc[0]);
        if (acc[1]) atomicAdd(&p.cat[0 * 3 + 2], acc[1]);
        if (acc[2]) atomicAdd(&p.cat[1 * 3 + 1], acc[2]);
        if (acc[3]) atomicAdd(&p.cat[1 * 3 + 2], acc[3]);
        if (n_target) atomicAdd(&p.cat[tcat * 3 + 0], n_target);
        if (acc[4]) atomicAdd(&p.cat[tcat * 3 + 1], acc[4]);
        if (acc[5]) atomicAdd(&p.cat[tcat * 3 + 2], acc[5]);
    }
}

// pcl_chunk_set_uniform_len: every record of the read file has the same length (nothing travels over PCIe for the plane)
__global__ void __launch_bounds__(PCL_BLOCK) k_fill_len(uint16_t *__restrict__ len, uint64_t n, uint16_t v) {
    uint32_t *w = reinterpret_cast<uint32_t *>(len);
    const uint32_t vv = (uint32_t)v | ((uint32_t)v << 16);
    const uint64_t nw = (n + 1) >> 1;                               // the plane is allocated with slack
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nw; i += (uint64_t)gridDim.x * blockDim.x) w[i] = vv;
}

// pcl_chunk_uniform_results_async: do all records of a read file carry the same fstatus and insert coordinates?
// out[0] = 1 if so (0 otherwise), out[1] = fstatus of record 0, out[2] = its tcoord.  out[0] must be 1 on entry.
__global__ void __launch_bounds__(PCL_BLOCK) k_uniform_check(const uint16_t *__restrict__ fstatus, const uint32_t *__restrict__ tcoord, uint64_t n,
                                                            uint32_t *out) {
    const uint32_t f0 = fstatus[0], t0 = tcoord ? tcoord[0] : 0u;
    bool same = true;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        same = same && fstatus[i] == f0 && (!tcoord || tcoord[i] == t0);
    if (!__all_sync(0xFFFFFFFFu, same) && (threadIdx.x & 31u) == 0) atomicExch(out, 0u);
    if (blockIdx.x == 0 && threadIdx.x == 0) { out[1] = f0; out[2] = t0; }
}

// counts in whitelist order for pcl_counts_fetch
__global__ void k_gather_counts(const uint32_t *__restrict__ counts, const uint32_t *__restrict__ index_to_slot,
                                unsigned long long *__restrict__ out, uint64_t n) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        out[i] = counts[index_to_slot[i]];
}

// pcl_chunk_upload_compact: rows of `src_bytes` (a multiple of 4) -> rows of `dst_bytes` (the device stride), the tail
// filled with 'N' like the host reader does.  One 32-bit word per thread, coalesced on both sides.
__global__ void __launch_bounds__(PCL_BLOCK) k_repack_rows(const uint32_t *__restrict__ src, uint32_t src_words,
                                                           uint32_t *__restrict__ dst, uint32_t dst_words, uint64_t n_rows) {
    const uint64_t total = n_rows * dst_words;
    for (uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t row = g / dst_words;
        const uint32_t k = (uint32_t)(g - row * dst_words);
        dst[g] = k < src_words ? __ldcs(src + row * src_words + k) : 0x4E4E4E4Eu;
    }
}

#endif  // PCL_KERNELS_CUH
