// inflate.cuh -- k_bgzf_inflate: BGZF members inflated on the device, one warp per member (phases in inflate_core.h).
//
// SURVEY.md §8(f) row 1.  Replaces the host inflate of open_file (seqspec/src/utils.rs:44-56) for BGZF input.
// Grid: the warps of all CTAs take members m, m + W, m + 2 W ... (neighbouring warps work on neighbouring members, so the
// compressed bytes and the text they write stay close in memory).  Per warp 5 KB of shared memory: the two decoding
// tables, the code description the dynamic header leaves, the queue of matches.  The CRC-32 table is built per CTA.
#ifndef PCL_INFLATE_CUH
#define PCL_INFLATE_CUH

#include "inflate_core.h"

#define INF_WARPS_PER_CTA 4u

static_assert(sizeof(pcl_bgzf_member) == sizeof(InfMember), "pcl_bgzf_member is the host view of InfMember");

struct InfDevExec {
    // the leading barrier keeps a lane from starting a phase (and overwriting the warp's state) while another lane still
    // reads what the previous phase left
    template <class F> __device__ __forceinline__ void each(F f) { __syncwarp(); f(threadIdx.x & 31u); __syncwarp(); }
};

__global__ void __launch_bounds__(INF_WARPS_PER_CTA * 32u) k_bgzf_inflate(const uint8_t *comp, const InfMember *members, uint32_t n_members,
                                                                         uint8_t *text, unsigned long long *err) {
    __shared__ InfWarp s_warp[INF_WARPS_PER_CTA];
    __shared__ uint32_t s_crc[256];
    for (uint32_t t = threadIdx.x; t < 256u; t += blockDim.x) s_crc[t] = inf_crc_table_entry(t);
    __syncthreads();
    const uint32_t warp = threadIdx.x >> 5;
    InfWarp &w = s_warp[warp];
    InfDevExec ex;
    for (uint32_t m = blockIdx.x * INF_WARPS_PER_CTA + warp; m < n_members; m += gridDim.x * INF_WARPS_PER_CTA) {
        const InfMember mem = members[m];
        const uint32_t e = inf_member(ex, w, s_crc, comp, mem, text + mem.dst_off);
        if (e && (threadIdx.x & 31u) == 0) atomicMin(err, ((unsigned long long)m << 8) | e);
    }
}

#endif  // PCL_INFLATE_CUH
