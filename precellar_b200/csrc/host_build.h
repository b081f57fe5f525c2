// host_build.h -- host-side construction of the device tables (pure C++, no CUDA calls):
//   * pcl_build_read : pcl_read_desc -> DRead + pcl_read_info   (the "compiled region table")
//   * pcl_build_table: whitelist byte strings -> bucketed open-addressing key table
// Used by pcl.cu (which uploads the results) and by the CPU unit-test harness in tests/emul.
#ifndef PCL_HOST_BUILD_H
#define PCL_HOST_BUILD_H

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "hd_core.h"

inline uint32_t pcl_key_bytes_for(const pcl_elem_desc &e) {
    if (e.max_len <= 15) return 4;
    if (e.min_len == 16 && e.max_len == 16) return 4;
    return 8;
}

inline int pcl_build_fail(std::string *err, int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (err) *err = buf;
    return code;
}

// Derive the device-side description of one read.  bc_region[j] = whitelist region of barcode j.
inline int pcl_build_read(int r, const pcl_read_desc &rd, DRead *out, pcl_read_info *info_out, int *bc_region,
                          std::string *err, bool qc = false) {
    if (rd.n_elems < 1 || rd.n_elems > PCL_MAX_ELEMS)
        return pcl_build_fail(err, PCL_ERR_UNSUPPORTED, "read %d: %d elements outside 1..%d", r, rd.n_elems, PCL_MAX_ELEMS);
    DRead d;
    memset(&d, 0, sizeof d);
    pcl_read_info info;
    memset(&info, 0, sizeof info);
    d.n_elems = rd.n_elems;
    d.is_reverse = rd.is_reverse ? 1 : 0;
    d.max_len = rd.max_len;
    bool fixed = true, payload = false, anchors_codable = true;
    uint32_t reach = 0, start = 0;
    bool buffered = false;     // the reference's buffer of unresolved elements is non-empty when the walk reaches this element
    for (int e = 0; e < rd.n_elems; ++e) {
        const pcl_elem_desc &el = rd.elems[e];
        DElem &de = d.elems[e];
        if (el.min_len > el.max_len) return pcl_build_fail(err, PCL_ERR_ARG, "read %d elem %d: min_len > max_len", r, e);
        de.min_len = el.min_len; de.max_len = el.max_len; de.kind = el.kind;
        de.is_anchor = el.is_fixed_seq ? 1 : 0;
        de.bc_j = -1; de.umi_j = -1;
        const bool fixed_len = el.min_len == el.max_len;
        if (el.is_fixed_seq) {
            if (el.seq_len > PCL_MAX_ANCHOR)
                return pcl_build_fail(err, PCL_ERR_UNSUPPORTED, "read %d elem %d: fixed sequence longer than %d", r, e, PCL_MAX_ANCHOR);
            de.seq_len = el.seq_len;
            memcpy(de.seq, el.sequence, el.seq_len);
            uint32_t k = 0;                                   // `mis as f64 <= 0.2 * len as f64` (segment.rs:267)
            while ((double)(k + 1) <= 0.2 * (double)el.seq_len) ++k;
            de.max_mis = (uint8_t)k;
            // code-space form of the anchor (register-resident general path)
            bool pure = el.seq_len <= 32;
            uint64_t ac = 0, arc = 0;
            for (uint32_t t = 0; t < el.seq_len && pure; ++t) {
                const uint32_t c = pcl_base_code((uint8_t)el.sequence[t], false);
                if (c > 3u) { pure = false; break; }
                ac |= (uint64_t)c << (2u * t);
                arc |= (uint64_t)(3u - c) << (2u * (el.seq_len - 1u - t));
            }
            de.a_code = ac; de.a_rc = arc;
            if (!pure) anchors_codable = false;
        }
        // A fixed sequence is searched for only while the buffer holds a variable run (segment.rs:251-254); a match
        // empties the buffer, so the fixed sequences behind it are taken at face value and their bytes never read.
        const bool may_anchor = buffered && el.is_fixed_seq && fixed_len;
        if (may_anchor) buffered = false; else if (!fixed_len) buffered = true;
        if (may_anchor) {
            payload = true;
            if (el.seq_len != el.max_len)
                return pcl_build_fail(err, PCL_ERR_UNSUPPORTED, "read %d elem %d: anchor sequence length %d != max_len %d", r, e, el.seq_len, el.max_len);
        }
        if (!fixed_len && e != rd.n_elems - 1) fixed = false;
        info.static_start[e] = (uint16_t)start;
        start += el.max_len;
        if (el.kind == PCL_KIND_BARCODE) {
            if (info.n_bc >= PCL_MAX_BC_PER_READ)
                return pcl_build_fail(err, PCL_ERR_UNSUPPORTED, "read %d: more than %d barcode elements", r, PCL_MAX_BC_PER_READ);
            if (el.region_slot < 0 || el.region_slot >= PCL_MAX_REGIONS)
                return pcl_build_fail(err, PCL_ERR_ARG, "read %d elem %d: bad region_slot", r, e);
            if (el.max_len > 255) return pcl_build_fail(err, PCL_ERR_UNSUPPORTED, "read %d elem %d: barcode longer than 255", r, e);
            de.bc_j = (int8_t)info.n_bc;
            bc_region[info.n_bc] = el.region_slot;
            info.bc_elem[info.n_bc] = (uint8_t)e;
            info.bc_key_bytes[info.n_bc] = (uint8_t)pcl_key_bytes_for(el);
            info.n_bc++;
            payload = true;
        } else if (el.kind == PCL_KIND_UMI) {
            if (info.n_umi >= PCL_MAX_UMI_PER_READ)
                return pcl_build_fail(err, PCL_ERR_UNSUPPORTED, "read %d: more than %d UMI elements", r, PCL_MAX_UMI_PER_READ);
            de.umi_j = (int8_t)info.n_umi;
            info.umi_elem[info.n_umi] = (uint8_t)e;
            info.umi_key_bytes[info.n_umi] = (uint8_t)(el.max_len <= 15 ? 4 : 8);
            info.n_umi++;
            payload = true;
        } else if (el.kind == PCL_KIND_TARGET) {
            info.has_target = 1;
        }
        if (el.kind == PCL_KIND_BARCODE || el.kind == PCL_KIND_UMI || may_anchor)
            if (start > reach) reach = start;                  // last byte split()/pack may touch
    }
    uint32_t need = payload ? reach : 0;
    if (rd.max_len && need > rd.max_len) need = rd.max_len;
    bool qual_only = false;
    if (qc && info.has_target) {                       // QcFastq reads every quality byte of the target segment
        if (rd.max_len == 0)
            return pcl_build_fail(err, PCL_ERR_UNSUPPORTED, "read %d: FASTQ QC of a read with a target needs Read.max_len > 0", r);
        if (rd.max_len > need) need = rd.max_len;
        qual_only = !payload;
    }
    uint32_t stride = need ? ((need + 15u) / 16u) * 16u : 0;
    if (stride > PCL_MAX_STRIDE)
        return pcl_build_fail(err, PCL_ERR_UNSUPPORTED, "read %d: %u payload bytes per record exceed %d", r, stride, PCL_MAX_STRIDE);
    if (start > 0xFFFF) return pcl_build_fail(err, PCL_ERR_UNSUPPORTED, "read %d: layout longer than 65535", r);
    d.stride = stride;
    d.payload_bytes = (uint16_t)(((need + 3u) / 4u) * 4u);
    // quality window: only barcode bases are read by the correction pass
    d.qoff = 0; d.qstride = stride;
    if (!qc && fixed && info.n_bc >= 1) {
        uint32_t lo = 0xFFFFFFFFu, hi = 0;
        for (int j = 0; j < info.n_bc; ++j) {
            const uint32_t s0 = info.static_start[info.bc_elem[j]], e0 = s0 + rd.elems[info.bc_elem[j]].max_len;
            if (s0 < lo) lo = s0;
            if (e0 > hi) hi = e0;
        }
        if (hi > stride) hi = stride;
        if (lo < hi) { d.qoff = lo; d.qstride = ((hi - lo + 15u) / 16u) * 16u; }
    }
    if (!info.n_bc && !qc) d.qstride = 0;               // no barcode: the quality plane is never read
    info.qual_offset = (uint16_t)d.qoff;
    info.qual_stride = (uint16_t)d.qstride;
    d.qual_only = qual_only ? 1 : 0;
    d.codes_ok = (payload && stride >= 16 && stride <= 64 && anchors_codable) ? 1 : 0;
    info.qual_only = d.qual_only;
    d.fixed_layout = fixed ? 1 : 0;
    d.n_bc = info.n_bc; d.n_umi = info.n_umi; d.has_target = info.has_target; d.needs_payload = payload ? 1 : 0;
    info.fixed_layout = d.fixed_layout;
    info.needs_payload = d.needs_payload;
    // register-resident path: see FastLayout
    {
        FastLayout &f = d.fast;
        const int ne = rd.n_elems;
        const pcl_elem_desc &last = rd.elems[ne - 1];
        const bool tail_var = last.min_len != last.max_len;
        bool ok = fixed && payload && (stride == 16 || stride == 32) && info.n_bc == 1 && info.n_umi <= 1;
        if (ok && tail_var && last.kind != PCL_KIND_TARGET) ok = false;         // a trailing variable barcode/UMI/other: generic path
        if (ok) {
            const pcl_elem_desc &b = rd.elems[info.bc_elem[0]];
            ok = b.min_len == b.max_len && b.max_len >= 1 && b.max_len <= 16 && info.static_start[info.bc_elem[0]] + b.max_len <= 32;
            if (ok && info.n_umi) {
                const pcl_elem_desc &u = rd.elems[info.umi_elem[0]];
                ok = u.min_len == u.max_len && u.max_len >= 1 && u.max_len <= 15 && info.static_start[info.umi_elem[0]] + u.max_len <= 32;
            }
            // targets other than a trailing variable one, or more than one target, stay on the generic path
            int n_t = 0;
            for (int e = 0; e < ne; ++e) n_t += rd.elems[e].kind == PCL_KIND_TARGET;
            if (n_t > 1 || (n_t == 1 && !(tail_var && last.kind == PCL_KIND_TARGET))) ok = false;
        }
        if (ok) {
            f.enabled = 1;
            f.bc_start = (uint8_t)info.static_start[info.bc_elem[0]];
            f.bc_len = (uint8_t)rd.elems[info.bc_elem[0]].max_len;
            if (info.n_umi) {
                f.umi_start = (uint8_t)info.static_start[info.umi_elem[0]];
                f.umi_len = (uint8_t)rd.elems[info.umi_elem[0]].max_len;
            }
            for (uint32_t i = 0; i < 32; ++i) {
                if (i >= f.bc_start && i < (uint32_t)f.bc_start + f.bc_len) f.bc_mask[i / 4] |= 0xFFu << (8 * (i % 4));
                if (f.umi_len && i >= f.umi_start && i < (uint32_t)f.umi_start + f.umi_len) f.umi_mask[i / 4] |= 0xFFu << (8 * (i % 4));
            }
            if (f.bc_len == 16 && f.bc_start == 0 && f.umi_len == 12 && f.umi_start == 16) f.spec = PCL_FAST_SPEC_B0_U16x12;
            else if (f.bc_len == 16 && f.bc_start == 0 && f.umi_len == 0) f.spec = PCL_FAST_SPEC_B0;
            else if (f.bc_len == 16 && f.bc_start == 8 && f.umi_len == 0) f.spec = PCL_FAST_SPEC_B8;
            if (tail_var) {
                f.has_tail_target = 1;
                f.t_start = info.static_start[ne - 1];
                f.t_min = last.min_len;
                f.t_max = last.max_len;
            }
        }
    }
    // AnchorPlan (see hd_core.h): [fixed ...] [one variable] [anchor] [fixed ...] [optional trailing variable]
    {
        AnchorPlan &pl = d.plan;
        const int ne = rd.n_elems;
        int v = -1;
        for (int e = 0; e < ne; ++e) if (rd.elems[e].min_len != rd.elems[e].max_len) { v = e; break; }
        bool ok = !fixed && payload && d.codes_ok && !qc && v >= 0 && v + 1 < ne;
        if (ok) {
            const pcl_elem_desc &ve = rd.elems[v], &ae = rd.elems[v + 1];
            ok = ve.max_len - ve.min_len <= 15 && ae.min_len == ae.max_len && ae.is_fixed_seq && ae.seq_len == ae.max_len &&
                 ae.seq_len >= 1 && ae.seq_len <= 16;
        }
        bool has_t = false;
        for (int e = v + 2; ok && e < ne; ++e) {
            if (rd.elems[e].min_len != rd.elems[e].max_len) {
                if (e == ne - 1) has_t = true; else ok = false;
            }
        }
        int n_t = 0;
        for (int e = 0; ok && e < ne; ++e) {
            const pcl_elem_desc &el = rd.elems[e];
            n_t += el.kind == PCL_KIND_TARGET;
            if (el.kind == PCL_KIND_BARCODE && (el.max_len > 16 || (el.min_len != el.max_len && el.max_len > 15))) ok = false;
            if (el.kind == PCL_KIND_UMI && el.max_len > 15) ok = false;
        }
        for (int j = 0; ok && j < info.n_bc; ++j) ok = info.bc_key_bytes[j] == 4;
        for (int j = 0; ok && j < info.n_umi; ++j) ok = info.umi_key_bytes[j] == 4;
        if (ok && n_t > 1) ok = false;
        if (ok) {
            uint32_t off = 0;
            for (int e = 0; e < ne; ++e) {
                if (off > 255) { ok = false; break; }
                pl.ustart[e] = (uint8_t)off;
                off += e == v ? rd.elems[e].min_len : rd.elems[e].max_len;       // the trailing variable element is last: its length is not needed
            }
        }
        if (ok) {
            pl.enabled = 1; pl.v = (uint8_t)v; pl.has_t = has_t ? 1 : 0; pl.t_elem = 0xFF;
            for (int e = 0; e < ne; ++e) if (rd.elems[e].kind == PCL_KIND_TARGET) pl.t_elem = (uint8_t)e;
            // straight-line path: every element fits whatever the anchor position when
            //   L >= start of the last element (variable element at its longest) + its minimal length (a trailing variable
            //   element is clipped at the record end, the others have min == max)
            const uint32_t slack = rd.elems[v].max_len - rd.elems[v].min_len;
            pl.n_pos = (uint8_t)(slack + 1);
            pl.a_start = (uint8_t)(pl.ustart[v] + rd.elems[v].min_len);
            pl.a_len = rd.elems[v + 1].seq_len;
            pl.a_maxmis = d.elems[v + 1].max_mis;
            const uint32_t last_end = pl.ustart[ne - 1] + slack + (has_t ? (uint32_t)std::max<int>(rd.elems[ne - 1].min_len, 1) : rd.elems[ne - 1].max_len);
            // the walk breaks at an element when max_off >= len (segment.rs:245-248): for the trailing variable element that
            // needs L > its latest start; last_end covers it because its minimal length counts as >= 1
            pl.full_len = (uint16_t)last_end;
            uint32_t touch = 0;                                  // bytes the barcode / UMI / anchor fields can reach
            for (int e = 0; e < ne; ++e) {
                const pcl_elem_desc &el = rd.elems[e];
                if (el.kind == PCL_KIND_BARCODE || el.kind == PCL_KIND_UMI || e == v + 1 || e == v)
                    touch = std::max<uint32_t>(touch, pl.ustart[e] + slack + (e == v ? rd.elems[e].min_len : el.max_len));
            }
            touch = std::min<uint32_t>(touch, stride);
            pl.n_words = (uint8_t)std::min<uint32_t>(16, (touch + 3) / 4);
            if (pl.n_pos > 16 || last_end > 0xFFFF) pl.full_len = 0xFFFF;     // never taken: general route only
            for (int j = 0; j < info.n_bc; ++j) pl.bc_elem[j] = info.bc_elem[j];
            for (int j = 0; j < info.n_umi; ++j) pl.umi_elem[j] = info.umi_elem[j];
        }
    }
    info.route = !d.needs_payload ? PCL_ROUTE_LENONLY
                 : d.fast.enabled ? (d.fast.spec != PCL_FAST_SPEC_NONE ? PCL_ROUTE_FAST_SPEC : PCL_ROUTE_FAST)
                 : d.plan.enabled ? ((!d.is_reverse && pcl_plan_geo_matches<PlanGeoSci3>(d)) ? PCL_ROUTE_PLAN_GEO : PCL_ROUTE_PLAN)
                 : PCL_ROUTE_WALK;
    *out = d;
    *info_out = info;
    return PCL_OK;
}

// Whitelist byte strings -> marker-form 64-bit keys (file order, duplicates kept), in parallel.  Also tells whether every
// entry has 16 bases / at most 15 (the two 32-bit table modes).
inline int pcl_pack_whitelist(const uint8_t *bytes, const uint64_t *offsets, uint64_t n, std::vector<uint64_t> *keys, bool *all16,
                              bool *all_le15, std::string *err) {
    keys->resize(n);
    unsigned T = std::thread::hardware_concurrency();
    T = std::max(1u, std::min(T ? T : 1u, 16u));
    if (n < 65536) T = 1;
    std::vector<int> rc(T, PCL_OK);
    std::vector<uint64_t> bad_at(T, 0);
    std::vector<uint8_t> a16(T, 1), a15(T, 1);
    auto work = [&](unsigned t) {
        for (uint64_t i = n * t / T, e = n * (t + 1) / T; i < e; ++i) {
            const uint64_t a = offsets[i], b = offsets[i + 1];
            if (b < a) { rc[t] = PCL_ERR_ARG; bad_at[t] = i; return; }
            const uint64_t L = b - a;
            if (L < 1 || L > PCL_MAX_KEY_LEN) { rc[t] = PCL_ERR_UNSUPPORTED; bad_at[t] = i; return; }
            uint32_t ninv, ipos;
            const uint64_t key = pcl_pack_key(bytes + a, 0, (uint32_t)L, false, &ninv, &ipos);
            if (ninv) { rc[t] = PCL_ERR_UNSUPPORTED + 100; bad_at[t] = i; return; }
            (*keys)[i] = key;
            if (L != 16) a16[t] = 0;
            if (L > 15) a15[t] = 0;
        }
    };
    if (T == 1) work(0);
    else {
        std::vector<std::thread> th;
        for (unsigned t = 0; t < T; ++t) th.emplace_back(work, t);
        for (auto &x : th) x.join();
    }
    *all16 = true; *all_le15 = true;
    for (unsigned t = 0; t < T; ++t) {
        if (rc[t] == PCL_ERR_ARG) return pcl_build_fail(err, PCL_ERR_ARG, "whitelist offsets not monotone at %llu", (unsigned long long)bad_at[t]);
        if (rc[t] == PCL_ERR_UNSUPPORTED) return pcl_build_fail(err, PCL_ERR_UNSUPPORTED, "whitelist entry %llu has a length outside 1..%d",
                                                                 (unsigned long long)bad_at[t], PCL_MAX_KEY_LEN);
        if (rc[t] != PCL_OK) return pcl_build_fail(err, PCL_ERR_UNSUPPORTED, "whitelist entry %llu contains a byte outside upper-case ACGT",
                                                    (unsigned long long)bad_at[t]);
        *all16 = *all16 && a16[t];
        *all_le15 = *all_le15 && a15[t];
    }
    return PCL_OK;
}

#define PCL_TABLE_LOAD_PCT 35

struct HostTable {
    uint32_t mode = 0, slots_per_bucket = 8;
    uint64_t n_buckets = 0, n_slots = 0, n_entries = 0, empty = 0;
    std::vector<uint32_t> tab32;
    std::vector<uint64_t> tab64;
    std::vector<uint32_t> ntab32[4];         // neighbourhood tables (32-bit modes): bucketed by hash(key & ~(0xFF << 8q))
    std::vector<uint32_t> index_to_slot;     // whitelist index (dedup file order) -> slot
    const void *data() const { return mode == PCL_TAB_K64 ? (const void *)tab64.data() : (const void *)tab32.data(); }
    size_t key_bytes() const { return mode == PCL_TAB_K64 ? 8 : 4; }
};

// Build the whitelist table of one region (WhitelistBuilder::new semantics: a set of byte strings,
// file order, duplicates dropped -- precellar/src/barcode.rs:393-407, seqspec/src/region.rs:459-469).
// Deterministic: every rank builds the same table, so counts[] can be all-reduced slot-wise.
inline int pcl_build_table(const uint8_t *bytes, const uint64_t *offsets, uint64_t n, HostTable *out, std::string *err) {
    std::vector<uint64_t> keys;
    keys.reserve(n);
    // set of the packed keys seen so far (marker-form keys are never 0): open addressing, load <= 0.5
    uint64_t set_cap = 16;
    while (set_cap < 2 * n + 2) set_cap <<= 1;
    std::vector<uint64_t> seen(set_cap, 0ull);
    auto seen_slot = [&](uint64_t key) -> uint64_t {
        uint64_t h = key * 0x9E3779B97F4A7C15ull;
        h ^= h >> 29;
        uint64_t i = h & (set_cap - 1);
        while (seen[i] != 0ull && seen[i] != key) i = (i + 1) & (set_cap - 1);
        return i;
    };
    bool all16 = true, all_le15 = true;
    for (uint64_t i = 0; i < n; ++i) {
        uint64_t a = offsets[i], b = offsets[i + 1];
        if (b < a) return pcl_build_fail(err, PCL_ERR_ARG, "whitelist offsets not monotone at %llu", (unsigned long long)i);
        uint64_t L = b - a;
        if (L < 1 || L > PCL_MAX_KEY_LEN)
            return pcl_build_fail(err, PCL_ERR_UNSUPPORTED, "whitelist entry %llu has length %llu (supported: 1..%d)",
                                  (unsigned long long)i, (unsigned long long)L, PCL_MAX_KEY_LEN);
        uint32_t ninv, ipos;
        uint64_t key = pcl_pack_key(bytes + a, 0, (uint32_t)L, false, &ninv, &ipos);
        if (ninv)
            return pcl_build_fail(err, PCL_ERR_UNSUPPORTED, "whitelist entry %llu contains a byte outside upper-case ACGT", (unsigned long long)i);
        const uint64_t si = seen_slot(key);
        if (seen[si] == 0ull) {
            seen[si] = key;
            keys.push_back(key);
            all16 &= (L == 16);
            all_le15 &= (L <= 15);
        }
    }
    HostTable t;
    const uint64_t m = keys.size();
    t.n_entries = m;
    t.mode = all_le15 ? PCL_TAB_K32_MARKER : (all16 ? PCL_TAB_K32_L16 : PCL_TAB_K64);
    t.slots_per_bucket = t.mode == PCL_TAB_K64 ? 4 : 8;
    const uint32_t spb = t.slots_per_bucket;
    // Load factor 0.35: a probe (and above all the unsuccessful neighbourhood probes of pass 2) ends at the first
    // bucket with a free slot, so full buckets cost extra random sectors.  Measured on 125 M v3 pairs: count / correct
    // 2.11 / 1.81 ms at 0.5, 2.04 / 1.71 ms at 0.35, no further gain at 0.25, and 2.9 / 4.1 ms at 0.8.
    t.n_buckets = (100 * m / PCL_TABLE_LOAD_PCT + spb - 1) / spb;
    if (t.n_buckets < 2) t.n_buckets = 2;
    t.n_slots = t.n_buckets * spb;
    if (t.n_slots >= (1ull << 29))
        return pcl_build_fail(err, PCL_ERR_UNSUPPORTED, "whitelist too large (%llu entries)", (unsigned long long)m);
    // Empty marker: a value that is not a key.  Marker-form keys (<= 31 bases) never have all bits set;
    // for marker-less 16-mers walk down from 0xFFFFFFFF until a free value is found (at most m steps).
    t.empty = t.mode == PCL_TAB_K64 ? 0xFFFFFFFFFFFFFFFFull : 0xFFFFFFFFull;
    if (t.mode == PCL_TAB_K32_L16)
        while (seen[seen_slot((1ull << 32) | t.empty)] != 0ull) --t.empty;
    t.index_to_slot.resize(m);
    if (t.mode == PCL_TAB_K64) t.tab64.assign(t.n_slots, t.empty); else t.tab32.assign(t.n_slots, (uint32_t)t.empty);
    for (uint64_t i = 0; i < m; ++i) {
        const uint64_t k = t.mode == PCL_TAB_K64 ? keys[i] : (keys[i] & 0xFFFFFFFFull);
        uint32_t b = t.mode == PCL_TAB_K64 ? pcl_hash_bucket(k, (uint32_t)t.n_buckets) : pcl_hash32_bucket((uint32_t)k, (uint32_t)t.n_buckets);
        for (;;) {
            bool placed = false;
            for (uint32_t s = 0; s < spb && !placed; ++s) {
                const uint64_t idx = (uint64_t)b * spb + s;
                if (t.mode == PCL_TAB_K64) {
                    if (t.tab64[idx] == t.empty) { t.tab64[idx] = k; placed = true; }
                } else {
                    if (t.tab32[idx] == (uint32_t)t.empty) { t.tab32[idx] = (uint32_t)k; placed = true; }
                }
                if (placed) t.index_to_slot[i] = (uint32_t)idx;
            }
            if (placed) break;
            b = (b + 1 == t.n_buckets) ? 0 : b + 1;
        }
    }
    if (t.mode != PCL_TAB_K64) {
        // the four neighbourhood tables are independent: one host thread each (same result as a serial build)
        auto build_q = [&](int q) {
            const uint32_t bm = 0xFFu << (8 * q);
            std::vector<uint32_t> &nt = t.ntab32[q];
            nt.assign(t.n_slots, (uint32_t)t.empty);
            for (uint64_t i = 0; i < m; ++i) {
                const uint32_t k = (uint32_t)keys[i];
                uint32_t b = pcl_hash32_bucket(k & ~bm, (uint32_t)t.n_buckets);
                for (;;) {
                    bool placed = false;
                    for (uint32_t s = 0; s < spb && !placed; ++s) {
                        const uint64_t idx = (uint64_t)b * spb + s;
                        if (nt[idx] == (uint32_t)t.empty) { nt[idx] = k; placed = true; }
                    }
                    if (placed) break;
                    b = (b + 1 == t.n_buckets) ? 0 : b + 1;
                }
            }
        };
        if (m > 100000) {
            std::thread th[3];
            for (int q = 1; q < 4; ++q) th[q - 1] = std::thread(build_q, q);
            build_q(0);
            for (auto &x : th) x.join();
        } else {
            for (int q = 0; q < 4; ++q) build_q(q);
        }
    }
    *out = std::move(t);
    return PCL_OK;
}

#endif  // PCL_HOST_BUILD_H
