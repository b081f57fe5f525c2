// emul.cpp -- CPU-side unit-test harness for the __host__ __device__ logic in
// precellar_b200/csrc/hd_core.h and host_build.h.
//
// TEST INFRASTRUCTURE ONLY.  It runs the per-record functions the CUDA kernels wrap (pcl_split,
// pcl_pack_key, pcl_probe, pcl_correct_one) record by record on the host so that their logic can be
// checked against the oracle in the CPU-only test tier.  It is not part of the product library and
// nothing in precellar_b200/ links or loads it.
#include <cfloat>
#include <cmath>
#include <string>
#include <vector>

#include "../../precellar_b200/csrc/host_build.h"
#include "../../precellar_b200/csrc/zhuff_core.h"
#include "../../precellar_b200/csrc/emit_core.h"
#include "../../precellar_b200/csrc/inflate_core.h"

extern "C" {

int emu_read_info(const pcl_read_desc *rd, pcl_read_info *info, uint32_t *stride, char *err, int errcap) {
    DRead d;
    int bcr[PCL_MAX_BC_PER_READ];
    std::string e;
    int rc = pcl_build_read(0, *rd, &d, info, bcr, &e);
    if (rc != PCL_OK) { snprintf(err, errcap, "%s", e.c_str()); return rc; }
    *stride = d.stride;
    return PCL_OK;
}

int emu_fast_enabled(const pcl_read_desc *rd) {
    DRead d;
    pcl_read_info info;
    int bcr[PCL_MAX_BC_PER_READ];
    std::string e;
    if (pcl_build_read(0, *rd, &d, &info, bcr, &e) != PCL_OK) return -1;
    return d.fast.enabled ? 1 : ((!d.needs_payload && d.fixed_layout) ? 2 : (d.plan.enabled ? 3 : 0));
}

struct emu_read_io {
    const uint8_t *seq, *qual;
    const uint16_t *len;
    uint16_t *fstatus;
    uint32_t *tcoord;
    void *bc_key[PCL_MAX_BC_PER_READ];
    void *umi_key[PCL_MAX_UMI_PER_READ];
    uint32_t *bcoord[PCL_MAX_BC_PER_READ];
    uint32_t *ucoord[PCL_MAX_UMI_PER_READ];
};

// counts_out[region]: n_entries values in whitelist order; scalars_out: total, mismatch per region;
// defects_out: per read.
int emu_run(const pcl_read_desc *reads, int n_reads, int n_regions, const uint8_t **wl_bytes, const uint64_t **wl_offs,
            const uint64_t *wl_n, emu_read_io *io, uint64_t n, int correct, double threshold, double max_ee, int use_fast,
            uint64_t **counts_out, uint64_t *scalars_out, uint64_t *defects_out, const uint64_t **counts_in,
            char *err, int errcap) {
    std::string e;
    std::vector<HostTable> ht(n_regions);
    std::vector<std::vector<uint32_t>> counts(n_regions);
    std::vector<unsigned long long> scal(2 * n_regions, 0);
    std::vector<DTable> dt(n_regions);
    std::vector<std::vector<uint32_t>> nbits(n_regions);
    for (int g = 0; g < n_regions; ++g) {
        int rc = pcl_build_table(wl_bytes[g], wl_offs[g], wl_n[g], &ht[g], &e);
        if (rc != PCL_OK) { snprintf(err, errcap, "%s", e.c_str()); return rc; }
        counts[g].assign(ht[g].n_slots, 0);
        dt[g].keys = ht[g].data();
        dt[g].counts = counts[g].data();
        dt[g].total = &scal[2 * g];
        dt[g].mismatch = &scal[2 * g + 1];
        dt[g].empty = ht[g].empty;
        dt[g].n_buckets = (uint32_t)ht[g].n_buckets;
        dt[g].mode = ht[g].mode;
        for (int q = 0; q < 4; ++q) dt[g].nkeys[q] = (use_fast == 2 || ht[g].mode == PCL_TAB_K64) ? nullptr : ht[g].ntab32[q].data();
        for (int q = 0; q < 4; ++q) dt[g].nbits[q] = nullptr;
        if (dt[g].nkeys[0] && use_fast != 3) {                 // mirrors k_build_nbits (use_fast == 3: neighbourhood tables without bitmaps)
            nbits[g].assign(4ull * PCL_NBITS_WORDS, 0u);
            for (uint64_t i = 0; i < ht[g].n_slots; ++i) {
                const uint32_t k = ht[g].tab32[i];
                if (k == (uint32_t)ht[g].empty) continue;
                for (uint32_t q = 0; q < 4; ++q) {
                    const uint32_t idx = pcl_partial_index(k, q);
                    nbits[g][(uint64_t)q * PCL_NBITS_WORDS + (idx >> 5)] |= 1u << (idx & 31u);
                }
            }
            for (int q = 0; q < 4; ++q) dt[g].nbits[q] = nbits[g].data() + (uint64_t)q * PCL_NBITS_WORDS;
        }
    }
    double lut[512];
    for (int q = 0; q < 256; ++q) {
        lut[q] = std::pow(10.0, -(((double)(q < 66 ? q : 66) - 33.0) / 10.0));
        lut[256 + q] = std::pow(10.0, -(((double)q - 33.0) / 10.0));
    }
    std::vector<DRead> dr(n_reads);
    std::vector<pcl_read_info> inf(n_reads);
    std::vector<std::vector<int>> bcr(n_reads, std::vector<int>(PCL_MAX_BC_PER_READ, -1));
    for (int r = 0; r < n_reads; ++r) {
        int rc = pcl_build_read(r, reads[r], &dr[r], &inf[r], bcr[r].data(), &e);
        if (rc != PCL_OK) { snprintf(err, errcap, "%s", e.c_str()); return rc; }
    }
    struct Work { unsigned long long ent; uint32_t key; };          // worklist entry (pcl_wl_pack) + 4-byte device key
    std::vector<std::vector<Work>> work(n_reads);
    // ---- pass 1 (mirrors k_count) ----
    for (int r = 0; r < n_reads; ++r) {
        const DRead &rd = dr[r];
        const pcl_read_info &in = inf[r];
        const bool rev = rd.is_reverse != 0;
        defects_out[r] = 0;
        const bool fast = use_fast && rd.fast.enabled && dt[bcr[r][0]].mode != PCL_TAB_K64;
        const bool lenonly = use_fast && !rd.needs_payload && rd.fixed_layout;
        if (fast) {                                   // mirrors k_count_fast
            const FastLayout &f = rd.fast;
            DTable &tab = dt[bcr[r][0]];
            const uint32_t bc_end = (uint32_t)f.bc_start + f.bc_len, umi_end = (uint32_t)f.umi_start + f.umi_len;
            const bool len_ok = tab.mode == PCL_TAB_K32_L16 ? f.bc_len == 16 : (tab.mode == PCL_TAB_K32_MARKER && f.bc_len <= 15);
            const uint32_t absent_rest = (PCL_BC_ABSENT << 7) | (PCL_BC_ABSENT << 10) | (PCL_BC_ABSENT << 13);
            for (uint64_t i = 0; i < n; ++i) {
                uint32_t L = io[r].len[i];
                if (rd.max_len && L > rd.max_len) L = rd.max_len;
                uint32_t w[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                memcpy(w, io[r].seq + i * (uint64_t)rd.stride, rd.stride);
                uint32_t bk, uk, bbad; bool ubad;
                pcl_fast_extract_spec(f, w, rev, &bk, &bbad, &uk, &ubad, false);         // the count kernels: no flag word
                {   // the specialised instances must agree with the mask-driven extraction, with and without the flag word
                    uint32_t bk2, uk2, bb2, bk3, uk3, bb3, bk4, uk4, bb4; bool ub2, ub3, ub4;
                    pcl_fast_extract(f, w, rev, &bk2, &bb2, &uk2, &ub2, false);
                    if (bk2 != bk || uk2 != uk || (bb2 != 0) != (bbad != 0) || ub2 != ubad) return PCL_ERR_INPUT;
                    pcl_fast_extract_spec(f, w, rev, &bk3, &bb3, &uk3, &ub3, true);
                    pcl_fast_extract(f, w, rev, &bk4, &bb4, &uk4, &ub4, true);
                    if (bk3 != bk4 || uk3 != uk4 || bb3 != bb4 || ub3 != ub4 || uk3 != uk || (bb3 != 0) != (bbad != 0)) return PCL_ERR_INPUT;
                }
                uint32_t fs = PCL_SPLIT_OK | absent_rest;
                if (f.has_tail_target) {
                    uint32_t tc = PCL_COORD_ABSENT;
                    if (f.t_start < L && (uint32_t)f.t_start + f.t_min <= L) {
                        const uint32_t hi = (uint32_t)f.t_start + f.t_max < L ? (uint32_t)f.t_start + f.t_max : L;
                        tc = ((uint32_t)f.t_start << 16) | (hi - f.t_start);
                        fs |= PCL_FSTATUS_TARGET;
                    }
                    io[r].tcoord[i] = tc;
                }
                uint32_t code = PCL_BC_ABSENT;
                if (bc_end <= L) {
                    ++*tab.total;
                    uint32_t slot = 0xFFFFFFFFu;
                    if (!bbad && len_ok) slot = pcl_probe32(tab, bk);
                    if (slot != 0xFFFFFFFFu) { code = PCL_BC_EXACT; tab.counts[slot]++; }
                    else {
                        code = PCL_BC_NO_MATCH; ++*tab.mismatch;
                        work[r].push_back({pcl_wl_pack(i, f.bc_start, f.bc_len, 0, bbad ? 1u : 0u, PCL_WL_REPACK_POS), bk});
                    }
                } else bk = 0;
                fs |= code << 4;
                ((uint32_t *)io[r].bc_key[0])[i] = bk;
                if (f.umi_len) ((uint32_t *)io[r].umi_key[0])[i] = umi_end <= L ? uk : PCL_KEY_ABSENT32;
                io[r].fstatus[i] = (uint16_t)fs;
            }
            continue;
        }
        if (lenonly) {                                // mirrors k_count_lenonly
            const uint32_t absent_all = (PCL_BC_ABSENT << 4) | (PCL_BC_ABSENT << 7) | (PCL_BC_ABSENT << 10) | (PCL_BC_ABSENT << 13);
            for (uint64_t i = 0; i < n; ++i) {
                uint32_t L = io[r].len[i];
                if (rd.max_len && L > rd.max_len) L = rd.max_len;
                uint32_t off = 0, tc = PCL_COORD_ABSENT, n_t = 0;
                for (int e = 0; e < rd.n_elems; ++e) {
                    const uint32_t mn = rd.elems[e].min_len, mx = rd.elems[e].max_len;
                    if (off >= L || off + mn > L) break;
                    const uint32_t hi = off + mx < L ? off + mx : L;
                    if (rd.elems[e].kind == PCL_KIND_TARGET) { tc = (off << 16) | (hi - off); ++n_t; }
                    off += mx;
                }
                uint32_t fs = absent_all | (n_t > 1u ? PCL_SPLIT_MULTI_TARGET : PCL_SPLIT_OK);
                if (n_t == 1u) fs |= PCL_FSTATUS_TARGET;
                if (rd.has_target) io[r].tcoord[i] = n_t == 1u ? tc : PCL_COORD_ABSENT;
                io[r].fstatus[i] = (uint16_t)fs;
            }
            continue;
        }
        for (uint64_t i = 0; i < n; ++i) {
            uint32_t L = io[r].len[i];
            if (rd.max_len && L > rd.max_len) L = rd.max_len;
            const uint8_t *rec = io[r].seq ? io[r].seq + i * (uint64_t)rd.stride : nullptr;
            bool plan = (use_fast == 1 || use_fast == 3) && rd.plan.enabled;                  // mirrors k_count<2>
            for (uint32_t j = 0; j < rd.n_bc; ++j) plan = plan && dt[bcr[r][j]].mode != PCL_TAB_K64;
            if (plan) {
                uint32_t w[16] = {0};
                memcpy(w, rec, rd.stride);
                PlanOut po, pg;
                memset(&po, 0, sizeof po); memset(&pg, 0, sizeof pg);
                if (rev) pcl_plan_record<true>(rd, w, L, pg); else pcl_plan_record<false>(rd, w, L, pg);
                const bool fastp = rev ? pcl_plan_record_fast<true>(rd, w, L, po) : pcl_plan_record_fast<false>(rd, w, L, po);
                if (fastp) {                                       // the straight-line route must agree with the walk
                    bool same = po.status == pg.status && po.tcoord == pg.tcoord;
                    for (uint32_t j = 0; j < rd.n_bc; ++j)
                        same = same && po.bcoord[j] == pg.bcoord[j] && po.bkey[j] == pg.bkey[j] && po.bninv[j] == pg.bninv[j] &&
                               (po.bninv[j] != 1u || po.bipos[j] == pg.bipos[j]);
                    for (uint32_t j = 0; j < rd.n_umi; ++j) same = same && po.ucoord[j] == pg.ucoord[j] && po.ukey[j] == pg.ukey[j];
                    if (!same) { snprintf(err, errcap, "plan fast path disagrees with the walk at record %llu", (unsigned long long)i); return PCL_ERR_INPUT; }
                } else po = pg;
                if (!rev && pcl_plan_geo_matches<PlanGeoSci3>(rd)) {      // ... and so must the compile-time geometry instance
                    PlanOut ps;
                    memset(&ps, 0, sizeof ps);
                    const bool fasts = pcl_plan_fast_g<false>(PlanGeoSci3(rd), w, L, ps);
                    bool same = fasts == fastp;
                    if (fasts && same) {
                        same = ps.status == pg.status && ps.tcoord == pg.tcoord;
                        for (uint32_t j = 0; j < rd.n_bc; ++j)
                            same = same && ps.bcoord[j] == pg.bcoord[j] && ps.bkey[j] == pg.bkey[j] && ps.bninv[j] == pg.bninv[j] &&
                                   (ps.bninv[j] != 1u || ps.bipos[j] == pg.bipos[j]);
                        for (uint32_t j = 0; j < rd.n_umi; ++j) same = same && ps.ucoord[j] == pg.ucoord[j] && ps.ukey[j] == pg.ukey[j];
                    }
                    if (!same) { snprintf(err, errcap, "PlanGeoSci3 disagrees with the walk at record %llu", (unsigned long long)i); return PCL_ERR_INPUT; }
                }
                uint32_t fs = po.status;
                if (po.status != PCL_SPLIT_OK) defects_out[r]++;
                if (rd.has_target) {
                    if (po.tcoord != PCL_COORD_ABSENT) fs |= PCL_FSTATUS_TARGET;
                    io[r].tcoord[i] = po.tcoord;
                }
                for (uint32_t j = 0; j < PCL_MAX_BC_PER_READ; ++j) {
                    if (j >= rd.n_bc) { fs |= PCL_BC_ABSENT << (4 + 3 * j); continue; }
                    const uint32_t c = po.bcoord[j];
                    io[r].bcoord[j][i] = c;
                    uint32_t code = PCL_BC_ABSENT, key = 0;
                    if (c != PCL_COORD_ABSENT) {
                        const uint32_t blen = c & 0xFFFFu, ninv = po.bninv[j], ipos = po.bipos[j];
                        key = po.bkey[j];
                        DTable &t = dt[bcr[r][j]];
                        uint32_t slot = 0xFFFFFFFFu;
                        if (ninv == 0u && (t.mode == PCL_TAB_K32_L16 ? blen == 16u : blen <= 15u)) slot = pcl_probe32(t, key);
                        ++*t.total;
                        if (slot != 0xFFFFFFFFu) { code = PCL_BC_EXACT; t.counts[slot]++; }
                        else {
                            code = PCL_BC_NO_MATCH; ++*t.mismatch;
                            if (ninv <= 1u) work[r].push_back({pcl_wl_pack(i, c >> 16, c & 0xFFFFu, j, ninv, ipos), key});
                        }
                    }
                    reinterpret_cast<uint32_t *>(io[r].bc_key[j])[i] = key;
                    fs |= code << (4 + 3 * j);
                }
                for (uint32_t j = 0; j < rd.n_umi; ++j) {
                    io[r].ucoord[j][i] = po.ucoord[j];
                    reinterpret_cast<uint32_t *>(io[r].umi_key[j])[i] = po.ukey[j];
                }
                io[r].fstatus[i] = (uint16_t)fs;
                continue;
            }
            SplitOut so;
            RecCodes rc;
            const bool codes = use_fast && rd.codes_ok;                    // mirrors k_count<true>
            if (codes) {
                uint32_t w[16] = {0};
                memcpy(w, rec, rd.stride);
                const uint32_t nbytes = L < rd.stride ? L : rd.stride;
                pcl_rec_codes(w, (int)((nbytes + 3u) >> 2), rev, rc);
                pcl_split_codes(rd, rc, L, so);
            } else {
                pcl_split(rd, rec, L, so);
            }
            uint32_t fs = so.status;
            const bool ok = so.status == PCL_SPLIT_OK;
            if (so.status == PCL_SPLIT_ANCHOR_NOT_FOUND || so.status == PCL_SPLIT_PATTERN_MISMATCH) defects_out[r]++;
            if (ok && so.tcoord != PCL_COORD_ABSENT) fs |= PCL_FSTATUS_TARGET;
            if (rd.has_target) io[r].tcoord[i] = ok ? so.tcoord : PCL_COORD_ABSENT;
            for (uint32_t j = 0; j < PCL_MAX_BC_PER_READ; ++j) {
                if (j >= rd.n_bc) { fs |= PCL_BC_ABSENT << (4 + 3 * j); continue; }
                const uint32_t c = ok ? so.seg[in.bc_elem[j]] : PCL_COORD_ABSENT;
                if (!rd.fixed_layout) io[r].bcoord[j][i] = c;
                uint32_t code = PCL_BC_ABSENT;
                uint64_t key = 0;
                if (c != PCL_COORD_ABSENT) {
                    uint32_t ninv, ipos;
                    key = codes ? pcl_pack_codes(rc, c >> 16, c & 0xFFFFu, rev, &ninv, &ipos) : pcl_pack_key(rec, c >> 16, c & 0xFFFFu, rev, &ninv, &ipos);
                    if (codes) {                                       // the code-space packer must agree with the byte packer
                        uint32_t n2, i2;
                        const uint64_t k2 = pcl_pack_key(rec, c >> 16, c & 0xFFFFu, rev, &n2, &i2);
                        if (k2 != key || n2 != ninv || (ninv == 1u && i2 != ipos)) return PCL_ERR_INPUT;
                    }
                    DTable &t = dt[bcr[r][j]];
                    uint32_t slot = 0xFFFFFFFFu;
                    if (ninv == 0) slot = pcl_probe(t, key, c & 0xFFFFu);
                    ++*t.total;
                    if (slot != 0xFFFFFFFFu) { code = PCL_BC_EXACT; t.counts[slot]++; }
                    else {
                        code = PCL_BC_NO_MATCH; ++*t.mismatch;
                        if (ninv <= 1u) work[r].push_back({pcl_wl_pack(i, c >> 16, c & 0xFFFFu, j, ninv, ipos), (uint32_t)key});
                    }
                }
                pcl_store_key(io[r].bc_key[j], in.bc_key_bytes[j], i, key);
                fs |= code << (4 + 3 * j);
            }
            for (uint32_t j = 0; j < rd.n_umi; ++j) {
                const uint32_t c = ok ? so.seg[in.umi_elem[j]] : PCL_COORD_ABSENT;
                if (!rd.fixed_layout) io[r].ucoord[j][i] = c;
                uint64_t key = PCL_KEY_ABSENT64;
                if (c != PCL_COORD_ABSENT) {
                    uint32_t ninv, ipos;
                    key = codes ? pcl_pack_codes(rc, c >> 16, c & 0xFFFFu, rev, &ninv) : pcl_pack_key(rec, c >> 16, c & 0xFFFFu, rev, &ninv, &ipos);
                    if (ninv) key = 0;
                }
                pcl_store_key(io[r].umi_key[j], in.umi_key_bytes[j], i, key);
            }
            io[r].fstatus[i] = (uint16_t)fs;
        }
    }
    for (int g = 0; g < n_regions; ++g) {
        for (uint64_t i = 0; i < ht[g].n_entries; ++i) counts_out[g][i] = counts[g][ht[g].index_to_slot[i]];
        scalars_out[2 * g] = scal[2 * g];
        scalars_out[2 * g + 1] = scal[2 * g + 1];
    }
    // multi-rank runs: pass 2 sees the all-reduced counts (the prior of likelihood1 is the GLOBAL count)
    if (counts_in)
        for (int g = 0; g < n_regions; ++g)
            if (counts_in[g])
                for (uint64_t i = 0; i < ht[g].n_entries; ++i) counts[g][ht[g].index_to_slot[i]] = (uint32_t)counts_in[g][i];
    // ---- pass 2 (mirrors k_enum_all + k_correct) ----
    if (correct) {
        const bool check_ee = max_ee < DBL_MAX;
        for (int r = 0; r < n_reads; ++r) {
            const DRead &rd = dr[r];
            const pcl_read_info &in = inf[r];
            if (!rd.n_bc) continue;
            const bool rev = rd.is_reverse != 0;
            if (check_ee) {
                work[r].clear();
                for (uint64_t i = 0; i < n; ++i) {
                    uint32_t fs = io[r].fstatus[i];
                    for (uint32_t j = 0; j < rd.n_bc; ++j) {
                        uint32_t code = PCL_BC_CODE(fs, j);
                        if (code != PCL_BC_EXACT && code != PCL_BC_NO_MATCH) continue;
                        uint32_t c;
                        if (rd.fixed_layout) {
                            uint32_t el = in.bc_elem[j];
                            uint32_t L = io[r].len[i];
                            if (rd.max_len && L > rd.max_len) L = rd.max_len;
                            uint32_t s = in.static_start[el], l = rd.elems[el].max_len;
                            if (rd.elems[el].min_len != rd.elems[el].max_len) l = (s + l < L ? s + l : L) - s;
                            c = (s << 16) | l;
                        } else c = io[r].bcoord[j][i];
                        work[r].push_back({pcl_wl_pack(i, c >> 16, c & 0xFFFFu, j, 0, 0), 0u});
                    }
                }
            }
            bool all32 = use_fast && !check_ee;
            for (uint32_t j = 0; j < rd.n_bc; ++j) all32 = all32 && dt[bcr[r][j]].mode != PCL_TAB_K64 && in.bc_key_bytes[j] == 4;
            for (const Work &wk : work[r]) {
                const uint64_t rec_i = PCL_WL_REC(wk.ent);
                const uint32_t w_start = PCL_WL_START(wk.ent), w_len = PCL_WL_LEN(wk.ent), w_j = PCL_WL_J(wk.ent);
                const uint8_t *rec = io[r].seq + rec_i * (uint64_t)rd.stride;
                const uint8_t *ql = io[r].qual + rec_i * (uint64_t)rd.qstride - rd.qoff;
                uint64_t out_key = 0;
                uint32_t code;
                if (all32) {                          // mirrors k_filter + k_resolve: key and N position come from the worklist
                    uint32_t out32 = 0, key32 = wk.key, ninv = PCL_WL_NINV(wk.ent), ipos = PCL_WL_IPOS(wk.ent);
                    if (PCL_WL_IS_REPACK(wk.ent)) {
                        uint32_t wds[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                        memcpy(wds, rec, rd.stride);
                        const uint32_t k2 = pcl_fast_repack(rd.fast, wds, rev, &ninv, &ipos);
                        if (k2 != key32) { key32 = k2; ((uint32_t *)io[r].bc_key[0])[rec_i] = k2; }
                    }
                    const uint64_t cand = pcl_filter32(dt[bcr[r][w_j]], key32, w_len, ninv, ipos);
                    code = cand ? pcl_resolve32(dt[bcr[r][w_j]], key32, w_len, cand, ql, w_start, rev, threshold, lut, &out32) : PCL_BC_NO_MATCH;
                    out_key = out32;
                } else {                              // mirrors k_correct
                    uint32_t ninv, ipos;
                    uint64_t key = pcl_pack_key(rec, w_start, w_len, rev, &ninv, &ipos);
                    bool is_exact = PCL_BC_CODE((uint32_t)io[r].fstatus[rec_i], w_j) == PCL_BC_EXACT;
                    code = pcl_correct_one(dt[bcr[r][w_j]], key, w_len, ninv, ipos, ql, w_start, rev, is_exact, threshold,
                                           max_ee, check_ee, lut, lut + 256, &out_key);
                }
                if (code == PCL_BC_CORRECTED) pcl_store_key(io[r].bc_key[w_j], in.bc_key_bytes[w_j], rec_i, out_key);
                if (code != PCL_BC_EXACT && code != PCL_BC_NO_MATCH) io[r].fstatus[rec_i] |= (uint16_t)(code << (4 + 3 * w_j));
            }
        }
    }
    return PCL_OK;
}


// ------------------------------------------------------------------------------------------
// zhuff_core.h on the host: the CTA of k_zh_blocks emulated phase by phase, thread by thread
// ------------------------------------------------------------------------------------------
}  // extern "C"
struct EmuExec {
    template <class F> void each(F f) { for (uint32_t t = 0; t < ZH_THREADS; ++t) f(t); }
};
extern "C" {

// src[0..n) -> one zstd frame (header + blocks) in dst; returns its size or -1 when dst is too small
// `shift` (0..63): the input is compressed from a copy that starts `shift` bytes behind a 64-byte boundary, so that both
// the 128-bit-load route (shift 0) and the byte route of the span loops are reachable from the tests
int64_t emu_zh_frame(const uint8_t *src_in, uint64_t n, uint8_t *dst, uint64_t cap, uint64_t *n_raw_blocks, uint32_t shift, int use_matches) {
    std::vector<uint8_t> copy(n + 192);
    uint8_t *src = copy.data() + ((64 - ((uintptr_t)copy.data() & 63)) & 63) + (shift & 63);
    if (n) memcpy(src, src_in, n);
    const uint64_t n_blocks = n ? (n + ZH_BLOCK_BYTES - 1) / ZH_BLOCK_BYTES : 1;
    if (cap < ZH_FRAME_HEADER_BYTES + n_blocks * (uint64_t)ZH_SLOT_BYTES) return -1;
    zh_put_frame_header(dst, n);
    uint64_t pos = ZH_FRAME_HEADER_BYTES, raw = 0;
    std::vector<uint8_t> slot(ZH_SLOT_BYTES);
    static ZhShared sh;
    static ZhScratch scratch;
    EmuExec ex;
    for (uint64_t b = 0; b < n_blocks; ++b) {
        const uint64_t lo = b * ZH_BLOCK_BYTES;
        const uint32_t len = (uint32_t)(n - lo < ZH_BLOCK_BYTES ? n - lo : ZH_BLOCK_BYTES);
        const uint32_t used = zh_compress_block(ex, sh, use_matches ? &scratch : nullptr, src + lo, len, b + 1 == n_blocks, slot.data());
        raw += ((slot[0] >> 1) & 3u) == 0;
        memcpy(dst + pos, slot.data(), used);
        pos += used;
    }
    if (n_raw_blocks) *n_raw_blocks = raw;
    return (int64_t)pos;
}


// ------------------------------------------------------------------------------------------
// emit_core.h on the host: k_emit_plan (sizes) -> exclusive scan -> k_emit_fill (bytes), group by group
// ------------------------------------------------------------------------------------------
struct emu_emit_read {
    const uint8_t *text;
    const uint32_t *nl;
    const uint16_t *len;
    const uint16_t *fstatus;
    const uint32_t *tcoord;
    const void *bc_key[PCL_MAX_BC_PER_READ];
    const uint32_t *bcoord[PCL_MAX_BC_PER_READ];
    const uint32_t *ucoord[PCL_MAX_UMI_PER_READ];
};

// returns 0, a PCL_EMIT_* code (> 0, *bad_group set) or -1 when an output buffer is too small
int emu_emit(const pcl_read_desc *reads, int n_reads, const emu_emit_read *io, uint64_t n, int correct, int paired,
             uint8_t *o1, uint64_t cap1, uint64_t *u1, uint8_t *o2, uint64_t cap2, uint64_t *u2, uint64_t *n_written,
             uint64_t *bad_group) {
    EmitParams p;
    memset(&p, 0, sizeof p);
    p.n_reads = n_reads; p.correct = correct; p.paired = paired; p.n = n;
    for (int r = 0; r < n_reads; ++r) {
        DRead d;
        pcl_read_info inf;
        int bcr[PCL_MAX_BC_PER_READ];
        std::string e;
        if (pcl_build_read(0, reads[r], &d, &inf, bcr, &e) != PCL_OK) return -2;
        EmitRead &er = p.rd[r];
        pcl_emit_describe(er, reads[r], inf);
        er.text = io[r].text; er.nl = io[r].nl; er.len = io[r].len; er.fstatus = io[r].fstatus; er.tcoord = io[r].tcoord;
        for (int j = 0; j < PCL_MAX_BC_PER_READ; ++j) { er.bc_key[j] = io[r].bc_key[j]; er.bcoord[j] = io[r].bcoord[j]; }
        for (int j = 0; j < PCL_MAX_UMI_PER_READ; ++j) er.ucoord[j] = io[r].ucoord[j];
    }
    std::vector<uint32_t> s1(n, 0), s2(n, 0);
    uint64_t t1 = 0, t2 = 0, written = 0;
    for (uint64_t i = 0; i < n; ++i) {
        const EmitDecision d = pcl_emit_decide(p, i);
        if (d.err) { *bad_group = i; return (int)d.err; }
        if (!d.write) continue;
        EmitCount c1;
        pcl_emit_r1(p, i, d, c1);
        s1[i] = c1.n; t1 += c1.n;
        if (paired) { EmitCount c2; pcl_emit_r2(p, i, d, c2); s2[i] = c2.n; t2 += c2.n; }
        ++written;
    }
    if (t1 > cap1 || t2 > cap2) return -1;
    uint64_t a1 = 0, a2 = 0;
    for (uint64_t i = 0; i < n; ++i) {
        if (!s1[i]) continue;
        const EmitDecision d = pcl_emit_decide(p, i);
        EmitSerial w1{o1 + a1};
        pcl_emit_r1(p, i, d, w1);
        if ((uint64_t)(w1.dst - (o1 + a1)) != s1[i]) return -3;
        a1 += s1[i];
        if (paired) {
            EmitSerial w2{o2 + a2};
            pcl_emit_r2(p, i, d, w2);
            if ((uint64_t)(w2.dst - (o2 + a2)) != s2[i]) return -3;
            a2 += s2[i];
        }
    }
    *u1 = t1; *u2 = t2; *n_written = written;
    return 0;
}


// ------------------------------------------------------------------------------------------
// inflate_core.h on the host: the warp of k_bgzf_inflate emulated phase by phase, lane by lane
// ------------------------------------------------------------------------------------------
}  // extern "C"
struct EmuWarpExec {
    int reverse;                 // lanes in descending order: a phase whose lanes depend on one another shows up as a difference
    template <class F> void each(F f) {
        if (reverse) for (uint32_t l = INF_LANES; l-- > 0;) f(l);
        else for (uint32_t l = 0; l < INF_LANES; ++l) f(l);
    }
};
extern "C" {

// comp[0 .. comp_bytes + INF_PAD_BYTES) readable; members[n]; out[out_cap].  Returns 0, or (member index + 1) << 8 | error
// of the first member that fails.
int64_t emu_inflate(const uint8_t *comp, uint64_t comp_bytes, const InfMember *members, uint32_t n, uint8_t *out, uint64_t out_cap,
                    int reverse) {
    uint32_t tab[256];
    for (uint32_t t = 0; t < 256u; ++t) tab[t] = inf_crc_table_entry(t);
    std::vector<InfWarp> state(1);               // per call: the tests drive several emulated inflaters from different threads
    InfWarp &w = state[0];
    EmuWarpExec ex{reverse};
    for (uint32_t i = 0; i < n; ++i) {
        const InfMember &m = members[i];
        if ((uint64_t)m.src_off + m.clen > comp_bytes || (uint64_t)m.dst_off + m.isize > out_cap) return ((int64_t)(i + 1) << 8) | 0xFF;
        const uint32_t e = inf_member(ex, w, tab, comp, m, out + m.dst_off);
        if (e) return ((int64_t)(i + 1) << 8) | e;
    }
    return 0;
}

uint32_t emu_inflate_warp_bytes() { return (uint32_t)sizeof(InfWarp); }

}  // extern "C"

// ------------------------------------------------------------------------------------------
// fastinflate.h (the host decoder behind pcl_gzsrc for ordinary gzip files): one raw DEFLATE stream, the input arriving
// in pieces of `piece` bytes, the output taken `take` bytes at a time
// ------------------------------------------------------------------------------------------
#include "../../precellar_b200/csrc/fastinflate.h"
extern "C" {
// returns the number of bytes produced, or -1 - (error: 1 decode error, 2 output too small, 3 stalled); *consumed = bytes
// of the stream (so that the caller can check where it ended)
int64_t emu_fast_inflate(const uint8_t *src, uint64_t n, uint8_t *dst, uint64_t cap, uint64_t piece, uint64_t take, uint64_t *consumed) {
    std::vector<uint8_t> buf(n + 64, 0);
    if (n) memcpy(buf.data(), src, n);
    std::vector<pcl_fi::Inflater> state(1);
    pcl_fi::Inflater *fi = &state[0];
    const uint8_t *in = buf.data();
    uint64_t arrived = piece < n ? piece : n, produced = 0;
    for (int guard = 0; guard < (1 << 28); ++guard) {
        const pcl_fi::Status st = fi->run(&in, buf.data() + arrived, arrived == n);
        while (fi->pending()) {
            uint64_t k = fi->pending() < take ? fi->pending() : take;
            if (produced + k > cap) return -3;
            memcpy(dst + produced, fi->rd, k);
            fi->rd += k;
            produced += k;
        }
        if (fi->window_full()) fi->slide();
        if (st == pcl_fi::FI_END) { *consumed = (uint64_t)(in - buf.data()); return (int64_t)produced; }
        if (st == pcl_fi::FI_ERROR) return -2;
        if (st == pcl_fi::FI_NEED_INPUT) {
            if (arrived == n) return -4;
            arrived = arrived + piece < n ? arrived + piece : n;
        }
    }
    return -4;
}
}  // extern "C"
