"""Field-by-field comparison of the CUDA path's outputs with the oracle's on the 10x v3 layout at sizes where
per-record Python work is not affordable (bench.py's parity block and the full-size GPU tests).
TEST / BENCH TOOLING."""
from __future__ import annotations

import numpy as np

from precellar_b200 import _ffi


def wl_keys_u32(wl: np.ndarray) -> np.ndarray:
    """[n,16] ASCII -> the 4-byte device keys (2 bits per base, first base most significant)."""
    code = ((wl >> 1) ^ (wl >> 2)) & 3
    k = np.zeros(len(wl), np.uint64)
    for p in range(wl.shape[1]):
        k = (k << np.uint64(2)) | code[:, p].astype(np.uint64)
    return k.astype(np.uint32)


def compare_v3(wl: np.ndarray, g1, g2, gcounts, ores, ocounts) -> dict:
    """g1/g2: ReadResults of rna-R1 / rna-R2; gcounts: (counts, mismatch, total); returns mismatch counts per field."""
    code = (g1.fstatus.astype(np.uint32) >> 4) & 7
    err = np.select([code == _ffi.BC_EXACT, code == _ffi.BC_CORRECTED, code == _ffi.BC_NO_MATCH,
                     code == _ffi.BC_LOW_CONF, code == _ffi.BC_EXCEED_EE], [0, 0, 1, 2, 3], 255).astype(np.uint8)
    keys = wl_keys_u32(wl)
    ok = ores.bc_index[0] >= 0
    out = {
        "split_status": int(((g1.fstatus & 3) != ores.fstatus[0]).sum() + ((g2.fstatus & 3) != ores.fstatus[1]).sum()),
        "barcode_status": int((err != ores.bc_err[0]).sum()),
        "assigned_set": int((((code == _ffi.BC_EXACT) | (code == _ffi.BC_CORRECTED)) != ok).sum()),
        "corrected_barcode": int((g1.bc_key[0][ok] != keys[ores.bc_index[0][ok]]).sum()),
        "insert_coords": int((g2.tcoord != ores.tcoord[1]).sum()),
        "whitelist_counts": int((np.asarray(gcounts[0], np.uint64) != ocounts[0]).sum()),
        "mismatch_total": int(gcounts[1] != ocounts[1]) + int(gcounts[2] != ocounts[2]),
    }
    # UMI: packed key of R1[16:28] (0 when it holds a non-ACGT byte) -- checked against the input bytes
    return out


def pack_keys(items) -> np.ndarray:
    """Whitelist entries (list[bytes] or [n,k] uint8) -> marker-form uint64 keys."""
    if isinstance(items, np.ndarray):
        code = ((items >> 1) ^ (items >> 2)) & 3
        k = np.ones(len(items), np.uint64)
        for p in range(items.shape[1]):
            k = (k << np.uint64(2)) | code[:, p].astype(np.uint64)
        return k
    lut = {65: 0, 67: 1, 71: 2, 84: 3}
    out = np.zeros(len(items), np.uint64)
    for i, s in enumerate(items):
        k = 1
        for b in s:
            k = (k << 2) | lut[b]
        out[i] = k
    return out


def compare_bulk(layout, infos, results, counts, ores, ocounts) -> dict:
    """Vectorised field-by-field comparison for any layout (no per-record Python work).
    Returns the number of mismatches per field; all zeros == bit-exact."""
    from precellar_b200.pipeline import widen_key
    out = {"split_status": 0, "insert_coords": 0, "barcode_status": 0, "assigned_set": 0, "corrected_barcode": 0,
           "segment_coords": 0, "whitelist_counts": 0, "mismatch_total": 0}
    slot = j_global = 0
    for r, (plan, info, res) in enumerate(zip(layout.reads, infos, results)):
        st = (res.fstatus & 3).astype(np.uint8)
        out["split_status"] += int((st != ores.fstatus[r]).sum())
        ok = st == 0
        if info.has_target:
            tc = np.where(ok, res.tcoord, np.uint32(_ffi.COORD_ABSENT))
            out["insert_coords"] += int((tc != ores.tcoord[r]).sum())
        jb = ju = 0
        for e, el in enumerate(plan.elems):
            if el.is_barcode():
                code = (res.fstatus.astype(np.uint32) >> (4 + 3 * jb)) & 7
                err = np.select([code == _ffi.BC_EXACT, code == _ffi.BC_CORRECTED, code == _ffi.BC_NO_MATCH,
                                 code == _ffi.BC_LOW_CONF, code == _ffi.BC_EXCEED_EE], [0, 0, 1, 2, 3], 255).astype(np.uint8)
                out["barcode_status"] += int((err != ores.bc_err[j_global]).sum())
                okb = ores.bc_index[j_global] >= 0
                out["assigned_set"] += int((((code == _ffi.BC_EXACT) | (code == _ffi.BC_CORRECTED)) != okb).sum())
                keys = pack_keys(layout.whitelists[layout.region_ids[plan.region_slots[e]]])
                mine = widen_key(res.bc_key[jb], el.min_len, el.max_len)
                out["corrected_barcode"] += int((mine[okb] != keys[ores.bc_index[j_global][okb]]).sum())
                if res.bcoord[jb] is not None:
                    out["segment_coords"] += int((res.bcoord[jb] != ores.bucoord[slot]).sum())
                jb += 1
                j_global += 1
                slot += 1
            elif el.is_umi():
                if res.ucoord[ju] is not None:
                    out["segment_coords"] += int((res.ucoord[ju] != ores.bucoord[slot]).sum())
                ju += 1
                slot += 1
    for g in range(len(layout.region_ids)):
        c, mm, tot = counts[g]
        oc, omm, otot = ocounts[g]
        out["whitelist_counts"] += int((np.asarray(c, np.uint64) != oc).sum())
        out["mismatch_total"] += int(mm != omm) + int(tot != otot)
    return out
