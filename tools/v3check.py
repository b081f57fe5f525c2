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
