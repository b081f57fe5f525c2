"""CPU restatement of the reference's per-barcode grouping (TEST INFRASTRUCTURE ONLY; see oracle/pyoracle.py).

    sort_by_barcode            precellar/src/fragment/mod.rs:434-455   sort_by(|a, b| a.barcode.cmp(&b.barcode)): byte-string order
    chunk_by(barcode)          precellar/src/fragment/mod.rs:359-363
    remove_duplicates          precellar/src/fragment/dedups.rs:320-340
        HashMap<FingerPrint, Fragment>: entry(fp).and_modify(count += 1).or_insert(Fragment::from_alignment(first))
    FingerPrint                precellar/src/fragment/dedups.rs:52-66    (reference_id, 5' coordinates, orientation, umi)

The alignment part of the fingerprint is produced by the aligner (out of scope); it enters as one opaque integer per
read.  Parity unpinned: the reference has no test for this function.
"""
from __future__ import annotations

import itertools
from typing import Dict, List, Optional, Sequence, Tuple


def group_reference(barcodes: Sequence[Optional[bytes]], umis: Sequence[Optional[bytes]],
                    fingerprints: Optional[Sequence[int]] = None) -> List[Tuple[bytes, Dict[tuple, List[int]]]]:
    """barcodes[i]: the CB of read i (corrected barcode) or None when it has none; umis[i]: its UMI string or None.
    Returns [(barcode, {(fingerprint, umi): [first read index, count]})] in barcode order."""
    idx = [i for i in range(len(barcodes)) if barcodes[i] is not None]
    idx.sort(key=lambda i: barcodes[i])                      # stable, like a merge of sorted runs that keeps input order
    out = []
    for bc, grp in itertools.groupby(idx, key=lambda i: barcodes[i]):
        frags: Dict[tuple, List[int]] = {}
        for i in grp:
            fp = (int(fingerprints[i]) if fingerprints is not None else 0, umis[i])
            if fp in frags:
                frags[fp][1] += 1                            # and_modify(|x| x.count += 1)
            else:
                frags[fp] = [i, 1]                           # or_insert(Fragment::from_alignment(ali))
        out.append((bc, frags))
    return out
