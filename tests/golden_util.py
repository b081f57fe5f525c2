"""Loader of tests/golden/hotpath_cases.json.gz (see tests/golden/make_golden.py)."""
import gzip
import json
import os

import numpy as np

from precellar_b200 import configs
from precellar_b200.pipeline import HostBatch

PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hotpath_cases.json.gz")


def load():
    with gzip.open(PATH, "rt") as fh:
        raw = json.load(fh)
    out = {}
    for name, c in raw.items():
        wl = {k: [x.encode() for x in v] for k, v in c["whitelists"].items()}
        if name == "10x_rna_v3":
            layout = configs.v3_layout(wl["wl"], 91)
        elif name == "10x_atac":
            layout = configs.atac_layout(wl["wl"], 50)
        elif name == "10x_multiome_atac":
            layout = configs.multiome_atac_layout(wl["wl"], 50)
        else:
            layout = configs.sci3_layout(wl["hp"], wl["rt"], 34, 100)
        full = []
        for r in c["reads"]:
            n = len(r["seq"])
            w = max(64, max(len(s) for s in r["seq"]))
            seq = np.full((n, w), ord("N"), np.uint8)
            qual = np.full((n, w), ord("!"), np.uint8)
            ln = np.zeros(n, np.uint16)
            for i, (s, q) in enumerate(zip(r["seq"], r["qual"])):
                seq[i, :len(s)] = np.frombuffer(s.encode(), np.uint8)
                qual[i, :len(q)] = np.frombuffer(q.encode("latin-1"), np.uint8)
                ln[i] = len(s)
            full.append(HostBatch(seq, qual, ln))
        out[name] = dict(layout=layout, full=full, lines=c["lines"],
                         counts={int(g): (np.array(v, np.uint64), c["mismatch_total"][g][0], c["mismatch_total"][g][1])
                                 for g, v in c["counts"].items()},
                         qc_cat=np.array(c["qc_cat"], np.uint64), qc_per_file=np.array(c["qc_per_file"], np.uint64))
    return out
