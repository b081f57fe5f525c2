"""Regenerates tests/golden/hotpath_cases.json.gz.

The reference is Rust and cannot be built in this image, so these are NOT outputs of the reference binary:
they are outputs of the oracle (oracle/oracle.cpp, itself pinned on the reference's own split vectors) frozen
at commit time, so that later changes to the oracle, the emulation harness or the CUDA path are all compared
against one fixed set of records.  Layouts are the BASELINE configurations (precellar_b200/configs.py).

    python tests/golden/make_golden.py
"""
import gzip
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from precellar_b200 import configs  # noqa: E402
from tests import util as U  # noqa: E402


def cases(rng):
    wl16 = U.gen_whitelist(rng, 400, 16)
    hp = U.gen_whitelist(rng, 96, [9, 10])
    rt = U.gen_whitelist(rng, 96, 10)
    return {
        "10x_rna_v3": (configs.v3_layout(wl16, 91), dict(wl=wl16)),
        "10x_atac": (configs.atac_layout(wl16, 50), dict(wl=wl16)),
        "10x_multiome_atac": (configs.multiome_atac_layout(wl16, 50), dict(wl=wl16)),
        "sci_rna_seq3": (configs.sci3_layout(hp, rt, 34, 100), dict(hp=hp, rt=rt)),
    }


def main():
    rng = np.random.default_rng(20261019)
    out = {}
    for name, (layout, wls) in cases(rng).items():
        infos, strides = U.emul_infos(layout)
        multi = len(layout.region_ids) > 1
        full = U.gen_reads(rng, layout, 300, strides, p_short=0.0 if multi else 0.04, p_tiny=0.03 if multi else 0.0,
                           read_len=[p.max_len for p in layout.reads])
        ores, ocounts = U.run_oracle(layout, full, correct=True, threshold=0.975)
        out[name] = dict(
            whitelists={k: [x.decode() for x in v] for k, v in wls.items()},
            reads=[dict(seq=[b.seq[i, :b.length[i]].tobytes().decode() for i in range(300)],
                        qual=[b.qual[i, :b.length[i]].tobytes().decode("latin-1") for i in range(300)]) for b in full],
            lines=ores.lines,
            counts={str(g): [int(x) for x in ocounts[g][0]] for g in ocounts},
            mismatch_total={str(g): [int(ocounts[g][1]), int(ocounts[g][2])] for g in ocounts},
            qc_cat=ores.qc_cat.astype(int).tolist(), qc_per_file=ores.qc_per_file.astype(int).tolist())
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "hotpath_cases.json.gz")
    with gzip.open(path, "wt") as fh:
        json.dump(out, fh)
    print(path, os.path.getsize(path))


if __name__ == "__main__":
    main()
