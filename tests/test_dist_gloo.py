"""CPU tier, world_size 2 over gloo: the N>1 host logic -- contiguous shards, per-rank count pass, sum of the
counts, correction against the GLOBAL counts -- reproduces the single-process oracle on the whole input."""
import os
import pickle
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, pickle
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, os.environ["PCL_ROOT"])
from tests import util as U
from tests.test_emul_parity import _layouts
from precellar_b200.dist import shard_bounds, exchange_unique_id
from precellar_b200.pipeline import HostBatch

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
uid = exchange_unique_id(lambda: bytes(range(128)), rank, world)          # the id travels intact
assert uid == bytes(range(128))
out = {}
for name in ("v3", "sci3", "atac"):
    layout = _layouts(np.random.default_rng(20240607))[name]
    rng = np.random.default_rng(99)
    infos, strides = U.emul_infos(layout)
    n = 1001
    multi = len(layout.region_ids) > 1
    full = U.gen_reads(rng, layout, n, strides, p_short=0.0 if multi else 0.03, p_tiny=0.03 if multi else 0.0)
    lo, hi = shard_bounds(n, world)[rank]
    mine = [HostBatch(b.seq[lo:hi], b.qual[lo:hi], b.length[lo:hi]) for b in full]
    dev = U.narrow(mine, strides, infos)
    # pass 1 on the shard, then ONE sum over ranks of counts (+ mismatch/total), as pcl_counts_allreduce does
    _, _, counts, defects = U.run_emul(layout, dev, correct=False)
    gcounts = {}
    for g in counts:
        c = torch.from_numpy(counts[g][0].astype(np.int64)); dist.all_reduce(c)
        s = torch.tensor([counts[g][1], counts[g][2]], dtype=torch.int64); dist.all_reduce(s)
        gcounts[g] = (c.numpy().astype(np.uint64), int(s[0]), int(s[1]))
    d = torch.from_numpy(defects.astype(np.int64)); dist.all_reduce(d)
    infos, results, _, _ = U.run_emul(layout, dev, correct=True, global_counts={g: gcounts[g][0] for g in gcounts})
    out[name] = dict(lo=lo, hi=hi, results=[(r.fstatus, r.tcoord, r.bc_key, r.umi_key, r.bcoord, r.ucoord) for r in results],
                     gcounts=gcounts, defects=d.numpy().astype(np.uint64))
with open(os.environ["PCL_OUT"] + f".{rank}", "wb") as fh:
    pickle.dump(out, fh)
dist.destroy_process_group()
'''


def test_two_ranks_reproduce_the_single_process_oracle():
    from tests import util as U
    from tests.test_emul_parity import _layouts
    from precellar_b200.pipeline import ReadResults
    with tempfile.TemporaryDirectory() as td:
        script = os.path.join(td, "worker.py")
        open(script, "w").write(WORKER)
        outp = os.path.join(td, "out")
        procs = []
        for r in range(2):
            env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29533",
                       PCL_ROOT=ROOT, PCL_OUT=outp)
            procs.append(subprocess.Popen([sys.executable, script], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT))
        logs = [p.communicate(timeout=600)[0].decode() for p in procs]
        assert all(p.returncode == 0 for p in procs), "\n".join(logs)
        parts = [pickle.load(open(f"{outp}.{r}", "rb")) for r in range(2)]
    for name in ("v3", "sci3", "atac"):
        layout = _layouts(np.random.default_rng(20240607))[name]
        rng = np.random.default_rng(99)
        infos, strides = U.emul_infos(layout)
        multi = len(layout.region_ids) > 1
        full = U.gen_reads(rng, layout, 1001, strides, p_short=0.0 if multi else 0.03, p_tiny=0.03 if multi else 0.0)
        ores, ocounts = U.run_oracle(layout, full)
        assert parts[0][name]["lo"] == 0 and parts[0][name]["hi"] == parts[1][name]["lo"] and parts[1][name]["hi"] == 1001
        results = []
        for r, info in enumerate(infos):
            cat = lambda k, j=None: np.concatenate([(p[name]["results"][r][k] if j is None else p[name]["results"][r][k][j]) for p in parts])
            a = parts[0][name]["results"][r]
            results.append(ReadResults(info, cat(0), None if a[1] is None else cat(1),
                                       [cat(2, j) for j in range(len(a[2]))], [cat(3, j) for j in range(len(a[3]))],
                                       [None if a[4][j] is None else cat(4, j) for j in range(len(a[4]))],
                                       [None if a[5][j] is None else cat(5, j) for j in range(len(a[5]))]))
        # both ranks hold the same global counts, and they equal the oracle's single-process counts
        for g in parts[0][name]["gcounts"]:
            assert np.array_equal(parts[0][name]["gcounts"][g][0], parts[1][name]["gcounts"][g][0])
        U.compare_with_oracle(layout, infos, full, results, parts[0][name]["gcounts"], parts[0][name]["defects"], ores, ocounts)
