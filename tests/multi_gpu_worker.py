"""Worker of tests/test_gpu_multi.py: one process per GPU (torchrun), NCCL.  Every rank runs the hot path on its
contiguous shard (pcl_count -> pcl_counts_allreduce -> pcl_correct); rank 0 gathers the result planes, concatenates
them in rank order and compares everything with the single-process oracle on the whole input."""
import os
import pickle
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from tests import util as U                                   # noqa: E402
from tests.test_emul_parity import _layouts                   # noqa: E402
from precellar_b200 import _ffi                               # noqa: E402
from precellar_b200.dist import shard_bounds, exchange_unique_id   # noqa: E402
from precellar_b200.pipeline import BarcodePipeline, HostBatch     # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", rank=rank, world_size=world)

    def make_uid():
        import ctypes as C
        buf = C.create_string_buffer(128)
        _ffi.check(None, _ffi.lib().pcl_comm_unique_id(buf))
        return buf.raw
    report = {}
    for name, n in (("v3", 200_000), ("sci3", 60_000), ("atac", 100_000)):
        layout = _layouts(np.random.default_rng(20240607))[name]
        rng = np.random.default_rng(4242)                      # the same data on every rank
        uid = exchange_unique_id(make_uid, rank, world)
        pipe = BarcodePipeline(layout, device=local, nranks=world, rank=rank, nccl_uid=uid)
        infos, strides = U.emul_infos(layout)
        for a, b in zip(pipe.infos, infos):
            assert bytes(a) == bytes(b)
        multi = len(layout.region_ids) > 1
        full = U.gen_reads(rng, layout, n, strides, p_short=0.0 if multi else 0.02, p_tiny=0.02 if multi else 0.0, hot=200)
        lo, hi = shard_bounds(n, world)[rank]
        mine = [HostBatch(b.seq[lo:hi], b.qual[lo:hi], b.length[lo:hi]) for b in full]
        dev = U.narrow(mine, strides, pipe.infos)
        chunks = []
        for clo, chi in ((0, (hi - lo) // 3), ((hi - lo) // 3, hi - lo)):          # two chunks per rank
            ch = pipe.new_chunk(chi - clo)
            ch.reset(chi - clo)
            for r, b in enumerate(dev):
                ch.upload(r, HostBatch(None if b.seq is None else np.ascontiguousarray(b.seq[clo:chi]),
                                       None if b.qual is None else np.ascontiguousarray(b.qual[clo:chi]),
                                       np.ascontiguousarray(b.length[clo:chi])))
            pipe.count(ch)
            chunks.append(ch)
        pipe.allreduce()
        for ch in chunks:
            pipe.correct(ch, 0.9)
        planes = []
        for r in range(len(layout.reads)):
            ps = [ch.fetch(r) for ch in chunks]
            cat = lambda xs: None if xs[0] is None else np.concatenate(xs)
            planes.append(dict(fstatus=cat([p.fstatus for p in ps]), tcoord=cat([p.tcoord for p in ps]),
                               bc_key=[cat([p.bc_key[j] for p in ps]) for j in range(len(ps[0].bc_key))],
                               umi_key=[cat([p.umi_key[j] for p in ps]) for j in range(len(ps[0].umi_key))],
                               bcoord=[cat([p.bcoord[j] for p in ps]) for j in range(len(ps[0].bcoord))],
                               ucoord=[cat([p.ucoord[j] for p in ps]) for j in range(len(ps[0].ucoord))]))
        counts = {g: pipe.counts(g) for g in range(len(layout.region_ids))}       # global after the allreduce
        qc = pipe.qc()
        blob = pickle.dumps(dict(planes=planes, counts=counts, qc=qc))
        gathered = [None] * world
        dist.all_gather_object(gathered, blob)
        if rank == 0:
            parts = [pickle.loads(b) for b in gathered]
            for p in parts[1:]:                                                     # every rank holds the same global counts
                for g in counts:
                    assert np.array_equal(p["counts"][g][0], counts[g][0]) and p["counts"][g][1:] == counts[g][1:]
            results = []
            for r in range(len(layout.reads)):
                res = pipe.alloc_results(r, n)
                cat = lambda key: None if parts[0]["planes"][r][key] is None else np.concatenate([p["planes"][r][key] for p in parts])
                catl = lambda key, j: (None if parts[0]["planes"][r][key][j] is None
                                       else np.concatenate([p["planes"][r][key][j] for p in parts]))
                res.fstatus, res.tcoord = cat("fstatus"), cat("tcoord")
                res.bc_key = [catl("bc_key", j) for j in range(len(res.bc_key))]
                res.umi_key = [catl("umi_key", j) for j in range(len(res.umi_key))]
                res.bcoord = [catl("bcoord", j) for j in range(len(res.bcoord))]
                res.ucoord = [catl("ucoord", j) for j in range(len(res.ucoord))]
                results.append(res)
            defects = sum(p["qc"][:, 1] for p in parts)
            ores, ocounts = U.run_oracle(layout, full, threshold=0.9)
            U.compare_with_oracle(layout, pipe.infos, full, results, counts, defects, ores, ocounts)
            codes = (results[[i for i, p in enumerate(layout.reads) if p.has_barcode][0]].fstatus >> 4) & 7
            report[name] = dict(n=n, world=world, corrected=int((codes == _ffi.BC_CORRECTED).sum()),
                                exact=int((codes == _ffi.BC_EXACT).sum()))
            assert report[name]["corrected"] > 0 and report[name]["exact"] > 0
        pipe.close()
    if rank == 0:
        print("MULTI_GPU_PARITY_OK", report, flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
