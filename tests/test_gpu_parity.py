"""GPU tier: the CUDA path through the C ABI against the oracle, bit-exact."""
import numpy as np
import pytest

from precellar_b200 import _ffi
from tests import util as U
from tests.test_emul_parity import NAMES, _layouts

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def layouts():
    return _layouts(np.random.default_rng(20240607))


def run_gpu(layout, dev_batches, correct=True, threshold=0.975, max_expected_errors=_ffi.F64_MAX, n_chunks=1):
    from precellar_b200.pipeline import BarcodePipeline, HostBatch
    pipe = BarcodePipeline(layout)
    try:
        n = len(dev_batches[0].length)
        bounds = np.linspace(0, n, n_chunks + 1).astype(int)
        chunks = []
        for c in range(n_chunks):
            lo, hi = bounds[c], bounds[c + 1]
            ch = pipe.new_chunk(max(hi - lo, 1))
            ch.reset(hi - lo)
            for r, b in enumerate(dev_batches):
                ch.upload(r, HostBatch(None if b.seq is None else np.ascontiguousarray(b.seq[lo:hi]),
                                       None if b.qual is None else np.ascontiguousarray(b.qual[lo:hi]),
                                       np.ascontiguousarray(b.length[lo:hi])))
            pipe.count(ch)
            chunks.append(ch)
        pipe.allreduce()
        if correct:
            for ch in chunks:
                pipe.correct(ch, threshold, max_expected_errors)
        parts = [[ch.fetch(r) for ch in chunks] for r in range(len(dev_batches))]
        results = []
        for r, ps in enumerate(parts):
            res = pipe.alloc_results(r, n)
            res.fstatus = np.concatenate([p.fstatus for p in ps])
            if res.tcoord is not None:
                res.tcoord = np.concatenate([p.tcoord for p in ps])
            res.bc_key = [np.concatenate([p.bc_key[j] for p in ps]) for j in range(len(res.bc_key))]
            res.umi_key = [np.concatenate([p.umi_key[j] for p in ps]) for j in range(len(res.umi_key))]
            res.bcoord = [None if res.bcoord[j] is None else np.concatenate([p.bcoord[j] for p in ps]) for j in range(len(res.bcoord))]
            res.ucoord = [None if res.ucoord[j] is None else np.concatenate([p.ucoord[j] for p in ps]) for j in range(len(res.ucoord))]
            results.append(res)
        counts = {g: pipe.counts(g) for g in range(len(layout.region_ids))}
        qc = pipe.qc()
        assert pipe.launch_count() > 0
        return pipe.infos, results, counts, qc[:, 1], qc[:, 0]
    finally:
        pipe.close()


@pytest.mark.parametrize("name", NAMES)
@pytest.mark.parametrize("mode", ["correct", "nocorrect", "ee", "lower"])
def test_gpu_matches_oracle(layouts, name, mode):
    layout = layouts[name]
    rng = np.random.default_rng(abs(hash((name, mode, "gpu"))) % (2 ** 31))
    infos, strides = U.emul_infos(layout)
    n = 3000
    multi = len(layout.region_ids) > 1
    full = U.gen_reads(rng, layout, n, strides, p_lower=0.02 if mode == "lower" else 0.0,
                       p_short=0.0 if multi else 0.03, p_tiny=0.03 if multi else 0.0)
    dev = U.narrow(full, strides)
    kw = dict(correct=mode != "nocorrect", threshold=0.975 if mode != "lower" else 0.6,
              max_expected_errors=1.5 if mode == "ee" else _ffi.F64_MAX)
    ginfos, results, counts, defects, nreads = run_gpu(layout, dev, n_chunks=3 if mode == "correct" else 1, **kw)
    for a, b in zip(ginfos, infos):
        assert bytes(a) == bytes(b)
    assert all(int(x) == n for x in nreads)
    ores, ocounts = U.run_oracle(layout, full, **kw)
    U.compare_with_oracle(layout, ginfos, full, results, counts, defects, ores, ocounts, correct=kw["correct"])


def test_empty_and_single_record(layouts):
    layout = layouts["v3"]
    infos, strides = U.emul_infos(layout)
    rng = np.random.default_rng(5)
    for n in (1, 31, 33):
        full = U.gen_reads(rng, layout, n, strides)
        ginfos, results, counts, defects, _ = run_gpu(layout, U.narrow(full, strides))
        ores, ocounts = U.run_oracle(layout, full)
        U.compare_with_oracle(layout, ginfos, full, results, counts, defects, ores, ocounts)


def test_unsupported_whitelists_are_rejected(layouts):
    from precellar_b200.pipeline import BarcodePipeline
    from tests.util import bc, make_layout
    for wl in ([b"ACGN"], [b"acgt"], [b"A" * 32]):
        layout = make_layout([dict(read_id="R1", elems=[bc("cb", len(wl[0]))], max_len=len(wl[0]))], {"cb": wl})
        with pytest.raises(_ffi.PclError) as ei:
            BarcodePipeline(layout)
        assert ei.value.code == _ffi.PCL_ERR_UNSUPPORTED
