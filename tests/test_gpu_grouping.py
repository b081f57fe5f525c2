"""GPU tier: per-barcode grouping (pcl_group_*: device radix sort + duplicate sets) against the restatement of the
reference's sort_by_barcode / chunk_by / remove_duplicates (oracle/grouping.py)."""
import numpy as np
import pytest

from oracle.grouping import group_reference
from precellar_b200 import _ffi
from tests import util as U
from tests.test_emul_parity import _layouts

pytestmark = pytest.mark.gpu


def _run(layout, full, fingerprint, correct=True, threshold=0.9):
    from precellar_b200.pipeline import BarcodePipeline
    infos, strides = U.emul_infos(layout)
    dev = U.narrow(full, strides, infos)
    pipe = BarcodePipeline(layout)
    try:
        n = len(dev[0].length)
        ch = pipe.new_chunk(max(n, 1))
        ch.reset(n)
        for r, b in enumerate(dev):
            ch.upload(r, b)
        pipe.count(ch)
        pipe.allreduce()
        if correct:
            pipe.correct(ch, threshold)
        g = pipe.group(ch, fingerprint=fingerprint)
        assert pipe.launch_count() > 0 or n == 0
    finally:
        pipe.close()
    return g


def _expected(layout, full, correct, threshold):
    """Barcode (CB) and UMI strings per read group from the oracle's annotate, as the aligner adapters would tag them."""
    ores, _ = U.run_oracle(layout, full, correct=correct, threshold=threshold)
    bcs, umis = [], []
    for ln in ores.lines:
        if not ln:
            bcs.append(None); umis.append(None)
            continue
        f = ln.split("\t")
        cb = f[2] if correct else f[0]                         # corrected barcode, or the raw one when nothing is corrected
        bcs.append(None if cb == "*" else cb.encode("latin-1"))
        umis.append(f[3].encode("latin-1") if f[3] else None)
    return bcs, umis


def _check(g, bcs, umis, fps, rev_umi=False):
    # read groups whose UMI holds a byte outside ACGTN are not grouped on the device: take them out of the expectation
    ok = lambda u: u is None or all(c in b"ACGTN" for c in u)
    skipped = [i for i in range(len(bcs)) if bcs[i] is not None and not ok(umis[i])]
    assert g.n_ungroupable == len(skipped)
    bcs = [b if ok(u) else None for b, u in zip(bcs, umis)]
    want = group_reference(bcs, umis, fps)
    assert g.n_assigned == sum(1 for b in bcs if b is not None) == len(g.order)
    assert len(g.bc_key) == len(want) and len(want) > 5
    assert sorted(g.order.tolist()) == [i for i in range(len(bcs)) if bcs[i] is not None]
    n_dup_sets = 0
    for b, (bc, frags) in enumerate(want):
        assert g.barcode(b) == bc
        f0, f1 = int(g.bc_frag_start[b]), int(g.bc_frag_start[b + 1])
        got = {}
        for f in range(f0, f1):
            recs = g.order[int(g.frag_start[f]):int(g.frag_start[f + 1])]
            assert len(recs) >= 1 and np.all(np.diff(recs.astype(np.int64)) > 0)        # ties keep input order
            first = int(recs[0])
            key = (int(fps[first]) if fps is not None else 0, umis[first])
            assert key not in got
            got[key] = [first, len(recs)]
            for i in recs:                                                              # the whole run shares the fingerprint
                assert (int(fps[i]) if fps is not None else 0, umis[i]) == key and bcs[i] == bc
            n_dup_sets += len(recs) > 1
        assert got == frags
    return n_dup_sets


@pytest.mark.parametrize("name,use_fp", [("v3", True), ("v3", False), ("sci3", True), ("atac", True), ("share", False)])
def test_grouping_matches_the_reference_semantics(name, use_fp):
    layout = _layouts(np.random.default_rng(20240607))[name]
    rng = np.random.default_rng(U.seed_of(name, "group", use_fp))
    infos, strides = U.emul_infos(layout)
    n = 30_000
    multi = len(layout.region_ids) > 1
    full = U.gen_reads(rng, layout, n, strides, p_short=0.0 if multi else 0.02, p_tiny=0.02 if multi else 0.0, hot=40)
    # few distinct UMIs so that duplicate sets exist; some with N, a few with a byte outside ACGTN
    for r, plan in enumerate(layout.reads):
        off = 0
        for e in plan.elems:
            if e.is_umi() and e.min_len == e.max_len and full[r].seq is not None and not any(x.min_len != x.max_len for x in plan.elems[:plan.elems.index(e)]):
                pool = rng.choice(np.frombuffer(b"ACGT", np.uint8), (12, e.max_len))
                pool[0, 0] = ord("N")
                pool[1, -1] = ord("N")
                pick = pool[rng.integers(0, len(pool), n)]
                full[r].seq[:, off:off + e.max_len] = pick
                odd = rng.random(n) < 0.002
                full[r].seq[odd, off] = ord("y")
            off += e.max_len
    fps = rng.integers(0, 6, n).astype(np.uint64) * np.uint64(0x0101010101010101) if use_fp else None
    g = _run(layout, full, fps)
    bcs, umis = _expected(layout, full, True, 0.9)
    n_dup = _check(g, bcs, umis, fps)
    if name in ("v3", "atac"):
        assert n_dup > 50                                       # the data must contain real duplicate sets


@pytest.mark.parametrize("name", ["v3", "sci3"])
def test_grouping_over_several_chunks_equals_one_chunk(name):
    """pcl_group_build_multi: a run split over three chunks (one of them empty) groups like the same records in one
    chunk - same order, fragments, barcodes - and like the reference restatement."""
    from precellar_b200.pipeline import BarcodePipeline, HostBatch
    layout = _layouts(np.random.default_rng(20240607))[name]
    rng = np.random.default_rng(U.seed_of(name, "group-multi"))
    infos, strides = U.emul_infos(layout)
    n = 20_000
    multi = len(layout.region_ids) > 1
    full = U.gen_reads(rng, layout, n, strides, p_short=0.0 if multi else 0.02, p_tiny=0.02 if multi else 0.0, hot=30)
    fps = rng.integers(0, 4, n).astype(np.uint64)
    dev = U.narrow(full, strides, infos)
    cut = [0, 7000, 7000, 15500, n]                                        # chunk 1 is empty
    sl = lambda b, lo, hi: HostBatch(None if b.seq is None else b.seq[lo:hi], None if b.qual is None else b.qual[lo:hi], b.length[lo:hi])
    pipe = BarcodePipeline(layout)
    try:
        chunks = []
        for lo, hi in zip(cut[:-1], cut[1:]):
            ch = pipe.new_chunk(max(hi - lo, 1))
            ch.reset(hi - lo)
            for r, b in enumerate(dev):
                ch.upload(r, sl(b, lo, hi))
            pipe.count(ch)
            chunks.append(ch)
        pipe.allreduce()
        for ch in chunks:
            pipe.correct(ch, 0.9)
        g = pipe.group(chunks, fingerprint=fps)
        with pytest.raises(_ffi.PclError):
            pipe.group([chunks[0], chunks[0]])
    finally:
        pipe.close()
    one = _run(layout, full, fps)
    assert g.n_records == n and g.n_assigned == one.n_assigned and g.n_ungroupable == one.n_ungroupable
    for a, b in ((g.order, one.order), (g.frag_start, one.frag_start), (g.bc_frag_start, one.bc_frag_start), (g.bc_key, one.bc_key)):
        np.testing.assert_array_equal(a, b)
    bcs, umis = _expected(layout, full, True, 0.9)
    _check(g, bcs, umis, fps)


def test_grouping_without_correction_and_empty_input():
    """Without pcl_correct only the exact whitelist members are assigned (a raw barcode outside the whitelist may hold
    non-ACGT bytes, which the packed keys do not keep); an empty chunk gives empty outputs."""
    layout = _layouts(np.random.default_rng(20240607))["v3"]
    rng = np.random.default_rng(3)
    infos, strides = U.emul_infos(layout)
    full = U.gen_reads(rng, layout, 5000, strides, hot=20)
    g = _run(layout, full, None, correct=False)
    bcs, umis = _expected(layout, full, False, 0.9)
    wl = set(layout.whitelists[layout.region_ids[0]])
    bcs = [b if b in wl else None for b in bcs]
    _check(g, bcs, umis, None)
    from precellar_b200.pipeline import HostBatch
    empty = [HostBatch(b.seq[:0], None if b.qual is None else b.qual[:0], b.length[:0]) if b.seq is not None else HostBatch(None, None, b.length[:0]) for b in full]
    g0 = _run(layout, empty, None)
    assert g0.n_assigned == 0 and len(g0.order) == 0 and len(g0.bc_key) == 0 and g0.frag_start.tolist() == [0]


def test_grouping_large_sort_is_a_sorted_permutation():
    """Size-independent properties at 5 M read pairs of the v3 configuration: the order is a permutation of the assigned
    records, barcodes ascend, every fragment's records share barcode and UMI, and fragment counts add up."""
    from precellar_b200.pipeline import BarcodePipeline
    from tools import workloads as W
    w = W.v3(200_000)
    n = 5_000_000
    pipe = BarcodePipeline(w.layout)
    try:
        ch = pipe.new_chunk(n)
        ch.reset(n)
        w.fill_chunk(pipe, ch, 0, n, 0)
        pipe.count(ch)
        pipe.allreduce()
        pipe.correct(ch, 0.975)
        res = ch.fetch(0)
        fps = (np.arange(n, dtype=np.uint64) * np.uint64(2654435761)) % np.uint64(3)
        g = pipe.group(ch, fps)
    finally:
        pipe.close()
    code = (res.fstatus.astype(np.uint32) >> 4) & 7
    assigned = (code == _ffi.BC_EXACT) | (code == _ffi.BC_CORRECTED)
    umi_ok = res.umi_key[0] != 0                                  # packable UMI == no byte outside ACGT (N is groupable, though)
    assert g.n_assigned + g.n_ungroupable == int(assigned.sum())
    order = g.order.astype(np.int64)
    assert len(np.unique(order)) == len(order) and assigned[order].all()
    keys = res.bc_key[0][order].astype(np.uint64)
    assert np.all(np.diff(keys.astype(np.int64)) >= 0)            # 16-mer keys: numeric order == string order
    assert len(g.bc_key) == len(np.unique(keys))
    counts = np.diff(g.frag_start.astype(np.int64))
    assert counts.min() >= 1 and counts.sum() == g.n_assigned
    # every fragment is homogeneous in (barcode, fingerprint, packed UMI) where the UMI is packable
    fid = np.repeat(np.arange(len(counts)), counts)
    first = order[g.frag_start[:-1].astype(np.int64)][fid]
    assert np.array_equal(keys, res.bc_key[0][first].astype(np.uint64)) and np.array_equal(fps[order], fps[first])
    both = umi_ok[order] & umi_ok[first]
    assert np.array_equal(res.umi_key[0][order][both], res.umi_key[0][first][both])
