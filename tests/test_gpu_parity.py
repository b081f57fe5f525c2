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
@pytest.mark.parametrize("mode", ["correct", "nocorrect", "ee", "lower", "correct-generic", "correct-direct", "lower-direct",
                                  "correct-oldfast", "lower-oldfast", "correct-nobits"])
def test_gpu_matches_oracle(layouts, name, mode, monkeypatch):
    # the specialised kernels + neighbourhood tables are the default; "-generic" forces the interpreter kernels,
    # "-direct" the 3*L direct probes, "-oldfast" the software-pipelined k_count_fast instead of k_count_spec,
    # "-nobits" the neighbourhood tables without the partial-key bitmaps, on the same data
    monkeypatch.setenv("PCL_FORCE_GENERIC", "1" if mode.endswith("-generic") else "0")
    monkeypatch.setenv("PCL_NO_NEIGHBOUR_TABLES", "1" if mode.endswith("-direct") else "0")
    monkeypatch.setenv("PCL_OLD_FAST", "1" if mode.endswith("-oldfast") else "0")
    monkeypatch.setenv("PCL_NO_NBITS", "1" if mode.endswith("-nobits") else "0")
    mode = mode.split("-")[0]
    layout = layouts[name]
    rng = np.random.default_rng(U.seed_of(name, mode, "gpu"))
    infos, strides = U.emul_infos(layout)
    n = 3000
    multi = len(layout.region_ids) > 1
    full = U.gen_reads(rng, layout, n, strides, p_lower=0.02 if mode == "lower" else 0.0,
                       p_short=0.0 if multi else 0.03, p_tiny=0.03 if multi else 0.0)
    dev = U.narrow(full, strides, infos)
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
        ginfos, results, counts, defects, _ = run_gpu(layout, U.narrow(full, strides, infos))
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


def test_synth_host_and_device_generate_the_same_bytes():
    from tools import synth as S
    from precellar_b200.configs import v3_layout
    from precellar_b200.pipeline import BarcodePipeline
    wl = S.make_whitelist(50_000, 16)
    syn = S.Synth(S.SynthSpec(50_000, n_cells=500), wl)
    n = 200_000
    pipe = BarcodePipeline(v3_layout(wl))
    try:
        ch = pipe.new_chunk(n)
        ch.reset(n)
        s1, q1, l1 = ch.device_ptrs(0)
        syn.fill_device(0, 1234, n, s1, q1, l1, ch.stream, qual_window=pipe.qual_window(0))
        pipe.sync()
        qo, qs = pipe.qual_window(0)
        assert (qo, qs) == (0, 16)                    # v3: only the 16 barcode qualities are uploaded
        seq_d, qual_d, len_d = np.empty((n, 32), np.uint8), np.empty((n, qs), np.uint8), np.empty(n, np.uint16)
        syn.memcpy_d2h(0, seq_d.ctypes.data, s1, n * 32)
        syn.memcpy_d2h(0, qual_d.ctypes.data, q1, n * qs)
        syn.memcpy_d2h(0, len_d.ctypes.data, l1, n * 2)
        seq_h, qual_h, len_h = syn.host_arrays(1234, n)
        assert np.array_equal(seq_d, seq_h) and np.array_equal(qual_d, qual_h[:, qo:qo + qs]) and np.array_equal(len_d, len_h)
    finally:
        pipe.close()


def test_v3_full_size_2p5M_pairs_6p8M_whitelist():
    """BASELINE configs[0]: 2.5 M synthetic read pairs, 6.8 M-entry whitelist, every field against the oracle."""
    from tools import synth as S
    from tools.v3check import compare_v3
    from oracle import pyoracle as O
    from precellar_b200.configs import v3_layout
    from precellar_b200.pipeline import BarcodePipeline, HostBatch
    n_wl, n = 6_800_000, 2_500_000
    wl = S.make_whitelist(n_wl, 16)
    syn = S.Synth(S.SynthSpec(n_wl), wl)
    seq, qual, length = syn.host_arrays(0, n)
    len2 = np.full(n, 91, np.uint16)
    len2[::1000] = 0                                   # a few empty / short inserts
    len2[1::1000] = 300                                # longer than max_len: truncated to 91
    layout = v3_layout(wl, 91)
    pipe = BarcodePipeline(layout)
    try:
        g1, g2 = pipe.run([pipe.host_planes(0, seq, qual, length), HostBatch(None, None, len2)], correct=True, threshold=0.975)
        gcounts = pipe.counts(0)
    finally:
        pipe.close()
    orc = O.Oracle(O.reads_from_layout(layout), {0: (wl.reshape(-1), np.arange(n_wl + 1, dtype=np.uint64) * np.uint64(16), n_wl)})
    orc.set_input(0, seq, qual, length, 32)
    orc.set_input(1, None, None, len2, 0)
    orc.count()
    ores = orc.annotate(correct=True, threshold=0.975, n_threads=16)
    fields = compare_v3(wl, g1, g2, gcounts, ores, orc.counts(0))
    assert sum(fields.values()) == 0, fields
    # UMI keys against the input bytes
    u = seq[:, 16:28]
    code = ((u >> 1) ^ (u >> 2)) & 3
    valid = np.isin(u, np.frombuffer(b"ACGT", np.uint8)).all(axis=1)
    k = np.ones(n, np.uint64)
    for p in range(12):
        k = (k << np.uint64(2)) | code[:, p].astype(np.uint64)
    assert np.array_equal(g1.umi_key[0], np.where(valid, k, 0).astype(np.uint32))
    code_bc = (g1.fstatus.astype(np.uint32) >> 4) & 7
    assert (code_bc == _ffi.BC_CORRECTED).sum() > 100_000 and (code_bc == _ffi.BC_EXACT).sum() > 1_500_000


@pytest.mark.parametrize("name", ["v3", "v2", "atac", "multiome_atac", "sci3", "share", "dsc", "k64_rev", "multi_umi"])
def test_qc_fastq_counters(layouts, name):
    """QcFastq (precellar/src/qc.rs:30-69): bases / Q30 bases / reads per category and per-file defects."""
    from precellar_b200.pipeline import BarcodePipeline, HostBatch
    layout = layouts[name]
    rng = np.random.default_rng(U.seed_of(name, "qc"))
    multi = len(layout.region_ids) > 1
    pipe = BarcodePipeline(layout, qc=True)
    try:
        strides = [pipe.stride(r) for r in range(len(layout.reads))]
        n = 2500
        full = U.gen_reads(rng, layout, n, strides, p_short=0.0 if multi else 0.03, p_tiny=0.03 if multi else 0.0,
                           read_len=[p.max_len for p in layout.reads])
        ch = pipe.new_chunk(n)
        ch.reset(n)
        for r in range(len(layout.reads)):
            ch.upload(r, pipe.host_planes(r, full[r].seq, full[r].qual, full[r].length))
        pipe.count(ch)
        pipe.qc_accumulate(ch)
        cat = pipe.qc_categories()
        per = pipe.qc()
    finally:
        pipe.close()
    ores, _ = U.run_oracle(layout, full, correct=False, render_lines=False)
    np.testing.assert_array_equal(cat, ores.qc_cat)
    np.testing.assert_array_equal(per, ores.qc_per_file)
    assert cat[1, 1] > 0 and cat[1, 2] > 0


def _bulk(layout, batches, n_threads=16, **kw):
    """Run the CUDA path and the oracle on host batches; returns compare_bulk's mismatch table."""
    from oracle import pyoracle as O
    from precellar_b200.pipeline import BarcodePipeline
    from tools.v3check import compare_bulk
    pipe = BarcodePipeline(layout)
    try:
        dev = [pipe.host_planes(r, b.seq, b.qual, b.length) if b.seq is not None else b for r, b in enumerate(batches)]
        results = pipe.run(dev, correct=True, threshold=kw.get("threshold", 0.975))
        counts = {g: pipe.counts(g) for g in range(len(layout.region_ids))}
        infos = pipe.infos
        qc = pipe.qc()
    finally:
        pipe.close()
    wl = {}
    for g, rid in enumerate(layout.region_ids):
        w = layout.whitelists[rid]
        if isinstance(w, np.ndarray):
            wl[g] = (w.reshape(-1), np.arange(len(w) + 1, dtype=np.uint64) * np.uint64(w.shape[1]), len(w))
        else:
            wl[g] = w
    orc = O.Oracle(O.reads_from_layout(layout), wl)
    for r, b in enumerate(batches):
        orc.set_input(r, b.seq, b.qual, b.length, 0 if b.seq is None else b.seq.shape[1])
    orc.count()
    ores = orc.annotate(correct=True, threshold=kw.get("threshold", 0.975), n_threads=n_threads)
    ocounts = {g: orc.counts(g) for g in range(len(layout.region_ids))}
    fields = compare_bulk(layout, infos, results, counts, ores, ocounts)
    assert np.array_equal(qc[:, 1], ores.qc_per_file[:, 1])
    return fields, results, ores


def test_atac_config_reverse_complemented_i2_737k_whitelist():
    """BASELINE configs[1] shape (10x scATAC: neg-strand 16 nt I2, 737 K whitelist, read triples), 5 M triples."""
    from tools import synth as S
    from precellar_b200.configs import atac_layout
    from precellar_b200.pipeline import HostBatch
    n_wl, n = 737_000, 5_000_000
    wl = S.make_whitelist(n_wl, 16, seed=0xA7AC)
    syn = S.Synth(S.SynthSpec(n_wl, bc_len=16, umi_len=0, stride=16, reverse=True, seed=0x5EED0002), wl)
    seq, qual, length = syn.host_arrays(0, n)
    l1 = np.full(n, 50, np.uint16)
    l2 = np.full(n, 50, np.uint16)
    l2[::777] = 0
    fields, results, ores = _bulk(atac_layout(wl, 50), [HostBatch(None, None, l1), HostBatch(seq, qual, length), HostBatch(None, None, l2)])
    assert sum(fields.values()) == 0, fields
    code = (results[1].fstatus.astype(np.uint32) >> 4) & 7
    assert (code == _ffi.BC_EXACT).mean() > 0.7 and (code == _ffi.BC_CORRECTED).sum() > 100_000


def test_multiome_atac_config_linker_plus_reverse_barcode():
    """BASELINE configs[2], ATAC side: I2 = 8 nt linker + neg-strand 16 nt barcode (24 nt), 737 K whitelist."""
    from tools import synth as S
    from precellar_b200.configs import multiome_atac_layout
    from precellar_b200.pipeline import HostBatch
    n_wl, n = 737_000, 3_000_000
    wl = S.make_whitelist(n_wl, 16, seed=0xA7AD)
    syn = S.Synth(S.SynthSpec(n_wl, bc_len=16, umi_len=0, prefix_len=8, stride=32, reverse=True, seed=0x5EED0003), wl)
    seq, qual, length = syn.host_arrays(0, n)
    l1 = np.full(n, 50, np.uint16)
    fields, results, ores = _bulk(multiome_atac_layout(wl, 50), [HostBatch(None, None, l1), HostBatch(seq, qual, length), HostBatch(None, None, l1)])
    assert sum(fields.values()) == 0, fields


def test_multiome_config_rna_and_atac_pipelines_resident_together():
    """BASELINE configs[2] as stated (10x_rna_atac.yaml): the RNA pipeline (v3 shape, 737 K gex list) and the ATAC
    pipeline (I2 = linker 8 + neg-strand CB16, 737 K atac list) live on ONE GPU at the same time -- both whitelist tables
    resident, their passes interleaved the way a caller with two assays would issue them -- and each equals the oracle
    (precellar/src/align/fastq.rs:120-151: one analyzer, one set of counts per assay)."""
    from precellar_b200.pipeline import BarcodePipeline
    from tools import workloads as W
    from tools.v3check import compare_bulk
    n = 2_000_000
    ws = [W.multiome_rna(), W.multiome_atac()]
    assert [len(w.layout.whitelists[w.layout.region_ids[0]]) for w in ws] == [737_000, 737_000]
    pipes = [BarcodePipeline(w.layout) for w in ws]
    try:
        fulls = [w.host_batches(0, n) for w in ws]
        chunks = []
        for w, pipe, full in zip(ws, pipes, fulls):           # both tables are loaded before any read is counted
            ch = pipe.new_chunk(n)
            ch.reset(n)
            for r, b in enumerate(w.device_batches(pipe, full)):
                ch.upload(r, b)
            chunks.append(ch)
        for pipe, ch in zip(pipes, chunks):                   # pass 1 of both, then the exchange, then pass 2 of both
            pipe.count(ch)
        for pipe in pipes:
            pipe.allreduce()
        for pipe, ch in zip(pipes, chunks):
            pipe.correct(ch, 0.975)
        for w, pipe, ch, full in zip(ws, pipes, chunks, fulls):
            results = [ch.fetch(r) for r in range(len(full))]
            counts = {0: pipe.counts(0)}
            _, ores, ocounts = W.run_oracle(w, full, 0.975, 16)
            fields = compare_bulk(w.layout, pipe.infos, results, counts, ores, ocounts)
            assert sum(fields.values()) == 0, (w.name, fields)
            assert np.array_equal(pipe.qc()[:, 1], ores.qc_per_file[:, 1])
            code = (results[w.bc_read].fstatus.astype(np.uint32) >> 4) & 7
            assert (code == _ffi.BC_EXACT).mean() > 0.7 and (code == _ffi.BC_CORRECTED).sum() > 50_000
    finally:
        for pipe in pipes:
            pipe.close()


def test_device_built_table_with_duplicates_and_both_build_paths(monkeypatch):
    """Whitelists of >= 100 K entries are deduplicated (first occurrence wins: IndexSet, region.rs:459-469) and hashed on
    the device (tablebuild.cuh); smaller ones on the host.  The same list with duplicates through both paths: counts in
    whitelist order, corrected barcodes and statuses equal the oracle."""
    from oracle import pyoracle as O
    from precellar_b200.configs import v3_layout
    from precellar_b200.pipeline import BarcodePipeline
    from tools import synth as S
    from tools.v3check import compare_bulk
    n_unique, n = 150_000, 1_000_000
    base = S.make_whitelist(n_unique, 16, seed=77)
    rng = np.random.default_rng(5)
    dup_src = rng.integers(0, n_unique, 30_000)
    order = np.concatenate([np.arange(n_unique), dup_src])
    rng.shuffle(order)
    wl = base[order]                                           # 180 K lines, 30 K of them repeats
    _, first = np.unique(order, return_index=True)
    uniq_in_file_order = wl[np.sort(first)]                    # what IndexSet keeps
    syn = S.Synth(S.SynthSpec(n_unique, n_cells=3000), uniq_in_file_order)
    seq, qual, length = syn.host_arrays(0, n)
    l2 = np.full(n, 91, np.uint16)
    layout_dup = v3_layout(wl, 91)
    layout_ref = v3_layout(uniq_in_file_order, 91)
    orc = O.Oracle(O.reads_from_layout(layout_dup), {0: (wl.reshape(-1), np.arange(len(wl) + 1, dtype=np.uint64) * np.uint64(16), n_unique)})
    orc.set_input(0, seq, qual, length, 32)
    orc.set_input(1, None, None, l2, 0)
    orc.count()
    ores = orc.annotate(correct=True, threshold=0.975, n_threads=16)
    ocounts = {0: orc.counts(0)}
    for min_device in ("1000", "100000000"):                   # device build, then host build
        monkeypatch.setenv("PCL_DEVICE_BUILD_MIN", min_device)
        pipe = BarcodePipeline(layout_dup)
        try:
            assert pipe.whitelist_size(0) == n_unique
            from precellar_b200.pipeline import HostBatch
            dev = [pipe.host_planes(0, seq, qual, length), HostBatch(None, None, l2)]
            results = pipe.run(dev, correct=True, threshold=0.975)
            counts = {0: pipe.counts(0)}
            fields = compare_bulk(layout_ref, pipe.infos, results, counts, ores, ocounts)
            assert sum(fields.values()) == 0, (min_device, fields)
        finally:
            pipe.close()


def test_sci3_config_variable_hairpin_and_anchor():
    """BASELINE configs[3] shape (sci-RNA-seq3: hairpin 9|10 + CAGAGC anchor + UMI 8 + RT 10), 3 M pairs through the
    element-table interpreter kernel; includes anchor mismatches (defect records)."""
    from tools import synth as S
    from precellar_b200.configs import sci3_layout
    from precellar_b200.pipeline import HostBatch
    hp = S.make_mixed_whitelist(238, 146, 7)
    rt = [bytes(r) for r in S.make_whitelist(384, 10, 99)]
    gen = S.SynthSci3(hp, rt)
    n = 3_000_000
    seq, qual, length = gen.host_arrays(0, n)
    l2 = np.full(n, 100, np.uint16)
    fields, results, ores = _bulk(sci3_layout(hp, rt, 34, 100), [HostBatch(seq, qual, length), HostBatch(None, None, l2)])
    assert sum(fields.values()) == 0, fields
    assert (ores.fstatus[0] != 0).sum() > 1000          # PatternMismatch records exist and agree
    assert (ores.bc_index[0] >= 0).mean() > 0.8 and (ores.bc_index[1] >= 0).mean() > 0.8


@pytest.mark.parametrize("name", ["all6", "len15", "mix15_16", "len31", "near_dups"])
@pytest.mark.parametrize("direct", [False, True])
def test_gpu_dense_and_boundary_whitelists(name, direct, monkeypatch):
    from tests.test_emul_parity import _dense_layouts
    monkeypatch.setenv("PCL_FORCE_GENERIC", "0")
    monkeypatch.setenv("PCL_NO_NEIGHBOUR_TABLES", "1" if direct else "0")
    layout = _dense_layouts(np.random.default_rng(77))[name]
    rng = np.random.default_rng(U.seed_of(name, "dense-gpu"))
    infos, strides = U.emul_infos(layout)
    full = U.gen_reads(rng, layout, 6000, strides, p_sub=0.05, p_n=0.02, hot=10 ** 6)
    ginfos, results, counts, defects, _ = run_gpu(layout, U.narrow(full, strides, infos), threshold=0.6, n_chunks=2)
    ores, ocounts = U.run_oracle(layout, full, threshold=0.6)
    U.compare_with_oracle(layout, ginfos, full, results, counts, defects, ores, ocounts)


def test_empty_chunk_and_reuse():
    """n = 0 chunks are legal; a chunk can be refilled (pcl_chunk_reset) with a different record count."""
    from precellar_b200.configs import v3_layout
    from precellar_b200.pipeline import BarcodePipeline, HostBatch
    rng = np.random.default_rng(3)
    wl = U.gen_whitelist(rng, 500, 16)
    layout = v3_layout(wl, 91)
    pipe = BarcodePipeline(layout)
    try:
        ch = pipe.new_chunk(5000)
        ch.reset(0)
        ch.upload(0, HostBatch(np.zeros((0, 32), np.uint8), np.zeros((0, 16), np.uint8), np.zeros(0, np.uint16)))
        ch.upload(1, HostBatch(None, None, np.zeros(0, np.uint16)))
        pipe.count(ch); pipe.allreduce(); pipe.correct(ch)
        assert pipe.counts(0)[2] == 0
        infos, strides = pipe.infos, [pipe.stride(r) for r in range(2)]
        for n in (5000, 1234):
            full = U.gen_reads(rng, layout, n, strides)
            dev = U.narrow(full, strides, infos)
            pipe.reset_counts()
            ch.reset(n)
            for r, b in enumerate(dev):
                ch.upload(r, b)
            pipe.count(ch); pipe.allreduce(); pipe.correct(ch)
            results = [ch.fetch(r) for r in range(2)]
            counts = {0: pipe.counts(0)}
            ores, ocounts = U.run_oracle(layout, full)
            U.compare_with_oracle(layout, infos, full, results, counts, pipe.qc()[:, 1], ores, ocounts)
    finally:
        pipe.close()


@pytest.mark.parametrize("name", ["v3", "fast12_tail", "v2", "atac", "sci3"])
def test_compact_upload_gives_the_same_results(layouts, name):
    """pcl_chunk_upload_compact: sequence rows cut down to pcl_read_payload_bytes (and one size in between) are widened
    on the device; every result plane and count equals the full-stride upload."""
    from precellar_b200.pipeline import BarcodePipeline, HostBatch
    layout = layouts[name]
    rng = np.random.default_rng(U.seed_of(name, "compact"))
    infos, strides = U.emul_infos(layout)
    n = 2500
    multi = len(layout.region_ids) > 1
    full = U.gen_reads(rng, layout, n, strides, p_short=0.0 if multi else 0.03, p_tiny=0.03 if multi else 0.0)
    dev = U.narrow(full, strides, infos)
    ref = run_gpu(layout, dev)

    def run(width_of):
        pipe = BarcodePipeline(layout)
        try:
            ch = pipe.new_chunk(n)
            ch.reset(n)
            for r, b in enumerate(dev):
                if b.seq is None:
                    ch.upload(r, b)
                    continue
                w = width_of(pipe, r)
                assert pipe.payload_bytes(r) <= w <= pipe.stride(r) and w % 4 == 0
                ch.upload_compact(r, HostBatch(np.ascontiguousarray(b.seq[:, :w]), b.qual, b.length))
            pipe.count(ch)
            pipe.allreduce()
            pipe.correct(ch, 0.975)
            return [ch.fetch(r) for r in range(len(dev))], {g: pipe.counts(g) for g in range(len(layout.region_ids))}
        finally:
            pipe.close()
    for width_of in (lambda p, r: p.payload_bytes(r), lambda p, r: min(p.stride(r), p.payload_bytes(r) + 4)):
        res, counts = run(width_of)
        for a, b in zip(res, ref[1]):
            assert np.array_equal(a.fstatus, b.fstatus)
            for x, y in zip(a.bc_key + a.umi_key, b.bc_key + b.umi_key):
                assert np.array_equal(x, y)
            for x, y in zip([a.tcoord] + a.bcoord + a.ucoord, [b.tcoord] + b.bcoord + b.ucoord):
                assert (x is None and y is None) or np.array_equal(x, y)
        for g in counts:
            assert np.array_equal(counts[g][0], ref[2][g][0]) and counts[g][1:] == ref[2][g][1:]
    # rows outside [payload_bytes, stride] or not a multiple of 4 are rejected
    pipe = BarcodePipeline(layout)
    try:
        r = next(i for i, b in enumerate(dev) if b.seq is not None)
        ch = pipe.new_chunk(8)
        ch.reset(8)
        bad = HostBatch(np.zeros((8, pipe.payload_bytes(r) - 4 if pipe.payload_bytes(r) > 4 else 2), np.uint8),
                        None if dev[r].qual is None else np.ascontiguousarray(dev[r].qual[:8]), dev[r].length[:8])
        with pytest.raises(_ffi.PclError):
            ch.upload_compact(r, bad)
    finally:
        pipe.close()


def test_gpu_random_anchor_plan_layouts():
    """The AnchorPlan kernel (k_count<2>) on random layouts of its shape, against the oracle."""
    import ctypes as C
    from tests.test_emul_parity import _random_plan_layout
    rng = np.random.default_rng(U.seed_of("random-plans-gpu"))
    L = U.emul_lib()
    done = tries = 0
    while done < 12 and tries < 300:
        tries += 1
        layout = _random_plan_layout(rng)
        if layout is None:
            continue
        try:
            infos, strides = U.emul_infos(layout)
        except _ffi.PclError:
            continue
        if L.emu_fast_enabled(C.byref(layout.descs()[0])) != 3:
            continue
        multi = len(layout.region_ids) > 1
        full = U.gen_reads(rng, layout, 600, strides, p_sub=0.04, p_n=0.02, p_lower=0.02 if layout.reads[0].is_reverse else 0.0,
                           p_short=0.0 if multi else 0.15, p_tiny=0.05 if multi else 0.0, p_anchor_mut=0.3)
        try:
            ores, ocounts = U.run_oracle(layout, full, threshold=0.8)
        except Exception:
            continue
        ginfos, results, counts, defects, _ = run_gpu(layout, U.narrow(full, strides, infos), threshold=0.8)
        U.compare_with_oracle(layout, ginfos, full, results, counts, defects, ores, ocounts)
        done += 1
    assert done >= 8


def test_uniform_length_upload_and_uniform_results(layouts):
    """pcl_chunk_set_uniform_len / pcl_chunk_uniform_results_async: a read file whose records all have one length sends no
    length plane, and an insert read's constant result planes come back as one pair -- same results as the plain route."""
    from precellar_b200.pipeline import BarcodePipeline, HostBatch
    layout = layouts["v3"]
    infos, strides = U.emul_infos(layout)
    rng = np.random.default_rng(11)
    full = U.gen_reads(rng, layout, 4000, strides, p_short=0.0)
    full[0].length[:] = 28
    full[1].length[:] = 91
    dev = U.narrow(full, strides, infos)
    ginfos, want, counts_w, _, _ = run_gpu(layout, dev)
    pipe = BarcodePipeline(layout)
    try:
        n = 4000
        ch = pipe.new_chunk(n)
        ch.reset(n)
        ch.set_uniform_len(0, 28)
        ch.upload_compact(0, HostBatch(np.ascontiguousarray(dev[0].seq[:, :28]), dev[0].qual, None))
        ch.set_uniform_len(1, 91)
        pipe.count(ch)
        pipe.allreduce()
        pipe.correct(ch, 0.975)
        import ctypes as C
        buf = np.zeros(4, np.uint32)
        ch.uniform_results_async(1, buf)
        got0 = ch.fetch(0)
        got1 = ch.fetch(1)
        assert buf[0] == 1 and (got1.fstatus == buf[1]).all() and (got1.tcoord == buf[2]).all()
        assert np.array_equal(got0.fstatus, want[0].fstatus) and np.array_equal(got0.bc_key[0], want[0].bc_key[0])
        assert np.array_equal(got0.umi_key[0], want[0].umi_key[0]) and np.array_equal(got1.tcoord, want[1].tcoord)
        # not uniform: one shorter insert
        lens = np.full(n, 91, np.uint16); lens[17] = 40
        ch.reset(n)
        ch.set_uniform_len(0, 28)
        ch.upload_compact(0, HostBatch(np.ascontiguousarray(dev[0].seq[:, :28]), dev[0].qual, None))
        ch.upload(1, HostBatch(None, None, lens))
        pipe.reset_counts()
        pipe.count(ch)
        ch.uniform_results_async(1, buf)
        pipe.sync()
        assert buf[0] == 0
    finally:
        pipe.close()
