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
@pytest.mark.parametrize("mode", ["correct", "nocorrect", "ee", "lower", "correct-generic", "correct-direct", "lower-direct"])
def test_gpu_matches_oracle(layouts, name, mode, monkeypatch):
    # the specialised kernels + neighbourhood tables are the default; "-generic" forces the interpreter kernels,
    # "-direct" the 3*L direct probes, on the same data
    monkeypatch.setenv("PCL_FORCE_GENERIC", "1" if mode.endswith("-generic") else "0")
    monkeypatch.setenv("PCL_NO_NEIGHBOUR_TABLES", "1" if mode.endswith("-direct") else "0")
    mode = mode.split("-")[0]
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
        syn.fill_device(0, 1234, n, s1, q1, l1, ch.stream)
        pipe.sync()
        seq_d, qual_d, len_d = np.empty((n, 32), np.uint8), np.empty((n, 32), np.uint8), np.empty(n, np.uint16)
        syn.memcpy_d2h(0, seq_d.ctypes.data, s1, n * 32)
        syn.memcpy_d2h(0, qual_d.ctypes.data, q1, n * 32)
        syn.memcpy_d2h(0, len_d.ctypes.data, l1, n * 2)
        seq_h, qual_h, len_h = syn.host_arrays(1234, n)
        assert np.array_equal(seq_d, seq_h) and np.array_equal(qual_d, qual_h) and np.array_equal(len_d, len_h)
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
        g1, g2 = pipe.run([HostBatch(seq, qual, length), HostBatch(None, None, len2)], correct=True, threshold=0.975)
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
    rng = np.random.default_rng(abs(hash((name, "qc"))) % (2 ** 31))
    multi = len(layout.region_ids) > 1
    pipe = BarcodePipeline(layout, qc=True)
    try:
        strides = [pipe.stride(r) for r in range(len(layout.reads))]
        n = 2500
        full = U.gen_reads(rng, layout, n, strides, p_short=0.0 if multi else 0.03, p_tiny=0.03 if multi else 0.0,
                           read_len=[p.max_len for p in layout.reads])
        dev = U.narrow(full, strides)
        ch = pipe.new_chunk(n)
        ch.reset(n)
        for r, b in enumerate(dev):
            if pipe.infos[r].qual_only:
                b = HostBatch(None, np.ascontiguousarray(full[r].qual[:, :strides[r]]) if full[r].qual.shape[1] >= strides[r]
                              else np.pad(full[r].qual, ((0, 0), (0, strides[r] - full[r].qual.shape[1])), constant_values=33), b.length)
            ch.upload(r, b)
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
