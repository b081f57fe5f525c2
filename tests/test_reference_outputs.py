"""The reference's own golden OUTPUT files (precellar/data/test{2,3,4}.out.gz, used by `test_fastq`,
precellar/src/align/fastq.rs:694-726) as a known-answer test: tests/golden/reference_outputs.json.gz holds input
records rebuilt from those lines (tests/golden/make_reference_outputs.py says how and what that can and cannot pin),
the three layouts compiled from the reference's YAML files, and the lines verbatim.  The oracle, the host emulation
of the device code and the CUDA path must all print exactly the reference's lines."""
import gzip
import json
import os

import numpy as np
import pytest

from precellar_b200.layout import CompiledLayout, ReadPlan
from precellar_b200.pipeline import HostBatch, assemble, render
from precellar_b200.seqspec import SegmentInfoElem
from tests import util as U

HERE = os.path.dirname(os.path.abspath(__file__))
with gzip.open(os.path.join(HERE, "golden", "reference_outputs.json.gz"), "rt") as fh:
    DOC = json.load(fh)
CASES = sorted(DOC)


def load_case(name):
    d = DOC[name]
    slot = {r: i for i, r in enumerate(d["region_ids"])}
    plans = []
    for r in d["reads"]:
        elems = [SegmentInfoElem(e["region_id"], e["region_type"], e["sequence_type"], e["sequence"], e["min_len"], e["max_len"])
                 for e in r["elems"]]
        plans.append(ReadPlan(r["read_id"], r["is_reverse"], r["max_len"], elems,
                              [slot[e.region_id] if e.is_barcode() else -1 for e in elems], []))
    wl = {k: [x.encode() for x in v] for k, v in d["whitelists"].items()}
    layout = CompiledLayout(d["modality"], plans, d["region_ids"], wl, {k: True for k in d["region_ids"]},
                            any(e.is_umi() for p in plans for e in p.elems))
    full = []
    for p in plans:
        recs = [s.encode() for s in d["inputs"][p.read_id]]
        w = ((max(len(s) for s in recs) + 15) // 16) * 16
        seq = np.full((len(recs), w), ord("N"), np.uint8)
        for i, s in enumerate(recs):
            seq[i, :len(s)] = np.frombuffer(s, np.uint8)
        full.append(HostBatch(seq, np.full_like(seq, ord("F")), np.array([len(s) for s in recs], np.uint16)))
    return layout, full, d["reference_lines"]


def show_fq(line: str) -> str:
    """pipeline.render's nine fields -> the four `show_fq` prints (fastq.rs:676-692)."""
    f = line.split("\t")
    return "\t".join([f[0], f[3], f[5], f[7]])


@pytest.mark.parametrize("name", CASES)
def test_layouts_are_what_the_survey_traced(name):
    layout, full, lines = load_case(name)
    kinds = {p.read_id: [(e.region_type, e.min_len, e.max_len) for e in p.elems] for p in layout.reads}
    if name == "test4":                                   # SURVEY Appendix A, sci-RNA-seq3 with max_len 36
        assert kinds["R1"] == [("barcode", 9, 10), ("linker", 6, 6), ("umi", 8, 8), ("barcode", 10, 10)]
        assert sorted({len(l.split("\t")[0]) for l in lines}) == [19, 20]
    if name == "test3":
        assert [k[0] for k in kinds["R1"]].count("barcode") == 3 and ("linker", 0, 4) in kinds["R1"]
    if name == "test2":
        assert [k for k in kinds["atac-I1"] if k[0] == "barcode"] == [("barcode", 8, 8)] * 3
    assert len(lines) == 250 and all(len(b.length) == 250 for b in full)
    if name in ("test3", "test4"):                        # dscATAC and sci-RNA-seq3 R1 run on the AnchorPlan kernel
        import ctypes as C
        d = layout.descs()
        assert U.emul_lib().emu_fast_enabled(C.byref(d[0])) == 3


@pytest.mark.parametrize("name", CASES)
def test_oracle_prints_the_reference_lines(name):
    layout, full, lines = load_case(name)
    ores, _ = U.run_oracle(layout, full, correct=False)
    assert [show_fq(l) for l in ores.lines] == lines


@pytest.mark.parametrize("use_fast", [0, 1])
@pytest.mark.parametrize("name", CASES)
def test_device_code_on_the_host_prints_the_reference_lines(name, use_fast):
    layout, full, lines = load_case(name)
    infos, strides = U.emul_infos(layout)
    infos, results, counts, defects = U.run_emul(layout, U.narrow(full, strides, infos), correct=False, use_fast=use_fast)
    g = assemble(layout, infos, full, results, correct=False)
    assert [show_fq(render(layout, full, g, i, correct=False)) for i in range(250)] == lines
    assert not defects.any()


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_cuda_path_prints_the_reference_lines(name):
    from tests.test_gpu_parity import run_gpu
    layout, full, lines = load_case(name)
    infos, strides = U.emul_infos(layout)
    ginfos, results, counts, defects, _ = run_gpu(layout, U.narrow(full, strides, infos), correct=False)
    g = assemble(layout, ginfos, full, results, correct=False)
    assert [show_fq(render(layout, full, g, i, correct=False)) for i in range(250)] == lines
    # with correction on, every barcode of these records is in its whitelist or one substitution away
    ginfos, results, counts, defects, _ = run_gpu(layout, U.narrow(full, strides, infos), correct=True, threshold=0.9)
    ores, ocounts = U.run_oracle(layout, full, correct=True, threshold=0.9)
    U.compare_with_oracle(layout, ginfos, full, results, counts, defects, ores, ocounts)
