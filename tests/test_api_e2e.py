"""GPU tier, end to end through the public API: seqspec YAML + FASTQ files on disk -> FastqProcessor /
make_fastq, against the oracle fed with the same records."""
import gzip
import os

import numpy as np
import pytest
import yaml

from oracle import pyoracle as O
from tests import util as U

pytestmark = pytest.mark.gpu


def region(rid, rtype, stype, seq, lo, hi, onlist=None):
    d = dict(region_id=rid, region_type=rtype, name=rid, sequence_type=stype, sequence=seq, min_len=lo, max_len=hi,
             onlist=None, regions=None)
    if onlist:
        d["onlist"] = dict(file_id=os.path.basename(onlist), filename=os.path.basename(onlist), filetype="txt",
                           filesize=0, url=onlist, urltype="local", md5=0, location="local")
    return d


def write_spec(td, hp, rt, forward_insert=False):
    if hp is not None:
        open(os.path.join(td, "hp.txt"), "wb").write(b"\n".join(hp))             # no trailing newline, like the fixtures
        with gzip.open(os.path.join(td, "rt.txt.gz"), "wb") as fh:
            fh.write(b"\r\n".join(rt) + b"\r\n")
    regions = [
        region("p5", "illumina_p5", "fixed", "AATGATACGGCGACCACCGAGATCTACAC", 29, 29),
        region("i5_primer", "truseq_read1", "fixed", "ACACTCTTTCCCTACACGACGCTCTTCCGATCT", 33, 33),
        region("hairpin_barcode", "barcode", "onlist", "N" * 10, 9, 10, "hp.txt"),
        region("linker", "linker", "fixed", "CAGAGC", 6, 6),
        region("umi", "umi", "random", "N" * 8, 8, 8),
        region("rt_barcode", "barcode", "onlist", "N" * 10, 10, 10, "rt.txt.gz"),
        region("poly_t", "poly_t", "random", "T" * 30, 30, 32),
        region("r2p", "custom_primer", "fixed", "GTCTCGTGGGCTCGGAGATG", 20, 20),
        region("cdna", "cdna", "random", "X", 1, 1000),
        region("me", "nextera_read2", "fixed", "CTGTCTCTTATACACATCTCCGAGCCCACGAGAC", 34, 34),
        region("p7", "illumina_p7", "fixed", "ATCTCGTATGCCGTCTTCTGCTTG", 24, 24),
    ]
    spec = dict(seqspec_version="0.3.0", assay_id="sci3-test", name="sci3-test", doi="", date="", description="",
                modalities=["rna"], lib_struct="", library_protocol="x", library_kit="x", sequence_protocol="x",
                sequence_kit="x",
                sequence_spec=[
                    dict(read_id="R1", name="R1", modality="rna", primer_id="i5_primer", min_len=34, max_len=34, strand="pos"),
                    dict(read_id="R2", name="R2", modality="rna", primer_id="r2p" if forward_insert else "me", min_len=40,
                         max_len=60, strand="pos" if forward_insert else "neg")],
                library_spec=[dict(region_id="rna", region_type="rna", name="rna", sequence_type="joined", sequence="",
                                   min_len=0, max_len=0, onlist=None, regions=regions)])
    path = os.path.join(td, "spec_fwd.yaml" if forward_insert else "spec.yaml")
    yaml.safe_dump(spec, open(path, "w"), sort_keys=False)
    return path


def stripped(nm: bytes) -> bytes:
    """strip_fq_suffix (fastq.rs:612-621): a trailing /1 or /2 of the NAME is dropped before annotate; the description stays."""
    parts = nm.split(b" ", 1)
    if len(parts[0]) > 2 and parts[0][-2:] in (b"/1", b"/2"):
        parts[0] = parts[0][:-2]
    return b" ".join(parts)


def write_fastq(path, names, batch, gz):
    if gz == "bgzf":                                               # BGZF: independent members, inflated on the device
        from tests.test_ingest import bgzf_bytes
        text = b"".join(b"@" + nm + b"\n" + batch.seq[i, :int(batch.length[i])].tobytes() + b"\n+\n" +
                        batch.qual[i, :int(batch.length[i])].tobytes() + b"\n" for i, nm in enumerate(names))
        open(path, "wb").write(bgzf_bytes(text, block=20000))
        return
    op = gzip.open if gz else open
    with op(path, "wb") as fh:
        for i, nm in enumerate(names):
            L = int(batch.length[i])
            fh.write(b"@" + nm + b"\n" + batch.seq[i, :L].tobytes() + b"\n+\n" + batch.qual[i, :L].tobytes() + b"\n")


@pytest.fixture(scope="module")
def dataset(tmp_path_factory):
    td = str(tmp_path_factory.mktemp("e2e"))
    rng = np.random.default_rng(11)
    hp = U.gen_whitelist(rng, 96, [9, 10])
    rt = U.gen_whitelist(rng, 96, 10)
    spec = write_spec(td, hp, rt)
    from precellar_b200 import Assay, compile_layout
    assay = Assay(spec)
    layout0 = compile_layout(assay, "rna")
    n = 5000
    full = U.gen_reads(rng, layout0, n, [64, 64], p_short=0.0, p_tiny=0.02, read_len=[34, 60])
    # the insert read of the synthetic generator is 1..60 long; make every sequence non-empty for a valid FASTQ
    for b in full:
        b.length[:] = np.maximum(b.length, 1)
    names1 = [f"read{i}/1 x:{i}".encode() for i in range(n)]
    names2 = [f"read{i}/2".encode() for i in range(n)]
    h = n // 2
    from precellar_b200.pipeline import HostBatch
    sl = lambda b, lo, hi: HostBatch(b.seq[lo:hi], b.qual[lo:hi], b.length[lo:hi])
    write_fastq(os.path.join(td, "R1_a.fq.gz"), names1[:h], sl(full[0], 0, h), True)      # two files per read: concatenated
    write_fastq(os.path.join(td, "R1_b.fq"), names1[h:], sl(full[0], h, n), False)
    write_fastq(os.path.join(td, "R2.fq.gz"), names2, full[1], True)
    assay.update_read("R1", fastq=[os.path.join(td, "R1_a.fq.gz"), os.path.join(td, "R1_b.fq")])
    assay.update_read("R2", fastq=os.path.join(td, "R2.fq.gz"))
    return dict(td=td, assay=assay, full=full, names1=names1, names2=names2, n=n)


def oracle_lines(assay, full, correct, threshold=0.975):
    from precellar_b200 import compile_layout
    layout = compile_layout(assay, "rna", require_files=True)
    ores, _ = U.run_oracle(layout, full, correct=correct, threshold=threshold)
    return layout, ores


def test_update_read_inferred_lengths(dataset):
    a = dataset["assay"]
    r1, r2 = a.sequence_spec["R1"], a.sequence_spec["R2"]
    L1, L2 = dataset["full"][0].length, dataset["full"][1].length
    assert (r1.min_len, r1.max_len) == (int(L1.min()), int(L1.max()))
    assert (r2.min_len, r2.max_len) == (int(L2.min()), int(L2.max()))


@pytest.mark.parametrize("device_parse", [False, True])
@pytest.mark.parametrize("correct", [False, True])
def test_gen_barcoded_fastq_matches_oracle(dataset, correct, device_parse):
    from precellar_b200.api import FastqProcessor
    proc = FastqProcessor(dataset["assay"], device_parse=device_parse).with_modality("rna").with_barcode_correct_prob(0.9)
    got = []
    for batch in proc.gen_barcoded_fastq(correct, chunk_size=100_000):      # several chunks
        got.extend(batch)
    layout, ores = oracle_lines(dataset["assay"], dataset["full"], correct, 0.9)
    want = [ln for ln in ores.lines if ln]
    assert len(got) == len(want) > 3000
    for rec, line in zip(got, want):
        f = line.split("\t")
        mine = [rec.barcode.raw.sequence, rec.barcode.raw.quality, rec.barcode.corrected if rec.barcode.corrected is not None else b"*",
                rec.umi.sequence if rec.umi else b"", rec.umi.quality if rec.umi else b"",
                rec.read1.sequence if rec.read1 else b"", rec.read1.quality if rec.read1 else b"",
                rec.read2.sequence if rec.read2 else b"", rec.read2.quality if rec.read2 else b""]
        assert [m.decode("latin-1") for m in mine] == f
    # the SAM tags an aligner adapter would attach (aligners.rs:347-372)
    t = got[0].tags()
    assert t["CR"] == got[0].barcode.raw.sequence and t["CY"] == got[0].barcode.raw.quality and t["UR"] == got[0].umi.sequence
    assert ("CB" in t) == (got[0].barcode.corrected is not None)
    # definitions: barcode/umi carry R1's definition line, read2 its own
    k = next(i for i, ln in enumerate(ores.lines) if ln)
    assert got[0].barcode.raw.definition == stripped(dataset["names1"][k]) == b"read%d x:%d" % (k, k)
    assert got[0].read2.definition == stripped(dataset["names2"][k]) == b"read%d" % k
    # QcFastq
    qc = proc.get_fastq_qc()
    cat = ores.qc_cat.astype(float)
    for j, nm in enumerate(("umi", "barcode", "read1", "read2")):
        if cat[j, 1] > 0:
            assert qc["frac_q30_bases"][nm] == cat[j, 2] / cat[j, 1]
        else:
            assert nm not in qc["frac_q30_bases"]
    per = ores.qc_per_file
    want_defect = {rid: per[r, 1] / per[r, 0] for r, rid in enumerate(("R1", "R2")) if per[r, 1] > 0}
    assert qc.get("frac_reads_with_misformed_structure", {}) == want_defect and want_defect


def test_gen_barcoded_fastq_on_two_gpus(dataset, tmp_path, monkeypatch):
    """FastqProcessor(devices=[0, 1]): chunks dealt out over two GPUs, one NCCL all-reduce of the counts inside the
    process, records in input order and equal to the oracle; make_fastq(devices=...) writes the same bytes as one GPU."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least two GPUs")
    from precellar_b200 import Assay, _zstd
    from precellar_b200 import api
    from precellar_b200.api import FastqProcessor, make_fastq
    monkeypatch.setattr(api, "MIN_CHUNK_RECORDS", 700)                       # 5000 records -> 8 chunks, 4 per GPU
    proc = FastqProcessor(dataset["assay"], devices=[0, 1]).with_modality("rna").with_barcode_correct_prob(0.9)
    got = []
    n_batches = 0
    for batch in proc.gen_barcoded_fastq(True, chunk_size=60_000):          # several chunks per GPU
        got.extend(batch)
        n_batches += 1
    assert n_batches >= 4
    layout, ores = oracle_lines(dataset["assay"], dataset["full"], True, 0.9)
    want = [ln for ln in ores.lines if ln]
    assert len(got) == len(want) > 3000
    for rec, line in zip(got, want):
        f = line.split("\t")
        mine = [rec.barcode.raw.sequence, rec.barcode.raw.quality, rec.barcode.corrected if rec.barcode.corrected is not None else b"*",
                rec.umi.sequence if rec.umi else b"", rec.umi.quality if rec.umi else b"",
                rec.read1.sequence if rec.read1 else b"", rec.read1.quality if rec.read1 else b"",
                rec.read2.sequence if rec.read2 else b"", rec.read2.quality if rec.read2 else b""]
        assert [m.decode("latin-1") for m in mine] == f
    per = ores.qc_per_file
    want_defect = {rid: per[r, 1] / per[r, 0] for r, rid in enumerate(("R1", "R2")) if per[r, 1] > 0}
    assert proc.get_fastq_qc().get("frac_reads_with_misformed_structure", {}) == want_defect
    # make_fastq: the forward-insert variant of the spec (R1.fq.zst only), one GPU against two
    td = dataset["td"]
    spec = write_spec(td, None, None, forward_insert=True)
    outs = []
    for k, devs in enumerate(([0], [0, 1])):
        a = Assay(spec)
        a.update_read("R1", fastq=[os.path.join(td, "R1_a.fq.gz"), os.path.join(td, "R1_b.fq")])
        a.update_read("R2", fastq=os.path.join(td, "R2.fq.gz"))
        out = str(tmp_path / f"out{k}")
        make_fastq(a, modality="rna", out_dir=out, correct_barcode=True, devices=devs)
        outs.append(_zstd.decompress_file(os.path.join(out, "R1.fq.zst")))
    assert outs[0] == outs[1] and len(outs[0]) > 100_000


@pytest.mark.parametrize("keep_text", [True, False])
@pytest.mark.parametrize("correct", [False, True])
def test_streamed_chunks_match_the_resident_run(dataset, monkeypatch, keep_text, correct):
    """Inputs larger than the device budget: the first chunks stay resident, the rest are counted through one reusable
    chunk and uploaded again in pass 2, recounted against the frozen counts (pcl_counts_freeze).  Records and QC equal
    the oracle's - the same as the all-resident run."""
    from precellar_b200 import api
    from precellar_b200.api import FastqProcessor
    monkeypatch.setattr(api, "MIN_CHUNK_RECORDS", 700)                       # 5000 records -> 8 chunks
    if not keep_text:
        monkeypatch.setattr(api, "KEEP_TEXT_BYTES", 0)                       # pass 2 re-reads the files
    proc = FastqProcessor(dataset["assay"]).with_modality("rna").with_barcode_correct_prob(0.9)
    probe = api.BarcodePipeline(api.compile_layout(dataset["assay"], "rna", require_files=True), device=0, qc=False)
    try:
        proc.resident_bytes = 3 * probe.chunk_bytes(700) + 1                 # three resident chunks, five streamed
    finally:
        probe.close()
    got = []
    for batch in proc.gen_barcoded_fastq(correct, chunk_size=60_000):
        got.extend(batch)
    assert proc.streamed_chunks == 5
    layout, ores = oracle_lines(dataset["assay"], dataset["full"], correct, 0.9)
    want = [ln for ln in ores.lines if ln]
    assert len(got) == len(want) > 3000
    for rec, line in zip(got, want):
        f = line.split("\t")
        mine = [rec.barcode.raw.sequence, rec.barcode.raw.quality, rec.barcode.corrected if rec.barcode.corrected is not None else b"*",
                rec.umi.sequence if rec.umi else b"", rec.umi.quality if rec.umi else b"",
                rec.read1.sequence if rec.read1 else b"", rec.read1.quality if rec.read1 else b"",
                rec.read2.sequence if rec.read2 else b"", rec.read2.quality if rec.read2 else b""]
        assert [m.decode("latin-1") for m in mine] == f
    qc = proc.get_fastq_qc()
    per = ores.qc_per_file
    want_defect = {rid: per[r, 1] / per[r, 0] for r, rid in enumerate(("R1", "R2")) if per[r, 1] > 0}
    assert qc.get("frac_reads_with_misformed_structure", {}) == want_defect and want_defect
    cat = ores.qc_cat.astype(float)
    for j, nm in enumerate(("umi", "barcode", "read1", "read2")):
        if cat[j, 1] > 0:
            assert qc["frac_q30_bases"][nm] == cat[j, 2] / cat[j, 1]


WRITERS = [("host", "zstd", 262144), ("device", "zstd", 262144), ("device", "device", 262144), ("device", "device", 700)]


@pytest.mark.parametrize("writer,compression,chunk_records", WRITERS)
def test_make_fastq(dataset, tmp_path, monkeypatch, writer, compression, chunk_records):
    """R1.fq.zst = corrected barcode + UMI + read1 (python/src/lib.rs:153-163); the insert is read on the pos strand.
    Host writer (records assembled on the host, libzstd level 9) and device writer (records assembled in HBM; zstd
    frames from libzstd or built on the GPU; one chunk or eight) must decompress to the same bytes."""
    import copy
    from precellar_b200 import Assay, _zstd, api, compile_layout
    from precellar_b200.api import make_fastq
    monkeypatch.setattr(api, "MIN_CHUNK_RECORDS", chunk_records)
    monkeypatch.setattr(api, "MAKE_FASTQ_CHUNK_BASES", 1_000_000 if chunk_records > 700 else 60_000)
    td = dataset["td"]
    spec = write_spec(td, None, None, forward_insert=True)           # whitelists are already on disk; only the YAML differs
    a = Assay(spec)
    a.update_read("R1", fastq=[os.path.join(td, "R1_a.fq.gz"), os.path.join(td, "R1_b.fq")])
    a.update_read("R2", fastq=os.path.join(td, "R2.fq.gz"))
    out = str(tmp_path / "out")
    make_fastq(a, modality="rna", out_dir=out, correct_barcode=True, writer=writer, compression=compression)
    layout = compile_layout(a, "rna", require_files=True)
    assert [p.is_reverse for p in layout.reads] == [False, False]
    ores, _ = U.run_oracle(layout, dataset["full"], correct=True, threshold=0.975)
    assert not os.path.exists(os.path.join(out, "R2.fq.zst"))      # is_paired_end needs targets on both strands (fastq.rs:314-329)
    r1 = _zstd.decompress_file(os.path.join(out, "R1.fq.zst")).split(b"\n")
    recs = [r1[i:i + 4] for i in range(0, len(r1) - 1, 4)]
    want = []
    for i, ln in enumerate(ores.lines):
        if not ln:
            continue
        f = ln.split("\t")
        if f[2] == "*":
            continue                                                # dropped: barcode not corrected
        want.append((dataset["names1"][i], (f[2] + f[3] + f[5]).encode("latin-1"), (f[1] + f[4] + f[6]).encode("latin-1")))
    assert len(recs) == len(want) > 1000
    for r, (nm, s, q) in zip(recs, want):
        assert r[0] == b"@" + stripped(nm) and r[1] == s and r[2] == b"+" and r[3] == q
    assert recs[0][0].startswith(b"@read") and b"/1" not in recs[0][0]


@pytest.mark.parametrize("writer", ["host", "device"])
def test_make_fastq_without_read1_fails_like_the_reference(dataset, tmp_path, writer):
    """sci-RNA-seq3 has no pos-strand insert: `record.read1.unwrap()` panics in the reference (python/src/lib.rs:160)."""
    from precellar_b200.api import make_fastq
    with pytest.raises(RuntimeError) as ei:
        make_fastq(dataset["assay"], modality="rna", out_dir=str(tmp_path / "o2"), correct_barcode=False, writer=writer)
    assert "read1" in str(ei.value)


@pytest.mark.parametrize("device_parse", [False, True])
def test_mismatched_names_are_rejected(dataset, tmp_path, device_parse):
    import copy
    from precellar_b200 import _ffi
    from precellar_b200.api import FastqProcessor
    from precellar_b200.pipeline import HostBatch
    bad = os.path.join(str(tmp_path), "R2_bad.fq")
    full, n = dataset["full"], 50
    names = [f"other{i}/2".encode() for i in range(n)]
    write_fastq(bad, names, HostBatch(full[1].seq[:n], full[1].qual[:n], full[1].length[:n]), False)
    a = copy.deepcopy(dataset["assay"])
    a.update_read("R2", fastq=bad, min_len=1, max_len=60)
    with pytest.raises(_ffi.PclError) as ei:
        for _ in FastqProcessor(a, device_parse=device_parse).with_modality("rna").gen_barcoded_fastq(True):
            pass
    assert ei.value.code == _ffi.PCL_ERR_INPUT


@pytest.mark.parametrize("writer,compression,chunk_records,container",
                         [w + (True if w[2] > 700 else False,) for w in WRITERS] +       # the multi-chunk case reads plain files (memory-mapped input)
                         [("device", "device", 262144, "bgzf"), ("device", "zstd", 700, "bgzf"), ("host", "zstd", 262144, "bgzf")])
def test_make_fastq_paired_end_atac_like(tmp_path, monkeypatch, writer, compression, chunk_records, container):
    """Paired-end layout (targets on both strands): R1.fq.zst and R2.fq.zst, neg-strand barcode read (I2).  Input as plain
    files, gzip files, or BGZF files (inflated on the device, csrc/inflate.cuh)."""
    from precellar_b200 import Assay, _zstd, api, compile_layout
    from precellar_b200.api import make_fastq
    monkeypatch.setattr(api, "MIN_CHUNK_RECORDS", chunk_records)
    monkeypatch.setattr(api, "MAKE_FASTQ_CHUNK_BASES", 1_000_000 if chunk_records > 700 else 60_000)
    td = str(tmp_path)
    rng = np.random.default_rng(21)
    wl = U.gen_whitelist(rng, 200, 16)
    open(os.path.join(td, "wl.txt"), "wb").write(b"\n".join(wl) + b"\n")
    regions = [
        region("p5", "illumina_p5", "fixed", "AATGATACGGCGACCACCGAGATCTACAC", 29, 29),
        region("cell_bc", "barcode", "onlist", "N" * 16, 16, 16, os.path.join(td, "wl.txt")),
        region("r1p", "nextera_read1", "fixed", "TCGTCGGCAGCGTCAGATGTGTATAAGAGACAG", 33, 33),
        region("gdna", "gdna", "random", "X", 1, 1000),
        region("r2p", "nextera_read2", "fixed", "CTGTCTCTTATACACATCTCCGAGCCCACGAGAC", 34, 34),
        region("p7", "illumina_p7", "fixed", "ATCTCGTATGCCGTCTTCTGCTTG", 24, 24),
    ]
    spec = dict(seqspec_version="0.3.0", assay_id="atac-test", name="x", doi="", date="", description="", modalities=["atac"],
                lib_struct="", library_protocol="x", library_kit="x", sequence_protocol="x", sequence_kit="x", sequence_spec=[],
                library_spec=[dict(region_id="atac", region_type="atac", name="atac", sequence_type="joined", sequence="",
                                   min_len=0, max_len=0, onlist=None, regions=regions)])
    path = os.path.join(td, "atac.yaml")
    yaml.safe_dump(spec, open(path, "w"), sort_keys=False)
    a = Assay(path)
    a.add_illumina_reads("atac", forward_strand_workflow=False)
    assert sorted(a.sequence_spec) == ["atac-I2", "atac-R1", "atac-R2"]
    for rid, ln in (("atac-R1", 40), ("atac-R2", 40)):
        a.update_read(rid, min_len=ln, max_len=ln)
    layout0 = compile_layout(a, "atac")
    n = 3000
    full = U.gen_reads(rng, layout0, n, [64] * 3, p_short=0.0, read_len=[p.max_len for p in layout0.reads])
    for b in full:
        b.length[:] = np.maximum(b.length, 1)
    order = [p.read_id for p in layout0.reads]
    for r, rid in enumerate(order):
        names = [f"q{i}/{1 if rid != 'atac-R2' else 2}".encode() for i in range(n)]
        gz = container
        write_fastq(os.path.join(td, rid + (".fq.gz" if gz else ".fq")), names, full[r], gz)
        a.update_read(rid, fastq=os.path.join(td, rid + (".fq.gz" if gz else ".fq")))
    out = os.path.join(td, "out")
    make_fastq(a, modality="atac", out_dir=out, correct_barcode=True, writer=writer, compression=compression)
    layout = compile_layout(a, "atac", require_files=True)
    ores, _ = U.run_oracle(layout, full, correct=True, threshold=0.975)
    r1 = _zstd.decompress_file(os.path.join(out, "R1.fq.zst")).split(b"\n")
    r2 = _zstd.decompress_file(os.path.join(out, "R2.fq.zst")).split(b"\n")
    want1, want2 = [], []
    for ln in ores.lines:
        if not ln:
            continue
        f = ln.split("\t")
        if f[2] == "*":
            continue
        want1.append(((f[2] + f[3] + f[5]).encode("latin-1"), (f[1] + f[4] + f[6]).encode("latin-1")))
        want2.append((f[7].encode("latin-1"), f[8].encode("latin-1")))
    assert all(r1[i] == r2[i] and r1[i].startswith(b"@q") and b"/" not in r1[i] for i in range(0, len(r1) - 1, 4))   # /1 and /2 stripped
    got1 = [(r1[i + 1], r1[i + 3]) for i in range(0, len(r1) - 1, 4)]
    got2 = [(r2[i + 1], r2[i + 3]) for i in range(0, len(r2) - 1, 4)]
    assert got1 == want1 and got2 == want2 and len(want1) > 1500


def test_update_read_verifies_a_sample_on_the_device(dataset, tmp_path, caplog):
    """Assay::verify (lib.rs:481-548): valid records and onlist matches of a sample, split with tolerances (0.2, 0.2)
    on the CUDA path; a read whose linker does not match the spec is reported."""
    import copy
    import logging
    a = copy.deepcopy(dataset["assay"])
    r1 = a.sequence_spec["R1"]
    got = a._verify(r1, 2000)
    full = dataset["full"]
    assert got is not None and got["total"] == 2000
    # the split of R1 with the (1.0, 0.2) tolerances of the hot path is at least as permissive as (0.2, 0.2)
    layout, ores = oracle_lines(dataset["assay"], full, False)
    ok_ref = float((ores.fstatus[0][:2000] == 0).mean()) * 100.0
    assert 50.0 < got["percent_valid"] <= ok_ref + 1e-9
    assert set(got["percent_matched"]) == {"hairpin_barcode", "rt_barcode"} and all(v > 50.0 for v in got["percent_matched"].values())
    # a spec whose linker is wrong: (nearly) nothing is valid any more and a warning is logged
    bad = copy.deepcopy(dataset["assay"])
    bad._region_map["linker"].sequence = "TTTTTT"
    with caplog.at_level(logging.WARNING):
        res = bad._verify(bad.sequence_spec["R1"], 2000)
    assert res["percent_valid"] < 10.0
    assert any("low percentage of valid records" in m for m in caplog.messages)


def test_make_fastq_auto_writer_falls_back_to_the_host_path(dataset, tmp_path, monkeypatch):
    """writer="auto": input the device parser rejects (blank lines between records: the host reader skips them) takes the
    host path, writer="device" reports the record; input that does not fit the device budget is streamed."""
    from precellar_b200 import Assay, _ffi, _zstd
    from precellar_b200.api import make_fastq
    td = dataset["td"]
    spec = write_spec(td, None, None, forward_insert=True)
    full, n = dataset["full"], 400
    r1, r2 = str(tmp_path / "R1_blank.fq"), str(tmp_path / "R2_blank.fq")
    for path, names, b in ((r1, dataset["names1"], full[0]), (r2, dataset["names2"], full[1])):
        with open(path, "wb") as fh:
            for i in range(n):
                L = int(b.length[i])
                fh.write(b"@" + names[i] + b"\n" + b.seq[i, :L].tobytes() + b"\n+\n" + b.qual[i, :L].tobytes() + b"\n")
                if i % 7 == 3:
                    fh.write(b"\n")                                   # stray blank line
    outs = {}
    for writer in ("host", "auto"):
        a = Assay(spec)
        a.update_read("R1", fastq=r1, min_len=34, max_len=34)
        a.update_read("R2", fastq=r2, min_len=40, max_len=60)
        out = str(tmp_path / writer)
        make_fastq(a, modality="rna", out_dir=out, correct_barcode=True, writer=writer)
        outs[writer] = _zstd.decompress_file(os.path.join(out, "R1.fq.zst"))
    assert outs["host"] == outs["auto"] and outs["host"].count(b"\n") > 400
    a = Assay(spec)
    a.update_read("R1", fastq=r1, min_len=34, max_len=34)
    a.update_read("R2", fastq=r2, min_len=40, max_len=60)
    with pytest.raises(_ffi.PclError):
        make_fastq(a, modality="rna", out_dir=str(tmp_path / "dev"), correct_barcode=True, writer="device")
    # device budget of 1 KB: the text cannot stay resident -> the device writer streams (two reads of the files, one chunk
    # of device memory), over several chunks here; same bytes as the host path
    from precellar_b200 import api
    monkeypatch.setattr(api, "MIN_CHUNK_RECORDS", 700)
    monkeypatch.setattr(api, "MAKE_FASTQ_CHUNK_BASES", 60_000)
    b = Assay(spec)
    b.update_read("R1", fastq=[os.path.join(td, "R1_a.fq.gz"), os.path.join(td, "R1_b.fq")])
    b.update_read("R2", fastq=os.path.join(td, "R2.fq.gz"))
    make_fastq(b, modality="rna", out_dir=str(tmp_path / "ref2"), correct_barcode=True, writer="host")
    want = _zstd.decompress_file(str(tmp_path / "ref2" / "R1.fq.zst"))
    monkeypatch.setenv("PCL_API_RESIDENT_BYTES", "1024")
    for k, comp in enumerate(("zstd", "device")):
        make_fastq(b, modality="rna", out_dir=str(tmp_path / f"dev{k}"), correct_barcode=True, writer="device", compression=comp)
        assert api.LAST_RUN["writer"] == "device" and api.LAST_RUN["streamed_chunks"] >= 7
        assert _zstd.decompress_file(str(tmp_path / f"dev{k}" / "R1.fq.zst")) == want and len(want) > 100_000
