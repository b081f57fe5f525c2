"""CPU tier: the seqspec model (Assay, add_illumina_reads, get_segments, whitelists) and the region-table
compiler, against the reference's own templates when the checkout is present, and against hand-built specs."""
import gzip
import os

import pytest
import yaml

from precellar_b200 import Assay, compile_layout, configs
from precellar_b200.seqspec import Onlist, SegmentInfo, SegmentInfoElem, read_lines

TEMPLATES = "/root/reference/seqspec_templates"
DATA = "/root/reference/precellar/data"
have_ref = pytest.mark.skipif(not os.path.isdir(TEMPLATES), reason="reference checkout not present")


def localize(assay, tmp_path, lists):
    """Point remote onlists at local files (the benchmarks do the same)."""
    for rid, items in lists.items():
        p = os.path.join(str(tmp_path), rid + ".txt")
        open(p, "wb").write(b"\n".join(items) + b"\n")
        assay._region_map[rid].onlist = Onlist(rid, rid + ".txt", "txt", 0, p, "local", "local", "0")


def sig(layout):
    return [(p.read_id, p.is_reverse, p.max_len,
             [(e.region_id, e.region_type, e.sequence_type, e.min_len, e.max_len) for e in p.elems], p.region_slots)
            for p in layout.reads]


@have_ref
def test_all_templates_parse():
    for f in sorted(os.listdir(TEMPLATES)):
        a = Assay(os.path.join(TEMPLATES, f))
        assert a.modalities()
        assert repr(a).splitlines()[0] == a.assay_id
        a2 = yaml.safe_load(a.to_yaml())
        assert a2["assay_id"] == a.assay_id


@have_ref
def test_v3_layout_from_template(tmp_path):
    a = Assay(os.path.join(TEMPLATES, "10x_rna_v3.yaml"))
    wl = [b"ACGTACGTACGTACGT", b"TTTTGGGGCCCCAAAA"]
    localize(a, tmp_path, {"rna-cell_barcode": wl})
    a.add_illumina_reads("rna")
    assert list(a.sequence_spec) == ["rna-R1", "rna-R2", "rna-I1"]
    i1 = a.sequence_spec["rna-I1"]
    assert (i1.primer_id, i1.strand, i1.min_len, i1.max_len) == ("rna-truseq_read2", "pos", 8, 8)
    # update_read(fastq=...) replaces the lengths by what the files hold: 28 / 91 in the synthetic data
    a.update_read("rna-R1", min_len=28, max_len=28)
    a.update_read("rna-R2", min_len=91, max_len=91)
    got = compile_layout(a, "rna")
    want = configs.v3_layout(wl, 91)
    assert sig(got) == sig(want)                       # rna-I1 holds only index7: no annotator (fastq.rs:460-471)
    assert got.whitelists["rna-cell_barcode"] == wl and got.contains_umi


@have_ref
def test_atac_layouts_from_template(tmp_path):
    a = Assay(os.path.join(TEMPLATES, "10x_atac.yaml"))
    wl = [b"ACGTACGTACGTACGT"]
    bc = [r for r in a.library_spec["atac"].subregions if r.is_barcode()][0].region_id
    localize(a, tmp_path, {bc: wl})
    a.add_illumina_reads("atac", forward_strand_workflow=False)
    i2 = a.sequence_spec["atac-I2"]
    assert (i2.strand, i2.min_len, i2.max_len) == ("neg", 16, 16)
    for rid in ("atac-R1", "atac-R2"):
        a.update_read(rid, min_len=50, max_len=50)
    got = compile_layout(a, "atac")
    want = configs.atac_layout(wl, 50)
    assert [(p.is_reverse, p.max_len, [(e.region_type, e.min_len, e.max_len) for e in p.elems]) for p in got.reads] == \
           [(p.is_reverse, p.max_len, [(e.region_type, e.min_len, e.max_len) for e in p.elems]) for p in want.reads]
    # forward-strand workflow: I2 is read from P5, forward (lib.rs:240-255)
    a.add_illumina_reads("atac", forward_strand_workflow=True)
    i2 = a.sequence_spec["atac-I2"]
    assert (i2.strand, i2.primer_id, i2.max_len) == ("pos", [r for r in a.library_spec["atac"].subregions if r.region_type == "illumina_p5"][0].region_id, 16)


@have_ref
def test_multiome_atac_layout_from_template(tmp_path):
    a = Assay(os.path.join(TEMPLATES, "10x_rna_atac.yaml"))
    assert sorted(a.modalities()) == ["atac", "rna"]
    for m in ("rna", "atac"):
        for r in a.library_spec[m].subregions:
            if r.is_barcode():
                localize(a, tmp_path, {r.region_id: [b"ACGTACGTACGTACGT"]})
    a.add_illumina_reads("atac", forward_strand_workflow=False)
    i2 = a.sequence_spec["atac-I2"]
    assert (i2.strand, i2.max_len) == ("neg", 24)      # 16 + the 8 nt linker (get_length(acc, true), lib.rs:176-185,257)
    got = compile_layout(a, "atac")
    plan = [p for p in got.reads if p.read_id == "atac-I2"][0]
    assert [(e.region_type, e.sequence_type, e.sequence, e.max_len) for e in plan.elems] == \
           [("linker", "fixed", "CGCGTCTG", 8), ("barcode", "onlist", "N" * 16, 16)]


@have_ref
def test_sci3_layout_and_fixture_whitelists(tmp_path):
    a = Assay(os.path.join(TEMPLATES, "sci_rna_seq3.yaml"))
    hp = read_lines(os.path.join(DATA, "sci_RNA_seq3-hp.txt.gz"))
    rt = read_lines(os.path.join(DATA, "sci_RNA_seq3-rt.txt.gz"))
    assert sorted({len(x) for x in hp}) == [9, 10] and len(hp) == 384 and {len(x) for x in rt} == {10} and len(rt) == 384
    localize(a, tmp_path, {"hairpin_barcode": hp, "rt_barcode": rt})
    a.update_read("R1", min_len=34, max_len=34)
    a.update_read("R2", min_len=100, max_len=100)
    got = compile_layout(a, "rna")
    want = configs.sci3_layout(hp, rt, 34, 100)
    assert sig(got) == sig(want)
    assert list(got.whitelists) == ["hairpin_barcode", "rt_barcode"]
    # the SHARE-seq fixtures have no final newline (BufRead::lines semantics)
    assert len(read_lines(os.path.join(DATA, "share_seq-round_1_bc.txt.gz"))) == 192


@have_ref
def test_reference_test_specs_parse():
    for f in ("test2.yaml", "test3.yaml", "test4.yaml"):
        a = Assay(os.path.join(DATA, f))
        m = a.modalities()[0]
        for r in a._region_map.values():                  # remote onlists whose file ships next to the spec
            if r.onlist is not None and r.onlist.urltype != "local" and os.path.exists(os.path.join(DATA, r.onlist.filename)):
                r.onlist.url, r.onlist.urltype = os.path.join(DATA, r.onlist.filename), "local"
        if f == "test3.yaml":
            # I1 reads 8 nt leftwards from the last region: nothing fits, get_segments is None and
            # get_segments_by_modality panics in the reference as well (seqspec/src/lib.rs:422-424)
            with pytest.raises(RuntimeError):
                a.get_segments_by_modality(m)
            a.delete_read("I1")
        segs = a.get_segments_by_modality(m)
        assert segs and all(len(info) > 0 for _, info in segs)
        wl = a.get_whitelists(m)
        assert all(len(v) > 0 for v in wl.values())


def test_truncate_max_and_segments(tmp_path):
    E = SegmentInfoElem
    info = SegmentInfo([E("b", "barcode", "onlist", "", 16, 16), E("u", "umi", "random", "", 12, 12),
                        E("t", "poly_t", "random", "", 10, 250), E("c", "cdna", "random", "", 1, 1000)], False)
    assert [e.region_id for e in info.truncate_max(28)] == ["b", "u"]
    assert [e.region_id for e in info.truncate_max(38)] == ["b", "u", "t"]
    assert [e.region_id for e in info.truncate_max(15)] == []


def test_structure_validation_and_lenient_yaml(tmp_path):
    leaf = dict(region_id="bc", region_type="BARCODE", name="x", sequence_type="onlist", sequence="NNNN", min_len=4,
                max_len=4, onlist=None, regions=None, parent_id="rna")
    primer = dict(region_id="p", region_type="TruSeq_Read1", name="p", sequence_type="fixed", sequence="ACGT", min_len=4,
                  max_len=4, onlist=None, regions=None)
    top = dict(region_id="rna", region_type="RNA", name="rna", sequence_type="joined", sequence="", min_len=0, max_len=0,
               onlist=None, regions=[primer, leaf])
    doc = dict(seqspec_version="0.3.0", assay_id="t", name="t", doi="", date="", description="", modalities=["RNA"],
               lib_struct="", library_protocol="x", library_kit="x", sequence_protocol="x", sequence_kit="x",
               sequence_spec=[], library_spec=[top])
    p = tmp_path / "ok.yaml"
    p.write_text("!Assay\n" + yaml.safe_dump(doc, sort_keys=False))
    a = Assay(str(p))
    assert a.modalities() == ["rna"] and a._region_map["bc"].is_barcode() and a._region_map["p"].is_sequencing_primer()
    a.update_read("R1", modality="rna", primer_id="p", is_reverse=False, min_len=4, max_len=4)
    assert [e.region_id for e in a.get_segments("R1")] == ["bc"]
    assert a.get_whitelists("rna") == {"bc": []}
    with pytest.raises(ValueError):
        a.update_read("R9", modality="rna", primer_id="nope", is_reverse=False, min_len=4, max_len=4)
    # nesting depth must be exactly modality -> leaf regions (lib.rs:75-90)
    deep = dict(top, regions=[dict(leaf, regions=[dict(leaf, region_id="inner")], sequence_type="joined")])
    p2 = tmp_path / "deep.yaml"
    p2.write_text(yaml.safe_dump(dict(doc, library_spec=[deep]), sort_keys=False))
    with pytest.raises(ValueError):
        Assay(str(p2))
    # save() writes relative paths and the file loads back
    out = tmp_path / "saved.yaml"
    a.save(str(out))
    assert Assay(str(out)).sequence_spec["R1"].primer_id == "p"
    a.delete_read("R1")
    assert "R1" not in a.sequence_spec
    a.delete_region("bc")
    assert "bc" not in a._region_map


def test_onlist_reading_rules(tmp_path):
    p = tmp_path / "wl.txt.gz"
    with gzip.open(str(p), "wb") as fh:
        fh.write(b"AAAA\r\nCCCC\nAAAA\nGGGG")            # CRLF, duplicate, no final newline
    o = Onlist("x", "wl.txt.gz", "txt", 0, str(p), "local", None, "0")
    assert o.read() == [b"AAAA", b"CCCC", b"GGGG"]
    with pytest.raises(FileNotFoundError):
        Onlist("x", "nope-remote.txt", "txt", 0, "https://example.org/x", "https", "remote", "0").read()
