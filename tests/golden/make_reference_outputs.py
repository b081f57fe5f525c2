"""Regenerates tests/golden/reference_outputs.json.gz from the reference's own golden OUTPUT files.

precellar/src/align/fastq.rs:694-726 (`test_fastq`) compares `gen_barcoded_fastq(false, 500)` on
data/test{2,3,4}.yaml with data/test{2,3,4}.out.gz (250 lines each: raw barcode, UMI, read1, read2).  The input FASTQ
files are not in the reference tree, so the test cannot run there -- but every output line determines the
barcode / UMI / insert bytes of its record, and the layouts fix where they sit.  This script rebuilds input records
that carry exactly those bytes (fixed sequences from the spec, 'A' filler where the spec says random), compiles the
three layouts from the reference's YAML files with this repository's seqspec model, and stores layouts, whitelists,
rebuilt inputs and the reference's lines verbatim.  tests/test_reference_outputs.py then requires the oracle, the
host emulation of the device code and the CUDA path to reproduce the reference's lines.

What this pins: seqspec YAML -> region table (3 real assays), split on a fixed multi-barcode layout (SHARE-seq),
on two anchor layouts (dscATAC with a 0-4 nt phase block, sci-RNA-seq3 with a 9|10 nt hairpin barcode), barcode
concatenation across segments, read1/read2 assignment and the dropping rules.  It cannot pin behaviour on bytes the
outputs do not show (mutated linkers, qualities).

    python tests/golden/make_reference_outputs.py        (needs /root/reference)
"""
import gzip
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from precellar_b200 import Assay, compile_layout  # noqa: E402

DATA = "/root/reference/precellar/data"


def load(spec):
    a = Assay(os.path.join(DATA, spec))
    m = a.modalities()[0]
    for r in a._region_map.values():                  # onlists ship next to the spec
        if r.onlist is not None and os.path.exists(os.path.join(DATA, r.onlist.filename)):
            r.onlist.url, r.onlist.urltype = os.path.join(DATA, r.onlist.filename), "local"
    if spec == "test3.yaml":
        a.delete_read("I1")                           # get_segments is None for it: the reference panics too (lib.rs:422-424)
    return a, m, compile_layout(a, m)


def lines_of(name):
    with gzip.open(os.path.join(DATA, name), "rt") as fh:
        return [ln.rstrip("\n") for ln in fh]


def rebuild(spec, layout, line):
    """Input records (one per read of the layout) that make the reference print `line`."""
    bc, umi, r1, r2 = line.split("\t")
    out = {}
    for plan in layout.reads:
        kinds = [e.region_type for e in plan.elems]
        if kinds in (["gdna"], ["cdna"]):                         # insert-only read: pos strand -> read1, neg -> read2
            out[plan.read_id] = r2 if plan.is_reverse else r1
            continue
        s, rest_bc = "", bc
        if spec == "test3.yaml":                                   # dscATAC: the phase block length follows from the read length
            total = 118
            fixed = sum(e.max_len for e in plan.elems if e.min_len == e.max_len)
            phase = total - fixed - len(r1)
            assert 0 <= phase <= 4, (phase, line)
        for e in plan.elems:
            if e.region_type == "barcode":
                n = e.max_len if e.min_len == e.max_len else len(bc) - sum(x.max_len for x in plan.elems
                                                                          if x.region_type == "barcode" and x is not e)
                s, rest_bc = s + rest_bc[:n], rest_bc[n:]
            elif e.region_type == "umi":
                s += umi
            elif e.region_type in ("gdna", "cdna"):
                s += r1
            elif e.sequence_type == "fixed":
                s += (e.sequence + "A" * e.max_len)[:e.max_len]     # the linker regions of test2 are longer than their sequence
            else:
                s += "A" * phase                                    # test3's phase block
        assert rest_bc == "", (spec, line)
        if plan.max_len and len(s) < plan.max_len and spec == "test4.yaml":
            s += "T" * (plan.max_len - len(s))                      # the poly-T that follows the RT barcode
        out[plan.read_id] = s
    return out


def main():
    doc = {}
    for spec, outs in (("test2.yaml", "test2.out.gz"), ("test3.yaml", "test3.out.gz"), ("test4.yaml", "test4.out.gz")):
        a, m, layout = load(spec)
        lines = lines_of(outs)
        recs = [rebuild(spec, layout, ln) for ln in lines]
        doc[spec.split(".")[0]] = dict(
            modality=layout.modality,
            reads=[dict(read_id=p.read_id, is_reverse=p.is_reverse, max_len=p.max_len,
                        elems=[dict(region_id=e.region_id, region_type=e.region_type, sequence_type=e.sequence_type,
                                    sequence=e.sequence, min_len=e.min_len, max_len=e.max_len) for e in p.elems])
                   for p in layout.reads],
            region_ids=layout.region_ids,
            whitelists={k: [x.decode() for x in v] for k, v in layout.whitelists.items()},
            inputs={p.read_id: [r[p.read_id] for r in recs] for p in layout.reads},
            reference_lines=lines)
        print(spec, len(lines), "lines;", {p.read_id: sorted({len(r[p.read_id]) for r in recs}) for p in layout.reads})
    path = os.path.join(ROOT, "tests", "golden", "reference_outputs.json.gz")
    with gzip.GzipFile(path, "wb", mtime=0) as fh:
        fh.write(json.dumps(doc, sort_keys=True).encode())
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
