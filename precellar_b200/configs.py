"""Read layouts of the BASELINE configurations, as the reference derives them from its templates.

Each function returns the ``CompiledLayout`` that ``compile_layout(Assay(template), modality)`` yields after
the README's ``add_illumina_reads`` / ``update_read`` calls (derivations traced in SURVEY.md appendix A;
tests/test_seqspec.py checks them against the templates whenever the reference checkout is present).
They exist so that benchmarks and GPU tests do not need the reference's YAML files on the box.
"""
from __future__ import annotations

from .layout import CompiledLayout, layout_from_elems
from .seqspec import SegmentInfoElem as E


def v3_layout(whitelist, r2_len: int = 91) -> CompiledLayout:
    """10x scRNA-seq v3 (seqspec_templates/10x_rna_v3.yaml): rna-R1 = [B16 U12] pos, rna-R2 = [cDNA] neg."""
    return layout_from_elems([
        dict(read_id="rna-R1", max_len=28, elems=[
            E("rna-cell_barcode", "barcode", "onlist", "N" * 16, 16, 16),
            E("rna-umi", "umi", "random", "X" * 12, 12, 12)]),
        dict(read_id="rna-R2", is_reverse=True, max_len=r2_len, elems=[
            E("rna-cDNA", "cdna", "random", "X", 1, 1000)]),
    ], {"rna-cell_barcode": whitelist}, "rna")


def atac_layout(whitelist, insert_len: int = 50) -> CompiledLayout:
    """10x scATAC-seq (10x_atac.yaml, forward_strand_workflow=False): atac-I2 = [B16] neg; R1 pos / R2 neg = [gDNA]."""
    return layout_from_elems([
        dict(read_id="atac-R1", max_len=insert_len, elems=[E("atac-gdna", "gdna", "random", "X", 1, 1000)]),
        dict(read_id="atac-I2", is_reverse=True, max_len=16, elems=[
            E("atac-cell_barcode", "barcode", "onlist", "N" * 16, 16, 16)]),
        dict(read_id="atac-R2", is_reverse=True, max_len=insert_len, elems=[E("atac-gdna", "gdna", "random", "X", 1, 1000)]),
    ], {"atac-cell_barcode": whitelist}, "atac")


def multiome_atac_layout(whitelist, insert_len: int = 50) -> CompiledLayout:
    """10x multiome ATAC side (10x_rna_atac.yaml): atac-I2 = [linker8 fixed, B16] neg (24 nt)."""
    return layout_from_elems([
        dict(read_id="atac-R1", max_len=insert_len, elems=[E("atac-gdna", "gdna", "random", "X", 1, 1000)]),
        dict(read_id="atac-I2", is_reverse=True, max_len=24, elems=[
            E("atac-linker", "linker", "fixed", "CGCGTCTG", 8, 8),
            E("atac-cell_barcode", "barcode", "onlist", "N" * 16, 16, 16)]),
        dict(read_id="atac-R2", is_reverse=True, max_len=insert_len, elems=[E("atac-gdna", "gdna", "random", "X", 1, 1000)]),
    ], {"atac-cell_barcode": whitelist}, "atac")


def sci3_layout(hp_whitelist, rt_whitelist, r1_len: int = 34, r2_len: int = 100) -> CompiledLayout:
    """sci-RNA-seq3 (sci_rna_seq3.yaml): R1 = [hairpin B 9-10, linker CAGAGC, UMI 8, RT B10] pos, R2 = [cDNA] neg."""
    return layout_from_elems([
        dict(read_id="R1", max_len=r1_len, elems=[
            E("hairpin_barcode", "barcode", "onlist", "N" * 10, 9, 10),
            E("linker", "linker", "fixed", "CAGAGC", 6, 6),
            E("umi", "umi", "random", "N" * 8, 8, 8),
            E("rt_barcode", "barcode", "onlist", "N" * 10, 10, 10)]),
        dict(read_id="R2", is_reverse=True, max_len=r2_len, elems=[E("cdna", "cdna", "random", "X", 1, 1000)]),
    ], {"hairpin_barcode": hp_whitelist, "rt_barcode": rt_whitelist}, "rna")
