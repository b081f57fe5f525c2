"""Pin the oracle's split() on the reference's own known-answer vectors.

Vectors transcribed verbatim from seqspec/src/read/segment.rs:617-705 (test_fixed, test_variable);
helper constructors mirror segment.rs:565-615.  Spaces in reads are removed like new_fq() does.
"""
import pytest

from oracle import pyoracle as O
from oracle.pyoracle import Elem


def bc(n):
    return Elem(O.BARCODE, n, n, False, "N", 0, "bc")


def umi(n):
    return Elem(O.UMI, n, n, False, "N", -1, "umi")


def cdna(a, b):
    return Elem(O.TARGET, a, b, False, "N", -1, "cdna")


def poly(a, b):
    return Elem(O.OTHER, a, b, False, "N", -1, "polya")


def linker(s):
    return Elem(O.OTHER, len(s), len(s), True, s, -1, "linker")


def var(a, b):
    return Elem(O.OTHER, a, b, False, "N", -1, "var")


VAR4 = [bc(4), var(1, 3), umi(2)]

GOLDEN = [
    # test_fixed (segment.rs:617-651)
    ("AAAAAA", [cdna(1, 5), linker("ATGT")], "[T]AAAAA"),
    ("AAAA TT CCC", [bc(4), umi(2), cdna(2, 4), poly(2, 4)], "[B]AAAA,[U]TT,[T]CCC"),
    ("AATT", [bc(4), umi(2), cdna(2, 4), poly(2, 4)], "[B]AATT"),
    ("AATT GG T", [bc(4), umi(2), cdna(2, 4), poly(2, 4)], "[B]AATT,[U]GG"),
    ("AAAA TT CCG GG", [bc(4), umi(2), cdna(2, 4), poly(2, 4)], "[B]AAAA,[U]TT,[TO]CCGGG"),
    ("AAAA TT CCCCCCC GGGGGGG AAAA TT", [bc(4), umi(2), cdna(2, 10), poly(2, 10), bc(4), umi(2)],
     "[B]AAAA,[U]TT,[TO]CCCCCCCGGGGGGGAAAATT"),
    # test_variable (segment.rs:653-706)
    ("AAAA A TT ACTG CCC", VAR4 + [linker("ACTG")], "[B]AAAA,[O]A,[U]TT,[O]ACTG"),
    ("AAAA A TT ACCGG CCCACTGG", VAR4 + [linker("ACTGG")], "[B]AAAA,[O]A,[U]TT,[O]ACCGG"),
    ("AAAA A TT ACCGG CCCACTG", VAR4 + [linker("ACTGG"), cdna(10, 100)], "[B]AAAA,[O]A,[U]TT,[O]ACCGG"),
    ("AAAA A TT ACCGG CCCACTG", VAR4 + [linker("ACTGG"), cdna(1, 100)],
     "[B]AAAA,[O]A,[U]TT,[O]ACCGG,[T]CCCACTG"),
    ("AAAA A TT ACCGG CC CA GGGG TTTT AT GGGG ACTGGGGGACTG",
     VAR4 + [linker("ACTGG"), var(2, 4), bc(2), linker("GGGG"), var(2, 4), bc(2), linker("GGGG"), cdna(1, 100)],
     "[B]AAAA,[O]A,[U]TT,[O]ACCGG,[O]CC,[B]CA,[O]GGGG,[O]TTTT,[B]AT,[O]GGGG,[T]ACTGGGGGACTG"),
]


@pytest.mark.parametrize("read,elems,expected", GOLDEN)
def test_split_golden(read, elems, expected):
    seq = "".join(read.split()).encode()
    st, text = O.split_str(elems, seq)
    assert st == O.SPLIT_OK
    assert text == expected


def test_split_errors():
    # interior anchor too dissimilar -> Err(PatternMismatch) (segment.rs:306-308)
    elems = VAR4 + [linker("ACTGG"), cdna(1, 100)]
    st, _ = O.split_str(elems, b"AAAAATTTTTTTTTTTTTTTTTTTTTTT")
    assert st == O.SPLIT_PATTERN_MISMATCH
    # needle of length 1 without an exact hit -> (None, 1) -> Err(AnchorNotFound) (segment.rs:508-512,309-311)
    st, _ = O.split_str([bc(2), var(1, 3), linker("G"), cdna(1, 50)], b"AATTTTTTTTTT")
    assert st == O.SPLIT_ANCHOR_NOT_FOUND


def test_truncate_max():
    # segment.rs:144-164 ; 10x v3 R1: [B16, U12, polyT 10-250, cDNA 1-1000] at 28 -> 2
    v3 = [bc(16), umi(12), poly(10, 250), cdna(1, 1000)]
    assert O.truncate_max(v3, 28) == 2
    assert O.truncate_max(v3, 38) == 3          # 28 + min 10 <= 38 keeps the poly run, then stops
    assert O.truncate_max([cdna(1, 1000)], 91) == 1
    assert O.truncate_max([bc(4), var(2, 4), bc(2), linker("GGGG"), cdna(1, 1000)], 50) == 5
    assert O.truncate_max([bc(16)], 8) == 0


def test_rev_compl_variants():
    # seqspec: upper-case only (seqspec/src/utils.rs:176-187); precellar: case-insensitive (utils.rs:138-149)
    assert O.rev_compl_seqspec(b"ACGTNacgt") == b"tgcaNACGT"
    assert O.rev_compl_precellar(b"ACGTNacgt") == b"ACGTNACGT"


def test_reverse_anchor_quirk():
    # neg-strand reads search the anchor in the rev-complemented window but apply the position as a
    # forward offset (segment.rs:260-272).
    elems = [bc(4), var(1, 3), linker("AAC")]
    # window = read[5..min(len, 7+3)] = read[5..10]; rev_compl(window) is searched for "AAC"
    seq = b"CCCCTGTTGG"
    st, text = O.split_str(elems, seq, is_reverse=True)
    assert st == O.SPLIT_OK
    # window "GTTGG" -> revcomp "CCAAC": exact at pos 2 -> offset_right = 5+2 = 7 -> var = read[4..7]
    assert text == "[B]CCCC,[O]TGT,[O]TGG"
