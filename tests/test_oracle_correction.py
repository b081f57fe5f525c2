"""Hand-computed cases for the counting / correction restatement.

The reference has no test for WhitelistBuilder / likelihood1 / correct_barcode ("parity unpinned",
SURVEY 8c); these cases are derived by hand from precellar/src/barcode.rs:160-184,311-358,464-468.
"""
import math

import numpy as np

from oracle import pyoracle as O

WL = [b"AAAA", b"AAAC", b"CAAA", b"GGGG", b"TTTT"]


def test_error_probability():
    assert O.error_probability(33) == 1.0
    assert O.error_probability(43) == 0.1
    assert O.error_probability(63) == math.pow(10.0, -3.0)
    assert O.error_probability(23) == 10.0         # q < 33: no clamp (barcode.rs:465-468)


def test_exact_match_is_one():
    code, idx, prob = O.correct_one(WL, [0] * 5, b"GGGG", b"####")
    assert (code, idx, prob) == (O.BC_OK, 3, 1.0)


def test_single_candidate_posterior_is_exactly_one():
    code, idx, prob = O.correct_one(WL, [5, 0, 0, 7, 0], b"GGGA", b"FFF#")
    assert code == O.BC_OK and idx == 3 and prob == 1.0


def test_two_candidates_order_and_tie():
    # query AAAG: pos 0 -> none (CAAG/GAAG/TAAG absent) ... pos 3 -> AAAA (A), AAAC (C) both present.
    # equal counts and same quality -> equal likelihood -> first in (pos, ACGT) order wins: AAAA
    code, idx, prob = O.correct_one(WL, [3, 3, 0, 0, 0], b"AAAG", b"FFFF", threshold=0.5)
    assert code == O.BC_OK and idx == 0 and prob == 0.5
    # threshold equality is accepted (>=) ; just above is LowConfidence
    code, idx, _ = O.correct_one(WL, [3, 3, 0, 0, 0], b"AAAG", b"FFFF", threshold=0.5000000001)
    assert code == O.BC_LOW_CONF and idx == -1
    # prior (1 + count) decides
    code, idx, prob = O.correct_one(WL, [1, 8, 0, 0, 0], b"AAAG", b"FFFF", threshold=0.8)
    assert code == O.BC_OK and idx == 1
    assert prob == (9.0 * O.error_probability(ord("F"))) / (2.0 * O.error_probability(ord("F")) + 9.0 * O.error_probability(ord("F")))


def test_quality_weighting_and_cap():
    # query CAAC: candidates AAAC (pos0 C->A) and CAAA (pos3 C->A).  Low quality at pos 0 favours AAAC.
    q = bytes([ord("#"), ord("F"), ord("F"), ord("F")])
    code, idx, prob = O.correct_one(WL, [0] * 5, b"CAAC", q, threshold=0.9)
    p0, p3 = O.error_probability(ord("#")), O.error_probability(66)   # "F" (70) is capped at 66
    assert code == O.BC_OK and idx == 1 and prob == p0 / (p0 + p3)
    # qualities above 'B' (66) are capped in the likelihood (barcode.rs:320)
    qcap = bytes([ord("B"), ord("~"), ord("~"), ord("~")])
    _, _, prob2 = O.correct_one(WL, [0] * 5, b"CAAC", qcap)
    assert prob2 == 0.5
    # ... but not in expected_errors (barcode.rs:167): sum of uncapped probabilities
    ee = sum(O.error_probability(c) for c in qcap)
    code, _, _ = O.correct_one(WL, [0] * 5, b"CAAC", qcap, threshold=0.5, max_expected_errors=ee)
    assert code == O.BC_EXCEED_EE                      # `>=`
    code, _, _ = O.correct_one(WL, [0] * 5, b"CAAC", qcap, threshold=0.5, max_expected_errors=np.nextafter(ee, 1))
    assert code == O.BC_OK


def test_n_positions():
    # one N: all four bases are tried there (barcode.rs:322-323)
    code, idx, prob = O.correct_one(WL, [0] * 5, b"GGNG", b"FF#F")
    assert code == O.BC_OK and idx == 3 and prob == 1.0
    # two N: no single substitution can reach the whitelist
    code, idx, prob = O.correct_one(WL, [0] * 5, b"GNNG", b"F##F")
    assert code == O.BC_NO_MATCH and idx == -1 and prob == 0.0
    # N at pos 3 of AAAN reaches AAAA and AAAC, first max wins
    code, idx, prob = O.correct_one(WL, [0] * 5, b"AAAN", b"FFF#", threshold=0.5)
    assert code == O.BC_OK and idx == 0 and prob == 0.5


def test_no_match():
    code, idx, prob = O.correct_one(WL, [9] * 5, b"ACGT", b"FFFF")
    assert code == O.BC_NO_MATCH and prob == 0.0


def test_fp64_sum_order():
    # three candidates with different likelihoods: the posterior must be best / ((l0 + l1) + l2)
    wl = [b"AC", b"CC", b"GA", b"GG"]
    cnt = [7, 1, 3, 0]
    q = bytes([40, 52])
    # query GC: pos0 -> AC (A), CC (C) ; pos1 -> GA (A), GG (G)
    p0, p1 = O.error_probability(40), O.error_probability(52)
    likes = [8.0 * p0, 2.0 * p0, 4.0 * p1, 1.0 * p1]
    total = 0.0
    for x in likes:
        total += x
    code, idx, prob = O.correct_one(wl, cnt, b"GC", q, threshold=0.1)
    assert idx == 0 and prob == likes[0] / total


def test_grouping_restatement_hand_case():
    """oracle/grouping.py on a hand-computed case: byte-string barcode order (a proper prefix first), first read of a
    duplicate set kept, counts per (fingerprint, UMI) -- fragment/mod.rs:359-363,434-455, fragment/dedups.rs:320-340."""
    from oracle.grouping import group_reference
    bcs = [b"ACGT", None, b"ACG", b"ACGT", b"TTTT", b"ACGT", b"ACG", b"ACGT"]
    umis = [b"AA", b"AA", b"CC", b"AA", None, b"AN", b"CC", b"AA"]
    fps = [7, 7, 1, 7, 0, 7, 2, 8]
    out = group_reference(bcs, umis, fps)
    assert [b for b, _ in out] == [b"ACG", b"ACGT", b"TTTT"]
    assert out[0][1] == {(1, b"CC"): [2, 1], (2, b"CC"): [6, 1]}
    assert out[1][1] == {(7, b"AA"): [0, 2], (7, b"AN"): [5, 1], (8, b"AA"): [7, 1]}
    assert out[2][1] == {(0, None): [4, 1]}
