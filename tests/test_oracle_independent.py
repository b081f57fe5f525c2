"""The C++ oracle (oracle/oracle.cpp) against a second restatement written independently from the Rust source
(tests/pyref.py): counting and correction must agree bit for bit -- code, corrected barcode and the fp64 posterior --
on randomised whitelists, barcodes and quality strings, including the cases the device code treats specially
(N and other non-ACGT bytes, lower case, qualities below '!' + 0 and above the cap, dense whitelists with several
candidates, ties)."""
import numpy as np
import pytest

from oracle import pyoracle as O
from tests import pyref as R
from tests import util as U

CODE = {R.OK: O.BC_OK, R.NO_MATCH: O.BC_NO_MATCH, R.LOW_CONFIDENCE: O.BC_LOW_CONF, R.EXCEED_EE: O.BC_EXCEED_EE}
ACGT = np.frombuffer(b"ACGT", np.uint8)


def random_case(rng, k, dense):
    if dense:                                    # a handful of seeds and all their single substitutions: many candidates, ties
        seeds = [bytes(rng.choice(ACGT, k)) for _ in range(3)]
        wl = set(seeds)
        for s in seeds:
            for pos in range(k):
                for c in b"ACGT":
                    if rng.random() < 0.5:
                        v = bytearray(s); v[pos] = c; wl.add(bytes(v))
        wl = sorted(wl)
        rng.shuffle(wl)
    else:
        wl = list(dict.fromkeys(bytes(rng.choice(ACGT, k)) for _ in range(int(rng.integers(1, 200)))))
    counts = [int(x) for x in rng.choice([0, 0, 1, 2, 7, 1000, 2 ** 33], len(wl))]
    return wl, counts


def random_query(rng, wl, k):
    base = bytearray(wl[int(rng.integers(len(wl)))]) if rng.random() < 0.85 else bytearray(rng.choice(ACGT, k))
    for _ in range(int(rng.choice([0, 1, 1, 1, 2, 3]))):
        pos = int(rng.integers(k))
        base[pos] = int(rng.choice(np.frombuffer(b"ACGTNacgtn.", np.uint8)))
    qual = bytes(int(x) for x in rng.choice([2, 20, 33, 35, 44, 58, 66, 67, 70, 73, 90, 126, 200, 255], k))
    return bytes(base), qual


@pytest.mark.parametrize("dense", [False, True])
@pytest.mark.parametrize("k", [1, 4, 9, 10, 16, 17, 31])
def test_correct_barcode_matches_independent_restatement(k, dense):
    rng = np.random.default_rng(U.seed_of("indep", k, dense))
    seen = set()
    for case in range(12):
        wl, counts = random_case(rng, k, dense)
        ref = R.Whitelist(wl)
        for b, c in zip(wl, counts):
            ref.barcode_counts.counts[b] = c
        for _ in range(120):
            bc, qual = random_query(rng, wl, k)
            thr = float(rng.choice([0.5, 0.9, 0.975, 1.0]))
            ee = float(rng.choice([np.inf, np.inf, 0.5, 3.0]))
            want_code, want_bc, want_prob = R.correct_barcode(ref, bc, qual, thr, ee)
            code, idx, prob = O.correct_one(wl, counts, bc, qual, threshold=thr, max_expected_errors=O.F64_MAX if np.isinf(ee) else ee)
            assert code == CODE[want_code], (bc, qual, wl)
            seen.add(want_code)
            if want_code == R.EXCEED_EE:
                continue
            assert prob == want_prob, (bc, qual, prob, want_prob)          # bit-exact fp64
            if want_code == R.OK:
                assert wl[idx] == want_bc
            else:
                assert idx == -1
    assert R.OK in seen
    if k >= 9 and not dense:
        assert R.NO_MATCH in seen
    if k >= 4:
        assert R.EXCEED_EE in seen and (R.LOW_CONFIDENCE in seen or not dense)


@pytest.mark.parametrize("is_reverse", [False, True])
def test_counts_match_independent_restatement(is_reverse):
    """WhitelistBuilder::add over the barcode of every read (barcode.rs:85-119), forward and neg-strand
    (rev_compl before the lookup, barcode.rs:110-111), through the oracle's whole-read driver."""
    rng = np.random.default_rng(U.seed_of("indep-count", is_reverse))
    wl = U.gen_whitelist(rng, 300, 16)
    layout = U.make_layout([dict(read_id="R1", elems=[U.bc("cb", 16), U.umi(10)], max_len=26, is_reverse=is_reverse)], {"cb": wl})
    full = U.gen_reads(rng, layout, 4000, [32], p_sub=0.05, p_n=0.03, p_lower=0.05 if is_reverse else 0.0, p_short=0.0)
    _, ocounts = U.run_oracle(layout, full, threshold=0.9)
    ref = R.Whitelist(wl)
    for i in range(4000):
        L = int(full[0].length[i])
        seq = full[0].seq[i, :min(L, 26)].tobytes()
        if len(seq) < 16:
            continue                                  # the barcode segment is absent: nothing is added
        bc = seq[:16]
        ref.add(R.rev_compl(bc) if is_reverse else bc)
    counts, mismatch, total = ocounts[0]
    assert total == ref.total_count and mismatch == ref.mismatch_count and ref.mismatch_count > 0
    assert [int(c) for c in counts] == [ref.barcode_counts.counts[b] for b in wl]


# ---------------------------------------------------------------------------------------------------
# split_with_tolerance: oracle vs the independent transcription, on random element tables and reads
# ---------------------------------------------------------------------------------------------------
LABEL = {O.OTHER: "O", O.BARCODE: "B", O.UMI: "U", O.TARGET: "T", O.PRIMER: "P"}


def random_elems(rng):
    elems = []
    for _ in range(int(rng.integers(1, 8))):
        kind = rng.random()
        rtype = int(rng.choice([O.OTHER, O.BARCODE, O.UMI, O.TARGET, O.PRIMER]))
        if kind < 0.35:                                   # fixed length, random sequence
            n = int(rng.integers(1, 9))
            elems.append(O.Elem(rtype, n, n))
        elif kind < 0.6:                                  # fixed sequence (anchor / linker)
            n = int(rng.integers(1, 8))
            seq = bytes(rng.choice(ACGT, n)).decode()
            elems.append(O.Elem(rtype, n, n, True, seq))
        else:                                             # variable length
            a = int(rng.integers(0, 6))
            b = a + int(rng.integers(1, 12))
            elems.append(O.Elem(rtype, a, b))
    return elems


def random_read(rng, elems, is_reverse):
    """Mostly a plausible instance of the element table (anchors present, sometimes mutated), sometimes noise."""
    if rng.random() < 0.15:
        return bytes(rng.choice(ACGT, int(rng.integers(0, 60))))
    out = bytearray()
    for e in elems:
        n = int(rng.integers(e.min_len, e.max_len + 1))
        if e.seq_fixed:
            s = bytearray(e.sequence.encode())
            if is_reverse:
                s = bytearray(R.seqspec_rev_compl(bytes(s)))
            if rng.random() < 0.3 and len(s):
                s[int(rng.integers(len(s)))] = int(rng.choice(np.frombuffer(b"ACGTNa", np.uint8)))
            out += s
        else:
            out += bytes(rng.choice(ACGT, n))
    cut = int(rng.integers(0, len(out) + 8)) if rng.random() < 0.4 else len(out) + int(rng.integers(0, 10))
    out = out[:cut] + bytes(rng.choice(ACGT, max(0, cut - len(out))))
    return bytes(out)


@pytest.mark.parametrize("is_reverse", [False, True])
@pytest.mark.parametrize("tols", [(1.0, 0.2), (0.5, 0.0), (0.0, 0.34)])
def test_split_matches_independent_restatement(is_reverse, tols):
    rng = np.random.default_rng(U.seed_of("indep-split", is_reverse, tols))
    seen = set()
    for _ in range(400):
        elems = random_elems(rng)
        ref_elems = [R.SegElem(LABEL[e.rtype], e.min_len, e.max_len, e.seq_fixed, e.sequence.encode()) for e in elems]
        for _ in range(12):
            read = random_read(rng, elems, is_reverse)
            st, text = O.split_str(elems, read, is_reverse, tols[0], tols[1])
            try:
                want = ("ok", R.render(R.split_with_tolerance(ref_elems, is_reverse, read, tols[0], tols[1])))
            except R.SplitError as e:
                want = (e.kind, "")
            except R.RustPanic:
                want = ("panic", "")
            got = {O.SPLIT_OK: "ok", O.SPLIT_ANCHOR_NOT_FOUND: "anchor_not_found", O.SPLIT_PATTERN_MISMATCH: "pattern_mismatch",
                   -100: "panic"}[st]
            assert (got, text if got == "ok" else "") == want, (elems, read, is_reverse)
            seen.add(got)
    assert {"ok", "pattern_mismatch"} <= seen
