"""Second, independent restatement of the reference's counting and correction, in plain Python, written from the Rust
source alone (precellar/src/barcode.rs) and sharing no code with oracle/oracle.cpp.  The reference has no test for
this part ("parity unpinned") and cannot be built here, so two independent transcriptions that agree
bit for bit on randomised inputs are the strongest pin available (tests/test_oracle_independent.py).

Test infrastructure only.
"""
import math

BC_MAX_QV = 66                                   # barcode.rs:18
BASE_OPTS = b"ACGT"                              # barcode.rs:19


def error_probability(qual: int) -> float:       # barcode.rs:464-468
    offset = 33.0
    return math.pow(10.0, -((float(qual) - offset) / 10.0))


class OligoFrequency:
    """barcode.rs:189-358: HashMap<Vec<u8>, usize> with the 1-mismatch posterior."""

    def __init__(self, counts: dict):
        self.counts = counts

    def likelihood1(self, query: bytes, qual: bytes):            # barcode.rs:311-342
        if query in self.counts:
            return query, 1.0
        best_option = None                                        # (likelihood, key)
        total_likelihood = 0.0
        query_bytes = bytearray(query)
        for pos, qv in enumerate(qual):
            qv = min(qv, BC_MAX_QV)
            existing = query_bytes[pos]
            for val in BASE_OPTS:
                if val != existing:
                    query_bytes[pos] = val
                    key = bytes(query_bytes)
                    if key in self.counts:
                        bc_count = 1 + self.counts[key]
                        likelihood = float(bc_count) * error_probability(qv)
                        if best_option is None or best_option[0] < likelihood:      # update_best_option, :345-358
                            best_option = (likelihood, key)
                        total_likelihood += likelihood
            query_bytes[pos] = existing
        if best_option is not None:
            return best_option[1], best_option[0] / total_likelihood
        return query, 0.0


class Whitelist:
    """WhitelistBuilder::{new, add, finish} for a non-empty list of valid barcodes (barcode.rs:382-460)."""

    def __init__(self, valid_barcodes):
        self.barcode_counts = OligoFrequency({bytes(b): 0 for b in valid_barcodes})
        assert self.barcode_counts.counts, "whitelist_exists == false is outside the path"
        self.mismatch_count = 0
        self.total_count = 0

    def add(self, barcode: bytes) -> None:                        # barcode.rs:410-426
        c = self.barcode_counts.counts
        if barcode in c:
            c[barcode] += 1
        else:
            self.mismatch_count += 1
        self.total_count += 1


OK, NO_MATCH, LOW_CONFIDENCE, EXCEED_EE = "ok", "no_match", "low_confidence", "exceed_expected_error"


def correct_barcode(wl: Whitelist, barcode: bytes, qual: bytes, bc_confidence_threshold: float,
                    max_expected_errors: float = float("inf")):      # barcode.rs:160-184 (max_mismatch = 1, fastq.rs:40)
    expected_errors = 0.0
    for q in qual:
        expected_errors += error_probability(q)
    if expected_errors >= max_expected_errors:
        return EXCEED_EE, None, expected_errors
    bc, prob = wl.barcode_counts.likelihood1(barcode, qual)
    if prob <= 0.0:
        return NO_MATCH, None, prob
    if prob >= bc_confidence_threshold:
        return OK, bc, prob
    return LOW_CONFIDENCE, None, prob


def rev_compl(seq: bytes) -> bytes:                                  # precellar/src/utils.rs:138-149
    m = {65: 84, 67: 71, 71: 67, 84: 65, 97: 84, 99: 71, 103: 67, 116: 65}
    return bytes(m.get(b, b) for b in reversed(seq))
