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


# ---------------------------------------------------------------------------------------------------
# SegmentInfo::split_with_tolerance, transcribed from seqspec/src/read/segment.rs:180-368,438-535 (independent of
# oracle/oracle.cpp).  Rust slice indexing / Option::unwrap failures are reported as RustPanic.
# ---------------------------------------------------------------------------------------------------
class SplitError(Exception):
    def __init__(self, kind, n=None):
        super().__init__(kind)
        self.kind, self.n = kind, n


class RustPanic(Exception):
    pass


class SegElem:
    """SegmentInfoElem: label = RegionType's Display letter (region.rs:321-335)."""

    def __init__(self, label, min_len, max_len, fixed_seq=False, sequence=b""):
        self.label, self.mn, self.mx, self.fixed_seq, self.sequence = label, min_len, max_len, fixed_seq, bytes(sequence)

    def is_fixed_len(self):
        return self.mn == self.mx


def seqspec_rev_compl(seq: bytes) -> bytes:                          # seqspec/src/utils.rs:176-187 (upper case only)
    m = {65: 84, 84: 65, 67: 71, 71: 67}
    return bytes(m.get(b, b) for b in reversed(seq))


def _get(seq: bytes, lo: int, hi: int) -> bytes:                     # slice.get(lo..hi).unwrap()
    if lo < 0 or lo > hi or hi > len(seq):
        raise RustPanic(f"range {lo}..{hi} of {len(seq)}")
    return seq[lo:hi]


def find_pattern(haystack: bytes, needle: bytes):                    # segment.rs:438-483 (KMP): first exact occurrence
    if not needle or not haystack or len(needle) > len(haystack):
        return None
    lps = [0] * len(needle)
    ln, i = 0, 1
    while i < len(needle):
        if needle[i] == needle[ln]:
            ln += 1
            lps[i] = ln
            i += 1
        elif ln != 0:
            ln = lps[ln - 1]
        else:
            lps[i] = 0
            i += 1
    i = j = 0
    while i < len(haystack):
        if needle[j] == haystack[i]:
            i += 1
            j += 1
        if j == len(needle):
            return i - j
        elif i < len(haystack) and needle[j] != haystack[i]:
            if j != 0:
                j = lps[j - 1]
            else:
                i += 1
    return None


USIZE_MAX = (1 << 64) - 1


def find_best_pattern_match(haystack: bytes, needle: bytes):          # segment.rs:496-535
    if not needle or not haystack or len(needle) > len(haystack):
        return None, USIZE_MAX
    pos = find_pattern(haystack, needle)
    if pos is not None:
        return pos, 0
    n, m = len(haystack), len(needle)
    if m == 1:
        return None, 1
    min_mismatches, best_pos = USIZE_MAX, None
    for i in range(0, max(n - m, 0) + 1):
        mismatches = 0
        for j in range(m):
            if haystack[i + j] != needle[j]:
                mismatches += 1
                if mismatches >= min_mismatches and min_mismatches != USIZE_MAX:
                    break
        if mismatches < min_mismatches:
            min_mismatches, best_pos = mismatches, i
    return best_pos, min_mismatches


def _consume_buf(buf, seq: bytes, result):                            # segment.rs:186-232
    n1 = len(result)
    while buf:
        elem = buf.pop()
        if not elem.is_fixed_len():
            buf.append(elem)
            break
        l = len(seq)
        i = l - elem.mx
        if i < 0:
            raise RustPanic("attempt to subtract with overflow")
        result.append((elem.label, seq[i:l]))
        seq = seq[:i]
    if buf:
        label = buf[0].label if len(buf) == 1 else "".join(e.label for e in buf)
        buf.clear()
        result.append((label, seq))
    n2 = len(result)
    if n2 > n1:
        result[n1:n2] = result[n1:n2][::-1]


def split_with_tolerance(elems, is_reverse: bool, read: bytes, linker_tolerance=1.0, anchor_tolerance=0.2):
    result = []
    min_offset = max_offset = 0
    buffer = []
    ln = len(read)
    for segment in elems:
        min_len, max_len = segment.mn, segment.mx
        if max_offset >= ln or min_offset + min_len > ln:
            break
        if min_len == max_len:
            if buffer:
                if segment.fixed_seq:
                    seq = _get(read, min_offset, min(ln, max_offset + max_len))
                    pos, mis = find_best_pattern_match(seqspec_rev_compl(seq) if is_reverse else seq, segment.sequence)
                    if pos is not None:
                        if float(mis) <= anchor_tolerance * float(len(segment.sequence)):
                            offset_left = min_offset - sum(s.mn for s in buffer)
                            offset_right = min_offset + pos
                            _consume_buf(buffer, _get(read, offset_left, offset_right), result)
                            min_offset = max_offset = offset_right
                            result.append((segment.label, _get(read, offset_right, offset_right + max_len)))
                        elif max_offset + max_len > ln:
                            break
                        else:
                            raise SplitError("pattern_mismatch", mis)
                    else:
                        raise SplitError("anchor_not_found")
                else:
                    buffer.append(segment)
            else:
                seq = _get(read, max_offset, max_offset + max_len)
                if linker_tolerance < 1.0 and segment.fixed_seq:
                    a = seqspec_rev_compl(seq) if is_reverse else seq
                    if len(a) != len(segment.sequence):
                        raise RustPanic("hamming_distance: lengths differ")
                    d = sum(1 for x, y in zip(a, segment.sequence) if x != y)
                    if float(d) > linker_tolerance * float(len(seq)):
                        raise SplitError("pattern_mismatch", d)
                result.append((segment.label, seq))
        else:
            buffer.append(segment)
        min_offset += min_len
        max_offset += max_len
    if buffer:
        offset_left = min_offset - sum(s.mn for s in buffer)
        _consume_buf(buffer, _get(read, offset_left, min(max_offset, ln)), result)
    return result


def render(segments) -> str:                                          # the unit tests' format, segment.rs:596-603
    return ",".join(f"[{lab}]{seq.decode('latin-1')}" for lab, seq in segments)
