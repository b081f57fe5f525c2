"""CPU tier: the device-side zstd writer (csrc/zhuff_core.h, run on the host by tests/emul) against libzstd's decoder.

The frames hold Huffman-coded literals only (no matches); whatever the coder does, a standard zstd decoder must
give back the input bytes."""
import ctypes as C

import numpy as np
import pytest

from precellar_b200 import _zstd
from tests import util as U


def zh_frame(data: bytes, shift: int = 0, matches: bool = True):
    """`shift` = 0: the input starts on a 64-byte boundary (the span loops take the 128-bit-load route, as in the CUDA
    kernels whose buffers are aligned); any other value: the byte route."""
    L = U.emul_lib()
    L.emu_zh_frame.restype = C.c_int64
    L.emu_zh_frame.argtypes = [C.c_char_p, C.c_uint64, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64), C.c_uint32, C.c_int]
    cap = 14 + (len(data) // 131072 + 2) * (131072 + 16)
    out = C.create_string_buffer(cap)
    raw = C.c_uint64(0)
    n = L.emu_zh_frame(data, len(data), out, cap, C.byref(raw), shift, int(matches))
    assert n > 0
    return out.raw[:n], raw.value


def fastq_text(rng, n, qual_levels=40):
    recs = []
    for i in range(n):
        L = int(rng.integers(30, 120))
        seq = rng.choice(np.frombuffer(b"ACGTN", np.uint8), L, p=[0.245, 0.245, 0.245, 0.245, 0.02]).tobytes()
        q = (33 + rng.integers(0, qual_levels, L)).astype(np.uint8).tobytes()
        recs.append(b"@inst:%d:FC:1:%d:%d extra\n" % (i % 7, i, int(rng.integers(0, 99999))) + seq + b"\n+\n" + q + b"\n")
    return b"".join(recs)


@pytest.mark.parametrize("n_records", [1, 7, 900, 5000])
def test_fastq_text_round_trip(n_records):
    rng = np.random.default_rng(n_records)
    data = fastq_text(rng, n_records)
    frame, raw_blocks = zh_frame(data)
    assert _zstd.decompress(frame) == data
    assert zh_frame(data, shift=5)[0] == frame                           # vector and byte routes of the span loops agree
    if len(data) >= 4096:
        assert raw_blocks == 0 and len(frame) < 0.7 * len(data)          # order-0 coding of FASTQ text


def test_exact_block_multiples_and_ragged_tails():
    rng = np.random.default_rng(5)
    base = fastq_text(rng, 6000)
    for n in (131072, 2 * 131072, 131072 + 1, 131072 + 1023, 131072 + 1024, 3 * 131072 - 1, 1024, 1023, 12, 1, 0):
        data = base[:n]
        frame, _ = zh_frame(data)
        assert _zstd.decompress(frame) == data, n
        assert zh_frame(data, shift=1 + n % 15)[0] == frame, n


def test_incompressible_and_degenerate_blocks_are_stored_raw():
    rng = np.random.default_rng(6)
    noise = rng.integers(0, 256, 300000, dtype=np.uint8).tobytes()         # bytes above 128: no direct weight table
    frame, raw_blocks = zh_frame(noise)
    assert raw_blocks == 3 and _zstd.decompress(frame) == noise
    flat = rng.integers(0, 128, 200000, dtype=np.uint8).tobytes()          # 7 bits of entropy per byte: a gain of 1/8 at best
    frame, _ = zh_frame(flat)
    assert _zstd.decompress(frame) == flat
    same = b"A" * 200000                                                   # one distinct byte: no Huffman tree
    frame, raw_blocks = zh_frame(same)
    assert raw_blocks == 2 and _zstd.decompress(frame) == same


def test_skewed_histograms_hit_the_length_limit():
    """Counts spread over 17 binary orders of magnitude give Huffman depths far above 11: the limiter must still produce a
    complete code (or fall back to a raw block) - the decoder rejects anything else."""
    rng = np.random.default_rng(7)
    for trial in range(12):
        k = int(rng.integers(3, 60))
        syms = rng.choice(np.arange(1, 128), k, replace=False)
        w = 2.0 ** (-rng.uniform(0.3, 1.6) * np.arange(k))
        data = rng.choice(syms, 200000 + trial * 997, p=w / w.sum()).astype(np.uint8).tobytes()
        frame, _ = zh_frame(data)
        assert _zstd.decompress(frame) == data, trial
    fib = [1, 1]
    while len(fib) < 24:
        fib.append(fib[-1] + fib[-2])                                       # Fibonacci counts: the deepest possible tree
    data = b"".join(bytes([40 + i]) * c for i, c in enumerate(fib))
    data = bytes(np.random.default_rng(8).permutation(np.frombuffer(data, np.uint8)))
    frame, _ = zh_frame(data)
    assert _zstd.decompress(frame) == data


def test_symbol_128_and_two_symbol_blocks():
    rng = np.random.default_rng(9)
    data = rng.choice(np.array([65, 128], np.uint8), 5000, p=[0.9, 0.1]).tobytes()
    frame, raw_blocks = zh_frame(data)
    assert raw_blocks == 0 and _zstd.decompress(frame) == data
    data = rng.choice(np.array([0, 1], np.uint8), 5000).tobytes()
    frame, _ = zh_frame(data)
    assert _zstd.decompress(frame) == data


def illumina_text(rng, n, read_len=100):
    recs = []
    qb = np.frombuffer(b"F:,#", np.uint8)
    for i in range(n):
        seq = rng.choice(np.frombuffer(b"ACGT", np.uint8), read_len).tobytes()
        q = rng.choice(qb, read_len, p=[0.88, 0.07, 0.04, 0.01]).tobytes()
        recs.append(b"@A00123:45:HXXXXXXXX:1:%d:%d:%d 1:N:0:ACGTACGT+TGCATGCA\n" % (1101 + i // 5000, 1000 + (i * 37) % 30000, 1000 + i % 40000)
                    + seq + b"\n+\n" + q + b"\n")
    return b"".join(recs)


def test_matches_against_the_previous_record_shrink_fastq_text():
    """Sequences (prefix / suffix matches of a line against the same line of the previous record, predefined FSE codes):
    libzstd must decode them, and on text with instrument-style names they must pay."""
    rng = np.random.default_rng(21)
    data = illumina_text(rng, 6000)
    with_m, raw_blocks = zh_frame(data)
    without, _ = zh_frame(data, matches=False)
    assert raw_blocks == 0 and _zstd.decompress(with_m) == data and _zstd.decompress(without) == data
    assert len(with_m) < 0.8 * len(without)
    assert zh_frame(data, shift=3)[0] == with_m
    for n in (131072, 131073, 262144 + 77, 5000, 1500):
        frame, _ = zh_frame(data[:n])
        assert _zstd.decompress(frame) == data[:n], n


def test_match_finder_corner_cases():
    rng = np.random.default_rng(22)
    cases = {
        "identical records": b"@same name here\nACGTACGTACGTACGTACGT\n+\nFFFFFFFFFFFFFFFFFFFF\n" * 4000,
        "more lines than the index holds": b"".join(b"ab%d\n" % (i % 7) for i in range(60000)),
        "few lines": b"x" * 70000 + b"\n" + b"y" * 70000 + b"\n",
        "lines longer than 64 K that repeat": (b"Q" * 30000 + b"\n" + b"ACGT" * 2000 + b"\n+\nzz\n") * 6,
        "no newline at all": bytes(rng.integers(65, 70, 200000, dtype=np.uint8)),
        "blank lines": b"\n" * 3000 + b"abcdefghijklmnop\n" * 3000,
        "prefix and suffix overlap": b"".join(b"@common-prefix-%d-common-suffix-of-the-name\nAC\n+\nFF\n" % (i % 10) for i in range(5000)),
        "long literal runs between matches": b"".join(b"@name-with-a-long-prefix:%d\n" % i + bytes(rng.integers(65, 91, 900, dtype=np.uint8)) + b"\n+\n"
                                                     + bytes(rng.integers(65, 91, 900, dtype=np.uint8)) + b"\n" for i in range(300)),
    }
    for name, data in cases.items():
        frame, _ = zh_frame(data)
        assert _zstd.decompress(frame) == data, name
        assert zh_frame(data, shift=7)[0] == frame, name
    # 64 KB+ identical lines: match lengths at the 16-bit limit
    line = bytes(rng.integers(65, 69, 66000, dtype=np.uint8)) + b"\n"
    data = (line + b"s\n+\nq\n") * 5
    frame, _ = zh_frame(data)
    assert _zstd.decompress(frame) == data


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_random_record_texts_round_trip(seed):
    """Random four-line records (shared name prefixes and suffixes of random length, empty and long lines, CRLF, truncated
    tails, small and large alphabets): whatever sequences and literals the coder picks, libzstd returns the input."""
    rng = np.random.default_rng(1000 + seed)
    for it in range(12):
        nrec = int(rng.integers(1, 2500))
        alpha = int(rng.choice([4, 8, 20, 60]))
        pre = bytes(rng.integers(65, 65 + alpha, int(rng.integers(0, 70)), dtype=np.uint8))
        suf = bytes(rng.integers(65, 65 + alpha, int(rng.integers(0, 40)), dtype=np.uint8))
        recs = []
        for i in range(nrec):
            L = int(rng.integers(0, 150))
            mid = bytes(rng.integers(48, 58, int(rng.integers(0, 9)), dtype=np.uint8))
            s = bytes(rng.integers(65, 65 + alpha, L, dtype=np.uint8))
            q = bytes(rng.choice(np.frombuffer(b"F:,#", np.uint8), L, p=[0.7, 0.1, 0.1, 0.1])) if rng.random() < 0.8 else b"F" * L
            eol = b"\r\n" if rng.random() < 0.1 else b"\n"
            recs.append(b"@" + pre + mid + suf + eol + s + eol + b"+" + eol + q + eol)
        data = b"".join(recs)
        if rng.random() < 0.3:
            data = data[:int(rng.integers(0, len(data) + 1))]
        frame, _ = zh_frame(data, shift=int(rng.integers(0, 16)))
        assert _zstd.decompress(frame) == data, (seed, it)
