"""CPU tier: the device-side zstd writer (csrc/zhuff_core.h, run on the host by tests/emul) against libzstd's decoder.

The frames hold Huffman-coded literals only (no matches); whatever the coder does, a standard zstd decoder must
give back the input bytes."""
import ctypes as C

import numpy as np
import pytest

from precellar_b200 import _zstd
from tests import util as U


def zh_frame(data: bytes, shift: int = 0):
    """`shift` = 0: the input starts on a 64-byte boundary (the span loops take the 128-bit-load route, as in the CUDA
    kernels whose buffers are aligned); any other value: the byte route."""
    L = U.emul_lib()
    L.emu_zh_frame.restype = C.c_int64
    L.emu_zh_frame.argtypes = [C.c_char_p, C.c_uint64, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64), C.c_uint32]
    cap = 14 + (len(data) // 131072 + 2) * (131072 + 16)
    out = C.create_string_buffer(cap)
    raw = C.c_uint64(0)
    n = L.emu_zh_frame(data, len(data), out, cap, C.byref(raw), shift)
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
