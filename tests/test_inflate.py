"""CPU tier: the device-side BGZF inflate (csrc/inflate_core.h, run on the host by tests/emul) against zlib.

Every member is inflated by the emulated warp (phase by phase, lanes in ascending and in descending order - a phase
whose lanes depended on one another would give different bytes) and compared with the text zlib compressed; the CRC-32
of the trailer is checked by the code under test.  Corrupted members must end in an error, never in wrong bytes.
"""
import ctypes as C
import zlib

import numpy as np
import pytest

from tests import util as U

PAD = 16


class Member(C.Structure):
    _fields_ = [("src_off", C.c_uint32), ("clen", C.c_uint32), ("dst_off", C.c_uint32), ("isize", C.c_uint32), ("crc", C.c_uint32)]


def deflate(data: bytes, level=6, strategy=zlib.Z_DEFAULT_STRATEGY, flush_every=0) -> bytes:
    co = zlib.compressobj(level, zlib.DEFLATED, -15, 9, strategy)
    if not flush_every:
        return co.compress(data) + co.flush()
    out = b""
    for i in range(0, len(data), flush_every):              # several blocks per member, with empty stored blocks between
        out += co.compress(data[i:i + flush_every]) + co.flush(zlib.Z_FULL_FLUSH if (i // flush_every) % 2 else zlib.Z_SYNC_FLUSH)
    return out + co.flush()


def run(parts, streams, reverse=0, crcs=None, isizes=None, gap=3):
    """parts[i]: expected bytes of member i, streams[i]: its DEFLATE stream.  Returns (rc, output)."""
    L = U.emul_lib()
    L.emu_inflate.restype = C.c_int64
    L.emu_inflate.argtypes = [C.c_char_p, C.c_uint64, C.POINTER(Member), C.c_uint32, C.c_void_p, C.c_uint64, C.c_int]
    comp = bytearray()
    members = (Member * max(1, len(parts)))()
    dst = 0
    for i, (p, s) in enumerate(zip(parts, streams)):
        comp += b"\xAA" * (gap + i % 4)                      # streams start at every alignment
        isize = len(p) if isizes is None else isizes[i]
        members[i] = Member(len(comp), len(s), dst, isize, (zlib.crc32(p) & 0xFFFFFFFF) if crcs is None else crcs[i])
        comp += s
        dst += isize
    n_comp = len(comp)
    comp += b"\xFF" * PAD
    out = np.full(dst + 64, 0xEE, np.uint8)
    rc = L.emu_inflate(bytes(comp), n_comp, members, len(parts), out.ctypes.data, dst, reverse)
    assert bytes(out[dst:]) == b"\xEE" * 64                  # nothing written behind the block
    return rc, bytes(out[:dst])


def fastq_text(rng, n, read_len=90):
    recs = []
    for i in range(n):
        seq = bytes(rng.choice(np.frombuffer(b"ACGTN", np.uint8), read_len, p=[.245, .245, .245, .245, .02]))
        q = bytes(rng.choice(np.frombuffer(b"F:,#", np.uint8), read_len, p=[.8, .1, .07, .03]))
        recs.append(b"@A00123:45:HXXXXXXXX:1:%d:%d:%d 1:N:0:ACGT\n%s\n+\n%s\n" % (1101 + i // 500, 1000 + 7 * i, 2000 + i, seq, q))
    return b"".join(recs)


def check(parts, streams):
    for rev in (0, 1):
        rc, out = run(parts, streams, reverse=rev)
        assert rc == 0, (rc >> 8, rc & 0xFF)
        assert out == b"".join(parts)


@pytest.mark.parametrize("level", [1, 6, 9])
def test_fastq_members_of_every_size(level):
    rng = np.random.default_rng(level)
    text = fastq_text(rng, 3000)
    sizes = [65536, 65280, 1, 2, 3, 257, 258, 259, 4096, 60000, 31, 32, 33, 1000]
    parts, pos = [], 0
    for s in sizes:
        parts.append(text[pos:pos + s])
        pos += s
    check(parts, [deflate(p, level) for p in parts])


def test_fixed_stored_and_multi_block_members():
    rng = np.random.default_rng(5)
    text = fastq_text(rng, 1200)
    noise = bytes(rng.integers(0, 256, 40000, dtype=np.uint8))
    parts = [text[:50000], text[50000:50300], noise, noise[:1], text[:65536], text[1000:40000], b"A", text[:20000] + noise[:20000]]
    streams = [deflate(parts[0], 6, zlib.Z_FIXED),            # fixed codes
               deflate(parts[1], 9, zlib.Z_FIXED),
               deflate(parts[2], 0),                          # stored blocks
               deflate(parts[3], 0),
               deflate(parts[4], 6, flush_every=5000),        # many blocks, empty stored blocks between them
               deflate(parts[5], 1, zlib.Z_HUFFMAN_ONLY),     # literals only: no distance code at all
               deflate(parts[6], 6),
               deflate(parts[7], 6, flush_every=7777)]
    check(parts, streams)


def test_overlapping_and_long_matches():
    rng = np.random.default_rng(6)
    parts = [b"\0" * 65536,                                   # distance 1, length 258 all the way
             b"AB" * 30000, b"ACG" * 20000, bytes(range(256)) * 200,
             (b"ACGT" * 64 + bytes(rng.integers(65, 70, 300, dtype=np.uint8))) * 100,
             b"".join(bytes([65 + (i % 7)]) * (i % 300) for i in range(400))[:65536],
             zlib.Z_RLE.to_bytes(1, "little") * 5]
    streams = [deflate(p, 9) for p in parts[:-1]] + [deflate(parts[-1], 6, zlib.Z_RLE)]
    check(parts, streams)
    parts = [bytes(rng.integers(0, 4, 65536, dtype=np.uint8)), bytes(rng.integers(0, 2, 65536, dtype=np.uint8))]
    check(parts, [deflate(parts[0], 6, zlib.Z_RLE), deflate(parts[1], 9)])


@pytest.mark.parametrize("seed", range(6))
def test_random_texts(seed):
    rng = np.random.default_rng(100 + seed)
    parts, streams = [], []
    for _ in range(12):
        n = int(rng.integers(1, 65537))
        kind = int(rng.integers(0, 4))
        if kind == 0:
            p = fastq_text(rng, n // 200 + 1, int(rng.integers(20, 151)))[:n]
        elif kind == 1:
            p = bytes(rng.integers(0, int(rng.integers(2, 257)), n, dtype=np.uint8))
        elif kind == 2:
            words = [bytes(rng.integers(97, 123, int(rng.integers(1, 12)), dtype=np.uint8)) for _ in range(50)]
            p = b" ".join(words[int(k)] for k in rng.integers(0, 50, n // 4 + 1))[:n]
        else:
            p = bytes(np.repeat(rng.integers(0, 256, n // 50 + 1, dtype=np.uint8), rng.integers(1, 100, n // 50 + 1)))[:n]
        if not p:
            p = b"x"
        parts.append(p)
        streams.append(deflate(p, int(rng.integers(0, 10)), [zlib.Z_DEFAULT_STRATEGY, zlib.Z_FILTERED, zlib.Z_RLE, zlib.Z_FIXED][int(rng.integers(0, 4))],
                               flush_every=int(rng.integers(0, 2)) * int(rng.integers(500, 30000))))
    check(parts, streams)


def test_crc_join_of_every_slice_shape():
    """Members of 0..70 bytes and around multiples of 32: slices of unequal length, empty slices, the empty member."""
    rng = np.random.default_rng(7)
    sizes = list(range(0, 71)) + [95, 96, 97, 1023, 1024, 1025, 65535]
    parts = [bytes(rng.integers(0, 256, s, dtype=np.uint8)) for s in sizes]
    check(parts, [deflate(p, 6) for p in parts])


def test_damaged_members_are_errors():
    rng = np.random.default_rng(8)
    text = fastq_text(rng, 400)[:30000]
    good = deflate(text, 6)
    rc, _ = run([text], [good], crcs=[(zlib.crc32(text) ^ 1) & 0xFFFFFFFF])
    assert rc == (1 << 8) | 10                                 # CRC
    rc, _ = run([text], [good], isizes=[len(text) + 5])
    assert rc == (1 << 8) | 9                                  # fewer bytes than ISIZE
    rc, _ = run([text], [good], isizes=[len(text) - 5])
    assert (rc >> 8) == 1 and (rc & 0xFF) == 7                 # more bytes than ISIZE
    rc, _ = run([text], [good[:len(good) // 2]])
    assert (rc >> 8) == 1 and (rc & 0xFF) != 0                 # truncated stream
    rc, _ = run([text], [b"\x07" + good])                      # reserved block type
    assert rc == (1 << 8) | 1
    rc, _ = run([text, text], [good, b"\x01\x05\x00\x00\x00abcde"])       # stored block whose NLEN is wrong, in the second member
    assert rc == (2 << 8) | 2
    bad = 0
    for k in range(300):                                       # random damage: an error or (rarely) the same text, never a crash
        s = bytearray(good)
        for _ in range(int(rng.integers(1, 4))):
            s[int(rng.integers(0, len(s)))] ^= 1 << int(rng.integers(0, 8))
        if k % 3 == 0:
            s = s[:int(rng.integers(1, len(s)))]
        for rev in (0, 1):
            rc, out = run([text], [bytes(s)], reverse=rev)
            assert rc != 0 or out == text
        bad += rc != 0
    assert bad >= 290
    for k in range(200):                                       # pure noise as a stream
        s = bytes(rng.integers(0, 256, int(rng.integers(1, 2000)), dtype=np.uint8))
        rc, _ = run([text], [s])
        assert rc != 0


# ---------------------------------------------------------------------------------------------------
# csrc/fastinflate.h: the host decoder behind pcl_gzsrc for ordinary gzip files (one DEFLATE stream per file), against zlib
# ---------------------------------------------------------------------------------------------------
def fast_inflate(stream: bytes, n_out: int, piece=1 << 30, take=1 << 30, tail=b""):
    L = U.emul_lib()
    L.emu_fast_inflate.restype = C.c_int64
    L.emu_fast_inflate.argtypes = [C.c_char_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.POINTER(C.c_uint64)]
    out = np.full(n_out + 64, 0xEE, np.uint8)
    used = C.c_uint64()
    data = stream + tail
    rc = L.emu_fast_inflate(data, len(data), out.ctypes.data, n_out, piece, take, C.byref(used))
    assert bytes(out[n_out:]) == b"\xEE" * 64
    return rc, bytes(out[:max(rc, 0)]), used.value


@pytest.mark.parametrize("level", [0, 1, 6, 9])
def test_fast_inflate_streams_of_every_kind(level):
    rng = np.random.default_rng(40 + level)
    text = fastq_text(rng, 12000)                              # ~ 3 MB: several windows of the decoder
    noise = bytes(rng.integers(0, 256, 200000, dtype=np.uint8))
    cases = [text, text[:1], b"", text[:700000] + noise + text[:900000], b"\0" * 1500000, b"AB" * 400000,
             bytes(rng.integers(0, 4, 800000, dtype=np.uint8))]
    for k, data in enumerate(cases):
        for strategy, flush in ((zlib.Z_DEFAULT_STRATEGY, 0), (zlib.Z_FIXED, 0), (zlib.Z_DEFAULT_STRATEGY, 30011), (zlib.Z_RLE, 0)):
            s = deflate(data, level, strategy, flush_every=flush)
            rc, out, used = fast_inflate(s, len(data), tail=b"TRAILER!")
            assert rc == len(data) and out == data and used == len(s), (k, strategy, flush)


@pytest.mark.parametrize("seed", range(4))
def test_fast_inflate_input_in_pieces_and_output_in_sips(seed):
    """The input arrives in pieces (the decoder asks for more with fewer than 2 KB left), the output is taken in small
    amounts; stored blocks span pieces and windows."""
    rng = np.random.default_rng(60 + seed)
    text = fastq_text(rng, 6000) + bytes(rng.integers(0, 256, 150000, dtype=np.uint8)) + fastq_text(rng, 2000)
    for level in (0, 1, 6):
        s = deflate(text, level, flush_every=int(rng.integers(0, 2)) * 50000)
        for piece, take in ((1 << 30, 7), (4096, 1 << 30), (2049, 100000), (int(rng.integers(2100, 70000)), int(rng.integers(1, 300000))),
                            (100, 1 << 30), (1, 65536)):
            rc, out, used = fast_inflate(s, len(text), piece, take, tail=b"12345678")
            assert rc == len(text) and out == text and used == len(s), (level, piece, take)


def test_fast_inflate_damaged_streams_are_errors():
    rng = np.random.default_rng(70)
    text = fastq_text(rng, 3000)
    good = deflate(text, 6)
    rc, _, _ = fast_inflate(good[:len(good) // 2], len(text))
    assert rc < 0                                              # truncated
    rc, _, _ = fast_inflate(b"\x07" + good, len(text))
    assert rc < 0                                              # reserved block type
    rc, _, _ = fast_inflate(b"\x01\x05\x00\x00\x00abcde", 5)
    assert rc < 0                                              # stored block whose NLEN is wrong
    rc, out, _ = fast_inflate(b"\x01\x05\x00\xfa\xffabcde", 5)
    assert rc == 5 and out == b"abcde"
    bad = 0
    for k in range(300):
        s = bytearray(good)
        for _ in range(int(rng.integers(1, 4))):
            s[int(rng.integers(0, len(s)))] ^= 1 << int(rng.integers(0, 8))
        if k % 3 == 0:
            s = s[:int(rng.integers(1, len(s)))]
        rc, out, _ = fast_inflate(bytes(s), len(text) + 70000, piece=[1 << 30, 5000][k % 2])
        if rc < 0:
            bad += 1
        else:                                                  # decodes to something: zlib must agree byte for byte (no CRC at this level)
            try:
                want = zlib.decompressobj(-15).decompress(bytes(s))
            except zlib.error:
                want = None
            assert want is None or out[:len(want)] == want[:len(out)]
    assert bad >= 100                                          # every truncated stream at least (the CRC is the caller's: pcl_gzsrc)
    for k in range(200):
        rc, _, _ = fast_inflate(bytes(rng.integers(0, 256, int(rng.integers(1, 3000)), dtype=np.uint8)), 1 << 20)
        assert rc < 0 or rc >= 0                               # noise: any outcome but a crash or an out-of-bounds write
