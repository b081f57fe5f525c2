"""CPU tier: the C++ FASTQ reader behind pcl_fastq_* (no GPU needed): plain / gzip / multi-member gzip / zstd input,
concatenated files, CRLF, SoA planes with a quality window, record text arena, name hashes and the cross-file
checks of BatchedFqReader (precellar/src/align/fastq.rs:400-448,612-621)."""
import gzip
import os

import numpy as np
import pytest

from precellar_b200 import _ffi, _zstd
from precellar_b200.ingest import FastqReader, GroupReader


def make_records(n, tag="/1"):
    rng = np.random.default_rng(4)
    out = []
    for i in range(n):
        L = int(rng.integers(1, 60))
        s = bytes(rng.choice(np.frombuffer(b"ACGTN", np.uint8), L))
        q = bytes(rng.choice(np.frombuffer(b"F:,#", np.uint8), L))
        out.append((f"read{i}{tag} extra:{i}".encode(), s, q))
    return out


def text_of(recs, eol=b"\n"):
    return b"".join(b"@" + d + eol + s + eol + b"+" + eol + q + eol for d, s, q in recs)


@pytest.fixture(scope="module")
def files(tmp_path_factory):
    td = str(tmp_path_factory.mktemp("ingest"))
    recs = make_records(3000)
    paths = {}
    open(os.path.join(td, "a.fq"), "wb").write(text_of(recs[:1000]))
    with open(os.path.join(td, "b.fq.gz"), "wb") as fh:                    # two gzip members in one file
        fh.write(gzip.compress(text_of(recs[1000:1500])))
        fh.write(gzip.compress(text_of(recs[1500:2000], b"\r\n")))
    w = _zstd.Writer(os.path.join(td, "c.fq.zst"), block=30000)
    w.write(text_of(recs[2000:]))
    w.close()
    paths["all"] = [os.path.join(td, f) for f in ("a.fq", "b.fq.gz", "c.fq.zst")]
    return td, recs, paths


def test_reads_every_format_and_concatenates(files):
    td, recs, paths = files
    r = FastqReader(paths["all"])
    got = 0
    while True:
        b = r.read(700, 32, keep_text=True, qual_window=(4, 16))
        if b is None:
            break
        for i in range(len(b.batch.length)):
            d, s, q = recs[got + i]
            # the definition comes back as BatchedFqReader hands it on: a trailing /1 or /2 of the NAME stripped (fastq.rs:612-621)
            nm, sep, desc = d.partition(b" ")
            if len(nm) > 2 and nm[-2:] in (b"/1", b"/2"):
                nm = nm[:-2]
            assert b.record(i) == (nm + sep + desc, s, q)
            L = len(s)
            assert b.batch.length[i] == L
            row = b.batch.seq[i].tobytes()
            assert row[:min(L, 32)] == s[:32] and row[min(L, 32):] == b"N" * (32 - min(L, 32))
            qrow = b.batch.qual[i].tobytes()                                 # quality bytes [4, 20) of the record
            want = q[4:20]
            assert qrow[:len(want)] == want and qrow[len(want):] == b"!" * (16 - len(want))
        got += len(b.batch.length)
    assert got == len(recs)
    r.close()


def test_text_arena_grows_when_records_outgrow_it(files):
    """The arena is sized from an estimate (bytes per record); long headers / long reads outgrow it: the reader stops in
    front of the record that does not fit, the arena is doubled and the same batch continues (pcl_fastq_text_full), on
    the buffer route (plain LF records) and the line route (CRLF members) alike.  Batches keep their size."""
    td, recs, paths = files
    r = FastqReader(paths["all"])
    got = 0
    sizes = []
    while True:
        b = r.read(700, 32, keep_text=True, text_bytes_per_record=1, qual_window=(4, 16))     # 700 bytes for 700 records
        if b is None:
            break
        sizes.append(len(b.batch.length))
        for i in range(len(b.batch.length)):
            d, s, q = recs[got + i]
            nm, sep, desc = d.partition(b" ")
            if len(nm) > 2 and nm[-2:] in (b"/1", b"/2"):
                nm = nm[:-2]
            assert b.record(i) == (nm + sep + desc, s, q)
            assert b.batch.length[i] == len(s)
            assert b.batch.seq[i].tobytes()[:min(len(s), 32)] == s[:32]
        got += len(b.batch.length)
    assert got == len(recs) and sizes == [700, 700, 700, 700, 200]
    r.close()


def test_group_reader_checks(files, tmp_path):
    td, recs, paths = files
    mates = [(d.replace(b"/1", b"/2"), s[::-1], q) for d, s, q in recs]
    p2 = str(tmp_path / "mate.fq")
    open(p2, "wb").write(text_of(mates))
    g = GroupReader([paths["all"], [p2]], [16, 0], [False, False])
    n = 0
    while True:
        got = g.next(1024)
        if got is None:
            break
        n += len(got[0].batch.length)
        assert got[1].batch.seq is None and np.array_equal(got[0].name_hash, got[1].name_hash)   # /1 and /2 stripped
    assert n == len(recs)
    g.close()
    # unequal record counts
    p3 = str(tmp_path / "short.fq")
    open(p3, "wb").write(text_of(mates[:2500]))
    g = GroupReader([paths["all"], [p3]], [16, 0], [False, False])
    with pytest.raises(_ffi.PclError) as ei:
        while g.next(1024) is not None:
            pass
    assert ei.value.code == _ffi.PCL_ERR_INPUT and "Unequal" in str(ei.value)
    # different names at the same position
    p4 = str(tmp_path / "other.fq")
    open(p4, "wb").write(text_of([(b"x" + d, s, q) for d, s, q in mates]))
    g = GroupReader([paths["all"], [p4]], [16, 0], [False, False])
    with pytest.raises(_ffi.PclError) as ei:
        g.next(100)
    assert "names mismatch" in str(ei.value)


def test_malformed_input_is_an_error(tmp_path):
    p = str(tmp_path / "bad.fq")
    open(p, "wb").write(b"@r1\nACGT\n+\nFFF\n")                                # quality shorter than the sequence
    with pytest.raises(_ffi.PclError):
        FastqReader([p]).read(10, 16)
    open(p, "wb").write(b"r1\nACGT\n+\nFFFF\n")
    with pytest.raises(_ffi.PclError):
        FastqReader([p]).read(10, 16)
    open(p, "wb").write(b"@r1\nACGT\n+\n")
    with pytest.raises(_ffi.PclError):
        FastqReader([p]).read(10, 16)
    with pytest.raises(_ffi.PclError):
        FastqReader([str(tmp_path / "missing.fq")])


# ---------------------------------------------------------------------------------------------------
# pcl_gzsrc_*: decompressed byte stream of a gzip file (BGZF blocks inflated in parallel)
# ---------------------------------------------------------------------------------------------------
def bgzf_bytes(data: bytes, block: int = 60000, eof_marker: bool = True) -> bytes:
    import struct
    import zlib
    out = bytearray()
    chunks = [data[i:i + block] for i in range(0, len(data), block)] + ([b""] if eof_marker else [])
    for c in chunks:
        co = zlib.compressobj(6, zlib.DEFLATED, -15)
        body = co.compress(c) + co.flush()
        bsize = 18 + len(body) + 8 - 1
        out += b"\x1f\x8b\x08\x04" + b"\0\0\0\0" + b"\x00\xff" + struct.pack("<H", 6) + b"BC" + struct.pack("<HH", 2, bsize)
        out += body + struct.pack("<II", zlib.crc32(c) & 0xFFFFFFFF, len(c))
    return bytes(out)


def gz_read_all(path, sizes, n_threads=4):
    from precellar_b200.textingest import GzSource
    src = GzSource(path, n_threads)
    out = bytearray()
    k = 0
    while True:
        buf = bytearray(sizes[k % len(sizes)])
        k += 1
        n = src.readinto(memoryview(buf))
        if n == 0:
            break
        assert n == len(buf) or src.readinto(memoryview(bytearray(16))) == 0       # short only at the end of the file
        out += buf[:n]
    is_bgzf = src.bgzf
    src.close()
    return bytes(out), is_bgzf


@pytest.mark.parametrize("sizes", [[1 << 20], [7, 100000, 1, 65536, 333]])
def test_gzsrc_bgzf_and_plain_gzip(tmp_path, sizes):
    rng = np.random.default_rng(8)
    data = text_of(make_records(20000)) + bytes(rng.integers(0, 256, 70000, dtype=np.uint8))
    p = str(tmp_path / "x.bgz")
    open(p, "wb").write(bgzf_bytes(data, block=int(rng.integers(2000, 65000))))
    got, is_bgzf = gz_read_all(p, sizes)
    assert is_bgzf and got == data
    open(p, "wb").write(bgzf_bytes(data, eof_marker=False))
    assert gz_read_all(p, sizes, n_threads=1) == (data, True)
    p2 = str(tmp_path / "y.gz")                                   # ordinary gzip: single member, then two members
    open(p2, "wb").write(gzip.compress(data))
    assert gz_read_all(p2, sizes) == (data, False)
    open(p2, "wb").write(gzip.compress(data[:50000]) + gzip.compress(data[50000:]))
    assert gz_read_all(p2, sizes) == (data, False)


def test_gzsrc_errors(tmp_path):
    data = text_of(make_records(3000))
    good = bgzf_bytes(data, block=5000)
    p = str(tmp_path / "bad.bgz")
    bad = bytearray(good)
    bad[len(bad) // 2] ^= 0x55                                     # a flipped bit inside some block
    open(p, "wb").write(bytes(bad))
    with pytest.raises(_ffi.PclError):
        gz_read_all(p, [1 << 20])
    open(p, "wb").write(good[:len(good) // 2])                      # cut inside a block
    with pytest.raises(_ffi.PclError):
        gz_read_all(p, [1 << 20])
    open(p, "wb").write(gzip.compress(data)[:-20])                  # truncated ordinary gzip
    with pytest.raises(_ffi.PclError):
        gz_read_all(p, [1 << 20])
    with pytest.raises(_ffi.PclError):
        from precellar_b200.textingest import GzSource
        GzSource(str(tmp_path / "missing.gz"))


# ---------------------------------------------------------------------------------------------------
# textingest.TextSource: the sliding, double-buffered text block in front of the device parser (host logic only;
# pinned memory is replaced by ordinary memory so that the tier needs no GPU)
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_text_source_reassembles_the_byte_stream(tmp_path, monkeypatch, seed):
    from concurrent.futures import ThreadPoolExecutor
    from precellar_b200 import textingest
    monkeypatch.setattr(textingest, "_pinned", lambda n: (np.empty(int(n), np.uint8), 0))
    monkeypatch.setattr(textingest, "_IO_SLICE", 4096)                 # plain files go through the sliced parallel reads
    rng = np.random.default_rng(seed)
    from precellar_b200 import _zstd
    parts = [text_of(make_records(int(rng.integers(200, 1500)))) for _ in range(5)]
    parts[1] = parts[1][:-1]                                           # a file without a final newline
    paths = []
    for k, data in enumerate(parts):
        p = str(tmp_path / f"f{k}")
        if k == 0:
            p += ".fq"; open(p, "wb").write(data)
        elif k == 1:
            p += ".fq.gz"; open(p, "wb").write(gzip.compress(data))
        elif k == 2:
            p += ".bgz"; open(p, "wb").write(bgzf_bytes(data, block=3000))
        elif k == 3:
            p += ".fq.zst"; open(p, "wb").write(_zstd.compress(data[:len(data) // 3]) + _zstd.compress(data[len(data) // 3:]))    # two frames, streamed
        else:
            p += ".fq"; open(p, "wb").write(data)
        paths.append(p)
    want = parts[0] + parts[1] + b"\n" + parts[2] + parts[3] + parts[4]   # the missing newline is supplied between the files
    cap = int(rng.integers(3000, 40000))
    src = textingest.TextSource(paths, cap)
    pool = ThreadPoolExecutor(max_workers=1)
    got = bytearray()
    try:
        src.prefetch(pool, cap)
        src.advance(0)
        for _ in range(100000):
            blk = src.block()
            if len(blk) == 0 and src.eof:
                break
            # sometimes consume everything, sometimes very little (the tail then outgrows the reserved gap and the
            # block is rebuilt, possibly in larger buffers), sometimes nothing at all
            u = rng.random()
            used = len(blk) if u < 0.3 else int(rng.integers(0, len(blk) + 1)) if u < 0.8 else int(rng.integers(0, min(len(blk), 50) + 1))
            got += blk[:used].tobytes()
            assert src.pending is None
            src.prefetch(pool, int(rng.integers(0, 2 * cap)))
            src.advance(used)
        else:
            raise AssertionError("the source did not reach the end of its input")
        assert bytes(got) == want
    finally:
        src.close()
        pool.shutdown()


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_mmap_source_walks_plain_files_without_copies(tmp_path, seed):
    """MmapSource: windows of the file maps, never spanning two files; only for uncompressed files that end in a newline."""
    from precellar_b200 import textingest
    rng = np.random.default_rng(seed)
    parts = [text_of(make_records(int(rng.integers(200, 1500)))) for _ in range(3)]
    paths = []
    for k, data in enumerate(parts):
        p = str(tmp_path / f"m{k}.fq")
        open(p, "wb").write(data)
        paths.append(p)
    assert textingest.MmapSource.eligible(paths)
    open(str(tmp_path / "nonl.fq"), "wb").write(parts[0][:-1])
    open(str(tmp_path / "z.fq.gz"), "wb").write(gzip.compress(parts[0]))
    open(str(tmp_path / "empty.fq"), "wb").close()
    for bad in ("nonl.fq", "z.fq.gz", "empty.fq"):
        assert not textingest.MmapSource.eligible([paths[0], str(tmp_path / bad)])
    src = textingest.MmapSource(paths, 0)
    assert abs(src.bytes_per_record(0.0) - len(b"".join(text_of([r]) for r in make_records(64))) / 64) < 1e-9
    got = bytearray()
    try:
        src.prefetch(None, 5000)
        src.advance(0)
        for _ in range(100000):
            blk = src.block()
            if len(blk) == 0 and src.eof:
                break
            u = rng.random()
            used = len(blk) if u < 0.4 else int(rng.integers(0, len(blk) + 1))
            got += blk[:used].tobytes()
            src.prefetch(None, int(rng.integers(0, 20000)))
            src.advance(used)
        else:
            raise AssertionError("the source did not reach the end of its input")
        assert bytes(got) == b"".join(parts)
    finally:
        del blk
        src.close()


# ---------------------------------------------------------------------------------------------------
# textingest.BgzfDeviceSource: the sliding block in DEVICE memory, fed by the device inflate.  Host logic only: the
# device buffers are numpy arrays and the kernel is the host run of csrc/inflate_core.h (tests/emul).
# ---------------------------------------------------------------------------------------------------
class FakeInflater:
    def __init__(self):
        self.slots = {}
        self.closed = False
        self.runs = 0

    def buffer(self, slot, nbytes):
        if nbytes == 0:
            self.slots.pop(slot, None)
            return 0
        if slot not in self.slots or len(self.slots[slot]) < nbytes:
            self.slots[slot] = np.full(nbytes, 0xEE, np.uint8)
        return self.slots[slot].ctypes.data

    def run(self, comp, members, n, slot, dst_off):
        import ctypes as C
        from tests import util as U
        L = U.emul_lib()
        L.emu_inflate.restype = C.c_int64
        L.emu_inflate.argtypes = [C.c_char_p, C.c_uint64, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint64, C.c_int]
        out = self.slots[slot][dst_off:]
        rc = L.emu_inflate(comp.tobytes() + b"\0" * 16, len(comp), C.addressof(members), n, out.ctypes.data, len(out), self.runs & 1)
        self.runs += 1
        if rc:
            raise _ffi.PclError(_ffi.PCL_ERR_INPUT, f"corrupt BGZF block: member {(rc >> 8) - 1}, error {rc & 0xFF}")

    def move(self, ss, so, ds, do, n):
        assert so + n <= len(self.slots[ss]) and do + n <= len(self.slots[ds])
        assert ss != ds or so + n <= do or do + n <= so
        self.slots[ds][do:do + n] = self.slots[ss][so:so + n]

    def fetch(self, slot, off, n):
        assert off + n <= len(self.slots[slot])
        return self.slots[slot][off:off + n].tobytes()

    def poke(self, slot, off, data):
        self.slots[slot][off:off + len(data)] = np.frombuffer(data, np.uint8)

    def close(self):
        self.closed = True


@pytest.mark.parametrize("seed", [0, 1, 2, 3, 4])
def test_bgzf_device_source_reassembles_the_byte_stream(tmp_path, seed):
    from concurrent.futures import ThreadPoolExecutor
    from precellar_b200 import textingest
    rng = np.random.default_rng(seed)
    parts = [text_of(make_records(int(rng.integers(200, 3000)))) for _ in range(4)]
    parts[1] = parts[1][:-1]                                           # a file without a final newline
    paths = []
    for k, data in enumerate(parts):
        p = str(tmp_path / f"f{k}.fq.gz")
        open(p, "wb").write(bgzf_bytes(data, block=int(rng.integers(500, 65000)), eof_marker=bool(k % 2)))
        paths.append(p)
    assert textingest.BgzfDeviceSource.eligible(paths)
    want = parts[0] + parts[1] + b"\n" + parts[2] + parts[3]
    cap = int(rng.integers(3000, 200000))
    backend = FakeInflater()
    src = textingest.BgzfDeviceSource(paths, cap, backend)
    pool = ThreadPoolExecutor(max_workers=1)
    got = bytearray()
    try:
        src.prefetch(pool, cap)
        src.advance(0)
        for _ in range(100000):
            blk = src.block()
            if len(blk) == 0 and src.eof:
                break
            assert blk.ptr == backend.slots[blk.slot].ctypes.data + blk.start
            text = bytes(blk)
            u = rng.random()
            used = len(blk) if u < 0.3 else int(rng.integers(0, len(blk) + 1)) if u < 0.8 else int(rng.integers(0, min(len(blk), 50) + 1))
            got += text[:used]
            src.prefetch(pool, int(rng.integers(0, 2 * cap)))
            src.advance(used)
        else:
            raise AssertionError("the source did not reach the end of its input")
        assert bytes(got) == want
        assert len(backend.slots) == 2                                 # rebuilt pairs were released
    finally:
        src.close()
        pool.shutdown()
    assert backend.closed


def test_bgzf_device_source_eligibility_and_errors(tmp_path):
    from concurrent.futures import ThreadPoolExecutor
    from precellar_b200 import textingest
    data = text_of(make_records(3000))
    good = bgzf_bytes(data, block=5000)
    p_bgzf, p_gz, p_plain = str(tmp_path / "a.gz"), str(tmp_path / "b.gz"), str(tmp_path / "c.fq")
    open(p_bgzf, "wb").write(good)
    open(p_gz, "wb").write(gzip.compress(data))
    open(p_plain, "wb").write(data)
    assert textingest.BgzfDeviceSource.eligible([p_bgzf])
    assert not textingest.BgzfDeviceSource.eligible([p_bgzf, p_gz])    # an ordinary gzip file among them: the host path
    assert not textingest.BgzfDeviceSource.eligible([p_plain])
    assert not textingest.BgzfDeviceSource.eligible([str(tmp_path / "missing.gz")])
    assert not textingest.BgzfDeviceSource.eligible([])
    pool = ThreadPoolExecutor(max_workers=1)

    def read_all(blob):
        p = str(tmp_path / "x.gz")
        open(p, "wb").write(blob)
        src = textingest.BgzfDeviceSource([p], 20000, FakeInflater())
        out = bytearray()
        try:
            src.prefetch(pool, 20000)
            src.advance(0)
            while not (len(src.block()) == 0 and src.eof):
                out += bytes(src.block())
                n = len(src.block())
                src.prefetch(pool, 20000)
                src.advance(n)
        finally:
            src.close()
        return bytes(out)

    assert read_all(good) == data
    bad = bytearray(good)
    bad[len(bad) // 2] ^= 0x55                                         # a flipped bit inside some member
    with pytest.raises(_ffi.PclError):
        read_all(bytes(bad))
    with pytest.raises(_ffi.PclError):
        read_all(good[:len(good) // 2])                                # cut inside a member
    with pytest.raises(_ffi.PclError):
        read_all(good[:20000] + gzip.compress(data))                   # an ordinary gzip member behind BGZF members
    pool.shutdown()


def test_gzsrc_ordinary_gzip_members_headers_and_refills(tmp_path, monkeypatch):
    """The fastinflate route of pcl_gzsrc: members with every optional header field, an empty member, input buffers of
    8 KB..1 MB (many refills, stored blocks and headers across them), odd read sizes; zlib route for comparison."""
    import struct
    import zlib
    rng = np.random.default_rng(9)
    text = text_of(make_records(30000))
    noise = bytes(rng.integers(0, 256, 200000, dtype=np.uint8))

    def member(data, level=6, flags=0, extra=b"", name=b"", comment=b""):
        co = zlib.compressobj(level, zlib.DEFLATED, -15)
        body = co.compress(data) + co.flush()
        head = b"\x1f\x8b\x08" + bytes([flags]) + b"\0\0\0\0\x00\x03"
        if flags & 4:
            head += struct.pack("<H", len(extra)) + extra
        if flags & 8:
            head += name + b"\0"
        if flags & 16:
            head += comment + b"\0"
        if flags & 2:
            head += struct.pack("<H", zlib.crc32(head) & 0xFFFF)
        return head + body + struct.pack("<II", zlib.crc32(data) & 0xFFFFFFFF, len(data) & 0xFFFFFFFF)

    parts = [text[:400000], b"", noise, text[400000:], b"x"]
    blob = (member(parts[0], 6, 8, name=b"reads_R1.fastq") + member(parts[1], 6) + member(parts[2], 0, 4 | 16, extra=b"ab\x03\x00xyz", comment=b"c" * 300) +
            member(parts[3], 1, 2 | 8 | 16, name=b"n", comment=b"") + member(parts[4], 9, 4, extra=b""))
    want = b"".join(parts)
    p = str(tmp_path / "m.gz")
    open(p, "wb").write(blob)
    for buf in (8192, 10000, 70001, 1 << 20, None):
        if buf is None:
            monkeypatch.delenv("PCL_GZSRC_BUF", raising=False)
        else:
            monkeypatch.setenv("PCL_GZSRC_BUF", str(buf))
        for sizes in ([1 << 20], [7, 100000, 1, 65536, 333], [3_000_000]):
            assert gz_read_all(p, sizes) == (want, False), (buf, sizes)
    monkeypatch.setenv("PCL_ZLIB_INFLATE", "1")
    assert gz_read_all(p, [1 << 20]) == (want, False)
    monkeypatch.delenv("PCL_ZLIB_INFLATE")
    monkeypatch.setenv("PCL_GZSRC_BUF", "8192")
    one = member(text[:300000], 6)
    for bad in (one[:-8] + struct.pack("<II", (zlib.crc32(text[:300000]) ^ 4) & 0xFFFFFFFF, 300000),      # CRC-32
                one[:-4] + struct.pack("<I", 299999),                                                       # ISIZE
                one + b"garbage after the member",                                                          # no gzip header
                one + one[:200],                                                                            # second member cut off
                one[:len(one) // 3],                                                                        # cut inside the stream
                one[:10], b"\x1f\x8b\x08"):                                                                 # cut inside / right behind the header
        open(p, "wb").write(bad)
        with pytest.raises(_ffi.PclError):
            gz_read_all(p, [50000])
    open(p, "wb").write(one + one)
    assert gz_read_all(p, [50000])[0] == text[:300000] * 2


class FakeChunk:
    """Stands in for pipeline.Chunk in TextGroupReader.next_into: the line index is numpy, the 'planes' are the bytes consumed."""

    def __init__(self, capacity, n_reads):
        self.capacity, self.n_reads = capacity, n_reads
        self.text = [np.zeros(0, np.uint8)] * n_reads
        self.taken = [bytearray() for _ in range(n_reads)]
        self.scanned_bytes = 0
        self.n = 0

    def text_scan(self, r, b):
        self.text[r] = np.frombuffer(bytes(b), np.uint8) if hasattr(b, "ptr") else np.array(b, copy=True)
        self.scanned_bytes += len(self.text[r])

    def text_lines(self, r):
        return int((self.text[r] == 10).sum())

    def reset(self, n):
        self.n = n

    def text_fill(self, r):
        used = int(np.flatnonzero(self.text[r] == 10)[4 * self.n - 1]) + 1
        self.taken[r] += self.text[r][:used].tobytes()
        return used

    def text_names_check(self):
        return -1


@pytest.mark.parametrize("target,ahead", [(300, 1 << 30), (300, 20000), (5000, 1 << 30), (50, 0)])
def test_text_group_reader_over_bgzf_device_sources(tmp_path, monkeypatch, target, ahead):
    """TextGroupReader.next_into over BgzfDeviceSource (emulated device): batches inflated far ahead of the chunk that is
    parsed, the reader looking at one chunk's worth at a time; every byte of every file arrives once, in order."""
    import types
    from precellar_b200 import textingest
    rng = np.random.default_rng(target + ahead % 7)
    n = 4000
    recs1, recs2 = make_records(n), make_records(n)
    recs2 = [(nm, s * 3, q * 3) for nm, s, q in recs2]                  # the second read's records are three times as long
    t1, t2 = text_of(recs1), text_of(recs2)
    cut1, cut2 = text_of(recs1[:1500]), text_of(recs2[:2500])
    files = [[str(tmp_path / "a1.gz"), str(tmp_path / "b1.gz")], [str(tmp_path / "a2.gz"), str(tmp_path / "b2.gz")]]
    open(files[0][0], "wb").write(bgzf_bytes(cut1, block=7000))
    open(files[0][1], "wb").write(bgzf_bytes(t1[len(cut1):], block=65000))
    open(files[1][0], "wb").write(bgzf_bytes(cut2[:-1], block=30000))     # no final newline: one is supplied
    open(files[1][1], "wb").write(bgzf_bytes(t2[len(cut2):], block=900))
    fakes = []

    def make(ctx):
        fakes.append(FakeInflater())
        return fakes[-1]
    monkeypatch.setattr(textingest, "DeviceInflater", make)
    monkeypatch.setattr(textingest.BgzfDeviceSource, "AHEAD_BYTES", ahead)
    monkeypatch.setenv("PCL_DEVICE_INFLATE", "1")
    rdr = textingest.TextGroupReader(types.SimpleNamespace(ctx=None), files, target, [40, 40])      # a poor guess of the record size
    assert [type(s).__name__ for s in rdr.sources] == ["BgzfDeviceSource"] * 2
    ch = FakeChunk(target, 2)
    total = 0
    try:
        while True:
            got = rdr.next_into(ch)
            if got == 0:
                break
            assert got <= target
            total += got
    finally:
        rdr.close()
    assert total == n and bytes(ch.taken[0]) == t1 and bytes(ch.taken[1]) == t2
    runs = sum(f.runs for f in fakes)
    if ahead == 1 << 30:                                               # a few launches per file, not one per chunk
        assert runs <= 8, runs
        assert ch.scanned_bytes < 4 * (len(t1) + len(t2))              # and no chunk indexed the whole buffered text
    assert all(f.closed for f in fakes)
