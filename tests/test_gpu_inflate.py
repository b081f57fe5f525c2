"""GPU tier: BGZF members inflated on the device (k_bgzf_inflate through pcl_inflater_*) against zlib, and the text
reader fed by it (textingest.BgzfDeviceSource) against the same reader fed by the host inflate."""
import ctypes as C
import os
import struct
import time
import zlib

import numpy as np
import pytest

from precellar_b200 import _ffi
from tests import util as U
from tests.test_inflate import deflate, fastq_text

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    L = _ffi.lib()
    h = C.c_void_p()
    _ffi.check(None, L.pcl_ctx_create(C.byref(h), 0))
    yield h
    L.pcl_ctx_destroy(h)


def member(text: bytes, stream: bytes, crc=None, isize=None) -> bytes:
    bsize = 18 + len(stream) + 8 - 1
    return (b"\x1f\x8b\x08\x04" + b"\0\0\0\0" + b"\x00\xff" + struct.pack("<H", 6) + b"BC" + struct.pack("<HH", 2, bsize) + stream +
            struct.pack("<II", (zlib.crc32(text) & 0xFFFFFFFF) if crc is None else crc, len(text) if isize is None else isize))


def inflate_on_device(ctx, blob: bytes, out_bytes: int) -> bytes:
    from precellar_b200.textingest import DeviceInflater, bgzf_walk
    inf = DeviceInflater(ctx)
    try:
        comp = np.frombuffer(blob, np.uint8)
        cap = 1 << 16
        members = (_ffi.BgzfMember * cap)()
        inf.buffer(0, out_bytes + 64)
        pos = got = 0
        while pos < len(comp):
            rc, n, used, out = bgzf_walk(comp[pos:], 1 << 40, members, cap)
            assert rc == 0 and used > 0
            inf.run(comp[pos:pos + used], members, n, 0, got)
            pos += used
            got += out
        assert got == out_bytes
        return inf.fetch(0, 0, got), inf.stats()
    finally:
        inf.close()


def test_members_of_every_kind_inflate_like_zlib(ctx):
    rng = np.random.default_rng(31)
    text = fastq_text(rng, 6000)
    noise = bytes(rng.integers(0, 256, 70000, dtype=np.uint8))
    parts, streams = [], []

    def add(p, s):
        parts.append(p)
        streams.append(s)
    pos = 0
    for size in [65536, 65280, 1, 2, 3, 31, 32, 33, 257, 258, 259, 4096, 60000, 1000] * 3:     # dynamic blocks at three levels
        p = text[pos:pos + size]
        pos += size
        add(p, deflate(p, [1, 6, 9][len(parts) % 3]))
    add(text[:50000], deflate(text[:50000], 6, zlib.Z_FIXED))
    add(noise[:40000], deflate(noise[:40000], 0))                                      # stored blocks
    add(text[:65536], deflate(text[:65536], 6, flush_every=5000))                      # many blocks per member
    add(text[1000:40000], deflate(text[1000:40000], 1, zlib.Z_HUFFMAN_ONLY))
    add(b"\0" * 65536, deflate(b"\0" * 65536, 9))                                      # distance 1, length 258
    add(b"AB" * 30000, deflate(b"AB" * 30000, 9))
    add(bytes(range(256)) * 200, deflate(bytes(range(256)) * 200, 9))
    add(bytes(rng.integers(0, 4, 65536, dtype=np.uint8)), deflate(bytes(rng.integers(0, 4, 65536, dtype=np.uint8)), 6, zlib.Z_RLE))
    parts[-1] = zlib.decompress(streams[-1], -15)
    blob = b"".join(member(p, s) for p, s in zip(parts, streams)) + member(b"", deflate(b"", 6))
    want = b"".join(parts)
    got, stats = inflate_on_device(ctx, blob, len(want))
    assert got == want
    assert stats["out_bytes"] == len(want) and stats["launches"] >= 1


def test_a_large_file_and_its_rate(ctx):
    rng = np.random.default_rng(32)
    text = fastq_text(rng, 150000, 100)                                                # ~ 40 MB of FASTQ text
    blob = b"".join(member(text[i:i + 65280], deflate(text[i:i + 65280], 6)) for i in range(0, len(text), 65280))
    t0 = time.perf_counter()
    got, stats = inflate_on_device(ctx, blob, len(text))
    wall = time.perf_counter() - t0
    assert got == text
    rate = len(text) / (stats["kernel_ms"] * 1e-3) / 1e9
    print(f"\nk_bgzf_inflate: {len(text) / 1e6:.1f} MB of text from {len(blob) / 1e6:.1f} MB in {stats['kernel_ms']:.2f} ms "
          f"({rate:.1f} GB/s of text; wall {wall * 1e3:.0f} ms with the copies)")
    out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(out_dir):
        with open(os.path.join(out_dir, "inflate_rate.json"), "w") as fh:
            fh.write('{"text_bytes": %d, "compressed_bytes": %d, "members": %d, "kernel_ms": %.3f, "text_GBps": %.2f, "wall_ms": %.1f}\n'
                     % (len(text), len(blob), (len(text) + 65279) // 65280, stats["kernel_ms"], rate, wall * 1e3))


def test_damaged_members_are_input_errors(ctx):
    rng = np.random.default_rng(33)
    text = fastq_text(rng, 300)[:30000]
    good = deflate(text, 6)
    for blob in (member(text, good, crc=(zlib.crc32(text) ^ 1) & 0xFFFFFFFF),
                 member(text, good, isize=len(text) + 5),
                 member(text, good[:len(good) // 2] + bytes(len(good) - len(good) // 2)),
                 member(text, b"\x07" + good),
                 member(text[:100], deflate(text[:100], 6)) + member(text, good[:-9] + b"\xff" * 9)):
        with pytest.raises(_ffi.PclError) as ei:
            inflate_on_device(ctx, blob, 40000)
        assert ei.value.code == _ffi.PCL_ERR_INPUT and "BGZF" in str(ei.value)


@pytest.mark.parametrize("target", [700, 20000])
def test_reader_on_bgzf_files_equals_the_host_inflate(tmp_path, monkeypatch, target):
    """TextGroupReader over BGZF files: members inflated on the device (BgzfDeviceSource) or by the host threads
    (PCL_DEVICE_INFLATE=0); every result plane and count must be equal.  Two files per read, one with CRLF line ends,
    one without a final newline; the small target forces many carried-over tails."""
    from precellar_b200.pipeline import BarcodePipeline
    from precellar_b200.textingest import TextGroupReader
    from tests.test_gpu_textparse import text_of, v3_like_layout
    from tests.test_ingest import bgzf_bytes
    rng = np.random.default_rng(41)
    layout, wl = v3_like_layout(rng)
    n = 20000
    full = U.gen_reads(rng, layout, n, [64, 256], p_short=0.02, p_n=0.02, read_len=[28, 200])
    for b in full:
        b.length[:] = np.maximum(b.length, 1)

    def recs_of(b, tag):
        return [(f"q{i}{tag}".encode(), b.seq[i, :int(b.length[i])].tobytes(), b.qual[i, :int(b.length[i])].tobytes()) for i in range(n)]
    r1, r2 = recs_of(full[0], "/1"), recs_of(full[1], "/2")
    td = str(tmp_path)

    def put(name, data, block):
        p = os.path.join(td, name)
        open(p, "wb").write(bgzf_bytes(data, block=block))
        return p
    files = [[put("a1.fq.gz", text_of(r1[:9000], b"\r\n"), 3000), put("b1.fq.gz", text_of(r1[9000:])[:-1], 65280)],
             [put("a2.fq.gz", text_of(r2[:12000])[:-1], 60000), put("b2.fq.gz", text_of(r2[12000:]), 1000)]]

    def run(device_inflate):
        monkeypatch.setenv("PCL_DEVICE_INFLATE", "1" if device_inflate else "0")
        pipe = BarcodePipeline(layout)
        try:
            rdr = TextGroupReader(pipe, files, target, [100, 300])
            kinds = {type(s).__name__ for s in rdr.sources}
            assert kinds == ({"BgzfDeviceSource"} if device_inflate else {"TextSource"})
            chunks = []
            while True:
                ch = pipe.new_chunk(target)
                if rdr.next_into(ch) == 0:
                    break
                pipe.count(ch)
                chunks.append(ch)
            assert rdr.records == n
            rdr.close()
            pipe.allreduce()
            outs = []
            for ch in chunks:
                pipe.correct(ch, 0.9)
                outs.append([ch.fetch(r) for r in range(2)])
            cat = lambda f: np.concatenate([f(o) for o in outs])
            res = dict(fs1=cat(lambda o: o[0].fstatus), fs2=cat(lambda o: o[1].fstatus), bc=cat(lambda o: o[0].bc_key[0]),
                       umi=cat(lambda o: o[0].umi_key[0]), tc=cat(lambda o: o[1].tcoord))
            return res, pipe.counts(0), pipe.qc()
        finally:
            pipe.close()
    a, ca, qa = run(False)
    b, cb, qb = run(True)
    for k in a:
        assert np.array_equal(a[k], b[k]), k
    assert np.array_equal(ca[0], cb[0]) and ca[1:] == cb[1:]
    assert np.array_equal(qa, qb)
    assert len(a["fs1"]) == n and (a["fs1"] >> 3).any()
