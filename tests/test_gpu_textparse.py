"""GPU tier: device-side FASTQ parse (pcl_text_*, csrc/textparse.cuh) against the host record reader
(pcl_fastq_read, itself pinned by tests/test_ingest.py) and a pure-Python parse of the same text: identical
sequence / quality-window / length planes and name hashes, chunking with carried-over tails, concatenated plain /
gzip files, CRLF, and the cross-file checks of BatchedFqReader (precellar/src/align/fastq.rs:400-448)."""
import gzip
import os

import numpy as np
import pytest

from tests import util as U

pytestmark = pytest.mark.gpu


def make_records(n, tag, rng, lo=1, hi=60):
    out = []
    for i in range(n):
        L = int(rng.integers(lo, hi + 1))
        s = bytes(rng.choice(np.frombuffer(b"ACGTN", np.uint8), L))
        q = bytes(rng.choice(np.frombuffer(b"F:,#@+", np.uint8), L))       # '@' and '+' also occur inside quality lines
        d = f"r{i}{tag}".encode() + (b" desc:%d" % i if i % 3 else b"")
        out.append((d, s, q))
    return out


def text_of(recs, eol=b"\n"):
    return b"".join(b"@" + d + eol + s + eol + b"+" + (d if i % 5 == 0 else b"") + eol + q + eol
                    for i, (d, s, q) in enumerate(recs))


def fnv1a(name: bytes) -> int:
    h = 1469598103934665603
    for c in name:
        h = ((h ^ c) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h


def expected_planes(recs, stride, qoff, qstride):
    n = len(recs)
    seq = np.full((n, stride), ord("N"), np.uint8)
    qual = np.full((n, max(qstride, 1)), ord("!"), np.uint8)
    length = np.zeros(n, np.uint16)
    nh = np.zeros(n, np.uint64)
    for i, (d, s, q) in enumerate(recs):
        c = min(len(s), stride)
        seq[i, :c] = np.frombuffer(s[:c], np.uint8)
        w = q[qoff:qoff + qstride]
        qual[i, :len(w)] = np.frombuffer(w, np.uint8)
        length[i] = min(len(s), 65535)
        nm = d.split(b" ")[0].split(b"\t")[0]
        if len(nm) > 2 and nm[-2:] in (b"/1", b"/2"):
            nm = nm[:-2]
        nh[i] = fnv1a(nm)
    return seq, qual[:, :qstride], length, nh


def d2h(ptr, nbytes, dtype, shape):
    from tools.synth import Synth
    out = np.empty(shape, dtype)
    assert out.nbytes == nbytes
    Synth.memcpy_d2h(0, out.ctypes.data, ptr, nbytes)
    return out


def v3_like_layout(rng, bc_len=16, umi_len=12):
    wl = U.gen_whitelist(rng, 500, bc_len)
    return U.make_layout([dict(read_id="R1", elems=[U.bc("cb", bc_len), U.umi(umi_len)], max_len=bc_len + umi_len),
                          dict(read_id="R2", elems=[U.cdna(1, 1000)], max_len=0, is_reverse=True)], {"cb": wl}), wl


@pytest.fixture(scope="module")
def pipe_and_layout():
    from precellar_b200.pipeline import BarcodePipeline
    rng = np.random.default_rng(5)
    layout, wl = v3_like_layout(rng)
    pipe = BarcodePipeline(layout)
    yield pipe, layout, wl
    pipe.close()


@pytest.mark.parametrize("eol", [b"\n", b"\r\n"])
def test_planes_match_host_reader(pipe_and_layout, eol):
    pipe, layout, _ = pipe_and_layout
    rng = np.random.default_rng(17)
    recs = make_records(5000, "/1", rng)
    text = np.frombuffer(text_of(recs, eol), np.uint8).copy()
    ch = pipe.new_chunk(8192)
    try:
        ch.text_scan(0, text)
        assert ch.text_lines(0) == 4 * len(recs)
        n = 4321
        ch.reset(n)
        used = ch.text_fill(0)
        assert used == len(text_of(recs[:n], eol))
        st, (qo, qs) = pipe.stride(0), pipe.qual_window(0)
        s_ptr, q_ptr, l_ptr = ch.device_ptrs(0)
        seq, qual, length, _ = expected_planes(recs[:n], st, qo, qs)
        assert np.array_equal(d2h(s_ptr, n * st, np.uint8, (n, st)), seq)
        assert np.array_equal(d2h(q_ptr, n * qs, np.uint8, (n, qs)), qual)
        assert np.array_equal(d2h(l_ptr, n * 2, np.uint16, (n,)), length)
        ends = ch.text_line_ends(0, 8)
        t = text.tobytes()
        assert [t[e] for e in ends] == [10] * 8 and t[:ends[0]].rstrip(b"\r") == b"@" + recs[0][0]
    finally:
        ch.close()
        pipe.chunks.remove(ch)


def write(path, data):
    if path.endswith(".gz"):
        with gzip.open(path, "wb") as fh:
            fh.write(data)
    else:
        with open(path, "wb") as fh:
            fh.write(data)


@pytest.mark.parametrize("target", [700, 4096, 100000])
def test_group_reader_matches_host_reader_end_to_end(tmp_path, target, monkeypatch):
    """Concatenated plain + gzip files, the last one without a trailing newline; chunk targets that force many
    carried-over tails; counts, keys and statuses must equal the host-reader path."""
    from precellar_b200 import textingest
    from precellar_b200.ingest import GroupReader
    if target != 700:
        monkeypatch.setattr(textingest, "_IO_SLICE", 10007)        # plain files are read in parallel slices
    from precellar_b200.pipeline import BarcodePipeline
    from precellar_b200.textingest import TextGroupReader
    rng = np.random.default_rng(23)
    layout, wl = v3_like_layout(rng)
    n = 20000
    full = U.gen_reads(rng, layout, n, [64, 256], p_short=0.02, p_n=0.02, read_len=[28, 200])
    for b in full:
        b.length[:] = np.maximum(b.length, 1)
    td = str(tmp_path)

    def recs_of(b, tag):
        return [(f"q{i}{tag}".encode(), b.seq[i, :int(b.length[i])].tobytes(), b.qual[i, :int(b.length[i])].tobytes())
                for i in range(n)]
    r1, r2 = recs_of(full[0], "/1"), recs_of(full[1], "/2")
    write(os.path.join(td, "a1.fq"), text_of(r1[:7000]))
    write(os.path.join(td, "b1.fq.gz"), text_of(r1[7000:15000], b"\r\n"))
    write(os.path.join(td, "c1.fq"), text_of(r1[15000:])[:-1])                   # no newline at the end of the file
    write(os.path.join(td, "a2.fq.gz"), text_of(r2[:11000]))
    write(os.path.join(td, "b2.fq"), text_of(r2[11000:]))
    files = [[os.path.join(td, x) for x in ("a1.fq", "b1.fq.gz", "c1.fq")], [os.path.join(td, x) for x in ("a2.fq.gz", "b2.fq")]]

    def run(device_parse):
        pipe = BarcodePipeline(layout)
        try:
            chunks = []
            if device_parse:
                rdr = TextGroupReader(pipe, files, target, [100, 300])
                while True:
                    ch = pipe.new_chunk(target)
                    got = rdr.next_into(ch)
                    if got == 0:
                        break
                    pipe.count(ch)
                    chunks.append(ch)
                assert rdr.records == n
                rdr.close()
            else:
                rdr = GroupReader(files, [pipe.stride(0), pipe.stride(1)], [False, False],
                                  qual_windows=[pipe.qual_window(0), pipe.qual_window(1)])
                while True:
                    got = rdr.next(target)
                    if got is None:
                        break
                    ch = pipe.new_chunk(len(got[0].batch.length))
                    ch.reset(len(got[0].batch.length))
                    for r in range(2):
                        ch.upload(r, got[r].batch)
                    pipe.count(ch)
                    pipe.sync()
                    chunks.append(ch)
                rdr.close()
            pipe.allreduce()
            outs = []
            for ch in chunks:
                pipe.correct(ch, 0.9)
                outs.append([ch.fetch(r) for r in range(2)])
            cat = lambda f: np.concatenate([f(o) for o in outs])
            res = dict(fs1=cat(lambda o: o[0].fstatus), fs2=cat(lambda o: o[1].fstatus), bc=cat(lambda o: o[0].bc_key[0]),
                       umi=cat(lambda o: o[0].umi_key[0]), tc=cat(lambda o: o[1].tcoord))
            counts = pipe.counts(0)
            return res, counts, pipe.qc()
        finally:
            pipe.close()
    a, ca, qa = run(False)
    b, cb, qb = run(True)
    for k in a:
        assert np.array_equal(a[k], b[k]), k
    assert np.array_equal(ca[0], cb[0]) and ca[1:] == cb[1:]
    assert np.array_equal(qa, qb)
    assert len(a["fs1"]) == n and (a["fs1"] >> 3).any()


def _one_chunk(pipe, files, target=1000, bpr=(100, 100)):
    from precellar_b200.textingest import TextGroupReader
    rdr = TextGroupReader(pipe, files, target, list(bpr))
    try:
        total = 0
        while True:
            ch = pipe.new_chunk(target)
            got = rdr.next_into(ch)
            ch.close()
            pipe.chunks.remove(ch)
            if got == 0:
                return total
            total += got
    finally:
        rdr.close()


def test_cross_file_checks_and_malformed_records(pipe_and_layout, tmp_path):
    from precellar_b200 import _ffi
    pipe, _, _ = pipe_and_layout
    rng = np.random.default_rng(3)
    td = str(tmp_path)
    r1, r2 = make_records(2500, "/1", rng), make_records(2500, "/2", rng)
    p1, p2 = os.path.join(td, "x1.fq"), os.path.join(td, "x2.fq")
    write(p1, text_of(r1)); write(p2, text_of(r2))
    assert _one_chunk(pipe, [[p1], [p2]]) == 2500
    # unequal record counts (fastq.rs:435-436)
    write(p2, text_of(r2[:2400]))
    with pytest.raises(_ffi.PclError, match="Unequal number of reads"):
        _one_chunk(pipe, [[p1], [p2]])
    # names differ in group 1234 (fastq.rs:438-442)
    bad = list(r2)
    bad[1234] = (b"zzz/2", bad[1234][1], bad[1234][2])
    write(p2, text_of(bad))
    with pytest.raises(_ffi.PclError, match="read names mismatch \\(group 1234\\)"):
        _one_chunk(pipe, [[p1], [p2]])
    # malformed records are reported with their index
    for mutate, msg in ((lambda t: t.replace(b"@r700/2", b"r700/2", 1), "does not start with '@'"),
                        (lambda t: t.replace(b"\n+\n", b"\n-\n", 1), "missing '\\+' line")):
        write(p2, mutate(text_of(r2)))
        with pytest.raises(_ffi.PclError, match=msg):
            _one_chunk(pipe, [[p1], [p2]])
    d, s, q = r2[10]
    broken = list(r2); broken[10] = (d, s, q + b"F")
    write(p2, text_of(broken))
    with pytest.raises(_ffi.PclError, match="record 10 .*lengths differ"):
        _one_chunk(pipe, [[p1], [p2]], target=4000, bpr=(200, 200))
    # a truncated last record
    write(p2, text_of(r2)[:-30])
    with pytest.raises(_ffi.PclError):
        _one_chunk(pipe, [[p1], [p2]])


def test_dense_lines_and_tiny_blocks(pipe_and_layout, tmp_path):
    """One-base records (more line ends than the index holds per block) and blocks smaller than one record still
    make progress and keep every record."""
    pipe, _, _ = pipe_and_layout
    rng = np.random.default_rng(9)
    td = str(tmp_path)
    n = 30000
    tiny1 = [(b"a", b"A", b"F")] * n
    tiny2 = [(b"a", b"C", b"F")] * n
    p1, p2 = os.path.join(td, "t1.fq"), os.path.join(td, "t2.fq")
    write(p1, text_of(tiny1)); write(p2, text_of(tiny2))
    assert _one_chunk(pipe, [[p1], [p2]], target=20000, bpr=(8, 8)) == n
    big1 = make_records(12, "/1", rng, 30000, 40000)             # one record is larger than the initial block
    big2 = make_records(12, "/2", rng, 10, 20)
    write(p1, text_of(big1)); write(p2, text_of(big2))
    assert _one_chunk(pipe, [[p1], [p2]], target=4, bpr=(100, 100)) == 12
