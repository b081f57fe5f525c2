"""CPU tier: the device writer's record assembly (csrc/emit_core.h, run on the host by tests/emul) against the records
the oracle renders - the make_fastq record loop of the reference (python/src/lib.rs:148-168)."""
import ctypes as C

import numpy as np
import pytest

from precellar_b200 import _ffi
from tests import util as U
from tests.test_emul_parity import _layouts


class _EmitIO(C.Structure):
    _fields_ = [("text", C.c_void_p), ("nl", C.c_void_p), ("len", C.c_void_p), ("fstatus", C.c_void_p), ("tcoord", C.c_void_p),
                ("bc_key", C.c_void_p * 4), ("bcoord", C.c_void_p * 4), ("ucoord", C.c_void_p * 4)]


def stripped(nm: bytes) -> bytes:
    parts = nm.split(b" ", 1)
    if len(parts[0]) > 2 and parts[0][-2:] in (b"/1", b"/2"):
        parts[0] = parts[0][:-2]
    return b" ".join(parts)


def fastq_text(names, batch, eol=b"\n"):
    out = bytearray()
    for i, nm in enumerate(names):
        L = int(batch.length[i])
        out += b"@" + nm + eol + batch.seq[i, :L].tobytes() + eol + b"+" + eol + batch.qual[i, :L].tobytes() + eol
    text = np.frombuffer(bytes(out), np.uint8)
    return text, np.flatnonzero(text == 10).astype(np.uint32)


def run_emit(layout, full, names_per_read, correct, paired, eol=b"\n"):
    infos, strides = U.emul_infos(layout)
    _, results, _, _ = U.run_emul(layout, U.narrow(full, strides, infos), correct=correct)
    n = len(full[0].length)
    n_reads = len(layout.reads)
    io = (_EmitIO * n_reads)()
    keep = []
    total = 0
    for r in range(n_reads):
        text, nl = fastq_text(names_per_read[r], full[r], eol)
        assert len(nl) == 4 * n
        ln = np.ascontiguousarray(full[r].length, np.uint16)
        keep += [text, nl, ln]
        total += len(text)
        res = results[r]
        io[r].text, io[r].nl, io[r].len = text.ctypes.data, nl.ctypes.data, ln.ctypes.data
        io[r].fstatus = res.fstatus.ctypes.data
        if res.tcoord is not None:
            io[r].tcoord = res.tcoord.ctypes.data
        for j, k in enumerate(res.bc_key):
            io[r].bc_key[j] = k.ctypes.data
            if res.bcoord[j] is not None:
                io[r].bcoord[j] = res.bcoord[j].ctypes.data
        for j in range(len(res.umi_key)):
            if res.ucoord[j] is not None:
                io[r].ucoord[j] = res.ucoord[j].ctypes.data
    cap = total + 8 * n + 64
    o1, o2 = np.zeros(cap, np.uint8), np.zeros(cap, np.uint8)
    u1, u2, nw, bad = C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_uint64()
    L = U.emul_lib()
    L.emu_emit.restype = C.c_int
    L.emu_emit.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64),
                           C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    rc = L.emu_emit(layout.descs(), n_reads, io, n, int(correct), int(paired), o1.ctypes.data, cap, C.byref(u1),
                    o2.ctypes.data, cap, C.byref(u2), C.byref(nw), C.byref(bad))
    return rc, o1[:u1.value].tobytes(), o2[:u2.value].tobytes(), nw.value, bad.value


def expected(layout, full, names_per_read, correct, paired):
    ores, _ = U.run_oracle(layout, full, correct=correct)
    bc_read = [r for r, p in enumerate(layout.reads) if p.has_barcode]
    r2_read = [r for r, p in enumerate(layout.reads) if p.is_reverse and any(e.is_target() for e in p.elems)]
    o1, o2, n = bytearray(), bytearray(), 0
    for i, ln in enumerate(ores.lines):
        if not ln:
            continue
        f = [x.encode("latin-1") for x in ln.split("\t")]
        if correct and f[2] == b"*":
            continue
        # the definition comes from the first read file that produced a barcode segment for this group
        first = next(r for r in bc_read if (ores.fstatus[r][i] & 3) == 0 and any(
            ((int(ores.fstatus[r][i]) >> (4 + 3 * j)) & 7) != _ffi.BC_ABSENT for j in range(4)))
        bseq = f[2] if correct else f[0]
        o1 += b"@" + stripped(names_per_read[first][i]) + b"\n" + bseq + f[3] + f[5] + b"\n+\n" + f[1] + f[4] + f[6] + b"\n"
        if paired:
            o2 += b"@" + stripped(names_per_read[r2_read[0]][i]) + b"\n" + f[7] + b"\n+\n" + f[8] + b"\n"
        n += 1
    return bytes(o1), bytes(o2), n


CASES = [("atac", True), ("multiome_atac", True), ("share", False), ("dsc", True), ("fast12_tail", False)]


@pytest.mark.parametrize("correct", [True, False])
@pytest.mark.parametrize("name,paired", CASES)
def test_emitted_records_equal_the_oracle(name, paired, correct):
    rng = np.random.default_rng(U.seed_of("emit", name, correct))
    layout = _layouts(rng)[name]
    n = 1500
    _, strides = U.emul_infos(layout)
    # short reads give the whitelists of a multi-barcode layout different totals (the reference asserts), and a defect R1
    # next to a barcode in R2 leaves a group without read1 (the reference panics): keep those layouts well-formed
    strict = len(layout.region_ids) > 1 or name == "fast12_tail"
    full = U.gen_reads(rng, layout, n, [max(s, 64) for s in strides], p_short=0.0 if strict else 0.03, p_tiny=0.0 if strict else 0.01,
                       p_anchor_mut=0.0 if name == "fast12_tail" else 0.15, read_len=[p.max_len for p in layout.reads])
    for b in full:
        b.length[:] = np.maximum(b.length, 1)
    names = [[(b"r%d/%d d:%d" % (i, 1 + (r == len(layout.reads) - 1), i)) if i % 3 else (b"x%d" % i) for i in range(n)]
             for r in range(len(layout.reads))]
    rc, o1, o2, nw, _ = run_emit(layout, full, names, correct, paired)
    assert rc == 0
    w1, w2, wn = expected(layout, full, names, correct, paired)
    assert nw == wn > 300
    assert o1 == w1
    assert o2 == (w2 if paired else b"")
    # CRLF input gives the same records
    rc, c1, c2, _, _ = run_emit(layout, full, names, correct, paired, eol=b"\r\n")
    assert rc == 0 and c1 == o1 and c2 == o2


def test_layout_without_read1_is_an_error_like_the_reference():
    """10x v3 has no pos-strand insert: `record.read1.unwrap()` panics (python/src/lib.rs:160)."""
    rng = np.random.default_rng(3)
    layout = _layouts(rng)["v3"]
    _, strides = U.emul_infos(layout)
    full = U.gen_reads(rng, layout, 200, [max(s, 64) for s in strides], read_len=[p.max_len for p in layout.reads])
    for b in full:
        b.length[:] = np.maximum(b.length, 1)
    names = [[b"r%d" % i for i in range(200)]] * 2
    rc, *_ = run_emit(layout, full, names, True, False)
    assert rc == 6                                               # PCL_EMIT_NO_READ1
