"""Shared helpers of the test-suite: hand-built layouts, synthetic reads, oracle / emulation drivers and
the structured comparison between the product's outputs and the oracle's."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Dict, List, Optional, Sequence

import numpy as np
import zlib

from oracle import pyoracle as O
from precellar_b200 import _ffi
from precellar_b200.layout import CompiledLayout, ReadPlan
from precellar_b200.pipeline import (Groups, HostBatch, ReadResults, assemble, decode_key_py, element_coords, render)
from precellar_b200.seqspec import SegmentInfoElem

HERE = os.path.dirname(os.path.abspath(__file__))
_PRIMERS = {"custom_primer", "truseq_read1", "truseq_read2", "nextera_read1", "nextera_read2", "illumina_p5", "illumina_p7"}


# ------------------------------------------------------------------------------------------------
# layouts
# ------------------------------------------------------------------------------------------------
def E(region_id, region_type, min_len, max_len=None, sequence_type="random", sequence=""):
    return SegmentInfoElem(region_id, region_type, sequence_type, sequence, min_len, min_len if max_len is None else max_len)


def bc(rid, n, m=None): return E(rid, "barcode", n, m, "onlist", "N" * n)
def umi(n, rid="umi"): return E(rid, "umi", n)
def cdna(a, b, rid="cdna"): return E(rid, "cdna", a, b)
def gdna(a, b, rid="gdna"): return E(rid, "gdna", a, b)
def poly(a, b, rid="polya"): return E(rid, "poly_a", a, b)
def linker(s, rid="linker"): return E(rid, "linker", len(s), len(s), "fixed", s)
def var(a, b, rid="var"): return E(rid, "linker", a, b)


def make_layout(reads: Sequence[dict], whitelists: Dict[str, List[bytes]], modality="rna") -> CompiledLayout:
    """reads: dicts with read_id, elems, is_reverse, max_len.  Region slots follow whitelist order."""
    region_ids = list(whitelists.keys())
    slot_of = {r: i for i, r in enumerate(region_ids)}
    plans = []
    for r in reads:
        elems = r["elems"]
        slots = [slot_of[e.region_id] if e.is_barcode() else -1 for e in elems]
        plans.append(ReadPlan(r["read_id"], bool(r.get("is_reverse", False)), int(r.get("max_len", 0)), list(elems), slots, []))
    return CompiledLayout(modality, plans, region_ids, {k: list(v) for k, v in whitelists.items()},
                          {k: True for k in region_ids}, any(e.is_umi() for p in plans for e in p.elems))


def oracle_reads(layout: CompiledLayout) -> List[O.ReadLayout]:
    return O.reads_from_layout(layout)


# ------------------------------------------------------------------------------------------------
# synthetic data
# ------------------------------------------------------------------------------------------------
_COMP = bytes.maketrans(b"ACGTacgt", b"TGCATGCA")


def revcomp(s: bytes) -> bytes:
    return s.translate(_COMP)[::-1]


def gen_whitelist(rng: np.random.Generator, n: int, k, unique=True) -> List[bytes]:
    ks = [k] if isinstance(k, int) else list(k)
    out, seen = [], set()
    while len(out) < n:
        L = ks[int(rng.integers(len(ks)))]
        s = bytes(rng.choice(np.frombuffer(b"ACGT", np.uint8), L))
        if unique and s in seen:
            continue
        seen.add(s)
        out.append(s)
    return out


QUAL_CHARS = np.frombuffer(b"FFFFFFFFFFFF:::,,#!K~5", np.uint8)


def gen_reads(rng: np.random.Generator, layout: CompiledLayout, n: int, stride_of: Sequence[int], *, p_sub=0.03,
              p_n=0.01, p_lower=0.0, p_short=0.02, p_tiny=0.0, p_anchor_mut=0.15, hot=32,
              read_len: Optional[Sequence[int]] = None
              ) -> List[HostBatch]:
    """Random reads that follow each read's element table loosely: whitelist barcodes (Zipf-ish over the
    first `hot` entries plus uniform), substitutions, N, optional lower case, truncated reads, mutated anchors."""
    acgt = np.frombuffer(b"ACGT", np.uint8)
    batches = []
    for r, plan in enumerate(layout.reads):
        stride = int(stride_of[r])
        total_max = sum(e.max_len for e in plan.elems)
        width = max(stride, min(total_max, 400), 1)
        seqs = np.full((n, width), ord("N"), np.uint8)
        quals = np.full((n, width), ord("!"), np.uint8)
        lens = np.zeros(n, np.uint16)
        for i in range(n):
            parts = []
            for e, slot in zip(plan.elems, plan.region_slots):
                L = e.max_len if e.min_len == e.max_len else int(rng.integers(e.min_len, min(e.max_len, e.min_len + 60) + 1))
                if e.is_barcode():
                    wl = layout.whitelists[layout.region_ids[slot]]
                    if rng.random() < 0.7:
                        s = wl[int(rng.integers(min(hot, len(wl))))]
                    elif rng.random() < 0.8:
                        s = wl[int(rng.integers(len(wl)))]
                    else:
                        s = bytes(rng.choice(acgt, L))
                    if not (e.min_len <= len(s) <= e.max_len):
                        s = bytes(rng.choice(acgt, L))
                    if plan.is_reverse:
                        s = revcomp(s)
                elif e.sequence_type == "fixed" and e.sequence:
                    s = e.sequence.encode()
                    if plan.is_reverse:
                        s = revcomp(s)
                    if rng.random() < p_anchor_mut:
                        b = bytearray(s)
                        for _ in range(int(rng.integers(1, 3))):
                            b[int(rng.integers(len(b)))] = int(rng.choice(acgt))
                        s = bytes(b)
                else:
                    s = bytes(rng.choice(acgt, L))
                parts.append(s)
            s = bytearray(b"".join(parts))
            for k in range(len(s)):
                u = rng.random()
                if u < p_sub:
                    s[k] = int(rng.choice(acgt))
                elif u < p_sub + p_n:
                    s[k] = ord("N")
                elif u < p_sub + p_n + p_lower:
                    s[k] = s[k] | 0x20
            L = len(s)
            if read_len is not None:
                L = min(L, int(read_len[r])) if read_len[r] else L
            if rng.random() < p_short:
                L = int(rng.integers(0, L + 1))
            if rng.random() < p_tiny:                                    # too short for any barcode of this read
                L = int(rng.integers(0, 4))
            L = min(L, width, len(s))
            seqs[i, :L] = np.frombuffer(bytes(s[:L]), np.uint8)
            quals[i, :L] = rng.choice(QUAL_CHARS, L)
            lens[i] = L
        batches.append(HostBatch(seqs, quals, lens))                     # full-width planes (oracle / rendering)
    return batches


def narrow(batches: Sequence[HostBatch], strides: Sequence[int], infos=None) -> List[HostBatch]:
    """The planes the device consumes: the first `stride` sequence bytes of every record and the quality
    window [qual_offset, qual_offset + qual_stride) given by pcl_read_info (None where a plane is not used)."""
    if infos is None:
        raise TypeError("narrow() needs the pcl_read_info list (quality windows)")
    out = []
    for b, st, info in zip(batches, strides, infos):
        def cut(a, lo, w, fill):
            if a is None or w == 0:
                return None
            o = np.full((a.shape[0], w), fill, np.uint8)
            hi = min(a.shape[1], lo + w)
            if hi > lo:
                o[:, :hi - lo] = a[:, lo:hi]
            return o
        seq = cut(b.seq, 0, st, ord("N")) if info.needs_payload else None
        qual = cut(b.qual, int(info.qual_offset), int(info.qual_stride), ord("!")) if (info.needs_payload or info.qual_only) else None
        out.append(HostBatch(seq, qual, b.length))
    return out


# ------------------------------------------------------------------------------------------------
# emulation harness (CPU run of the device functions)
# ------------------------------------------------------------------------------------------------
class _EmuIO(C.Structure):
    _fields_ = [("seq", C.c_void_p), ("qual", C.c_void_p), ("len", C.c_void_p), ("fstatus", C.c_void_p),
                ("tcoord", C.c_void_p), ("bc_key", C.c_void_p * 4), ("umi_key", C.c_void_p * 4),
                ("bcoord", C.c_void_p * 4), ("ucoord", C.c_void_p * 4)]


_emul = None


def emul_lib():
    global _emul
    if _emul is None:
        src = os.path.join(HERE, "emul", "emul.cpp")
        # PCL_EMUL_SANITIZE=1: build the host copy of the device code with ASan + UBSan (run pytest with
        # LD_PRELOAD=$(gcc -print-file-name=libasan.so)); compute-sanitizer is not available on the GPU pool, so this is
        # how out-of-bounds accesses in the __host__ __device__ record logic are hunted
        san = os.environ.get("PCL_EMUL_SANITIZE") == "1"
        so = os.path.join(HERE, "emul", "libemul_san.so" if san else "libemul.so")
        deps = [src] + [os.path.join(HERE, "..", "precellar_b200", "csrc", f) for f in ("hd_core.h", "host_build.h", "zhuff_core.h", "emit_core.h", "inflate_core.h", "fastinflate.h")] + \
               [os.path.join(HERE, "..", "include", "precellar_b200.h")]
        if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps if os.path.exists(d)):
            subprocess.check_call(["g++", "-O1" if san else "-O2", "-g", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off"] +
                                  (["-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-fno-omit-frame-pointer"] if san else []) +
                                  ["-o", so, src])
        _emul = C.CDLL(so)
    return _emul


def emul_infos(layout: CompiledLayout):
    descs = layout.descs()
    infos, strides = [], []
    for r in range(len(layout.reads)):
        info = _ffi.ReadInfo()
        stride = C.c_uint32()
        err = C.create_string_buffer(512)
        rc = emul_lib().emu_read_info(C.byref(descs[r]), C.byref(info), C.byref(stride), err, 512)
        if rc != 0:
            raise _ffi.PclError(rc, err.value.decode())
        infos.append(info)
        strides.append(stride.value)
    return infos, strides


def alloc_results(info: _ffi.ReadInfo, n: int) -> ReadResults:
    kd = lambda b: np.uint32 if b == 4 else np.uint64
    return ReadResults(
        info=info, fstatus=np.zeros(n, np.uint16), tcoord=np.zeros(n, np.uint32) if info.has_target else None,
        bc_key=[np.zeros(n, kd(info.bc_key_bytes[j])) for j in range(info.n_bc)],
        umi_key=[np.zeros(n, kd(info.umi_key_bytes[j])) for j in range(info.n_umi)],
        bcoord=[None if info.fixed_layout else np.zeros(n, np.uint32) for _ in range(info.n_bc)],
        ucoord=[None if info.fixed_layout else np.zeros(n, np.uint32) for _ in range(info.n_umi)])


def run_emul(layout: CompiledLayout, batches: Sequence[HostBatch], correct=True, threshold=0.975,
             max_expected_errors=_ffi.F64_MAX, use_fast=True, global_counts=None):
    """Returns (infos, results, counts{slot: (counts, mismatch, total)}, defects)."""
    L = emul_lib()
    infos, strides = emul_infos(layout)
    n = len(batches[0].length)
    n_reads = len(layout.reads)
    descs = layout.descs()
    io = (_EmuIO * n_reads)()
    results, keep = [], []
    for r in range(n_reads):
        res = alloc_results(infos[r], n)
        results.append(res)
        b = batches[r]
        ln = np.ascontiguousarray(b.length, np.uint16)
        keep.append(ln)
        if infos[r].needs_payload:
            assert b.seq is not None and b.seq.shape[1] == strides[r], "pass narrow(batches, strides, infos)"
            io[r].seq = b.seq.ctypes.data
            if b.qual is not None:
                assert b.qual.shape[1] == infos[r].qual_stride
                io[r].qual = b.qual.ctypes.data
        io[r].len = ln.ctypes.data
        io[r].fstatus = res.fstatus.ctypes.data
        if res.tcoord is not None:
            io[r].tcoord = res.tcoord.ctypes.data
        for j in range(infos[r].n_bc):
            io[r].bc_key[j] = res.bc_key[j].ctypes.data
            if res.bcoord[j] is not None:
                io[r].bcoord[j] = res.bcoord[j].ctypes.data
        for j in range(infos[r].n_umi):
            io[r].umi_key[j] = res.umi_key[j].ctypes.data
            if res.ucoord[j] is not None:
                io[r].ucoord[j] = res.ucoord[j].ctypes.data
    n_regions = len(layout.region_ids)
    blobs, offs, wl_n, cnts = [], [], (C.c_uint64 * max(n_regions, 1))(), []
    bp, op, cp = (C.c_void_p * max(n_regions, 1))(), (C.c_void_p * max(n_regions, 1))(), (C.c_void_p * max(n_regions, 1))()
    for g, rid in enumerate(layout.region_ids):
        blob, off = O.pack_strings(layout.whitelists[rid])
        blobs.append(blob); offs.append(off)
        bp[g], op[g] = blob.ctypes.data, off.ctypes.data
        wl_n[g] = len(layout.whitelists[rid])
        c = np.zeros(max(len(dict.fromkeys(layout.whitelists[rid])), 1), np.uint64)
        cnts.append(c)
        cp[g] = c.ctypes.data
    scal = np.zeros(2 * max(n_regions, 1), np.uint64)
    defects = np.zeros(n_reads, np.uint64)
    gin = None
    if global_counts is not None:                    # {slot: uint64 counts in whitelist order} from an allreduce
        gin = (C.c_void_p * max(n_regions, 1))()
        gkeep = [np.ascontiguousarray(global_counts[g], np.uint64) for g in range(n_regions)]
        for g in range(n_regions):
            gin[g] = gkeep[g].ctypes.data
    err = C.create_string_buffer(512)
    L.emu_run.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64,
                          C.c_int, C.c_double, C.c_double, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_char_p, C.c_int]
    rc = L.emu_run(descs, n_reads, n_regions, bp, op, wl_n, io, n, int(correct), threshold, max_expected_errors,
                   int(use_fast), cp, scal.ctypes.data, defects.ctypes.data, gin, err, 512)
    if rc != 0:
        raise _ffi.PclError(rc, err.value.decode())
    counts = {}
    for g, rid in enumerate(layout.region_ids):
        m = len(dict.fromkeys(layout.whitelists[rid]))
        counts[g] = (cnts[g][:m], int(scal[2 * g + 1]), int(scal[2 * g]))
    return infos, results, counts, defects


# ------------------------------------------------------------------------------------------------
# oracle driver + comparison
# ------------------------------------------------------------------------------------------------
def run_oracle(layout: CompiledLayout, batches: Sequence[HostBatch], correct=True,
               threshold=0.975, max_expected_errors=_ffi.F64_MAX, render_lines=True, n_threads=1):
    orc = O.Oracle(oracle_reads(layout), {g: layout.whitelists[rid] for g, rid in enumerate(layout.region_ids)})
    for r, b in enumerate(batches):
        if b.seq is not None:
            orc.set_input(r, b.seq, b.qual, b.length, b.seq.shape[1])
        else:
            orc.set_input(r, None, None, b.length, 0)
    orc.count()
    res = orc.annotate(correct=correct, threshold=threshold, max_expected_errors=max_expected_errors,
                       n_threads=n_threads, render=render_lines)
    counts = {g: orc.counts(g) for g in range(len(layout.region_ids))}
    return res, counts


_CODE_TO_ERR = {_ffi.BC_EXACT: O.BC_OK, _ffi.BC_CORRECTED: O.BC_OK, _ffi.BC_NO_MATCH: O.BC_NO_MATCH,
                _ffi.BC_LOW_CONF: O.BC_LOW_CONF, _ffi.BC_EXCEED_EE: O.BC_EXCEED_EE, _ffi.BC_ABSENT: O.BC_ABSENT}


def compare_with_oracle(layout: CompiledLayout, infos, batches, results: Sequence[ReadResults], counts, defects,
                        ores: O.OracleResult, ocounts, correct=True, check_lines=True):
    """Bit-exact comparison of every observable output.  Raises AssertionError with a description."""
    n = len(batches[0].length)
    g = assemble(layout, infos, batches, results, correct=correct)
    slot = 0
    j_global = 0
    for r, (plan, info, res) in enumerate(zip(layout.reads, infos, results)):
        st = (res.fstatus & 3).astype(np.uint8)
        np.testing.assert_array_equal(st, ores.fstatus[r], err_msg=f"split status of read {plan.read_id}")
        ok = st == 0
        if info.has_target:
            tc = np.where(ok, res.tcoord, np.uint32(_ffi.COORD_ABSENT))
            np.testing.assert_array_equal(tc, ores.tcoord[r], err_msg=f"target coords of read {plan.read_id}")
            tp = (res.fstatus & _ffi.FSTATUS_TARGET) != 0
            np.testing.assert_array_equal(tp, ores.tcoord[r] != _ffi.COORD_ABSENT)
        else:
            assert np.all(ores.tcoord[r] == _ffi.COORD_ABSENT)
    # emit slots: barcode / UMI elements of all reads, file-major, element order
    bc_by_key = {(b["read"], b["elem"]): b for b in g.bc}
    umi_by_key = {(u["read"], u["elem"]): u for u in g.umi}
    for r, plan in enumerate(layout.reads):
        for e, el in enumerate(plan.elems):
            if el.is_barcode():
                b = bc_by_key[(r, e)]
                np.testing.assert_array_equal(b["coord"], ores.bucoord[slot], err_msg=f"barcode coords {plan.read_id}:{el.region_id}")
                code = b["code"]
                if correct:
                    err = np.vectorize(_CODE_TO_ERR.get, otypes=[np.uint8])(code)
                else:
                    err = np.where(code == _ffi.BC_ABSENT, O.BC_ABSENT, O.BC_OK).astype(np.uint8)
                np.testing.assert_array_equal(err, ores.bc_err[j_global], err_msg=f"barcode status {plan.read_id}:{el.region_id}")
                wl = list(dict.fromkeys(layout.whitelists[layout.region_ids[b["region"]]]))
                index_of = {s: k for k, s in enumerate(wl)}
                sel = (code == _ffi.BC_EXACT) | (code == _ffi.BC_CORRECTED) if correct else (code == _ffi.BC_EXACT)
                idx = np.full(n, -1, np.int64)
                keys = b["key"][sel]
                uniq, inv = np.unique(keys, return_inverse=True)
                uidx = np.array([index_of.get(decode_key_py(int(k)), -999) for k in uniq], np.int64) if len(uniq) else np.zeros(0, np.int64)
                idx[sel] = uidx[inv] if len(uniq) else []
                np.testing.assert_array_equal(idx, ores.bc_index[j_global], err_msg=f"corrected barcode {plan.read_id}:{el.region_id}")
                j_global += 1
                slot += 1
            elif el.is_umi():
                u = umi_by_key[(r, e)]
                np.testing.assert_array_equal(u["coord"], ores.bucoord[slot], err_msg=f"UMI coords {plan.read_id}:{el.region_id}")
                slot += 1
    # group flags
    flags = np.zeros(n, np.uint8)
    any_ok = np.any(~g.defect, axis=0)
    flags |= g.has_barcode.astype(np.uint8)
    flags |= (g.corrected_some.astype(np.uint8) << 1)
    has_umi = np.zeros(n, bool)
    for u in g.umi:
        has_umi |= u["present"]
    flags |= has_umi.astype(np.uint8) << 2
    flags |= (g.read1["owner"] >= 0).astype(np.uint8) << 3
    flags |= (g.read2["owner"] >= 0).astype(np.uint8) << 4
    flags = np.where(any_ok, flags, 0).astype(np.uint8)
    np.testing.assert_array_equal(flags, ores.g_flags, err_msg="group flags")
    # counts
    for gi in range(len(layout.region_ids)):
        c, mm, tot = counts[gi]
        oc, omm, otot = ocounts[gi]
        np.testing.assert_array_equal(np.asarray(c, np.uint64), oc, err_msg=f"whitelist counts region {gi}")
        assert (int(mm), int(tot)) == (omm, otot), f"mismatch/total of region {gi}: {(mm, tot)} vs {(omm, otot)}"
    np.testing.assert_array_equal(np.asarray(defects, np.uint64), ores.qc_per_file[:, 1], err_msg="num_defect")
    if check_lines and ores.lines is not None:
        for i in range(n):
            mine = render(layout, batches, g, i, correct=correct)
            assert mine == ores.lines[i], f"record {i}:\n  got      {mine!r}\n  expected {ores.lines[i]!r}"
    return g


def seed_of(*parts):
    """Stable per-case RNG seed (str hashes differ between interpreter runs)."""
    return zlib.crc32("/".join(str(p) for p in parts).encode()) & 0x7FFFFFFF
