"""ctypes front-end of the CPU parity oracle (oracle/liboracle.so).

TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and the cpu_baseline /
``--impl reference`` legs of bench.py.  The product package (precellar_b200/) never imports it.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

OTHER, BARCODE, UMI, TARGET, PRIMER = 0, 1, 2, 3, 4
SPLIT_OK, SPLIT_ANCHOR_NOT_FOUND, SPLIT_PATTERN_MISMATCH = 0, 1, 2
BC_OK, BC_NO_MATCH, BC_LOW_CONF, BC_EXCEED_EE, BC_ABSENT = 0, 1, 2, 3, 255
F64_MAX = 1.7976931348623157e308


def build(force: bool = False) -> str:
    """Compile oracle/liboracle.so with the committed Makefile (gcc only)."""
    src = os.path.join(_HERE, "oracle.cpp")
    hdr = os.path.join(_HERE, "oracle.h")
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.getmtime(p) > os.path.getmtime(_LIB_PATH) for p in (src, hdr))
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-B" if force else "-s", "liboracle.so"])
    return _LIB_PATH


class _Elem(C.Structure):
    _fields_ = [("rtype", C.c_uint8), ("seq_fixed", C.c_uint8), ("pad", C.c_uint8 * 2),
                ("min_len", C.c_uint32), ("max_len", C.c_uint32), ("region", C.c_int32),
                ("sequence", C.c_char_p)]


class _Out(C.Structure):
    _fields_ = [("fstatus", C.c_void_p), ("tcoord", C.c_void_p), ("bucoord", C.c_void_p),
                ("bc_index", C.c_void_p), ("bc_err", C.c_void_p), ("g_flags", C.c_void_p)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.orc_create.restype = C.c_void_p
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_last_error.restype = C.c_char_p
        L.orc_last_error.argtypes = [C.c_void_p]
        L.orc_add_read.argtypes = [C.c_void_p, C.c_int, C.c_uint32, C.POINTER(_Elem), C.c_int]
        L.orc_set_whitelist.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_uint64]
        L.orc_set_input.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint64]
        L.orc_set_keep_records.argtypes = [C.c_void_p, C.c_int]
        L.orc_count.argtypes = [C.c_void_p]
        L.orc_get_counts.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_annotate.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_int, C.c_uint64, C.POINTER(_Out)]
        L.orc_get_qc.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_render.restype = C.c_int64
        L.orc_render.argtypes = [C.c_void_p, C.c_uint64, C.c_char_p, C.c_uint64]
        L.orc_fastq_load.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_char_p), C.c_int, C.c_uint32]
        L.orc_n_records.restype = C.c_uint64
        L.orc_n_records.argtypes = [C.c_void_p, C.c_int]
        L.orc_make_fastq.restype = C.c_int64
        L.orc_make_fastq.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.POINTER(C.c_int)]
        L.orc_split_str.argtypes = [C.POINTER(_Elem), C.c_int, C.c_int, C.c_char_p, C.c_char_p, C.c_uint32,
                                    C.c_double, C.c_double, C.c_char_p, C.c_uint32]
        L.orc_truncate_max.argtypes = [C.POINTER(_Elem), C.c_int, C.c_uint32]
        L.orc_rev_compl_seqspec.argtypes = [C.c_char_p, C.c_uint32, C.c_char_p]
        L.orc_rev_compl_precellar.argtypes = [C.c_char_p, C.c_uint32, C.c_char_p]
        L.orc_error_probability.restype = C.c_double
        L.orc_error_probability.argtypes = [C.c_uint8]
        L.orc_correct_one.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_char_p, C.c_char_p,
                                      C.c_uint32, C.c_double, C.c_double, C.POINTER(C.c_int64), C.POINTER(C.c_double)]
        _lib = L
    return _lib


@dataclass
class Elem:
    """SegmentInfoElem (seqspec/src/read/segment.rs:371-378), region_type reduced to its class."""
    rtype: int
    min_len: int
    max_len: int
    seq_fixed: bool = False
    sequence: str = ""
    region: int = -1
    region_id: str = ""


@dataclass
class ReadLayout:
    """One sequenced read: its SegmentInfo after truncate_max, strand and reader truncation."""
    elems: List[Elem]
    is_reverse: bool = False
    max_len: int = 0
    read_id: str = ""


_PRIMER_TYPES = {"custom_primer", "truseq_read1", "truseq_read2", "nextera_read1", "nextera_read2", "illumina_p5",
                 "illumina_p7"}


def reads_from_layout(layout) -> List["ReadLayout"]:
    """Oracle read tables from a compiled layout (duck-typed: .reads[].elems[] with seqspec field names)."""
    out = []
    for p in layout.reads:
        elems = []
        for e, slot in zip(p.elems, p.region_slots):
            rt = (BARCODE if e.region_type == "barcode" else UMI if e.region_type == "umi"
                  else TARGET if e.region_type in ("gdna", "cdna") else PRIMER if e.region_type in _PRIMER_TYPES else OTHER)
            elems.append(Elem(rt, e.min_len, e.max_len, e.sequence_type == "fixed", e.sequence, slot, e.region_id))
        out.append(ReadLayout(elems, p.is_reverse, p.max_len, p.read_id))
    return out


def _elem_array(elems: Sequence) -> "C.Array[_Elem]":
    arr = (_Elem * max(1, len(elems)))()
    for i, e in enumerate(elems):
        arr[i].rtype = int(e.rtype)
        arr[i].seq_fixed = 1 if e.seq_fixed else 0
        arr[i].min_len = int(e.min_len)
        arr[i].max_len = int(e.max_len)
        arr[i].region = int(e.region)
        arr[i].sequence = (e.sequence or "").encode()
    return arr


def pack_strings(items: Sequence[bytes]):
    """list of byte strings -> (blob uint8[], offsets uint64[n+1])"""
    offs = np.zeros(len(items) + 1, dtype=np.uint64)
    if len(items):
        offs[1:] = np.cumsum([len(x) for x in items], dtype=np.uint64)
    blob = np.frombuffer(b"".join(items), dtype=np.uint8).copy() if len(items) else np.zeros(0, np.uint8)
    return blob, offs


def split_str(elems: Sequence[Elem], seq: bytes, is_reverse: bool = False, linker_tol: float = 1.0,
              anchor_tol: float = 0.2):
    """SegmentInfo::split rendered like the reference's unit tests; returns (status, text)."""
    buf = C.create_string_buffer(4096)
    qual = b"@" * len(seq)
    st = lib().orc_split_str(_elem_array(elems), len(elems), int(is_reverse), seq, qual, len(seq),
                             linker_tol, anchor_tol, buf, 4096)
    return st, buf.value.decode()


def truncate_max(elems: Sequence[Elem], length: int) -> int:
    return lib().orc_truncate_max(_elem_array(elems), len(elems), length)


def rev_compl_seqspec(s: bytes) -> bytes:
    out = C.create_string_buffer(len(s) + 1)
    lib().orc_rev_compl_seqspec(s, len(s), out)
    return out.raw[:len(s)]


def rev_compl_precellar(s: bytes) -> bytes:
    out = C.create_string_buffer(len(s) + 1)
    lib().orc_rev_compl_precellar(s, len(s), out)
    return out.raw[:len(s)]


def error_probability(q: int) -> float:
    return lib().orc_error_probability(q)


def correct_one(whitelist: Sequence[bytes], counts: Optional[Sequence[int]], bc: bytes, qual: bytes,
                threshold: float = 0.975, max_expected_errors: float = F64_MAX):
    """correct_barcode on an ad-hoc table; returns (BC_* code, whitelist index, likelihood1 posterior)."""
    blob, offs = pack_strings(list(whitelist))
    cnt = np.asarray(counts if counts is not None else [0] * len(whitelist), dtype=np.uint64)
    idx = C.c_int64(-1)
    prob = C.c_double(0.0)
    code = lib().orc_correct_one(blob.ctypes.data, offs.ctypes.data, cnt.ctypes.data, len(whitelist), bc, qual,
                                 len(bc), threshold, max_expected_errors, C.byref(idx), C.byref(prob))
    return code, idx.value, prob.value


@dataclass
class OracleResult:
    n: int
    fstatus: np.ndarray      # [n_files, n] uint8
    tcoord: np.ndarray       # [n_files, n] uint32
    bucoord: np.ndarray      # [n_slots, n] uint32
    bc_index: np.ndarray     # [n_bcsegs, n] int64
    bc_err: np.ndarray       # [n_bcsegs, n] uint8
    g_flags: np.ndarray      # [n] uint8
    qc_per_file: np.ndarray  # [n_files, 2] uint64 (num_reads, num_defect)
    qc_cat: np.ndarray       # [4, 3] uint64 rows umi, barcode, read1, read2; cols num_reads,total,q30
    lines: Optional[List[str]] = field(default=None)


class Oracle:
    """Drives orc_* for one (assay, modality): layouts + whitelists + SoA inputs."""

    def __init__(self, reads: Sequence, whitelists: dict):
        """reads: objects with .elems/.is_reverse/.max_len ; whitelists: region slot -> list[bytes]"""
        self.L = lib()
        self.h = C.c_void_p(self.L.orc_create())
        self.reads = list(reads)
        self._keep = []
        for r in self.reads:
            arr = _elem_array(r.elems)
            self._keep.append(arr)
            rc = self.L.orc_add_read(self.h, int(r.is_reverse), int(r.max_len), arr, len(r.elems))
            if rc < 0:
                raise RuntimeError(self.error())
        self.n_files = len(self.reads)
        self.n_slots = sum(1 for r in self.reads for e in r.elems if e.rtype in (BARCODE, UMI))
        self.n_bcsegs = sum(1 for r in self.reads for e in r.elems if e.rtype == BARCODE)
        self.wl_sizes = {}
        for region, items in whitelists.items():
            if isinstance(items, tuple):            # (blob uint8[], offsets uint64[n+1], n_unique): large lists
                blob, offs, n_unique = items
                blob = np.ascontiguousarray(blob, np.uint8)
                offs = np.ascontiguousarray(offs, np.uint64)
                n_items = len(offs) - 1
            else:
                items = list(items)
                blob, offs = pack_strings(items)
                n_items, n_unique = len(items), len(dict.fromkeys(items))
            rc = self.L.orc_set_whitelist(self.h, int(region), blob.ctypes.data, offs.ctypes.data, n_items)
            if rc < 0:
                raise RuntimeError(self.error())
            self.wl_sizes[int(region)] = n_unique
        self.n = None

    def error(self) -> str:
        return self.L.orc_last_error(self.h).decode()

    def close(self):
        if self.h:
            self.L.orc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_input(self, file: int, seq, qual, length: np.ndarray, stride: int):
        length = np.ascontiguousarray(length, dtype=np.uint16)
        self._keep.append((seq, qual, length))
        n = len(length)
        self.n = n
        sp = seq.ctypes.data if seq is not None else None
        qp = qual.ctypes.data if qual is not None else None
        rc = self.L.orc_set_input(self.h, file, sp, qp, length.ctypes.data, stride, n)
        if rc < 0:
            raise RuntimeError(self.error())

    def fastq_load(self, file: int, paths: Sequence[str], stride: int) -> int:
        """Read the concatenated FASTQ files (plain / gzip) of one read into planes owned by the oracle."""
        arr = (C.c_char_p * len(paths))(*[p.encode() for p in paths])
        rc = self.L.orc_fastq_load(self.h, file, arr, len(paths), stride)
        if rc < 0:
            raise RuntimeError(self.error())
        self.n = int(self.L.orc_n_records(self.h, file))
        return self.n

    def make_fastq(self, correct: bool, r1_path: str, r2_path: Optional[str], level: int = 9, threads: int = 8):
        """make_fastq's record loop over the records kept by annotate(render=True); returns (records, zstd workers)."""
        w = C.c_int(0)
        n = self.L.orc_make_fastq(self.h, int(correct), r1_path.encode(), r2_path.encode() if r2_path else None, level, threads, C.byref(w))
        if n < 0:
            raise RuntimeError(self.error())
        return int(n), w.value

    def count(self):
        rc = self.L.orc_count(self.h)
        if rc < 0:
            raise RuntimeError(self.error())

    def counts(self, region: int):
        n = self.wl_sizes.get(region, 0)
        cnt = np.zeros(max(n, 1), dtype=np.uint64)
        mm = C.c_uint64(0)
        tot = C.c_uint64(0)
        self.L.orc_get_counts(self.h, region, cnt.ctypes.data, C.byref(mm), C.byref(tot))
        return cnt[:n], mm.value, tot.value

    def annotate(self, correct: bool = True, threshold: float = 0.975, max_expected_errors: float = F64_MAX,
                 n_threads: int = 1, chunk_bases: int = 1_000_000, render: bool = False,
                 outputs: bool = True, keep: bool = False) -> OracleResult:
        n = self.n
        self.L.orc_set_keep_records(self.h, int(render or keep))
        res = OracleResult(
            n=n,
            fstatus=np.zeros((self.n_files, n), np.uint8), tcoord=np.zeros((self.n_files, n), np.uint32),
            bucoord=np.zeros((max(self.n_slots, 1), n), np.uint32),
            bc_index=np.zeros((max(self.n_bcsegs, 1), n), np.int64),
            bc_err=np.zeros((max(self.n_bcsegs, 1), n), np.uint8), g_flags=np.zeros(n, np.uint8),
            qc_per_file=np.zeros((self.n_files, 2), np.uint64), qc_cat=np.zeros((4, 3), np.uint64))
        out = _Out(res.fstatus.ctypes.data, res.tcoord.ctypes.data, res.bucoord.ctypes.data,
                   res.bc_index.ctypes.data, res.bc_err.ctypes.data, res.g_flags.ctypes.data)
        rc = self.L.orc_annotate(self.h, int(correct), threshold, max_expected_errors, n_threads, chunk_bases,
                                 C.byref(out) if outputs else None)
        if rc < 0:
            raise RuntimeError(self.error())
        self.L.orc_get_qc(self.h, res.qc_per_file.ctypes.data, res.qc_cat.ctypes.data)
        res.bucoord = res.bucoord[:self.n_slots]
        res.bc_index = res.bc_index[:self.n_bcsegs]
        res.bc_err = res.bc_err[:self.n_bcsegs]
        if render:
            buf = C.create_string_buffer(1 << 16)
            lines = []
            for i in range(n):
                k = self.L.orc_render(self.h, i, buf, 1 << 16)
                if k < 0:
                    raise RuntimeError("render failed")
                lines.append(buf.raw[:k].decode("latin-1"))
            res.lines = lines
        return res
