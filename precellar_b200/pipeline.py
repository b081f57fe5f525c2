"""Driver of the device hot path over the C ABI, and host-side assembly of the per-read-group records.

``BarcodePipeline`` plays the role of the reference's ``BarcodeAnalyzer`` + ``AnnotatedFastqReader``
(precellar/src/barcode.rs:51-185, precellar/src/align/fastq.rs:278-389) with the per-read work done by the
CUDA library:  count (pass 1) -> allreduce -> correct (pass 2) -> fetch.
``assemble`` rebuilds what ``AnnotatedFastq`` exposes (barcode raw/corrected, umi, read1, read2;
fastq.rs:528-604) from the device outputs and the caller's host batch.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import _ffi
from ._ffi import check, lib
from .layout import CompiledLayout


@dataclass
class HostBatch:
    """SoA planes of one read file: record i is seq[i, :length[i]] (row stride = pcl_read_stride)."""
    seq: Optional[np.ndarray]      # [n, stride] uint8 or None when the read needs lengths only
    qual: Optional[np.ndarray]
    length: np.ndarray             # [n] uint16, untruncated record lengths


@dataclass
class ReadResults:
    info: _ffi.ReadInfo
    fstatus: np.ndarray
    tcoord: Optional[np.ndarray]
    bc_key: List[np.ndarray]
    umi_key: List[np.ndarray]
    bcoord: List[Optional[np.ndarray]]
    ucoord: List[Optional[np.ndarray]]


@dataclass
class GroupResult:
    """Per-barcode grouping of one chunk (pcl_group_*).  Fragment f = records order[frag_start[f]:frag_start[f+1]] (a
    duplicate set; the first one is kept, the length is Fragment.count); barcode b = fragments
    bc_frag_start[b]:bc_frag_start[b+1], barcodes in byte-string order."""
    n_records: int
    n_assigned: int
    n_ungroupable: int
    order: np.ndarray
    frag_start: np.ndarray
    bc_frag_start: np.ndarray
    bc_key: np.ndarray

    def barcode(self, b: int) -> bytes:
        buf = C.create_string_buffer(40)
        n = lib().pcl_group_key_decode(int(self.bc_key[b]), buf)
        if n < 0:
            raise ValueError("not a group key")
        return buf.raw[:n]


class Chunk:
    def __init__(self, pipe: "BarcodePipeline", capacity: int):
        self.pipe = pipe
        self.capacity = int(capacity)
        h = C.c_void_p()
        check(pipe.ctx, lib().pcl_chunk_create(pipe.ctx, self.capacity, C.byref(h)))
        self.h = h
        self.n = 0
        self._keep = []

    def reset(self, n: int) -> None:
        check(self.pipe.ctx, lib().pcl_chunk_reset(self.h, int(n)))
        self.n = int(n)
        self._keep = []

    def upload(self, read_slot: int, batch: HostBatch) -> None:
        """Asynchronous H2D of one read file's planes (pinned memory recommended)."""
        stride = self.pipe.stride(read_slot)
        length = np.ascontiguousarray(batch.length, dtype=np.uint16)
        assert len(length) == self.n, "all read files of a chunk must hold the same number of records"
        sp = qp = None
        info = self.pipe.infos[read_slot]
        if info.needs_payload:
            assert batch.seq is not None, "this read needs its sequence plane"
            assert batch.seq.dtype == np.uint8 and batch.seq.shape == (self.n, stride) and batch.seq.flags.c_contiguous
            sp = batch.seq.ctypes.data
        if (info.needs_payload or info.qual_only) and info.qual_stride:
            # the quality plane covers record bytes [qual_offset, qual_offset + qual_stride) only (the barcode bases)
            assert batch.qual is not None, "this read needs its quality plane"
            assert batch.qual.dtype == np.uint8 and batch.qual.shape == (self.n, info.qual_stride) and batch.qual.flags.c_contiguous, \
                f"quality plane must be [n, {info.qual_stride}] (record bytes from {info.qual_offset})"
            qp = batch.qual.ctypes.data
        self._keep.append((batch, length))
        check(self.pipe.ctx, lib().pcl_chunk_upload(self.h, read_slot, sp, qp, length.ctypes.data))

    def set_uniform_len(self, read_slot: int, length: int) -> None:
        """All records of this read file are `length` long: no length plane crosses PCIe (``upload`` may then get a
        batch whose ``length`` is None)."""
        check(self.pipe.ctx, lib().pcl_chunk_set_uniform_len(self.h, read_slot, int(length)))

    def uniform_results_async(self, read_slot: int, out: np.ndarray) -> None:
        """Stream-ordered: out (uint32[4], pinned) = [uniform?, fstatus, tcoord, 0] of an insert read's result planes."""
        assert out.dtype == np.uint32 and len(out) >= 4
        check(self.pipe.ctx, lib().pcl_chunk_uniform_results_async(self.h, read_slot, out.ctypes.data))

    def upload_compact(self, read_slot: int, batch: HostBatch) -> None:
        """``upload`` with sequence rows narrower than the device stride (any multiple of 4 bytes from
        ``BarcodePipeline.payload_bytes`` up): fewer bytes over PCIe, widened on the device."""
        length = None if batch.length is None else np.ascontiguousarray(batch.length, dtype=np.uint16)
        assert length is None or len(length) == self.n
        info = self.pipe.infos[read_slot]
        if not info.needs_payload:
            return self.upload(read_slot, batch)
        assert batch.seq is not None and batch.seq.dtype == np.uint8 and batch.seq.flags.c_contiguous and batch.seq.shape[0] == self.n
        qp = None
        if info.qual_stride:
            assert batch.qual is not None and batch.qual.shape == (self.n, info.qual_stride) and batch.qual.flags.c_contiguous
            qp = batch.qual.ctypes.data
        self._keep.append((batch, length))
        check(self.pipe.ctx, lib().pcl_chunk_upload_compact(self.h, read_slot, batch.seq.ctypes.data, batch.seq.shape[1], qp,
                                                            length.ctypes.data if length is not None else None))

    # -- device-side FASTQ parse (pcl_text_*) ----------------------------------------------------------
    def text_scan(self, read_slot: int, text: np.ndarray, n_bytes: Optional[int] = None) -> None:
        """Copy a block of uncompressed FASTQ text (uint8, starting at a record boundary) to the device and index
        its lines.  Asynchronous: ``text`` must stay untouched until ``text_lines`` returns."""
        if hasattr(text, "ptr"):                     # textingest.DevBlock: the text is in device memory already
            n = len(text) if n_bytes is None else int(n_bytes)
            check(self.pipe.ctx, lib().pcl_text_scan_device(self.h, read_slot, text.ptr if n else None, n))
            return
        assert text.dtype == np.uint8 and text.flags.c_contiguous
        n = len(text) if n_bytes is None else int(n_bytes)
        check(self.pipe.ctx, lib().pcl_text_scan(self.h, read_slot, text.ctypes.data if n else None, n))

    def text_lines(self, read_slot: int) -> int:
        v = C.c_uint64()
        check(self.pipe.ctx, lib().pcl_text_lines(self.h, read_slot, C.byref(v)))
        return v.value

    def text_fill(self, read_slot: int) -> int:
        """Fill the read's planes with the first ``n`` records of the scanned text; returns the bytes consumed."""
        v = C.c_uint64()
        check(self.pipe.ctx, lib().pcl_text_fill(self.h, read_slot, C.byref(v)))
        return v.value

    def text_keep(self, read_slot: int) -> None:
        """Keep this chunk's own device copy of the text consumed by ``text_fill`` (for ``emit_fastq``)."""
        check(self.pipe.ctx, lib().pcl_text_keep(self.h, read_slot))

    def emit_fastq(self, correct_barcode: bool, paired: bool, compress: bool):
        """The chunk's make_fastq records, assembled (and with ``compress`` entropy-coded into one zstd frame per file)
        on the device.  Returns (r1, r2, info): uint8 views of pinned memory that stay valid until the second next call
        on this pipeline (None when nothing was written) and the byte / record counts."""
        res = _ffi.EmitResult()
        check(self.pipe.ctx, lib().pcl_emit_fastq(self.h, int(correct_barcode), int(paired), int(compress), C.byref(res)))
        def view(p, n):
            if not p or not n:
                return None
            a = np.frombuffer((C.c_uint8 * int(n)).from_address(p), dtype=np.uint8)
            a.flags.writeable = False            # the bytes belong to the library (and are recycled two calls later)
            return a
        return view(res.r1, res.r1_bytes), view(res.r2, res.r2_bytes), {
            "n_written": int(res.n_written), "r1_raw_bytes": int(res.r1_raw_bytes), "r2_raw_bytes": int(res.r2_raw_bytes)}

    def text_names_check(self) -> int:
        v = C.c_int64()
        check(self.pipe.ctx, lib().pcl_text_names_check(self.h, C.byref(v)))
        return v.value

    def text_line_ends(self, read_slot: int, n_lines: int) -> np.ndarray:
        out = np.empty(int(n_lines), np.uint32)
        check(self.pipe.ctx, lib().pcl_text_line_ends(self.h, read_slot, out.ctypes.data, int(n_lines)))
        return out

    def device_ptrs(self, read_slot: int):
        s, q, l = C.c_void_p(), C.c_void_p(), C.c_void_p()
        check(self.pipe.ctx, lib().pcl_chunk_device_ptrs(self.h, read_slot, C.byref(s), C.byref(q), C.byref(l)))
        return s.value, q.value, l.value

    @property
    def stream(self) -> int:
        return lib().pcl_chunk_stream(self.h)

    def timer_start(self) -> None:
        check(self.pipe.ctx, lib().pcl_timer_start(self.h))

    def timer_stop_ms(self) -> float:
        ms = C.c_double()
        check(self.pipe.ctx, lib().pcl_timer_stop_ms(self.h, C.byref(ms)))
        return ms.value

    def fetch(self, read_slot: int, out: Optional[ReadResults] = None, *, wait: bool = True,
              planes: Optional[Sequence[str]] = None) -> ReadResults:
        """D2H of one read file's results into (optionally caller-provided) host arrays.
        wait=False only enqueues the copies (``BarcodePipeline.sync`` completes them).  `planes` restricts the copy
        to a subset of {"fstatus", "tcoord", "bc_key", "umi_key", "coords"}: everything except the barcode codes in
        fstatus and bc_key is final right after ``count``."""
        info = self.pipe.infos[read_slot]
        n = self.n
        if out is None:
            out = self.pipe.alloc_results(read_slot, n)
        want = (lambda k: True) if planes is None else (lambda k: k in planes)
        rr = _ffi.ReadResults()
        if want("fstatus"):
            rr.fstatus = out.fstatus.ctypes.data
        if info.has_target and want("tcoord"):
            rr.tcoord = out.tcoord.ctypes.data
        for j in range(info.n_bc):
            if want("bc_key"):
                rr.bc_key[j] = out.bc_key[j].ctypes.data
            if not info.fixed_layout and want("coords"):
                rr.bcoord[j] = out.bcoord[j].ctypes.data
        for j in range(info.n_umi):
            if want("umi_key"):
                rr.umi_key[j] = out.umi_key[j].ctypes.data
            if not info.fixed_layout and want("coords"):
                rr.ucoord[j] = out.ucoord[j].ctypes.data
        fn = lib().pcl_chunk_fetch if wait else lib().pcl_chunk_fetch_async
        check(self.pipe.ctx, fn(self.h, read_slot, C.byref(rr)))
        return out

    def close(self) -> None:
        if self.h:
            lib().pcl_chunk_destroy(self.h)
            self.h = None


class BarcodePipeline:
    """One GPU's instance of the hot path for one (assay, modality)."""

    def __init__(self, layout: CompiledLayout, device: int = 0, nranks: int = 1, rank: int = 0,
                 nccl_uid: Optional[bytes] = None, qc: bool = False, verify: bool = False):
        L = lib()
        ctx = C.c_void_p()
        rc = L.pcl_ctx_create(C.byref(ctx), int(device))
        if rc < 0:
            raise _ffi.PclError(rc, (L.pcl_last_error(None) or b"").decode())
        self.ctx = ctx
        self.layout = layout
        self.device = device
        self.qc_enabled = bool(qc)
        if qc:                                          # QcFastq needs the quality plane of insert reads too
            check(ctx, L.pcl_qc_enable(ctx, 1))
        if verify:                                      # Assay::verify: split_with_tolerance(read, 0.2, 0.2)
            check(ctx, L.pcl_verify_enable(ctx, 1))
        self._descs = layout.descs()
        check(ctx, L.pcl_layout_set(ctx, self._descs, len(layout.reads)))
        self.infos: List[_ffi.ReadInfo] = []
        for r in range(len(layout.reads)):
            info = _ffi.ReadInfo()
            check(ctx, L.pcl_read_info_get(ctx, r, C.byref(info)))
            self.infos.append(info)
        self.wl_bytes: Dict[int, List[bytes]] = {}
        for slot, rid in enumerate(layout.region_ids):
            wl = layout.whitelists[rid]
            if isinstance(wl, np.ndarray):             # [n, k] uint8: large equal-length list, no Python objects
                n, k = wl.shape
                self.load_whitelist_packed(slot, wl.reshape(-1), np.arange(n + 1, dtype=np.uint64) * np.uint64(k))
            elif getattr(wl, "unique", False) and hasattr(wl, "packed"):     # seqspec.OnlistItems: deduplicated, packed once
                blob, offs = wl.packed()
                self.load_whitelist_packed(slot, blob, offs)
                self.wl_bytes[slot] = wl
            else:
                self.load_whitelist(slot, list(dict.fromkeys(wl)))
        if nranks > 1:
            buf = C.create_string_buffer(nccl_uid, 128)
            check(ctx, L.pcl_comm_init(ctx, nranks, rank, buf))
        self.nranks, self.rank = nranks, rank
        self.chunks: List[Chunk] = []

    # -- configuration ---------------------------------------------------------------------------
    @staticmethod
    def new_unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        rc = lib().pcl_comm_unique_id(buf)
        if rc < 0:
            raise _ffi.PclError(rc, (lib().pcl_last_error(None) or b"").decode())
        return buf.raw

    def load_whitelist(self, slot: int, items: Sequence[bytes]) -> None:
        items = list(items)
        offs = np.zeros(len(items) + 1, dtype=np.uint64)
        if items:
            offs[1:] = np.cumsum([len(x) for x in items], dtype=np.uint64)
        blob = np.frombuffer(b"".join(items), dtype=np.uint8) if items else np.zeros(1, np.uint8)
        check(self.ctx, lib().pcl_whitelist_load(self.ctx, slot, blob.ctypes.data, offs.ctypes.data, len(items)))
        self.wl_bytes[slot] = items

    def load_whitelist_packed(self, slot: int, blob: np.ndarray, offsets: np.ndarray) -> None:
        """Large lists: one contiguous byte blob + uint64 offsets, no Python objects."""
        blob = np.ascontiguousarray(blob, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        check(self.ctx, lib().pcl_whitelist_load(self.ctx, slot, blob.ctypes.data, offsets.ctypes.data, len(offsets) - 1))

    def stride(self, read_slot: int) -> int:
        return lib().pcl_read_stride(self.ctx, read_slot)

    def qual_window(self, read_slot: int):
        """(offset, stride) of the quality plane: row i = quality bytes [offset, offset + stride) of record i."""
        info = self.infos[read_slot]
        return int(info.qual_offset), int(info.qual_stride)

    def host_planes(self, read_slot: int, seq_full: np.ndarray, qual_full: np.ndarray, length: np.ndarray) -> "HostBatch":
        """Cut full-width [n, W] sequence / quality arrays down to the planes this read uploads."""
        info = self.infos[read_slot]
        st, (qo, qs) = self.stride(read_slot), self.qual_window(read_slot)

        def cut(a, lo, w, fill):
            if a is None or w == 0:
                return None
            out = np.full((a.shape[0], w), fill, np.uint8)
            hi = min(a.shape[1], lo + w)
            if hi > lo:
                out[:, :hi - lo] = a[:, lo:hi]
            return out
        seq = cut(seq_full, 0, st, ord("N")) if info.needs_payload else None
        qual = cut(qual_full, qo, qs, ord("!")) if (info.needs_payload or info.qual_only) else None
        return HostBatch(seq, qual, length)

    def payload_bytes(self, read_slot: int) -> int:
        """Record bytes the path can touch, rounded up to 4: the narrowest row ``Chunk.upload_compact`` accepts."""
        return lib().pcl_read_payload_bytes(self.ctx, read_slot)

    def whitelist_size(self, slot: int) -> int:
        return lib().pcl_whitelist_size(self.ctx, slot)

    # -- chunks --------------------------------------------------------------------------------------
    def new_chunk(self, capacity: int) -> Chunk:
        c = Chunk(self, capacity)
        self.chunks.append(c)
        return c

    def alloc_results(self, read_slot: int, n: int) -> ReadResults:
        info = self.infos[read_slot]
        kd = lambda b: np.uint32 if b == 4 else np.uint64
        return ReadResults(
            info=info, fstatus=np.empty(n, np.uint16), tcoord=np.empty(n, np.uint32) if info.has_target else None,
            bc_key=[np.empty(n, kd(info.bc_key_bytes[j])) for j in range(info.n_bc)],
            umi_key=[np.empty(n, kd(info.umi_key_bytes[j])) for j in range(info.n_umi)],
            bcoord=[None if info.fixed_layout else np.empty(n, np.uint32) for _ in range(info.n_bc)],
            ucoord=[None if info.fixed_layout else np.empty(n, np.uint32) for _ in range(info.n_umi)])

    # -- stages --------------------------------------------------------------------------------------
    def reset_counts(self) -> None:
        check(self.ctx, lib().pcl_counts_reset(self.ctx))

    def count(self, chunk: Chunk) -> None:
        check(self.ctx, lib().pcl_count(self.ctx, chunk.h))

    def allreduce(self) -> None:
        check(self.ctx, lib().pcl_counts_allreduce(self.ctx))

    def correct(self, chunk: Chunk, threshold: float = 0.975, max_expected_errors: float = _ffi.F64_MAX) -> None:
        check(self.ctx, lib().pcl_correct(self.ctx, chunk.h, float(threshold), float(max_expected_errors)))

    def freeze_counts(self, on: bool = True) -> None:
        """After the all-reduce: ``count`` then recomputes a re-uploaded chunk without changing counts / totals / QC."""
        check(self.ctx, lib().pcl_counts_freeze(self.ctx, int(on)))

    def chunk_bytes(self, capacity: int) -> int:
        return int(lib().pcl_chunk_bytes(self.ctx, int(capacity)))

    def mem_info(self):
        f, t = C.c_uint64(), C.c_uint64()
        check(self.ctx, lib().pcl_mem_info(self.ctx, C.byref(f), C.byref(t)))
        return f.value, t.value

    def sync(self) -> None:
        check(self.ctx, lib().pcl_sync(self.ctx))

    def counts(self, slot: int):
        """(counts in whitelist order, mismatch_count, total_count) -- precellar/src/barcode.rs:362-366."""
        n = self.whitelist_size(slot)
        cnt = np.zeros(max(n, 1), np.uint64)
        mm, tot = C.c_uint64(), C.c_uint64()
        check(self.ctx, lib().pcl_counts_fetch(self.ctx, slot, cnt.ctypes.data, C.byref(mm), C.byref(tot)))
        return cnt[:n], mm.value, tot.value

    def qc(self) -> np.ndarray:
        """[n_reads, 2]: num_reads, num_defect per read file (QcFastq, precellar/src/qc.rs:23-24)."""
        out = np.zeros((len(self.layout.reads), 2), np.uint64)
        check(self.ctx, lib().pcl_qc_fetch(self.ctx, out.ctypes.data))
        return out

    def qc_accumulate(self, chunk: Chunk) -> None:
        """QcFastq::update over a counted chunk (bases / Q30 bases per category)."""
        check(self.ctx, lib().pcl_qc_accumulate(self.ctx, chunk.h))

    def qc_categories(self) -> np.ndarray:
        """[4, 3]: rows umi, barcode, read1, read2; columns num_reads, num_total_bases, num_q30_bases."""
        out = np.zeros((4, 3), np.uint64)
        check(self.ctx, lib().pcl_qc_fetch_categories(self.ctx, out.ctypes.data))
        return out

    def qc_json(self) -> dict:
        """QcFastq::to_json (precellar/src/qc.rs:91-121)."""
        per, cat = self.qc(), self.qc_categories()
        out = {}
        defect = {p.read_id: float(per[r, 1]) / float(per[r, 0]) for r, p in enumerate(self.layout.reads) if per[r, 1] > 0}
        if defect:
            out["frac_reads_with_misformed_structure"] = defect
        names = ("umi", "barcode", "read1", "read2")
        out["frac_q30_bases"] = {names[k]: float(cat[k, 2]) / float(cat[k, 1]) for k in range(4) if cat[k, 1] > 0}
        return out

    # -- downstream grouping ---------------------------------------------------------------------------
    def group(self, chunk, fingerprint: Optional[np.ndarray] = None) -> "GroupResult":
        """Sort the assigned read groups of a chunk - or of a list of chunks, numbered one after the other - by (barcode,
        fingerprint, UMI) on the device and mark the duplicate sets: sort_by_barcode + chunk_by(barcode) +
        remove_duplicates of the reference (precellar/src/fragment/mod.rs:359-363,434-455, fragment/dedups.rs:320-340).
        ``fingerprint``: one uint64 per read group with the aligner's part of the FingerPrint (reference id, 5'
        coordinates, orientation), or None."""
        chunks = list(chunk) if isinstance(chunk, (list, tuple)) else [chunk]
        fp = None
        if fingerprint is not None:
            fp = np.ascontiguousarray(fingerprint, dtype=np.uint64)
            assert len(fp) == sum(c.n for c in chunks)
        h = C.c_void_p()
        arr = (C.c_void_p * len(chunks))(*[c.h for c in chunks])
        check(self.ctx, lib().pcl_group_build_multi(self.ctx, arr, len(chunks), fp.ctypes.data if fp is not None else None, C.byref(h)))
        try:
            info = _ffi.GroupInfo()
            check(self.ctx, lib().pcl_group_info_get(h, C.byref(info)))
            order = np.empty(info.n_assigned, np.uint32)
            frag_start = np.empty(info.n_fragments + 1, np.uint32)
            bc_frag_start = np.empty(info.n_barcodes + 1, np.uint32)
            bc_key = np.empty(info.n_barcodes, np.uint64)
            check(self.ctx, lib().pcl_group_fetch(h, order.ctypes.data, frag_start.ctypes.data, bc_frag_start.ctypes.data, bc_key.ctypes.data))
        finally:
            lib().pcl_group_destroy(h)
        return GroupResult(int(info.n_records), int(info.n_assigned), int(info.n_ungroupable), order, frag_start, bc_frag_start, bc_key)

    # -- instrumentation -----------------------------------------------------------------------------
    def profile_enable(self, on: bool = True) -> None:
        check(self.ctx, lib().pcl_profile_enable(self.ctx, int(on)))

    def profile_reset(self) -> None:
        check(self.ctx, lib().pcl_profile_reset(self.ctx))

    def profile_get(self, k: int):
        ms, n = C.c_double(), C.c_uint64()
        check(self.ctx, lib().pcl_profile_get(self.ctx, k, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def launch_count(self) -> int:
        return lib().pcl_launch_count(self.ctx)

    # -- one-call convenience: the whole path over host batches ------------------------------------------
    def run(self, batches: Sequence[HostBatch], correct: bool = True, threshold: float = 0.975,
            max_expected_errors: float = _ffi.F64_MAX, chunk: Optional[Chunk] = None) -> List[ReadResults]:
        n = len(batches[0].length)
        own = chunk is None
        if own:
            chunk = self.new_chunk(max(n, 1))
        chunk.reset(n)
        for r, b in enumerate(batches):
            chunk.upload(r, b)
        self.count(chunk)
        self.allreduce()
        if correct:
            self.correct(chunk, threshold, max_expected_errors)
        res = [chunk.fetch(r) for r in range(len(batches))]
        if own:
            chunk.close()
            self.chunks.remove(chunk)
        return res

    def close(self) -> None:
        for c in self.chunks:
            c.close()
        self.chunks = []
        if self.ctx:
            lib().pcl_ctx_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---------------------------------------------------------------------------------------------------
# host-side assembly of AnnotatedFastq-equivalent records
# ---------------------------------------------------------------------------------------------------
_RC = bytes.maketrans(b"ACGTacgt", b"TGCATGCA")     # precellar::utils::rev_compl (precellar/src/utils.rs:138-149)


def rev_compl(seq: bytes) -> bytes:
    return seq.translate(_RC)[::-1]


def widen_key(key: np.ndarray, min_len: int, max_len: int) -> np.ndarray:
    """4-byte keys of 16-base elements have their marker bit dropped on the device; restore it."""
    k = key.astype(np.uint64)
    if key.dtype == np.uint32 and min_len == 16 and max_len == 16:
        k = k | np.uint64(1 << 32)
    return k


def decode_key(key: int) -> bytes:
    buf = C.create_string_buffer(40)
    n = lib().pcl_key_decode(int(key), buf)
    if n < 0:
        raise ValueError(f"not a packed key: {key:#x}")
    return buf.raw[:n]


def decode_key_py(key: int) -> bytes:
    """Pure-Python twin of pcl_key_decode (usable without the library)."""
    top = int(key).bit_length() - 1
    if key <= 0 or top % 2:
        raise ValueError(f"not a packed key: {key:#x}")
    L = top // 2
    return bytes(b"ACGT"[(key >> (2 * (L - 1 - p))) & 3] for p in range(L))


def element_coords(layout: CompiledLayout, infos: Sequence[_ffi.ReadInfo], r: int, res: ReadResults,
                   batch: HostBatch, elem: int, coord: Optional[np.ndarray], present: np.ndarray) -> np.ndarray:
    """(start << 16 | len) of one barcode/UMI element per record; COORD_ABSENT where not present.
    Variable layouts carry coordinates from the device; fixed layouts use the static table."""
    if coord is not None:
        return coord
    plan = layout.reads[r]
    info = infos[r]
    el = plan.elems[elem]
    start = int(info.static_start[elem])
    L = batch.length.astype(np.int64)
    if plan.max_len > 0:
        L = np.minimum(L, plan.max_len)
    if el.min_len == el.max_len:
        ln = np.full(len(L), el.max_len, np.int64)
    else:                                               # trailing variable element: [start, min(start+max, len))
        ln = np.minimum(start + el.max_len, L) - start
    c = ((start << 16) | ln).astype(np.uint32)
    return np.where(present, c, np.uint32(_ffi.COORD_ABSENT))


@dataclass
class Groups:
    """Vectorised view of the joined records (AnnotatedFastq after `join`, fastq.rs:571-603)."""
    n: int
    has_barcode: np.ndarray            # kept by process_chunk (fastq.rs:381-385)
    corrected_some: np.ndarray
    bc: List[dict] = field(default_factory=list)      # per barcode segment (join order): read, elem, code, key, coord
    umi: List[dict] = field(default_factory=list)     # per UMI element: read, elem, present, coord, key
    read1: Optional[dict] = None       # read, coord
    read2: Optional[dict] = None
    defect: Optional[np.ndarray] = None  # [n_reads, n] bool


def assemble(layout: CompiledLayout, infos: Sequence[_ffi.ReadInfo], batches: Sequence[HostBatch],
             results: Sequence[ReadResults], correct: bool = True) -> Groups:
    n = len(batches[0].length)
    g = Groups(n=n, has_barcode=np.zeros(n, bool), corrected_some=np.ones(n, bool))
    defect = np.zeros((len(layout.reads), n), bool)
    r1_owner = np.full(n, -1, np.int64)
    r2_owner = np.full(n, -1, np.int64)
    r1c = np.full(n, _ffi.COORD_ABSENT, np.uint32)
    r2c = np.full(n, _ffi.COORD_ABSENT, np.uint32)
    for r, (plan, info, res, batch) in enumerate(zip(layout.reads, infos, results, batches)):
        fs = res.fstatus.astype(np.uint32)
        st = fs & 3
        if np.any(st == _ffi.SPLIT_MULTI_TARGET):
            raise _ffi.PclError(_ffi.PCL_ERR_INPUT, "Multiple target regions found in one fastq record!")   # fastq.rs:505-506
        defect[r] = st != 0
        for j in range(info.n_bc):
            e = info.bc_elem[j]
            code = (fs >> (4 + 3 * j)) & 7
            present = code != _ffi.BC_ABSENT
            el = plan.elems[e]
            key = widen_key(res.bc_key[j], el.min_len, el.max_len)
            coord = element_coords(layout, infos, r, res, batch, e, res.bcoord[j], present)
            g.bc.append(dict(read=r, elem=e, code=code, key=key, coord=coord, present=present,
                             region=plan.region_slots[e]))
            g.has_barcode |= present
            if correct:
                ok = (code == _ffi.BC_EXACT) | (code == _ffi.BC_CORRECTED)
                g.corrected_some &= ok | ~present
        for j in range(info.n_umi):
            e = info.umi_elem[j]
            k = res.umi_key[j]
            absent = _ffi.KEY_ABSENT32 if k.dtype == np.uint32 else _ffi.KEY_ABSENT64
            present = k != k.dtype.type(absent)
            coord = element_coords(layout, infos, r, res, batch, e, res.ucoord[j], present)
            g.umi.append(dict(read=r, elem=e, present=present, coord=coord, key=k))
        if info.has_target:
            tp = (fs & _ffi.FSTATUS_TARGET) != 0
            if plan.is_reverse:                                       # fastq.rs:510-514
                if np.any(tp & (r2_owner >= 0)):
                    raise _ffi.PclError(_ffi.PCL_ERR_INPUT, "Read2 already exists")    # fastq.rs:596-599
                r2_owner[tp] = r
                r2c[tp] = res.tcoord[tp]
            else:
                if np.any(tp & (r1_owner >= 0)):
                    raise _ffi.PclError(_ffi.PCL_ERR_INPUT, "Read1 already exists")    # fastq.rs:588-591
                r1_owner[tp] = r
                r1c[tp] = res.tcoord[tp]
    g.corrected_some &= g.has_barcode
    g.read1 = dict(owner=r1_owner, coord=r1c)
    g.read2 = dict(owner=r2_owner, coord=r2c)
    g.defect = defect
    return g


def _slice(batch: HostBatch, i: int, coord: int):
    s, l = coord >> 16, coord & 0xFFFF
    seq = batch.seq[i, s:s + l].tobytes() if batch.seq is not None else b"N" * l
    qual = batch.qual[i, s:s + l].tobytes() if batch.qual is not None else b"!" * l
    return seq, qual


def render(layout: CompiledLayout, batches: Sequence[HostBatch], g: Groups, i: int, correct: bool = True) -> str:
    """One joined record as text:
    bc_seq, bc_qual, corrected|*, umi_seq, umi_qual, read1 seq/qual, read2 seq/qual (tab separated);
    the empty string for a group the reference drops (no barcode)."""
    if not g.has_barcode[i]:
        return ""
    bc_seq = bc_qual = corr = b""
    for b in g.bc:
        if not b["present"][i]:
            continue
        plan = layout.reads[b["read"]]
        s, q = _slice(batches[b["read"]], i, int(b["coord"][i]))
        if plan.is_reverse:                                           # rev_compl_fastq_record (utils.rs:132-136)
            s, q = rev_compl(s), q[::-1]
        bc_seq += s
        bc_qual += q
        if g.corrected_some[i]:
            code = int(b["code"][i])
            corr += s if (not correct or code == _ffi.BC_EXACT) else decode_key_py(int(b["key"][i]))
    umi_seq = umi_qual = b""
    for r in range(len(layout.reads)):
        last = None                                                   # `umi = Some(fq)`: the last UMI segment of a file wins (fastq.rs:502)
        for u in g.umi:
            if u["read"] == r and u["present"][i]:
                last = u
        if last is not None:
            s, q = _slice(batches[r], i, int(last["coord"][i]))
            if layout.reads[r].is_reverse:
                s, q = rev_compl(s), q[::-1]
            umi_seq += s
            umi_qual += q
    out = [bc_seq, bc_qual, corr if g.corrected_some[i] else b"*", umi_seq, umi_qual]
    for rd in (g.read1, g.read2):
        o = int(rd["owner"][i])
        if o >= 0:
            s, q = _slice(batches[o], i, int(rd["coord"][i]))
            out += [s, q]
        else:
            out += [b"", b""]
    return b"\t".join(out).decode("latin-1")
