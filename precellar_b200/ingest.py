"""Host ingest: FASTQ files -> SoA planes ready for ``pcl_chunk_upload`` (C++ reader in csrc/ingest.cpp).

Mirrors ``Read::open`` / ``FastqReader`` (seqspec/src/read.rs:83-155) and the per-file half of
``BatchedFqReader`` (precellar/src/align/fastq.rs:392-449).  Decoding one file is single-threaded like the
reference's; the files of a read-group are decoded concurrently (ctypes releases the GIL).
"""
from __future__ import annotations

import ctypes as C
from concurrent.futures import ThreadPoolExecutor
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

from . import _ffi
from ._ffi import lib
from .pipeline import HostBatch


@dataclass
class FastqBatch:
    batch: HostBatch                 # planes with the requested stride (seq/qual None when stride == 0)
    name_hash: np.ndarray            # uint64 per record
    text: Optional[np.ndarray]       # arena: definition, sequence, quality of every record (keep_text)
    off: Optional[np.ndarray]        # uint64[3n+1] offsets into text

    def record(self, i: int):
        """(definition, sequence, quality) bytes of record i (needs keep_text)."""
        o = self.off
        t = self.text
        return (t[o[3 * i]:o[3 * i + 1]].tobytes(), t[o[3 * i + 1]:o[3 * i + 2]].tobytes(),
                t[o[3 * i + 2]:o[3 * i + 3]].tobytes())


class FastqReader:
    """The concatenated FASTQ files of one read."""

    def __init__(self, paths: Sequence[str]):
        arr = (C.c_char_p * len(paths))(*[p.encode() for p in paths])
        h = C.c_void_p()
        rc = lib().pcl_fastq_open(arr, len(paths), C.byref(h))
        if rc < 0:
            raise _ffi.PclError(rc, f"cannot open file: {paths[0] if paths else ''}")
        self.h = h
        self.paths = list(paths)

    def read(self, max_records: int, stride: int, keep_text: bool = False, text_bytes_per_record: int = 512,
             qual_only: bool = False, qual_window=None) -> Optional[FastqBatch]:
        n = int(max_records)
        qoff, qstride = qual_window if qual_window is not None else (0, stride)
        seq = np.empty((n, stride), np.uint8) if (stride and not qual_only) else None
        qual = np.empty((n, qstride), np.uint8) if qstride else None
        length = np.empty(n, np.uint16)
        nh = np.empty(n, np.uint64)
        text = off = None
        used = C.c_uint64(0)
        if keep_text:
            text = np.empty(n * text_bytes_per_record, np.uint8)
            off = np.zeros(3 * n + 1, np.uint64)
        # The arena is sized from `text_bytes_per_record`; records with long headers or untruncated long reads may
        # outgrow it: pcl_fastq_read then stops in front of the record that does not fit, the arena is doubled and the
        # same batch continues behind the records already delivered.
        g = 0
        at = lambda a, k: (a.ctypes.data + k * a.strides[0]) if a is not None else None
        while True:
            got = lib().pcl_fastq_read(self.h, n - g, stride, qoff, qstride, at(seq, g), at(qual, g), at(length, g), at(nh, g),
                                       text.ctypes.data if keep_text else None, len(text) if keep_text else 0,
                                       at(off, 3 * g) if keep_text else None, C.byref(used))
            if got < 0:
                raise _ffi.PclError(int(got), lib().pcl_fastq_error(self.h).decode())
            g += int(got)
            if not (keep_text and lib().pcl_fastq_text_full(self.h)):
                break
            bigger = np.empty(2 * len(text) + 4096, np.uint8)
            bigger[:used.value] = text[:used.value]
            text = bigger
        if g == 0:
            return None
        hb = HostBatch(None if seq is None else seq[:g], None if qual is None else qual[:g], length[:g])
        # (the arena is sized for the worst case: keep only what was written)
        return FastqBatch(hb, nh[:g], text[:used.value].copy() if keep_text else None, off[:3 * g + 1].copy() if keep_text else None)

    def close(self):
        if self.h:
            lib().pcl_fastq_close(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class GroupReader:
    """One record per file per read-group, like BatchedFqReader (fastq.rs:392-449): equal record counts and
    equal names (after stripping /1 /2) across the files are enforced."""

    def __init__(self, files_per_read: Sequence[Sequence[str]], strides: Sequence[int], qual_only: Sequence[bool],
                 keep_text: bool = False, text_bytes: Optional[Sequence[int]] = None, qual_windows=None):
        self.readers = [FastqReader(f) for f in files_per_read]
        self.strides = list(strides)
        self.qual_only = list(qual_only)
        self.keep_text = keep_text
        self.qual_windows = list(qual_windows) if qual_windows is not None else [(0, st) for st in self.strides]
        self.text_bytes = list(text_bytes) if text_bytes is not None else [2304] * len(self.readers)
        self.pool = ThreadPoolExecutor(max_workers=max(1, len(self.readers)))

    def next(self, max_records: int) -> Optional[List[FastqBatch]]:
        futs = [self.pool.submit(r.read, max_records, st, self.keep_text, tb, qo, qw)
                for r, st, qo, tb, qw in zip(self.readers, self.strides, self.qual_only, self.text_bytes, self.qual_windows)]
        out = [f.result() for f in futs]
        if all(b is None for b in out):
            return None
        if any(b is None for b in out) or len({len(b.batch.length) for b in out}) != 1:
            raise _ffi.PclError(_ffi.PCL_ERR_INPUT, "Unequal number of reads in the chunk")        # fastq.rs:435-436
        for b in out[1:]:
            if not np.array_equal(b.name_hash, out[0].name_hash):
                raise _ffi.PclError(_ffi.PCL_ERR_INPUT, "read names mismatch")                      # fastq.rs:438-442
        return out

    def close(self):
        for r in self.readers:
            r.close()
        self.pool.shutdown(wait=False)
