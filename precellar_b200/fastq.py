"""FASTQ ingest helpers of the host layer (configuration-time, small inputs).

``infer_length`` mirrors ``Read::infer_length`` (seqspec/src/read.rs:158-172).  Bulk ingest into pinned
SoA batches lives in ``ingest.py``.
"""
from __future__ import annotations

from typing import Iterator, List, Sequence, Tuple

from .seqspec import open_file


def iter_fastq(paths: Sequence[str]) -> Iterator[Tuple[bytes, bytes, bytes]]:
    """(definition line without '@', sequence, quality) of 4-line FASTQ records; files are concatenated
    (read.rs:148-149)."""
    for p in paths:
        with open_file(p) as fh:
            while True:
                name = fh.readline()
                if not name:
                    break
                seq = fh.readline()
                plus = fh.readline()
                qual = fh.readline()
                if not qual and not plus:
                    raise ValueError(f"truncated FASTQ record in {p}")
                yield name.rstrip(b"\r\n")[1:], seq.rstrip(b"\r\n"), qual.rstrip(b"\r\n")


def infer_length(paths: Sequence[str], n: int) -> Tuple[int, int]:
    if not paths:
        raise RuntimeError("No fastq files found.")          # .expect("No fastq files found.")
    try:                                                     # the C++ record reader: lengths only, no planes
        from .ingest import FastqReader
        rdr = FastqReader(list(paths))
        try:
            b = rdr.read(int(n), 0)
        finally:
            rdr.close()
        if b is not None and len(b.batch.length):
            return int(b.batch.length.min()), int(b.batch.length.max())
    except Exception:                                        # records it does not take (> 65535 bases ...): the plain loop below
        pass
    lo, hi = 1 << 62, 0
    for k, (_, seq, _) in enumerate(iter_fastq(paths)):
        if k >= n:
            break
        lo, hi = min(lo, len(seq)), max(hi, len(seq))
    return lo, hi
