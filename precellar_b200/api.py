"""Public API of the path, mirroring precellar's: ``FastqProcessor`` and ``make_fastq``.

    reference                                                     here
    FastqProcessor::{new, with_modality, with_barcode_correct_prob, gen_barcoded_fastq, get_fastq_qc}
        precellar/src/align/fastq.rs:20-154                       class FastqProcessor
    AnnotatedFastq / Barcode                                      dataclasses below
        precellar/src/align/fastq.rs:528-604
    precellar.make_fastq(assay, *, modality, out_dir, correct_barcode=False)
        python/src/lib.rs:117-171                                 make_fastq

The per-read work (split, whitelist count, correction, QC counters) runs in the CUDA library; this module
only moves batches in and turns the result planes back into records.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Iterator, List, Optional, Sequence, Union

import numpy as np

from . import _ffi
from .ingest import FastqBatch, GroupReader
from .layout import CompiledLayout, compile_layout, is_paired_end
from .pipeline import BarcodePipeline, Chunk, assemble, decode_key_py, rev_compl
from .seqspec import Assay, parse_modality


# Stage times of the last run (seconds), filled when PCL_API_TIMING=1: where an API-level run spends its wall clock.
TIMING: dict = {}


class _timed:
    def __init__(self, name):
        self.name = name

    def __enter__(self):
        import time
        self.t0 = time.perf_counter()

    def __exit__(self, *a):
        import time
        if os.environ.get("PCL_API_TIMING") == "1":
            TIMING[self.name] = TIMING.get(self.name, 0.0) + time.perf_counter() - self.t0


# Pass 2 needs the text of every record again.  The reference decodes the files a second time; here the text read in
# pass 1 stays in host memory when the (estimated) decoded size of all files is below this many bytes, so that gzip
# inputs are inflated once.  Larger inputs are re-read like the reference does.
KEEP_TEXT_BYTES = int(os.environ.get("PCL_KEEP_TEXT_BYTES", 8 << 30))

# Device chunks hold at least this many read groups whatever ``chunk_size`` asks for (results do not depend on the
# batching; small chunks only multiply launches and allocations).  Tests lower it to exercise multi-chunk runs.
MIN_CHUNK_RECORDS = 262144


@dataclass
class FastqRecord:
    """name + description (the FASTQ definition line without '@'), sequence, quality."""
    definition: bytes
    sequence: bytes
    quality: bytes

    @property
    def name(self) -> bytes:
        return self.definition.split(None, 1)[0] if self.definition else b""

    def to_bytes(self) -> bytes:
        return b"@" + self.definition + b"\n" + self.sequence + b"\n+\n" + self.quality + b"\n"


@dataclass
class Barcode:                                   # fastq.rs:528-532
    raw: FastqRecord
    corrected: Optional[bytes]


@dataclass
class AnnotatedFastq:                            # fastq.rs:549-556
    barcode: Optional[Barcode]
    umi: Optional[FastqRecord]
    read1: Optional[FastqRecord]
    read2: Optional[FastqRecord]

    def tags(self) -> dict:
        """The SAM tags the aligner adapters attach to every alignment of this read
        (add_cell_barcode / add_umi, precellar/src/align/aligners.rs:347-372): CB corrected barcode (absent when not
        corrected), CR / CY raw barcode and its qualities, UR / UY UMI and its qualities."""
        out = {}
        if self.barcode is not None:
            if self.barcode.corrected is not None:
                out["CB"] = self.barcode.corrected
            out["CR"] = self.barcode.raw.sequence
            out["CY"] = self.barcode.raw.quality
        if self.umi is not None:
            out["UR"] = self.umi.sequence
            out["UY"] = self.umi.quality
        return out


def _as_assays(assay) -> List[Assay]:
    """extract_assays (python/src/pyseqspec.rs:12-27): a path, an Assay, or a list of either."""
    items = assay if isinstance(assay, (list, tuple)) else [assay]
    return [a if isinstance(a, Assay) else Assay(os.fspath(a)) for a in items]


# chunk_size make_fastq hands to gen_barcoded_fastq (python/src/lib.rs:133: 1000000 bases); tests lower it together with
# MIN_CHUNK_RECORDS to run the writers over several chunks
MAKE_FASTQ_CHUNK_BASES = 1_000_000

# What the last make_fastq of this process did: {"writer": "device" | "host", "compression": ..., "streamed_chunks": n}
LAST_RUN: dict = {}


class FastqProcessor:
    """Barcode counting, correction and annotation of the FASTQ files named by one or more assays."""

    def __init__(self, assay, device: int = 0, device_parse: bool = False, devices: Optional[Sequence[int]] = None):
        """``devices``: several GPUs of the box.  The record stream is dealt out chunk by chunk (chunk k goes to
        ``devices[k % len(devices)]``), every GPU counts its chunks against its own copy of the whitelist tables, the
        counts are summed with one NCCL all-reduce (``pcl_counts_allreduce``, one rank per GPU inside this process) and
        every GPU corrects its own chunks; records come out in input order.  ``device`` is the single-GPU form."""
        self.assays = _as_assays(assay)
        self.devices = [int(d) for d in devices] if devices else [int(device)]
        if len(set(self.devices)) != len(self.devices):
            raise ValueError("devices must be distinct")
        self.device = self.devices[0]
        # pass 1 reads the FASTQ text on the device (textingest.TextGroupReader) instead of the host record reader;
        # stricter dialect: no blank lines between records
        self.device_parse = bool(device_parse)
        self.current_modality: Optional[str] = None
        self.barcode_correct_prob = 0.975        # fastq.rs:39
        self.mismatch_in_barcode = 1             # fastq.rs:40 (no setter in the reference either)
        self._qc = None
        self._layouts: dict = {}
        self.compute_qc = True
        # device bytes per GPU the resident chunks may take (None: 60 % of the free memory); batches past it are
        # streamed - counted in pass 1, uploaded and recounted against the frozen counts in pass 2
        self.resident_bytes: Optional[int] = None
        self.streamed_chunks = 0

    def with_modality(self, modality: str) -> "FastqProcessor":
        self.current_modality = parse_modality(modality)
        return self

    def with_barcode_correct_prob(self, prob: float) -> "FastqProcessor":
        self.barcode_correct_prob = float(prob)
        return self

    def modality(self) -> str:
        if self.current_modality is None:
            raise RuntimeError("modality not set, please call set_modality first")
        return self.current_modality

    def get_fastq_qc(self) -> dict:
        """QcFastq::to_json of everything processed so far (precellar/src/qc.rs:91-121)."""
        if self._qc is None:
            raise RuntimeError("fastq qc not found")
        return self._qc

    def _layout(self, assay) -> CompiledLayout:
        """compile_layout reads the whitelist files: once per (assay, modality)."""
        key = (id(assay), self.modality())
        if key not in self._layouts:
            self._layouts[key] = compile_layout(assay, self.modality(), require_files=True)
        return self._layouts[key]

    def is_paired_end(self) -> bool:
        vals = {is_paired_end(self._layout(a)) for a in self.assays}
        if len(vals) != 1:
            raise RuntimeError("Not all readers are with the same paired-end status")
        return vals.pop()

    # ------------------------------------------------------------------------------------------
    def gen_barcoded_fastq(self, correct_barcode: bool, chunk_size: int = 1_000_000) -> Iterator[List[AnnotatedFastq]]:
        """Batches of annotated records, assays one after the other (fastq.rs:109-153, 243-255)."""
        for layout, pipe, text, results in self._gen_chunks(correct_barcode, chunk_size):
            yield self._annotate(layout, pipe, text, results, correct_barcode)

    def _gen_chunks(self, correct_barcode: bool, chunk_size: int, emit: Optional[dict] = None):
        """(layout, pipeline, record text per read file, fetched result planes) per device chunk.  With ``emit``
        ({"paired": bool, "compress": bool}, the device writer of make_fastq) the last two are None and what
        ``Chunk.emit_fastq`` returned."""
        modality = self.modality()
        per_file = None
        cats = np.zeros((4, 3), np.uint64)
        read_ids: List[str] = []
        for assay in self.assays:
            with _timed("compile_layout+whitelists"):
                layout = self._layout(assay)
            for p in layout.reads:
                if p.has_barcode and not p.files:
                    raise RuntimeError(f"Failed to open FASTQ file: {p.read_id}")             # barcode.rs:88-90
            # reads of the modality that carry barcodes but no files make BarcodeAnalyzer::new panic too
            for read, info in assay.get_segments_by_modality(modality):
                if any(e.is_barcode() for e in info) and not read.fastq_paths():
                    raise RuntimeError(f"Failed to open FASTQ file: {read.read_id}")
            with _timed("pipeline_create"):
                pipes = self._make_pipes(layout)
            pipe = pipes[0]
            if os.environ.get("PCL_API_KERNEL_MS") == "1":               # diagnostic: every kernel bracketed by events
                pipe.profile_enable(True)
            try:
                if emit is not None:
                    for got in self._emit_assay(pipe, layout, correct_barcode, chunk_size, emit):
                        yield layout, pipe, None, got
                for text, results in (() if emit is not None else self._run_assay(pipes, layout, correct_barcode, chunk_size)):
                    yield layout, pipe, text, results
                q = sum(pp.qc() for pp in pipes)
                cats += sum(pp.qc_categories() for pp in pipes) if self.compute_qc else 0
                if per_file is None:
                    per_file, read_ids = q.copy(), [p.read_id for p in layout.reads]
                else:                                                                      # QcFastq::extend, keyed by read id
                    for r, p in enumerate(layout.reads):
                        if p.read_id in read_ids:
                            per_file[read_ids.index(p.read_id)] += q[r]
                        else:
                            read_ids.append(p.read_id)
                            per_file = np.vstack([per_file, q[r:r + 1]])
            finally:
                if os.environ.get("PCL_API_KERNEL_MS") == "1":
                    names = {_ffi.K_COUNT: "count", _ffi.K_COUNT_LENONLY: "count_lenonly", _ffi.K_FILTER: "filter", _ffi.K_CORRECT: "resolve",
                             _ffi.K_QC: "qc", _ffi.K_TEXT_INDEX: "text_index", _ffi.K_TEXT_FILL: "text_fill", _ffi.K_EMIT: "emit",
                             _ffi.K_ZSTD: "zstd_blocks"}
                    try:
                        TIMING["kernel_ms"] = {nm: round(pipe.profile_get(k)[0], 3) for k, nm in names.items()}
                    except _ffi.PclError:
                        pass
                for pp in pipes:
                    pp.close()
            out = {}
            defect = {rid: float(per_file[k, 1]) / float(per_file[k, 0]) for k, rid in enumerate(read_ids) if per_file[k, 1] > 0}
            if defect:
                out["frac_reads_with_misformed_structure"] = defect
            names = ("umi", "barcode", "read1", "read2")
            out["frac_q30_bases"] = {names[k]: float(cats[k, 2]) / float(cats[k, 1]) for k in range(4) if cats[k, 1] > 0}
            self._qc = out

    def _make_pipes(self, layout: CompiledLayout) -> List[BarcodePipeline]:
        """One pipeline (ctx, whitelist tables, NCCL rank) per GPU.  ncclCommInitRank waits for all ranks, so the
        pipelines are created by one thread each (the ctypes calls release the GIL)."""
        if len(self.devices) == 1:
            return [BarcodePipeline(layout, device=self.devices[0], qc=self.compute_qc)]
        import threading
        uid = BarcodePipeline.new_unique_id()
        G = len(self.devices)
        out: List[Optional[BarcodePipeline]] = [None] * G
        errs: List[BaseException] = []

        def make(g):
            try:
                out[g] = BarcodePipeline(layout, device=self.devices[g], nranks=G, rank=g, nccl_uid=uid, qc=self.compute_qc)
            except BaseException as e:          # noqa: BLE001 - re-raised below
                errs.append(e)
        ths = [threading.Thread(target=make, args=(g,)) for g in range(G)]
        for t in ths:
            t.start()
        for t in ths:
            t.join()
        if errs:
            for pp in out:
                if pp is not None:
                    pp.close()
            raise errs[0]
        return out          # type: ignore[return-value]

    @staticmethod
    def _allreduce_all(pipes: Sequence[BarcodePipeline]) -> None:
        """The collective of every rank, issued by one thread per rank (a single thread driving several NCCL ranks
        one after the other may deadlock)."""
        if len(pipes) == 1:
            pipes[0].allreduce()
            return
        import threading
        errs: List[BaseException] = []

        def go(pp):
            try:
                pp.allreduce()
            except BaseException as e:          # noqa: BLE001
                errs.append(e)
        ths = [threading.Thread(target=go, args=(pp,)) for pp in pipes]
        for t in ths:
            t.start()
        for t in ths:
            t.join()
        if errs:
            raise errs[0]

    def _run_assay(self, pipes: Sequence[BarcodePipeline], layout: CompiledLayout, correct_barcode: bool, chunk_size: int):
        pipe = pipes[0]
        G = len(pipes)
        n_reads = len(layout.reads)
        strides = [pipe.stride(r) for r in range(n_reads)]
        qonly = [bool(pipe.infos[r].qual_only) for r in range(n_reads)]
        files = [p.files for p in layout.reads]
        bases = sum((p.max_len or 100) for p in layout.reads)
        # the reference batches >= chunk_size bases (fastq.rs:406-445); results do not depend on the batching, so the
        # device chunks are at least 256 K read groups to keep launches and allocations few
        per_chunk = max(MIN_CHUNK_RECORDS, int(chunk_size) // max(bases, 1))

        # pass 1 (BarcodeAnalyzer::new): every batch is uploaded, counted and kept resident on the device
        chunks: List[Chunk] = []
        sizes: List[int] = []
        owners: List[BarcodePipeline] = []
        if self.device_parse:
            if G > 1:
                raise NotImplementedError("device_parse drives one GPU (the text staging buffers belong to one ctx)")
            from .textingest import TextGroupReader
            trdr = TextGroupReader(pipe, files, per_chunk, [2 * (p.max_len or 150) + 80 for p in layout.reads])
            try:
                while True:
                    ch = pipe.new_chunk(per_chunk)
                    n = trdr.next_into(ch)
                    if n == 0:
                        ch.close()
                        pipe.chunks.remove(ch)
                        break
                    pipe.count(ch)
                    if self.compute_qc:
                        pipe.qc_accumulate(ch)
                    chunks.append(ch)
                    sizes.append(n)
                    owners.append(pipe)
            finally:
                trdr.close()
        est = 0
        for fs_ in files:
            for f in fs_:
                try:
                    est += os.path.getsize(f) * (5 if f.endswith((".gz", ".zst", ".bgz")) else 1)
                except OSError:
                    pass
        keep = (not self.device_parse) and est <= KEEP_TEXT_BYTES
        kept: List[Sequence[FastqBatch]] = []
        text_bytes = [2 * max(p.max_len, 1024) + 256 for p in layout.reads]
        rdr = GroupReader(files, strides, qonly, keep_text=keep, text_bytes=text_bytes if keep else None,
                          qual_windows=[pipe.qual_window(r) for r in range(n_reads)])
        # chunks stay resident on their GPU up to a budget; the batches past it are "streamed": counted through one
        # reusable chunk per GPU now, and uploaded again in pass 2 - which is what the reference does for every batch
        # (two reads of the files, fastq.rs:183-199 then fastq.rs:268-331)
        budget = [self._resident_budget(pp) for pp in pipes]
        used = [0] * G
        spare: List[Optional[Chunk]] = [None] * G
        self.streamed_chunks = 0
        try:
            while not self.device_parse:
                with _timed("pass1_host_read"):
                    got = rdr.next(per_chunk)
                if got is None:
                    break
                n = len(got[0].batch.length)
                g = len(chunks) % G              # chunk k lives on GPU k mod G
                own = pipes[g]
                need = own.chunk_bytes(n)
                if used[g] + need <= budget[g]:
                    ch = own.new_chunk(n)
                    used[g] += need
                else:
                    if spare[g] is None:
                        spare[g] = own.new_chunk(per_chunk)
                    ch = None
                    self.streamed_chunks += 1
                tgt = ch if ch is not None else spare[g]
                tgt.reset(n)
                for r in range(n_reads):
                    tgt.upload(r, got[r].batch)
                own.count(tgt)
                if self.compute_qc:
                    own.qc_accumulate(tgt)
                with _timed("pass1_device_wait"):
                    own.sync()                   # the host planes of this batch go out of scope
                if keep:
                    kept.append(got)
                chunks.append(ch)
                sizes.append(n)
                owners.append(own)
        finally:
            rdr.close()
        self._allreduce_all(pipes)               # every GPU now holds the counts of the whole input
        # "All whitelists should have the same total read count" (barcode.rs:125-129)
        totals = {pipe.counts(g)[2] for g in range(len(layout.region_ids))}
        if len(totals) > 1:
            raise _ffi.PclError(_ffi.PCL_ERR_INPUT, "All whitelists should have the same total read count")
        self.num_reads = totals.pop() if totals else 0

        # pass 2 (AnnotatedFastqReader): correct, fetch, and rebuild the records from a second read of the files
        streaming = any(ch is None for ch in chunks)
        rdr = None
        if not keep:                             # streamed batches need their sequence / quality planes again
            rdr = GroupReader(files, strides if streaming else [0] * n_reads, qonly if streaming else [False] * n_reads,
                              keep_text=True, qual_windows=[pipe.qual_window(r) for r in range(n_reads)] if streaming else [(0, 0)] * n_reads,
                              text_bytes=text_bytes)
        try:
            if streaming:                        # recounting a batch must not move the counts the correction reads
                for pp in pipes:
                    pp.freeze_counts(True)
            if correct_barcode:                  # every GPU corrects its resident chunks while the host re-reads the files
                for ch, own in zip(chunks, owners):
                    if ch is not None:
                        own.correct(ch, self.barcode_correct_prob)
            for k, (ch, n, own) in enumerate(zip(chunks, sizes, owners)):
                if keep:
                    text, kept[k] = kept[k], None
                else:
                    text = rdr.next(n)
                assert text is not None and len(text[0].batch.length) == n
                if ch is None:
                    tgt = spare[pipes.index(own)]
                    tgt.reset(n)
                    for r in range(n_reads):
                        tgt.upload(r, text[r].batch)
                    own.count(tgt)
                    if correct_barcode:
                        own.correct(tgt, self.barcode_correct_prob)
                else:
                    tgt = ch
                with _timed("pass2_fetch"):
                    results = [tgt.fetch(r) for r in range(n_reads)]
                yield text, results
                if ch is not None:
                    ch.close()
                    own.chunks.remove(ch)
        finally:
            if rdr is not None:
                rdr.close()
            if streaming:
                for pp in pipes:
                    pp.freeze_counts(False)

    def _emit_assay(self, pipe: BarcodePipeline, layout: CompiledLayout, correct_barcode: bool, chunk_size: int, emit: dict):
        """make_fastq with the device writer: the FASTQ text is parsed on the GPU and stays there (``Chunk.text_keep``),
        pass 2 corrects the resident chunks and every chunk's records are assembled - and with emit["compress"]
        entropy-coded into zstd frames - in HBM; the host appends finished bytes to the files.

        Input whose text does not fit the device budget is streamed instead, the way the reference works (two reads of
        the files, fastq.rs:183-199 then 268-331): pass 1 only counts, through one reusable chunk; pass 2 parses the files
        again and counts each chunk against the frozen counts, corrects and emits it.  Device memory is O(chunk)."""
        from .textingest import TextGroupReader
        n_reads = len(layout.reads)
        files = [p.files for p in layout.reads]
        bases = sum((p.max_len or 100) for p in layout.reads)
        per_chunk = max(MIN_CHUNK_RECORDS, int(chunk_size) // max(bases, 1))
        resident = self._device_writer_resident(pipe)
        self.streamed_chunks = 0
        bpr = [2 * (p.max_len or 150) + 80 for p in layout.reads]
        chunks: List[Chunk] = []
        spare: Optional[Chunk] = None
        with _timed("text_reader_open"):
            trdr = TextGroupReader(pipe, files, per_chunk, bpr)
        try:
            while True:
                with _timed("chunk_create"):
                    ch = pipe.new_chunk(per_chunk) if (resident or spare is None) else spare
                with _timed("pass1_text_parse"):
                    n = trdr.next_into(ch)
                if n == 0:
                    if resident:
                        ch.close()
                        pipe.chunks.remove(ch)
                    else:
                        spare = ch
                    break
                with _timed("pass1_keep_count"):
                    if resident:
                        for r in range(n_reads):
                            ch.text_keep(r)
                    pipe.count(ch)
                    if self.compute_qc:
                        pipe.qc_accumulate(ch)
                if resident:
                    chunks.append(ch)
                else:
                    spare = ch
                    self.streamed_chunks += 1
        finally:
            with _timed("text_reader_close"):
                trdr.close()
        with _timed("allreduce_counts"):
            pipe.allreduce()
            totals = {pipe.counts(g)[2] for g in range(len(layout.region_ids))}
        if len(totals) > 1:
            raise _ffi.PclError(_ffi.PCL_ERR_INPUT, "All whitelists should have the same total read count")
        self.num_reads = totals.pop() if totals else 0
        if resident:
            if correct_barcode:
                with _timed("pass2_correct_launch"):
                    for ch in chunks:
                        pipe.correct(ch, self.barcode_correct_prob)
            for ch in chunks:
                with _timed("emit"):
                    got = ch.emit_fastq(correct_barcode, emit["paired"], emit["compress"])
                yield got
                with _timed("chunk_close"):
                    ch.close()
                    pipe.chunks.remove(ch)
            return
        # streamed: the second read of the files
        pipe.freeze_counts(True)                 # recounting a chunk must not move the counts the correction reads
        trdr = TextGroupReader(pipe, files, per_chunk, bpr)
        try:
            while True:
                with _timed("pass2_text_parse"):
                    n = trdr.next_into(spare)
                if n == 0:
                    break
                for r in range(n_reads):
                    spare.text_keep(r)
                pipe.count(spare)
                if correct_barcode:
                    pipe.correct(spare, self.barcode_correct_prob)
                with _timed("emit"):
                    got = spare.emit_fastq(correct_barcode, emit["paired"], emit["compress"])
                yield got
        finally:
            trdr.close()
            pipe.freeze_counts(False)
            spare.close()
            pipe.chunks.remove(spare)

    def device_writer_obstacle(self) -> Optional[str]:
        """Why make_fastq cannot assemble its output on the device for these assays (None: it can)."""
        if len(self.devices) != 1:
            return "the device writer drives one GPU"
        est = 0
        for assay in self.assays:
            layout = self._layout(assay)
            for p in layout.reads:
                if not p.files:
                    return f"read {p.read_id} has no FASTQ files"
                for f in p.files:
                    try:
                        est += os.path.getsize(f) * (6 if f.endswith((".gz", ".zst", ".bgz")) else 1)
                    except OSError:
                        return f"cannot stat {f}"
        self._text_estimate = est
        return None

    def _device_writer_resident(self, pipe: BarcodePipeline) -> bool:
        """Whether the text of all files may stay on the device between the passes (else the device writer streams)."""
        est = getattr(self, "_text_estimate", 0)
        fixed = self.resident_bytes if self.resident_bytes is not None else os.environ.get("PCL_API_RESIDENT_BYTES")
        if fixed not in (None, ""):
            budget = int(fixed)
        else:
            free, _total = pipe.mem_info()
            budget = int(free * 0.45)
        return est <= budget

    def _resident_budget(self, pipe: BarcodePipeline) -> int:
        """Bytes of device memory the resident chunks of one GPU may take: `resident_bytes` (or the environment's
        PCL_API_RESIDENT_BYTES), else 60 % of what is free after the whitelist tables were built."""
        fixed = self.resident_bytes if self.resident_bytes is not None else os.environ.get("PCL_API_RESIDENT_BYTES")
        if fixed is not None and str(fixed) != "":
            return int(fixed)
        free, _total = pipe.mem_info()
        return int(free * 0.6)

    @staticmethod
    def _annotate(layout: CompiledLayout, pipe: BarcodePipeline, text: Sequence[FastqBatch], results,
                  correct: bool) -> List[AnnotatedFastq]:
        batches = [t.batch for t in text]
        g = assemble(layout, pipe.infos, batches, results, correct=correct)
        out: List[AnnotatedFastq] = []
        n_reads = len(layout.reads)
        for i in np.flatnonzero(g.has_barcode):                       # groups without a barcode are dropped (fastq.rs:381-385)
            i = int(i)
            recs = [t.record(i) for t in text]
            bc_def = None
            bseq = bqual = corr = b""
            for b in g.bc:
                if not b["present"][i]:
                    continue
                r = b["read"]
                c = int(b["coord"][i])
                s0, ln = c >> 16, c & 0xFFFF
                s, q = recs[r][1][s0:s0 + ln], recs[r][2][s0:s0 + ln]
                if layout.reads[r].is_reverse:
                    s, q = rev_compl(s), q[::-1]
                if bc_def is None:
                    bc_def = recs[r][0]
                bseq += s
                bqual += q
                if g.corrected_some[i]:
                    corr += s if (not correct or int(b["code"][i]) == _ffi.BC_EXACT) else decode_key_py(int(b["key"][i]))
            umi = None
            for r in range(n_reads):
                last = None
                for u in g.umi:
                    if u["read"] == r and u["present"][i]:
                        last = u
                if last is None:
                    continue
                c = int(last["coord"][i])
                s0, ln = c >> 16, c & 0xFFFF
                s, q = recs[r][1][s0:s0 + ln], recs[r][2][s0:s0 + ln]
                if layout.reads[r].is_reverse:
                    s, q = rev_compl(s), q[::-1]
                if umi is None:
                    umi = FastqRecord(recs[r][0], s, q)
                else:                                                  # extend_fastq_record (fastq.rs:606-610)
                    umi.sequence += s
                    umi.quality += q
            reads = []
            for rd in (g.read1, g.read2):
                o = int(rd["owner"][i])
                if o < 0:
                    reads.append(None)
                else:
                    c = int(rd["coord"][i])
                    s0, ln = c >> 16, c & 0xFFFF
                    reads.append(FastqRecord(recs[o][0], recs[o][1][s0:s0 + ln], recs[o][2][s0:s0 + ln]))
            out.append(AnnotatedFastq(Barcode(FastqRecord(bc_def, bseq, bqual), corr if g.corrected_some[i] else None),
                                      umi, reads[0], reads[1]))
        return out


def make_fastq(assay, *, modality: str, out_dir, correct_barcode: bool = False, device: int = 0,
               devices: Optional[Sequence[int]] = None, writer: str = "auto", compression: str = "zstd") -> None:
    """Generate consolidated fastq files from the sequencing specification.

    The barcodes and UMIs are concatenated to the read 1 sequence; fixed sequences and linkers are removed
    (python/src/lib.rs:114-171).  Writes ``R1.fq.zst`` and, for paired-end layouts, ``R2.fq.zst``; with
    ``correct_barcode`` reads whose barcode cannot be corrected are dropped and the corrected sequence is written.

    ``writer``: "device" parses the FASTQ text and assembles the output records on the GPU (one GPU; the text stays
    resident between the passes when it fits, else the files are read twice like the reference does; the stricter
    FASTQ dialect of the device parser: no blank lines between records), "host" reads and assembles records on the
    host (several GPUs), "auto" picks the device writer when it applies.
    ``compression``: "zstd" = libzstd level 9 on the host like the reference (seqspec/src/utils.rs:33-36); "device" =
    zstd frames built on the GPU (Huffman-coded literals and FSE-coded matches against the previous record: every zstd
    decoder reads them, the files are about 1.3-1.4 times the size of level 9) - only with the device writer.  The decompressed bytes are the same
    in every combination.
    """
    from . import _zstd
    if writer not in ("auto", "device", "host") or compression not in ("zstd", "device"):
        raise ValueError("writer is 'auto', 'device' or 'host'; compression is 'zstd' or 'device'")
    proc = FastqProcessor(assay, device=device, devices=devices).with_modality(modality)
    out_dir = os.fspath(out_dir)
    os.makedirs(out_dir, exist_ok=True)
    with _timed("is_paired_end"):
        paired = proc.is_paired_end()
    asked_writer = writer
    if writer != "host":
        with _timed("writer_choice"):
            why = proc.device_writer_obstacle()
        if why is not None and (writer == "device" or compression == "device"):
            raise RuntimeError(f"make_fastq: the device writer does not apply: {why}")
        writer = "device" if why is None else "host"
    elif compression == "device":
        raise ValueError("compression='device' needs the device writer")
    if writer == "device":
        state = {"wrote": False}
        try:
            _make_fastq_device(proc, out_dir, correct_barcode, paired, compression == "device", state)
            LAST_RUN.update(writer="device", compression=compression, streamed_chunks=proc.streamed_chunks)
            return
        except (_ffi.PclError, RuntimeError) as e:
            # the device parser's FASTQ dialect is stricter than the host reader's (no blank lines between records): in
            # "auto" mode an input it rejects before anything was written gets the host path's verdict instead
            code = getattr(e, "code", getattr(e.__cause__, "code", None))
            if asked_writer != "auto" or compression == "device" or state["wrote"] or code != _ffi.PCL_ERR_INPUT:
                raise
        proc = FastqProcessor(assay, device=device, devices=devices).with_modality(modality)
    LAST_RUN.update(writer="host", compression="zstd", streamed_chunks=0)
    # level 9 like the reference (seqspec/src/utils.rs:33-36); it asks for 8 worker threads, here every core up to 16 helps
    zt = max(1, min(16, os.cpu_count() or 8))
    w1 = _zstd.Writer(os.path.join(out_dir, "R1.fq.zst"), threads=zt)
    w2 = _zstd.Writer(os.path.join(out_dir, "R2.fq.zst"), threads=zt) if paired else None
    import ctypes as C
    L = _ffi.lib()
    try:
        for layout, pipe, text, results in proc._gen_chunks(correct_barcode, MAKE_FASTQ_CHUNK_BASES):
            n = len(text[0].batch.length)
            hr = (_ffi.HostRead * len(text))()
            keep = []
            for r, (t, res) in enumerate(zip(text, results)):
                ln = np.ascontiguousarray(t.batch.length, np.uint16)
                keep.append(ln)
                hr[r].text, hr[r].off, hr[r].len = t.text.ctypes.data, t.off.ctypes.data, ln.ctypes.data
                hr[r].res.fstatus = res.fstatus.ctypes.data
                if res.tcoord is not None:
                    hr[r].res.tcoord = res.tcoord.ctypes.data
                for j, k in enumerate(res.bc_key):
                    hr[r].res.bc_key[j] = k.ctypes.data
                    if res.bcoord[j] is not None:
                        hr[r].res.bcoord[j] = res.bcoord[j].ctypes.data
                for j, k in enumerate(res.umi_key):
                    if res.ucoord[j] is not None:
                        hr[r].res.ucoord[j] = res.ucoord[j].ctypes.data
            cap1 = int(sum(len(t.text) for t in text)) + 8 * n + 64
            cap2 = cap1 if paired else 0
            o1 = np.empty(cap1, np.uint8)
            o2 = np.empty(max(cap2, 1), np.uint8)
            u1, u2, nw = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0)
            with _timed("assemble"):
                _ffi.check(pipe.ctx, L.pcl_make_fastq_batch(pipe.ctx, hr, n, int(correct_barcode), int(paired), o1.ctypes.data, cap1,
                                                            C.byref(u1), o2.ctypes.data if paired else None, cap2,
                                                            C.byref(u2) if paired else None, C.byref(nw)))
            with _timed("zstd_submit"):
                w1.write(o1[:u1.value])
                if w2 is not None:
                    w2.write(o2[:u2.value])
    except _ffi.PclError as e:
        if "unwrap" in str(e):
            raise RuntimeError(str(e)) from e
        raise
    finally:
        with _timed("zstd_drain"):
            w1.close()
            if w2 is not None:
                w2.close()


def _make_fastq_device(proc: FastqProcessor, out_dir: str, correct_barcode: bool, paired: bool, device_zstd: bool,
                       state: Optional[dict] = None) -> None:
    """The device writer: every chunk comes back as finished bytes - zstd frames built on the GPU (appended to the files
    by a writer thread while the next chunk is emitted), or FASTQ text that goes through the host's level-9 encoder."""
    from collections import deque
    from concurrent.futures import ThreadPoolExecutor
    from . import _zstd
    p1, p2 = os.path.join(out_dir, "R1.fq.zst"), os.path.join(out_dir, "R2.fq.zst")
    zt = max(1, min(16, os.cpu_count() or 8))
    if device_zstd:
        f1, f2 = open(p1, "wb"), (open(p2, "wb") if paired else None)
        pool, pending = ThreadPoolExecutor(max_workers=1), deque()
    else:
        w1 = _zstd.Writer(p1, threads=zt)
        w2 = _zstd.Writer(p2, threads=zt) if paired else None

    def append(a, b):
        if a is not None:
            f1.write(a)
        if b is not None and f2 is not None:
            f2.write(b)
    try:
        gen = proc._gen_chunks(correct_barcode, MAKE_FASTQ_CHUNK_BASES, emit={"paired": paired, "compress": device_zstd})
        while True:
            if device_zstd:
                while len(pending) > 1:          # the pinned result buffers alternate: the bytes of call k live until call k + 2
                    pending.popleft().result()
            try:
                _layout, _pipe, _none, (r1, r2, _info) = next(gen)
            except StopIteration:
                break
            if state is not None:
                state["wrote"] = True
            if device_zstd:
                pending.append(pool.submit(append, r1, r2))
            else:
                with _timed("zstd_submit"):
                    if r1 is not None:
                        w1.write(r1)             # read-only views: the writer copies the pieces it queues
                    if r2 is not None and w2 is not None:
                        w2.write(r2)
    except _ffi.PclError as e:
        if "unwrap" in str(e):
            raise RuntimeError(str(e)) from e
        raise
    finally:
        with _timed("zstd_drain"):
            if device_zstd:
                try:
                    while pending:
                        pending.popleft().result()
                finally:
                    pool.shutdown(wait=True)
                    with _timed("file_close"):
                        f1.close()
                        if f2 is not None:
                            f2.close()
            else:
                w1.close()
                if w2 is not None:
                    w2.close()
