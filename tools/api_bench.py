"""End to end through the API the reference exposes: FASTQ files (plain / gzip / BGZF) -> make_fastq -> R1.fq.zst, R2.fq.zst.

Bench tooling (bench.py `api_e2e`).  The workload is the 10x scATAC shape of BASELINE configs[1] (I2 = 16 nt cell barcode on
the neg strand, R1 / R2 = 50 nt gDNA, 737 K whitelist) because make_fastq needs a pos-strand insert: on 10x v3 (configs[0])
`record.read1.unwrap()` panics in the reference (python/src/lib.rs:160) and raises here.

  ours        precellar_b200.api.make_fastq (python/src/lib.rs:117-171), device writer: FASTQ text -> H2D -> parse on the GPU
              -> count -> all-reduce -> correct -> records assembled in HBM -> D2H -> zstd (level 9, host threads), or with
              compression="device" zstd frames built on the GPU -> D2H -> file
  reference   the oracle's file-level driver: FASTQ parse (one thread per file) -> serial count pass -> parallel annotate
              -> the record loop of make_fastq -> zstd (level 9, 8 workers)       [the CPU restatement, kind "port"]
Both arms write the same bytes (checked).
"""
from __future__ import annotations

import hashlib
import os
import struct
import time
import zlib
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import yaml

from tools import workloads as W


def _region(rid, rtype, stype, seq, lo, hi, onlist=None):
    d = dict(region_id=rid, region_type=rtype, name=rid, sequence_type=stype, sequence=seq, min_len=lo, max_len=hi, onlist=None, regions=None)
    if onlist:
        d["onlist"] = dict(file_id=os.path.basename(onlist), filename=os.path.basename(onlist), filetype="txt", filesize=0, url=onlist,
                           urltype="local", md5=0, location="local")
    return d


def write_atac_spec(td: str, whitelist: np.ndarray) -> str:
    wl_path = os.path.join(td, "atac_wl.txt")
    with open(wl_path, "wb") as fh:
        fh.write(np.concatenate([whitelist, np.full((len(whitelist), 1), 10, np.uint8)], axis=1).tobytes())
    regions = [
        _region("p5", "illumina_p5", "fixed", "AATGATACGGCGACCACCGAGATCTACAC", 29, 29),
        _region("cell_bc", "barcode", "onlist", "N" * 16, 16, 16, wl_path),
        _region("r1p", "nextera_read1", "fixed", "TCGTCGGCAGCGTCAGATGTGTATAAGAGACAG", 33, 33),
        _region("gdna", "gdna", "random", "X", 1, 1000),
        _region("r2p", "nextera_read2", "fixed", "CTGTCTCTTATACACATCTCCGAGCCCACGAGAC", 34, 34),
        _region("p7", "illumina_p7", "fixed", "ATCTCGTATGCCGTCTTCTGCTTG", 24, 24),
    ]
    spec = dict(seqspec_version="0.3.0", assay_id="atac-bench", name="x", doi="", date="", description="", modalities=["atac"],
                lib_struct="", library_protocol="x", library_kit="x", sequence_protocol="x", sequence_kit="x", sequence_spec=[],
                library_spec=[dict(region_id="atac", region_type="atac", name="atac", sequence_type="joined", sequence="",
                                   min_len=0, max_len=0, onlist=None, regions=regions)])
    path = os.path.join(td, "atac.yaml")
    yaml.safe_dump(spec, open(path, "w"), sort_keys=False)
    return path


def write_fastq(path: str, seq: np.ndarray, qual: np.ndarray, L: int, tag: str) -> None:
    n = len(seq)
    nl = np.full((n, 1), 10, np.uint8)
    names = np.char.add(np.char.add("@r", np.arange(n).astype(str)), tag).astype("S")
    w = names.dtype.itemsize
    name_arr = np.frombuffer(names.tobytes(), np.uint8).reshape(n, w).copy()
    name_arr[name_arr == 0] = ord(" ")                       # pad with blanks: part of the description
    plus = np.full((n, 1), ord("+"), np.uint8)
    rec = np.concatenate([name_arr, nl, seq[:, :L], nl, plus, nl, qual[:, :L], nl], axis=1)
    with open(path, "wb") as fh:
        fh.write(rec.tobytes())


def gzip_file(src: str, dst: str, bgzf: bool) -> None:
    """Level-1 gzip: one member, or BGZF (independent <= 64 KB members with their sizes in the header)."""
    with open(src, "rb") as fi, open(dst, "wb") as fo:
        if not bgzf:
            co = zlib.compressobj(1, zlib.DEFLATED, 31)
            while True:
                c = fi.read(1 << 24)
                if not c:
                    break
                fo.write(co.compress(c))
            fo.write(co.flush())
            return
        while True:
            c = fi.read(65280)
            co = zlib.compressobj(1, zlib.DEFLATED, -15)
            body = co.compress(c) + co.flush()
            fo.write(b"\x1f\x8b\x08\x04\0\0\0\0\x00\xff" + struct.pack("<H", 6) + b"BC" + struct.pack("<HH", 2, 18 + len(body) + 8 - 1))
            fo.write(body + struct.pack("<II", zlib.crc32(c) & 0xFFFFFFFF, len(c)))
            if not c:
                break


def make_inputs(td: str, n: int):
    """FASTQ files of n read triples of the ATAC workload in three encodings; returns (workload, spec path, {fmt: {read_id: path}})."""
    w = W.atac()
    full = w.host_batches(0, n)
    rng = np.random.default_rng(1)
    ins = rng.choice(np.frombuffer(b"ACGT", np.uint8), (n, 50))
    insq = np.full((n, 50), ord("F"), np.uint8)
    plain = {"atac-I2": os.path.join(td, "I2.fq"), "atac-R1": os.path.join(td, "R1.fq"), "atac-R2": os.path.join(td, "R2.fq")}
    write_fastq(plain["atac-I2"], full[1].seq, full[1].qual, 16, "/1")
    write_fastq(plain["atac-R1"], ins, insq, 50, "/1")
    write_fastq(plain["atac-R2"], ins[:, ::-1].copy(), insq, 50, "/2")
    files = {"plain": plain, "gzip": {k: v + ".gz" for k, v in plain.items()}, "bgzf": {k: v + ".bgz" for k, v in plain.items()}}
    with ThreadPoolExecutor(6) as ex:
        jobs = [ex.submit(gzip_file, v, files["gzip"][k], False) for k, v in plain.items()]
        jobs += [ex.submit(gzip_file, v, files["bgzf"][k], True) for k, v in plain.items()]
        for j in jobs:
            j.result()
    spec = write_atac_spec(td, w.layout.whitelists[w.layout.region_ids[0]])
    return w, spec, files


def assay_for(spec: str, paths: dict):
    from precellar_b200 import Assay
    a = Assay(spec)
    a.add_illumina_reads("atac", forward_strand_workflow=False)
    for rid in ("atac-R1", "atac-R2", "atac-I2"):
        a.update_read(rid, fastq=paths[rid])                 # infers min_len / max_len from the first records
    return a


def _digest(path: str) -> str:
    from precellar_b200 import _zstd
    return hashlib.sha256(_zstd.decompress_file(path)).hexdigest()


def run_ours(spec: str, paths: dict, out_dir: str, devices=None, writer: str = "auto", compression: str = "zstd") -> float:
    from precellar_b200.api import make_fastq
    t0 = time.perf_counter()
    make_fastq(assay_for(spec, paths), modality="atac", out_dir=out_dir, correct_barcode=True, devices=devices, writer=writer,
               compression=compression)
    return time.perf_counter() - t0


def run_reference(w, spec: str, paths: dict, out_dir: str, n_threads: int):
    """The reference algorithm end to end on the host (oracle port).  Returns (seconds, records written, zstd workers)."""
    from oracle import pyoracle as O
    from precellar_b200 import compile_layout
    os.makedirs(out_dir, exist_ok=True)
    t0 = time.perf_counter()
    layout = compile_layout(assay_for(spec, paths), "atac", require_files=True)      # the same YAML / update_read work as the API arm
    wl = layout.whitelists[layout.region_ids[0]]
    items = (wl.reshape(-1), np.arange(len(wl) + 1, dtype=np.uint64) * np.uint64(wl.shape[1]), len(wl)) if isinstance(wl, np.ndarray) else list(wl)
    orc = O.Oracle(O.reads_from_layout(layout), {0: items})
    with ThreadPoolExecutor(len(layout.reads)) as ex:                                  # one decoding thread per file
        jobs = [ex.submit(orc.fastq_load, r, list(p.files), 64) for r, p in enumerate(layout.reads)]
        ns = [j.result() for j in jobs]
    assert len(set(ns)) == 1
    orc.n = ns[0]
    orc.count()                                                                        # BarcodeAnalyzer::new, serial
    orc.annotate(correct=True, threshold=0.975, n_threads=n_threads, chunk_bases=1_000_000, outputs=False, keep=True)
    written, workers = orc.make_fastq(True, os.path.join(out_dir, "R1.fq.zst"), os.path.join(out_dir, "R2.fq.zst"), 9, 8)
    dt = time.perf_counter() - t0
    orc.close()
    return dt, written, workers


def api_e2e(td: str, n: int, host_cores: int, devices=None, formats=("plain", "gzip", "bgzf")) -> dict:
    """Per input format: make_fastq as called by default (device writer when it applies, libzstd level 9 like the
    reference), make_fastq(compression="device") (zstd frames built on the GPU) and the CPU port, same work each."""
    w, spec, files = make_inputs(td, n)
    out = {"workload": f"{w.name}: {n} read triples, FASTQ files -> make_fastq(correct_barcode=True) -> R1.fq.zst + R2.fq.zst",
           "unit": "read-triples/s", "formats": {},
           "note": "wall clock around Assay + update_read x 3 + make_fastq; best of two runs after one warm-up; outputs compared "
                   "after decompression with libzstd; 'device_zstd' = make_fastq(compression='device'): frames of Huffman-coded "
                   "literals built by k_zh_blocks (larger files, no host encoder)"}
    names = ("R1.fq.zst", "R2.fq.zst")
    for fmt in formats:
        o_ours, o_dev, o_ref = os.path.join(td, f"ours_{fmt}"), os.path.join(td, f"dev_{fmt}"), os.path.join(td, f"ref_{fmt}")
        run_ours(spec, files[fmt], o_ours, devices)                                    # warm-up: CUDA context, page cache
        t_ours = min(run_ours(spec, files[fmt], o_ours, devices) for _ in range(2))
        t_ref, written, workers = run_reference(w, spec, files[fmt], o_ref, host_cores)
        ref_digest = [_digest(os.path.join(o_ref, f)) for f in names]
        same = [_digest(os.path.join(o_ours, f)) for f in names] == ref_digest
        row = {"ours_s": t_ours, "ours_value": n / t_ours, "reference_s": t_ref, "reference_value": n / t_ref,
               "speedup": t_ref / t_ours, "records_written": written, "outputs_identical": bool(same),
               "input_bytes": int(sum(os.path.getsize(p) for p in files[fmt].values())),
               "out_bytes": [os.path.getsize(os.path.join(o_ours, f)) for f in names],
               "reference_out_bytes": [os.path.getsize(os.path.join(o_ref, f)) for f in names],
               "reference_zstd_workers": workers, "reference_threads": host_cores}
        if devices is None or len(devices) == 1:
            run_ours(spec, files[fmt], o_dev, devices, compression="device")
            t_dev = min(run_ours(spec, files[fmt], o_dev, devices, compression="device") for _ in range(2))
            row["device_zstd"] = {"ours_s": t_dev, "ours_value": n / t_dev, "speedup": t_ref / t_dev,
                                  "out_bytes": [os.path.getsize(os.path.join(o_dev, f)) for f in names],
                                  "outputs_identical": bool([_digest(os.path.join(o_dev, f)) for f in names] == ref_digest)}
        out["formats"][fmt] = row
    return out
