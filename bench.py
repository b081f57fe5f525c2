#!/usr/bin/env python
"""Benchmark of the barcode hot path (BASELINE.json metric: barcode-corrected read-pairs/s).

    python bench.py --gpus N --steps K --warmup W            # our CUDA path (one rank per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host cores

Workload: 10x scRNA-seq v3 (16 nt CB + 12 nt UMI on R1, 91 nt cDNA on R2), 6.8 M-entry synthetic whitelist,
125 M synthetic read pairs per GPU (= the 1 B-pair / 8 GPU configuration of BASELINE.json, weak scaling),
barcodes Zipf-sampled over 10 000 cells with injected substitution errors (tools/synth).

A step = one pass of the whole path over the rank's shard:
    reset counts -> pass 1 (split + pack + whitelist probe + histogram) -> allreduce of the counts
    -> pass 2 (1-Hamming correction with the fp64 posterior).
`value` times it with the inputs resident in HBM; `e2e` times the same through the public API from pinned
host buffers (H2D of every input plane and D2H of every result plane inside the timed region).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_WHITELIST = 6_800_000
R2_LEN = 91
THRESHOLD = 0.975          # make_fastq default (precellar/src/align/fastq.rs:39)
# Bytes of the IMPLEMENTED algorithm per unit of work (DESIGN.md section 3 has the derivation; every kernel's `frac` in
# the JSON line is these bytes / its event-timed duration / the measured copy bandwidth):
#   count kernel of a read with barcodes, per record: the payload bytes the layout can touch + 2 (length) + one 32-byte
#       probe sector per barcode + the result words (fstatus 2, insert coordinates 4, keys 4/8 each, segment coordinates
#       4 each for variable layouts) + 12 per missed barcode (worklist entry + packed key)
#   count kernel of an insert read (lengths only), per record: 2 + 2 + 4
#   k_filter, per missed barcode: 12 in + 8 out + four partial-key lookups of one sector each (answered by the bitmap
#       word or by the neighbourhood bucket): 148
#   k_resolve, per missed barcode: 20 (entry, key, candidates) + one sector each of qualities, main table and counts
#       + 6 of result words: 122
B_FILTER_PER_MISS = 12 + 8 + 4 * 32
B_RESOLVE_PER_MISS = 20 + 3 * 32 + 6
B_WORKLIST_PER_MISS = 12


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


class ClockSampler:
    """SM clock / throttle reasons of one GPU sampled during a timed region (NVML)."""

    def __init__(self, index: int, period: float = 0.02):
        self.index, self.period = index, period
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
                 0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
                 0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, nm in names.items():
                    if r & bit and nm != "gpu_idle":
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(self.period)

    def start(self):
        if self.nv:
            self._stop.clear()
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()

    def stop(self):
        if self._thr:
            self._stop.set()
            self._thr.join()
            self._thr = None

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def pinned(nbytes: int, dtype, shape):
    from precellar_b200 import _ffi
    p = _ffi.lib().pcl_host_alloc(int(nbytes))
    if not p:
        raise MemoryError(f"pinned allocation of {nbytes} bytes failed")
    buf = (C.c_uint8 * int(nbytes)).from_address(p)
    arr = np.frombuffer(buf, dtype=dtype).reshape(shape)
    return arr, p


def count_bytes_per_record(pipe, r: int) -> float:
    """Implemented-algorithm bytes of pass 1 for one record of read r, without the worklist (see the table above)."""
    info = pipe.infos[r]
    if not info.needs_payload:
        return 2 + 2 + (4 if info.has_target else 0)
    b = pipe.payload_bytes(r) + 2 + 32 * info.n_bc + 2 + (4 if info.has_target else 0)
    b += sum(info.bc_key_bytes[j] for j in range(info.n_bc)) + sum(info.umi_key_bytes[j] for j in range(info.n_umi))
    if not info.fixed_layout:
        b += 4 * (info.n_bc + info.n_umi)
    return b


def measure_device_ingest(pipe, p1: str, p2: str, n_pairs: int) -> dict:
    """FASTQ text -> chunk planes parsed on the GPU (pcl_text_*): from pinned host memory (PCIe + parse kernels) and
    from the plain files on disk through textingest.TextGroupReader (adds the host's file reads)."""
    from precellar_b200 import _ffi
    from precellar_b200.textingest import TextGroupReader
    L = _ffi.lib()
    texts, ptrs = [], []
    for p in (p1, p2):
        sz = os.path.getsize(p)
        arr, ptr = pinned(sz, np.uint8, (sz,))
        with open(p, "rb", buffering=0) as fh:
            got = 0
            mv = memoryview(arr)
            while got < sz:
                got += fh.readinto(mv[got:])
        texts.append(arr)
        ptrs.append(ptr)
    ch = pipe.new_chunk(n_pairs)
    out = {}
    try:
        best = None
        for it in range(4):
            if it == 1:
                L.pcl_profile_reset(pipe.ctx)
                L.pcl_profile_enable(pipe.ctx, 1)
            t0 = time.perf_counter()
            for r in range(2):
                ch.text_scan(r, texts[r])
            n = min(ch.text_lines(r) // 4 for r in range(2))
            ch.reset(n)
            used = [ch.text_fill(r) for r in range(2)]
            bad = ch.text_names_check()
            dt = time.perf_counter() - t0
            assert n == n_pairs and bad == -1 and used == [len(t) for t in texts]
            if it >= 1:
                best = dt if best is None else min(best, dt)
        ms_i, ms_f, k = C.c_double(), C.c_double(), C.c_uint64()
        L.pcl_profile_get(pipe.ctx, _ffi.K_TEXT_INDEX, C.byref(ms_i), C.byref(k))
        L.pcl_profile_get(pipe.ctx, _ffi.K_TEXT_FILL, C.byref(ms_f), C.byref(k))
        L.pcl_profile_enable(pipe.ctx, 0)
        L.pcl_profile_reset(pipe.ctx)
        nbytes = sum(len(t) for t in texts)
        out["text_bytes_per_pair"] = nbytes / n_pairs
        out["pinned_text_pairs_per_s"] = n_pairs / best
        out["pinned_text_GBps"] = nbytes / best / 1e9
        out["kernels_ms_per_call"] = {"index": ms_i.value / 3, "fill_and_names": ms_f.value / 3}
        out["kernels_text_GBps"] = nbytes / ((ms_i.value + ms_f.value) / 3 * 1e-3) / 1e9
    finally:
        ch.close()
        pipe.chunks.remove(ch)
        for ptr in ptrs:
            L.pcl_host_free(ptr)
    per = 1 << 19
    rdr = TextGroupReader(pipe, [[p1], [p2]], per, [80, 210])
    chs = [pipe.new_chunk(per) for _ in range(2)]
    try:
        t0 = time.perf_counter()
        got = k = 0
        while True:
            n = rdr.next_into(chs[k & 1])
            if n == 0:
                break
            got += n
            k += 1
        dt = time.perf_counter() - t0
        assert got == n_pairs
        out["plain_files_pairs_per_s"] = n_pairs / dt
    finally:
        rdr.close()
        for c in chs:
            c.close()
            pipe.chunks.remove(c)
    # the same through gzip: ordinary single-member files (one zlib stream per file) and BGZF (blocks inflated in parallel)
    import struct
    import zlib

    def bgzf(src, dst, block=65280):
        with open(src, "rb") as fi, open(dst, "wb") as fo:
            while True:
                c = fi.read(block)
                co = zlib.compressobj(1, zlib.DEFLATED, -15)
                body = co.compress(c) + co.flush()
                fo.write(b"\x1f\x8b\x08\x04\0\0\0\0\x00\xff" + struct.pack("<H", 6) + b"BC" + struct.pack("<HH", 2, 18 + len(body) + 8 - 1))
                fo.write(body + struct.pack("<II", zlib.crc32(c) & 0xFFFFFFFF, len(c)))
                if not c:
                    break
    variants = []
    if os.path.exists(p1 + ".gz") and os.path.exists(p2 + ".gz"):
        variants.append(("gzip_files_pairs_per_s", p1 + ".gz", p2 + ".gz"))
    bgzf(p1, p1 + ".bgz")
    bgzf(p2, p2 + ".bgz")
    # BGZF twice: members inflated on the device (k_bgzf_inflate, the default) and by the host threads (PCL_DEVICE_INFLATE=0)
    variants.append(("bgzf_files_pairs_per_s", p1 + ".bgz", p2 + ".bgz"))
    variants.append(("bgzf_files_host_inflate_pairs_per_s", p1 + ".bgz", p2 + ".bgz"))
    from precellar_b200 import textingest
    for key, a, b in variants:
        host_inflate = key == "bgzf_files_host_inflate_pairs_per_s"
        saved = os.environ.get("PCL_DEVICE_INFLATE")
        if host_inflate:
            os.environ["PCL_DEVICE_INFLATE"] = "0"
        for kk in textingest.INFLATED_STATS:
            textingest.INFLATED_STATS[kk] = 0
        try:
            rdr = TextGroupReader(pipe, [[a], [b]], per, [80, 210])
        finally:
            if host_inflate:
                os.environ.pop("PCL_DEVICE_INFLATE", None)
                if saved is not None:
                    os.environ["PCL_DEVICE_INFLATE"] = saved
        chs = [pipe.new_chunk(per) for _ in range(2)]
        try:
            t0 = time.perf_counter()
            got = k = 0
            while True:
                n = rdr.next_into(chs[k & 1])
                if n == 0:
                    break
                got += n
                k += 1
            dt = time.perf_counter() - t0
            assert got == n_pairs
            out[key] = n_pairs / dt
        finally:
            rdr.close()
            for c in chs:
                c.close()
                pipe.chunks.remove(c)
        st = textingest.INFLATED_STATS
        if key == "bgzf_files_pairs_per_s" and st["launches"]:
            out["bgzf_device_inflate"] = {"kernel": "k_bgzf_inflate", "launches": int(st["launches"]), "text_bytes": int(st["out_bytes"]),
                                          "kernel_ms_sum": st["kernel_ms"], "text_GBps_per_launch_stream": st["out_bytes"] / st["kernel_ms"] / 1e6,
                                          "note": "one warp per <= 64 KB member; latency-bound (a launch lasts as long as one member takes its warp), "
                                                  "so a launch holds up to 256 MB of text - a whole file here; the files' launches overlap on their own streams"}
    out["inflate_threads_per_file"] = max(1, min(8, (os.cpu_count() or 2) // 2))
    out["note"] = ("raw text -> device; line index and SoA planes (R1 32 B rows + 16 B quality window + lengths, R2 lengths, "
                   "name hashes) built by k_text_* kernels; host cores only move bytes")
    return out


def measure_ingest(syn, n_pairs: int, tmpdir: str, pipe=None) -> dict:
    """Host ingest reported separately from the device path: FASTQ text (plain and gzip) -> pinned-ready SoA planes
    through the C++ reader (csrc/ingest.cpp), one decoding thread per file like the reference's readers."""
    import gzip
    import shutil
    from precellar_b200.ingest import GroupReader
    seq, qual, length = syn.host_arrays(0, n_pairs)
    rng = np.random.default_rng(1)
    ins = rng.choice(np.frombuffer(b"ACGT", np.uint8), (n_pairs, R2_LEN))
    insq = np.full((n_pairs, R2_LEN), ord("F"), np.uint8)
    def fq(path, s, q, L, tag):
        nl = np.full((n_pairs, 1), 10, np.uint8)
        names = np.char.add(np.char.add("@r", np.arange(n_pairs).astype(str)), tag).astype("S")
        w = names.dtype.itemsize
        name_arr = np.frombuffer(names.tobytes(), np.uint8).reshape(n_pairs, w).copy()
        name_arr[name_arr == 0] = ord(" ")              # pad with blanks: part of the description, ignored by the name check
        plus = np.full((n_pairs, 1), ord("+"), np.uint8)
        rec = np.concatenate([name_arr, nl, s[:, :L], nl, plus, nl, q[:, :L], nl], axis=1)
        with open(path, "wb") as fh:
            fh.write(rec.tobytes())
    p1, p2 = os.path.join(tmpdir, "R1.fq"), os.path.join(tmpdir, "R2.fq")
    fq(p1, seq, qual, 28, "/1")
    fq(p2, ins, insq, R2_LEN, "/2")
    out = {"pairs": n_pairs, "threads": 2, "note": "FASTQ text -> SoA planes (R1 32 B rows + lengths, R2 lengths), one decoder thread per file"}
    for kind in ("plain", "gzip"):
        a, b = p1, p2
        if kind == "gzip":
            for p in (p1, p2):
                with open(p, "rb") as fi, gzip.open(p + ".gz", "wb", compresslevel=1) as fo:
                    shutil.copyfileobj(fi, fo, 1 << 24)
            a, b = p1 + ".gz", p2 + ".gz"
        rd = GroupReader([[a], [b]], [32, 0], [False, False])
        t0 = time.perf_counter()
        got = 0
        while True:
            g = rd.next(1 << 20)
            if g is None:
                break
            got += len(g[0].batch.length)
        dt = time.perf_counter() - t0
        rd.close()
        assert got == n_pairs
        out[f"{kind}_fastq_pairs_per_s"] = n_pairs / dt
    if pipe is not None:
        out["device_parse"] = measure_device_ingest(pipe, p1, p2, n_pairs)
    return out


# ------------------------------------------------------------------------------------------------------------------
# device-resident measurement of one or several pipelines that live on the GPU at the same time
# ------------------------------------------------------------------------------------------------------------------
def load_traffic():
    """DRAM bytes per unit of work of every kernel, from the committed ncu capture of this build (profiles/ncu_traffic.json;
    `ncu --set full`, dram__bytes_read.sum + dram__bytes_write.sum).  Not measurable inside an un-profiled run."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        return json.load(open(path))
    except Exception:
        return {}


def kernel_table(pipe, w, n, prof, steps, mis_per_region, peak, traffic_key):
    """Per-kernel {ms, implemented-algorithm bytes, GB/s, frac of the measured copy peak, ncu DRAM traffic} of one pipeline."""
    from precellar_b200 import _ffi
    infos = pipe.infos
    n_miss = float(sum(mis_per_region))
    bytes_count = sum(count_bytes_per_record(pipe, r) for r in range(len(infos)) if infos[r].needs_payload) * n + B_WORKLIST_PER_MISS * n_miss
    bytes_len = sum(count_bytes_per_record(pipe, r) for r in range(len(infos)) if not infos[r].needs_payload) * n
    t = lambda k: prof[k][0] / steps                  # all launches of that kind in one step (e.g. both insert reads of a triple)
    rows = {
        "count": {"ms": t(_ffi.K_COUNT), "alg_bytes": bytes_count, "unit_bytes": bytes_count / n, "per": "read group"},
        "count_lenonly": {"ms": t(_ffi.K_COUNT_LENONLY), "alg_bytes": bytes_len, "unit_bytes": bytes_len / n, "per": "read group"},
        "filter": {"ms": t(_ffi.K_FILTER), "alg_bytes": B_FILTER_PER_MISS * n_miss, "unit_bytes": B_FILTER_PER_MISS, "per": "missed barcode"},
        "resolve": {"ms": t(_ffi.K_CORRECT), "alg_bytes": B_RESOLVE_PER_MISS * n_miss, "unit_bytes": B_RESOLVE_PER_MISS, "per": "missed barcode"},
    }
    tr = load_traffic().get(traffic_key, {})
    out = {}
    for name, kv in rows.items():
        if kv["ms"] <= 0:
            continue
        gbs = kv["alg_bytes"] / (kv["ms"] / 1e3) / 1e9
        units = n if kv["per"] == "read group" else n_miss
        tb = tr.get(name)
        traffic = (float(tb) * units) if tb is not None else None
        out[name] = {"ms_per_launch": kv["ms"], "alg_bytes_per_launch": kv["alg_bytes"], "alg_bytes_per_unit": kv["unit_bytes"],
                     "unit": kv["per"], "achieved_gbs": gbs, "frac": gbs / peak, "traffic": traffic,
                     # the same fraction on the DRAM bytes ncu counted: what the memory system really moved per second
                     # (below `frac` where table sectors hit in L2, above it where sectors are fetched more than once)
                     "frac_traffic": (traffic / (kv["ms"] / 1e3) / 1e9 / peak) if traffic is not None else None}
    return out


class Resident:
    """One workload resident on one GPU: pipeline + one chunk of n read groups generated in place."""

    def __init__(self, w, n, device, world=1, rank=0, uid=None):
        from precellar_b200.pipeline import BarcodePipeline
        t0 = time.perf_counter()
        self.w, self.n = w, n
        self.pipe = BarcodePipeline(w.layout, device=device, nranks=world, rank=rank, nccl_uid=uid)
        self.setup_s = time.perf_counter() - t0
        self.chunk = self.pipe.new_chunk(n)
        self.chunk.reset(n)
        w.fill_chunk(self.pipe, self.chunk, rank * n, n, device)
        self.pipe.sync()

    def step(self):
        p = self.pipe
        p.reset_counts()
        p.count(self.chunk)
        p.allreduce()
        p.correct(self.chunk, THRESHOLD)

    def close(self):
        self.pipe.close()


def measure_resident(res_list, steps, warmup, barrier, max_over_ranks, clocks=None, sustained_s=0.0):
    """Time `steps` steps of all pipelines in res_list together (CUDA events on the first pipeline's chunk stream when
    there is one pipeline; host clock between device syncs when several share the GPU)."""
    from precellar_b200 import _ffi
    for _ in range(max(warmup, 3)):
        for r in res_list:
            r.step()
    for r in res_list:
        r.pipe.sync()
        r.pipe.profile_enable(True)
        r.pipe.profile_reset()
    l0 = [r.pipe.launch_count() for r in res_list]
    barrier()
    if clocks:
        clocks.start()
    if len(res_list) == 1:
        res_list[0].chunk.timer_start()
        for _ in range(steps):
            res_list[0].step()
        ms_total = res_list[0].chunk.timer_stop_ms()
    else:
        t0 = time.perf_counter()
        for _ in range(steps):
            for r in res_list:
                r.step()
        for r in res_list:
            r.pipe.sync()
        ms_total = (time.perf_counter() - t0) * 1e3
    barrier()
    launches = sum(r.pipe.launch_count() - a for r, a in zip(res_list, l0))
    prof = [{k: r.pipe.profile_get(k) for k in (_ffi.K_COUNT, _ffi.K_CORRECT, _ffi.K_COUNT_LENONLY, _ffi.K_FILTER)} for r in res_list]
    for r in res_list:
        r.pipe.profile_enable(False)
    ms_step = max_over_ranks(ms_total) / steps
    sustained = None
    if sustained_s > 0:                      # a longer region (the contract's K steps are a few tens of milliseconds)
        k = max(steps, int(sustained_s * 1e3 / max(ms_step, 1e-3)))
        barrier()
        if len(res_list) == 1:
            res_list[0].chunk.timer_start()
            for _ in range(k):
                res_list[0].step()
            ms = res_list[0].chunk.timer_stop_ms()
        else:
            t0 = time.perf_counter()
            for _ in range(k):
                for r in res_list:
                    r.step()
            for r in res_list:
                r.pipe.sync()
            ms = (time.perf_counter() - t0) * 1e3
        barrier()
        sustained = {"steps": k, "ms_per_step": max_over_ranks(ms) / k}
    if clocks:
        clocks.stop()
    return ms_step, prof, launches, sustained


def parity_prefix(w, pipe, ns, device, world, rank, host_cores, gather):
    """The first `ns` read groups of the workload, sharded over the ranks like the real job: count -> all-reduce ->
    correct on every rank's shard, results gathered to rank 0 in rank order and compared field by field with the oracle
    (the reference algorithm on the host) run on the whole prefix.  Returns (parity dict, cpu seconds) on rank 0."""
    from precellar_b200.dist import shard_bounds
    from tools.v3check import compare_bulk
    lo, hi = shard_bounds(ns, world)[rank]
    full = w.host_batches(lo, hi - lo)                 # the generator is a function of the record index: same bytes as resident
    dev = w.device_batches(pipe, full)
    ch = pipe.new_chunk(max(hi - lo, 1))
    pipe.reset_counts()
    ch.reset(hi - lo)
    for r, b in enumerate(dev):
        ch.upload(r, b)
    pipe.count(ch)
    pipe.allreduce()
    pipe.correct(ch, THRESHOLD)
    mine = [ch.fetch(r) for r in range(len(dev))]
    counts = {g: pipe.counts(g) for g in range(len(w.layout.region_ids))}
    defects = pipe.qc()[:, 1].copy()
    ch.close()
    pipe.chunks.remove(ch)
    parts = gather(dict(res=[dict(fstatus=m.fstatus, tcoord=m.tcoord, bc_key=m.bc_key, umi_key=m.umi_key, bcoord=m.bcoord, ucoord=m.ucoord)
                             for m in mine], defects=defects))
    if rank != 0:
        return None, None
    cat = lambda xs: None if xs[0] is None else np.concatenate(xs)
    results = []
    for r in range(len(dev)):
        res = pipe.alloc_results(r, 0)
        ps = [p["res"][r] for p in parts]
        res.fstatus, res.tcoord = cat([p["fstatus"] for p in ps]), cat([p["tcoord"] for p in ps])
        res.bc_key = [cat([p["bc_key"][j] for p in ps]) for j in range(len(ps[0]["bc_key"]))]
        res.umi_key = [cat([p["umi_key"][j] for p in ps]) for j in range(len(ps[0]["umi_key"]))]
        res.bcoord = [cat([p["bcoord"][j] for p in ps]) for j in range(len(ps[0]["bcoord"]))]
        res.ucoord = [cat([p["ucoord"][j] for p in ps]) for j in range(len(ps[0]["ucoord"]))]
        results.append(res)
    from tools.workloads import run_oracle
    whole = w.host_batches(0, ns) if world > 1 else full
    cpu_s, ores, ocounts = run_oracle(w, whole, THRESHOLD, host_cores)
    fields = compare_bulk(w.layout, pipe.infos, results, counts, ores, ocounts)
    fields["defect_counts"] = int((sum(p["defects"] for p in parts) != ores.qc_per_file[:, 1]).sum())
    return {"checked_read_groups": int(ns), "ranks": world, "mismatches": int(sum(fields.values())), "fields": fields}, cpu_s


def main():
    # stdout carries exactly one JSON line: libraries that print to fd 1 (e.g. NCCL's version banner) go to stderr
    json_fd = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(os.dup(2), "w")

    def emit(obj):
        os.write(json_fd, (json.dumps(obj) + "\n").encode())

    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pairs-per-gpu", type=int, default=125_000_000)
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--e2e-chunks", type=int, default=8)
    ap.add_argument("--cpu-sample", type=int, default=2_500_000)
    ap.add_argument("--sustained-s", type=float, default=1.0)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true", help="skip the oracle legs (parity, cpu_baseline) and the host ingest measurements")
    ap.add_argument("--no-other", action="store_true", help="skip the other BASELINE configurations")
    ap.add_argument("--other-n", type=int, default=0, help="read groups per other configuration (default: BASELINE sizes)")
    ap.add_argument("--ingest-sample", type=int, default=1_000_000)
    ap.add_argument("--api-sample", type=int, default=2_000_000, help="read triples of the API end-to-end leg (0: skip)")
    args = ap.parse_args()

    rank, local_rank, world = env_int("RANK", 0), env_int("LOCAL_RANK", 0), env_int("WORLD_SIZE", 1)
    n = args.pairs_per_gpu
    host_cores = os.cpu_count() or 1
    workload = (f"10x scRNA-seq v3 (16nt CB + 12nt UMI, R2 {R2_LEN}nt), {N_WHITELIST} whitelist, "
                f"{n} synthetic read pairs per GPU (1B-pair / 8-GPU config)")

    from tools import workloads as W

    # ------------------------------------------------------------------ reference arm (host cores only)
    if args.impl == "reference":
        if rank != 0:
            return 0
        w = W.v3(N_WHITELIST, R2_LEN)
        # bounded sample: calibrate on 200k pairs, then size the per-step sample so the whole run takes ~2 minutes
        nc = min(200_000, n)
        tcal, _, _ = W.run_oracle(w, w.host_batches(0, nc), THRESHOLD, host_cores, outputs=False)
        rate = nc / tcal
        total_steps = args.steps + args.warmup
        ns = int(max(50_000, min(args.cpu_sample, n, 110.0 * rate / max(total_steps, 1))))
        full = w.host_batches(0, ns)
        times = []
        for s in range(total_steps):
            dt, _, _ = W.run_oracle(w, full, THRESHOLD, host_cores, outputs=False)
            if s >= args.warmup:
                times.append(dt)
        ms = 1e3 * float(np.mean(times))
        v = ns / (ms / 1e3)
        cfg = {"workload": workload, "sample_pairs_per_step": ns, "l2": "inputs >> L2 (host arm)"}
        emit({
            "impl": "reference", "metric": "barcode-corrected read-pairs/sec", "value": v, "unit": "read-pairs/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/f64", "data": "synthetic",
            "config": cfg,
            "cpu_baseline": {"value": v, "unit": "read-pairs/s", "cores": host_cores, "kind": "port",
                             "sample": f"{ns} read pairs per step: serial count pass + {host_cores}-thread correct pass "
                                       "(C++ restatement of the reference algorithm, in-memory SoA input)"},
            "e2e": {"value": v, "unit": "read-pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0})
        return 0

    # ------------------------------------------------------------------ our arm
    import pickle
    import torch
    import torch.distributed as dist
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def gather(obj):
        if world == 1:
            return [obj]
        out = [None] * world
        dist.all_gather_object(out, pickle.dumps(obj))
        return [pickle.loads(b) for b in out]

    from precellar_b200 import _ffi
    from precellar_b200.pipeline import BarcodePipeline, HostBatch
    from precellar_b200.dist import exchange_unique_id

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak = float(json.load(open(peaks_path))["hbm_gbs"])
        peak_src = "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
    else:
        peak, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"

    uid = exchange_unique_id(BarcodePipeline.new_unique_id, rank, world)
    w3 = W.v3(N_WHITELIST, R2_LEN)
    R = Resident(w3, n, local_rank, world, rank, uid)
    pipe, big, syn = R.pipe, R.chunk, w3.gen
    t_setup = R.setup_s                               # ctx + layout + whitelist tables (host build and upload) + NCCL init
    first = rank * n

    clocks = ClockSampler(local_rank)
    ms_step, prof, launches, sustained = measure_resident([R], args.steps, args.warmup, barrier, max_over_ranks, clocks, args.sustained_s)
    value = n * world / (ms_step / 1e3)
    cnt, mismatch, total = pipe.counts(0)
    assert total == n * world, (total, n, world)                 # the allreduce summed every rank's shard
    assert int(cnt.sum()) + mismatch == total
    f_mis_global = mismatch / total
    kern = kernel_table(pipe, w3, n, prof[0], args.steps, [f_mis_global * n], peak, "v3")
    dom = max(kern, key=lambda k: kern[k]["ms_per_launch"])
    step_alg_bytes = sum(v["alg_bytes_per_launch"] for v in kern.values())
    roofline = {
        "bound": "hbm", "kernel": dom, "achieved": kern[dom]["achieved_gbs"], "peak": peak, "unit": "GB/s",
        "frac": kern[dom]["frac"], "traffic": kern[dom]["traffic"], "peak_source": peak_src,
        "algorithmic_bytes_per_launch": kern[dom]["alg_bytes_per_launch"], "ms_per_launch": kern[dom]["ms_per_launch"],
        "byte_model": "bytes of the implemented algorithm (bench.py header / DESIGN.md section 3); traffic = ncu dram bytes of the committed capture",
        "kernels": kern,
        "whole_step": {"alg_bytes_per_pair": step_alg_bytes / n, "achieved_gbs": step_alg_bytes / (ms_step / 1e3) / 1e9,
                       "frac": step_alg_bytes / (ms_step / 1e3) / 1e9 / peak,
                       "survey_model_bytes_per_pair": 108 + 1536 * f_mis_global,
                       "note": "frac uses the implemented-algorithm bytes; SURVEY 8d's 108 + 1536 f_mis charges 48 probes per miss, which "
                               "the neighbourhood tables replace by four (kept for reference only)"},
    }

    # ------------------------------------------------------------------ downstream grouping of the resident shard
    grouping = None
    if not args.no_other:
        import ctypes as C
        L = _ffi.lib()
        best = None
        for _ in range(3):
            pipe.sync()
            t0 = time.perf_counter()
            h = C.c_void_p()
            _ffi.check(pipe.ctx, L.pcl_group_build(pipe.ctx, big.h, None, C.byref(h)))       # blocks until the outputs exist
            dt = time.perf_counter() - t0
            gi = _ffi.GroupInfo()
            L.pcl_group_info_get(h, C.byref(gi))
            L.pcl_group_destroy(h)
            best = dt if best is None else min(best, dt)
        grouping = {"read_groups": n, "assigned": int(gi.n_assigned), "barcodes": int(gi.n_barcodes), "duplicate_sets": int(gi.n_fragments),
                    "ms": best * 1e3, "read_groups_per_s": n / best,
                    "note": "pcl_group_build on the resident shard after pass 2: stable device radix sort on (corrected barcode, UMI), "
                            "barcode boundaries and duplicate sets; host clock around the blocking call (allocations included)"}

    # ------------------------------------------------------------------ end to end from pinned host memory
    e2e = None
    if not args.no_e2e:
        s1, q1, l1 = big.device_ptrs(0)
        _, _, l2 = big.device_ptrs(1)
        nc = max(1, args.e2e_chunks)
        bounds = np.linspace(0, n, nc + 1).astype(np.int64)
        cap = int(np.max(np.diff(bounds)))
        qo, qs = pipe.qual_window(0)             # the quality plane only covers the barcode bases (16 of 28 bytes)
        st1 = pipe.payload_bytes(0)              # compact upload rows: the 28 bytes of R1 the layout uses (device rows are 32)
        seq_h, _p1 = pinned(n * st1, np.uint8, (n, st1))
        qual_h, _p2 = pinned(n * qs, np.uint8, (n, qs))
        len_h, _p3 = pinned(n * 2, np.uint16, (n,))
        len2_h, _p4 = pinned(n * 2, np.uint16, (n,))
        fs1_h, _p5 = pinned(n * 2, np.uint16, (n,))
        bck_h, _p6 = pinned(n * 4, np.uint32, (n,))
        umi_h, _p7 = pinned(n * 4, np.uint32, (n,))
        fs2_h, _p8 = pinned(n * 2, np.uint16, (n,))
        tc2_h, _p9 = pinned(n * 4, np.uint32, (n,))
        # same bytes as the resident shard: copy them out of the device planes (generator parity is tested separately)
        dst1 = pipe.stride(0)
        for lo in range(0, n, 1 << 24):          # device rows (32 B) -> host rows (28 B), in slices
            hi = min(n, lo + (1 << 24))
            tmp = np.empty((hi - lo, dst1), np.uint8)
            syn.memcpy_d2h(local_rank, tmp.ctypes.data, s1 + lo * dst1, (hi - lo) * dst1)
            seq_h[lo:hi] = tmp[:, :st1]
        del tmp
        syn.memcpy_d2h(local_rank, qual_h.ctypes.data, q1, n * qs)
        syn.memcpy_d2h(local_rank, len_h.ctypes.data, l1, n * 2)
        syn.memcpy_d2h(local_rank, len2_h.ctypes.data, l2, n * 2)
        chunks = [pipe.new_chunk(cap) for _ in range(nc)]
        from precellar_b200.pipeline import ReadResults
        outs = []
        for c in range(nc):
            lo, hi = int(bounds[c]), int(bounds[c + 1])
            r1 = ReadResults(pipe.infos[0], fs1_h[lo:hi], None, [bck_h[lo:hi]], [umi_h[lo:hi]], [None], [None])
            r2 = ReadResults(pipe.infos[1], fs2_h[lo:hi], tc2_h[lo:hi], [], [], [], [])
            outs.append((r1, r2))
        # Batch metadata an ingest stage has for free (min / max record length while parsing): when a read file's records
        # all have one length, its length plane stays on the host (pcl_chunk_set_uniform_len) and the result planes of an
        # insert read come back as one (fstatus, tcoord) pair per chunk (pcl_chunk_uniform_results_async).
        uni1 = int(len_h.min()) == int(len_h.max())
        uni2 = int(len2_h.min()) == int(len2_h.max())
        L1u, L2u = int(len_h[0]), int(len2_h[0])
        uni_res, _p10 = pinned(nc * 16, np.uint32, (nc, 4))
        h2d = n * (st1 + qs + (0 if uni1 else 2) + (0 if uni2 else 2))
        d2h = n * (2 + 4 + 4 + (0 if uni2 else 2 + 4)) + (16 * nc if uni2 else 0)

        def step_e2e():
            pipe.reset_counts()
            for c, ch in enumerate(chunks):
                lo, hi = int(bounds[c]), int(bounds[c + 1])
                ch.reset(hi - lo)
                if uni1:
                    ch.set_uniform_len(0, L1u)
                ch.upload_compact(0, HostBatch(seq_h[lo:hi], qual_h[lo:hi], None if uni1 else len_h[lo:hi]))
                if uni2:
                    ch.set_uniform_len(1, L2u)
                else:
                    ch.upload(1, HostBatch(None, None, len2_h[lo:hi]))
                pipe.count(ch)
                # UMI keys and insert coordinates are final after pass 1: they travel back while later chunks upload
                ch.fetch(0, outs[c][0], wait=False, planes=("umi_key",))
                if uni2:
                    ch.uniform_results_async(1, uni_res[c])
                else:
                    ch.fetch(1, outs[c][1], wait=False)
            pipe.allreduce()
            for c, ch in enumerate(chunks):
                pipe.correct(ch, THRESHOLD)
                ch.fetch(0, outs[c][0], wait=False, planes=("fstatus", "bc_key"))
            pipe.sync()
            if uni2:
                for c, ch in enumerate(chunks):
                    if uni_res[c, 0] != 1:                    # not constant after all: the planes themselves
                        ch.fetch(1, outs[c][1])

        for _ in range(2):
            step_e2e()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            step_e2e()
        pipe.sync()
        barrier()
        dt = max_over_ranks(time.perf_counter() - t0) / args.e2e_steps
        e2e = {"value": n * world / dt, "unit": "read-pairs/s", "h2d_bytes_per_step": int(h2d) * world,
               "d2h_bytes_per_step": int(d2h) * world, "h2d_bytes_per_step_per_gpu": int(h2d),
               "d2h_bytes_per_step_per_gpu": int(d2h), "ms_per_step": dt * 1e3, "chunks": nc,
               "uniform_lengths": [bool(uni1), bool(uni2)],
               "note": "pinned host SoA planes (R1 rows of the 28 bytes the layout uses, widened to 32 on the device) -> H2D -> count -> allreduce -> correct -> D2H of every result plane; "
                       "read files whose records all have one length send no length plane, and an insert read's constant result planes come back as one (fstatus, tcoord) pair per chunk"}
        for ch in chunks:
            ch.close()
            pipe.chunks.remove(ch)
        assert not uni2 or (uni_res[:, 0] == 1).all()
        for ptr in (_p1, _p2, _p3, _p4, _p5, _p6, _p7, _p8, _p9, _p10):
            _ffi.lib().pcl_host_free(ptr)
        del seq_h, qual_h, len_h, len2_h, fs1_h, bck_h, umi_h, fs2_h, tc2_h, outs, uni_res

    # ------------------------------------------------------------------ parity (every rank takes part) + CPU baseline
    parity = cpu_baseline = None
    if not args.no_cpu:
        ns = min(args.cpu_sample, n)
        parity, cpu_s = parity_prefix(w3, pipe, ns, local_rank, world, rank, host_cores, gather)
        if rank == 0:
            cpu_baseline = {"value": ns / cpu_s, "unit": "read-pairs/s", "cores": host_cores, "kind": "port",
                            "sample": f"first {ns} read pairs of the workload, one pass: serial count + {host_cores}-thread "
                                      "correct (C++ restatement of the reference algorithm; in-memory SoA input, no gzip)"}

    ingest = None
    if rank == 0 and world == 1 and not args.no_cpu and args.ingest_sample > 0:
        import tempfile
        try:
            with tempfile.TemporaryDirectory() as td:
                ingest = measure_ingest(syn, args.ingest_sample, td, pipe)
        except Exception as exc:                 # an extra like the API leg: a failure here must not take the headline line with it
            import traceback
            traceback.print_exc(file=sys.stderr)
            ingest = {"error": f"{type(exc).__name__}: {exc}"}
    R.close()

    # ------------------------------------------------------------------ the other BASELINE configurations (one GPU)
    other = None
    if world == 1 and not args.no_other:
        other = []
        specs = [("configs[1] 10x scATAC", [W.atac()], 100_000_000, "atac"),
                 ("configs[2] 10x multiome: RNA and ATAC pipelines resident together", [W.multiome_rna(), W.multiome_atac()], 50_000_000, "multiome"),
                 ("configs[3] sci-RNA-seq3", [W.sci3()], 100_000_000, "sci3")]
        for title, ws, n_cfg, key in specs:
            n_cfg = args.other_n or n_cfg
            rs = [Resident(wk, n_cfg, local_rank) for wk in ws]
            ms, profs, ln, _ = measure_resident(rs, args.steps, args.warmup, barrier, max_over_ranks)
            ent = {"config": title, "read_groups_per_pipeline": n_cfg, "pipelines": [], "ms_per_step": ms,
                   "value": n_cfg * len(rs) / (ms / 1e3), "unit": "read-groups/s", "gpu_launches": int(ln),
                   "timing": "CUDA events" if len(rs) == 1 else "host clock between device syncs (two contexts, kernels of both in flight)"}
            for k, (r, pf) in enumerate(zip(rs, profs)):
                tots = [r.pipe.counts(g) for g in range(len(r.w.layout.region_ids))]
                kt = kernel_table(r.pipe, r.w, n_cfg, pf, args.steps, [t[1] for t in tots], peak, key if len(rs) == 1 else f"{key}{k}")
                pent = {"workload": r.w.name, "unit": r.w.unit, "f_mis": [t[1] / max(t[2], 1) for t in tots],
                        "whitelist_tables_resident": [int(r.pipe.whitelist_size(g)) for g in range(len(tots))],
                        "defect_records": [int(x) for x in r.pipe.qc()[:, 1]], "setup_s": r.setup_s, "kernels": kt}
                if not args.no_cpu:
                    par, cpu_s = parity_prefix(r.w, r.pipe, min(args.cpu_sample, n_cfg), local_rank, 1, 0, host_cores, gather)
                    pent["parity"] = par
                    pent["cpu_baseline"] = {"value": par["checked_read_groups"] / cpu_s, "unit": "read-groups/s", "cores": host_cores, "kind": "port"}
                ent["pipelines"].append(pent)
            other.append(ent)
            for r in rs:
                r.close()

    # ------------------------------------------------------------------ through the reference's API: files -> make_fastq -> .fq.zst
    api = None
    if rank == 0 and world == 1 and not args.no_cpu and args.api_sample > 0:
        import tempfile
        from tools import api_bench
        try:
            with tempfile.TemporaryDirectory() as td:
                api = api_bench.api_e2e(td, args.api_sample, host_cores)
        except Exception as exc:                 # the API leg is an extra: a failure there must not take the headline line with it
            import traceback
            traceback.print_exc(file=sys.stderr)
            api = {"error": f"{type(exc).__name__}: {exc}"}

    if rank == 0:
        out = {
            "metric": "barcode-corrected read-pairs/sec", "value": value, "unit": "read-pairs/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8 (2-bit packed keys) + f64 posterior", "data": "synthetic",
            "config": {"workload": workload, "pairs_per_gpu": n, "whitelist": N_WHITELIST, "n_cells": 10000,
                       "f_mis": f_mis_global, "threshold": THRESHOLD, "parallelism": f"reads sharded over {world} GPU(s), "
                       "whitelist replicated, one NCCL allreduce of the counts",
                       "l2": "inputs (6.3 GB per GPU) are far larger than L2; no flush needed"},
            "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": int(launches),
            "clocks": clocks.summary(), "sustained": sustained, "parity": parity, "host_cores": host_cores, "host_ingest": ingest,
            "other_configs": other, "grouping": grouping, "api_e2e": api,
            "setup_s": {"pipeline_create_incl_whitelist_tables": t_setup},
        }
        emit(out)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
