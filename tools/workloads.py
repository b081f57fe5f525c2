"""The BASELINE.json configurations as synthetic workloads (bench / test tooling).

A ``Workload`` couples a compiled read layout with the deterministic generator of its reads (tools/synth: the same bytes
on the host and on the device), so that bench.py and the GPU tests can
  * fill a resident device chunk in place                      (``fill_chunk``),
  * produce the full-width host rows of any record range       (``host_batches``: oracle input, upload source).

    v3(n_wl)         configs[0] / configs[4]   10x scRNA-seq v3: R1 = CB16 + UMI12, R2 = cDNA 91
    atac()           configs[1]                10x scATAC: I2 = CB16 neg strand, R1 / R2 = gDNA 50, 737 K whitelist
    multiome()       configs[2]                RNA (v3 shape, 737 K gex list) + ATAC (I2 = linker 8 + CB16 neg strand, 737 K atac list)
    sci3()           configs[3]                sci-RNA-seq3: R1 = hairpin 9|10 + CAGAGC + UMI 8 + RT 10, R2 = cDNA 100
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, List, Optional

import numpy as np

from precellar_b200 import configs
from precellar_b200.pipeline import HostBatch
from tools import synth as S


@dataclass
class Workload:
    name: str
    layout: object
    bc_read: int                         # the read file that holds the barcode(s)
    insert_lens: dict                    # read slot -> fixed length of the insert reads
    gen: object                          # Synth / SynthSci3
    unit: str = "read-pairs"

    def fill_chunk(self, pipe, ch, first: int, n: int, device: int) -> None:
        """Generate records [first, first + n) straight into the chunk's device planes."""
        s, q, l = ch.device_ptrs(self.bc_read)
        if isinstance(self.gen, S.SynthSci3):
            self.gen.fill_device(device, first, n, s, q, l, ch.stream)
        else:
            self.gen.fill_device(device, first, n, s, q, l, ch.stream, qual_window=pipe.qual_window(self.bc_read))
        for slot, ln in self.insert_lens.items():
            _, _, lp = ch.device_ptrs(slot)
            S.Synth.fill_u16_device(device, lp, ln, n, ch.stream)

    def host_batches(self, first: int, n: int) -> List[HostBatch]:
        """Full-width rows (stride of the generator) of records [first, first + n): what the oracle reads."""
        seq, qual, length = self.gen.host_arrays(first, n)
        out: List[Optional[HostBatch]] = [None] * len(self.layout.reads)
        out[self.bc_read] = HostBatch(seq, qual, length)
        for slot, ln in self.insert_lens.items():
            out[slot] = HostBatch(None, None, np.full(n, ln, np.uint16))
        return out

    def device_batches(self, pipe, full: List[HostBatch]) -> List[HostBatch]:
        """The planes pcl_chunk_upload expects (device stride, quality window) cut from full-width rows."""
        return [pipe.host_planes(r, b.seq, b.qual, b.length) if b.seq is not None else b for r, b in enumerate(full)]


def v3(n_wl: int = 6_800_000, r2_len: int = 91, seed: int = S.SEED, name: Optional[str] = None) -> Workload:
    wl = S.make_whitelist(n_wl, 16, seed)
    gen = S.Synth(S.SynthSpec(n_wl, bc_len=16, umi_len=12, stride=32, n_cells=10000, seed=seed), wl)
    return Workload(name or f"10x_rna_v3 (CB16 + UMI12, R2 {r2_len} nt, {n_wl} whitelist)", configs.v3_layout(wl, r2_len), 0, {1: r2_len}, gen)


def atac(n_wl: int = 737_000) -> Workload:
    wl = S.make_whitelist(n_wl, 16, seed=0xA7AC)
    gen = S.Synth(S.SynthSpec(n_wl, bc_len=16, umi_len=0, stride=16, reverse=True, seed=0x5EED0002), wl)
    return Workload(f"10x_atac (I2 = CB16 neg strand, R1/R2 gDNA 50 nt, {n_wl} whitelist)", configs.atac_layout(wl, 50), 1,
                    {0: 50, 2: 50}, gen, unit="read-triples")


def multiome_atac(n_wl: int = 737_000) -> Workload:
    wl = S.make_whitelist(n_wl, 16, seed=0xA7AD)
    gen = S.Synth(S.SynthSpec(n_wl, bc_len=16, umi_len=0, prefix_len=8, stride=32, reverse=True, seed=0x5EED0003), wl)
    return Workload(f"10x_multiome ATAC (I2 = linker 8 + CB16 neg strand, {n_wl} atac whitelist)", configs.multiome_atac_layout(wl, 50), 1,
                    {0: 50, 2: 50}, gen, unit="read-triples")


def multiome_rna(n_wl: int = 737_000) -> Workload:
    return v3(n_wl, 91, seed=0x5EED0004, name=f"10x_multiome RNA (v3 shape, {n_wl} gex whitelist)")


def sci3() -> Workload:
    hp = S.make_mixed_whitelist(238, 146, 7)
    rt = [bytes(r) for r in S.make_whitelist(384, 10, 99)]
    gen = S.SynthSci3(hp, rt)
    return Workload("sci_rna_seq3 (R1 = hairpin 9|10 + CAGAGC + UMI 8 + RT 10, R2 cDNA 100 nt, 384 + 384 whitelists)",
                    configs.sci3_layout(hp, rt, 34, 100), 0, {1: 100}, gen)


def run_oracle(w: Workload, full: List[HostBatch], threshold: float, n_threads: int, outputs: bool = True):
    """The reference algorithm (oracle port: serial count pass, parallel correct pass) on host rows.
    Returns (seconds, OracleResult, counts per region).  TEST / BASELINE leg only."""
    import time
    from oracle import pyoracle as O
    layout = w.layout
    wls = {}
    for slot, rid in enumerate(layout.region_ids):
        items = layout.whitelists[rid]
        if isinstance(items, np.ndarray):
            n, k = items.shape
            wls[slot] = (items.reshape(-1), np.arange(n + 1, dtype=np.uint64) * np.uint64(k), n)
        else:
            wls[slot] = list(items)
    orc = O.Oracle(O.reads_from_layout(layout), wls)
    for r, b in enumerate(full):
        orc.set_input(r, b.seq, b.qual, b.length, b.seq.shape[1] if b.seq is not None else 0)
    t0 = time.perf_counter()
    orc.count()                                                     # BarcodeAnalyzer::new, serial
    res = orc.annotate(correct=True, threshold=threshold, n_threads=n_threads, chunk_bases=1_000_000, outputs=outputs)
    dt = time.perf_counter() - t0
    counts = {g: orc.counts(g) for g in range(len(layout.region_ids))}
    orc.close()
    return dt, res, counts
