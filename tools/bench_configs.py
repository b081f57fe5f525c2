"""Device-resident throughput of the other BASELINE configurations (ATAC, multiome ATAC, sci-RNA-seq3).
These are parity-test cases rather than bench lines; this script only records where their kernels stand.

    python tools/bench_configs.py [--n 100000000] [--steps 5]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from precellar_b200 import _ffi, configs  # noqa: E402
from precellar_b200.pipeline import BarcodePipeline  # noqa: E402
from tools import synth as S  # noqa: E402


def run(name, layout, fill, n, steps):
    pipe = BarcodePipeline(layout)
    ch = pipe.new_chunk(n)
    ch.reset(n)
    fill(pipe, ch)
    pipe.sync()

    def step():
        pipe.reset_counts()
        pipe.count(ch)
        pipe.allreduce()
        pipe.correct(ch, 0.975)

    for _ in range(3):
        step()
    pipe.sync()
    pipe.profile_enable(True)
    pipe.profile_reset()
    ch.timer_start()
    for _ in range(steps):
        step()
    ms = ch.timer_stop_ms() / steps
    prof = {k: pipe.profile_get(k) for k in (_ffi.K_COUNT, _ffi.K_COUNT_LENONLY, _ffi.K_CORRECT)}
    tot = [pipe.counts(g) for g in range(len(layout.region_ids))]
    out = dict(config=name, n=n, ms_per_step=ms, groups_per_s=n / ms * 1e3,
               k_count_ms=prof[_ffi.K_COUNT][0] / steps, k_count_lenonly_ms=prof[_ffi.K_COUNT_LENONLY][0] / steps,
               k_correct_ms=prof[_ffi.K_CORRECT][0] / steps,
               f_mis=[t[1] / max(t[2], 1) for t in tot], defects=[int(x) for x in pipe.qc()[:, 1]])
    pipe.close()
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=100_000_000)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    n = a.n

    def fill_len(pipe, ch, slot, v):
        _, _, l = ch.device_ptrs(slot)
        S.Synth.fill_u16_device(0, l, v, n, ch.stream)

    if a.only in ("", "atac"):
        wl = S.make_whitelist(737_000, 16, seed=0xA7AC)
        syn = S.Synth(S.SynthSpec(737_000, bc_len=16, umi_len=0, stride=16, reverse=True, seed=0x5EED0002), wl)

        def fill(pipe, ch):
            s, q, l = ch.device_ptrs(1)
            syn.fill_device(0, 0, n, s, q, l, ch.stream, qual_window=pipe.qual_window(1))
            fill_len(pipe, ch, 0, 50)
            fill_len(pipe, ch, 2, 50)
        run("10x_atac (I2 16nt neg strand, 737K whitelist, read triples)", configs.atac_layout(wl, 50), fill, n, a.steps)
    if a.only in ("", "multiome"):
        wl = S.make_whitelist(737_000, 16, seed=0xA7AD)
        syn = S.Synth(S.SynthSpec(737_000, bc_len=16, umi_len=0, prefix_len=8, stride=32, reverse=True, seed=0x5EED0003), wl)

        def fill(pipe, ch):
            s, q, l = ch.device_ptrs(1)
            syn.fill_device(0, 0, n, s, q, l, ch.stream, qual_window=pipe.qual_window(1))
            fill_len(pipe, ch, 0, 50)
            fill_len(pipe, ch, 2, 50)
        run("10x_multiome ATAC side (I2 = linker8 + 16nt neg strand)", configs.multiome_atac_layout(wl, 50), fill, n, a.steps)
    if a.only in ("", "sci3"):
        hp = S.make_mixed_whitelist(238, 146, 7)
        rt = [bytes(r) for r in S.make_whitelist(384, 10, 99)]
        gen = S.SynthSci3(hp, rt)

        def fill(pipe, ch):
            s, q, l = ch.device_ptrs(0)
            gen.fill_device(0, 0, n, s, q, l, ch.stream)
            fill_len(pipe, ch, 1, 100)
        run("sci_rna_seq3 (hairpin 9|10 + CAGAGC + UMI8 + RT10, AnchorPlan kernel)", configs.sci3_layout(hp, rt, 34, 100), fill, n, a.steps)


if __name__ == "__main__":
    main()
