"""GPU tier: examples/v3_demo (C++, the C ABI only -- no Python host layer, no torch) on FASTQ files, against the
Python host layer on the same files: identical status counts, whitelist totals and checksums of every result plane."""
import json
import os
import subprocess

import numpy as np
import pytest

from tests import util as U

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
M64 = (1 << 64) - 1


def mix(h, v):
    return (h ^ ((v + 0x9E3779B97F4A7C15 + ((h << 6) & M64) + (h >> 2)) & M64)) & M64


def test_cpp_demo_matches_python_host_layer(tmp_path):
    from precellar_b200 import _ffi
    from precellar_b200.build import build_examples
    from precellar_b200.ingest import GroupReader
    from precellar_b200.pipeline import BarcodePipeline
    from tests.test_api_e2e import write_fastq
    exe = build_examples()
    rng = np.random.default_rng(31)
    wl = U.gen_whitelist(rng, 3000, 16)
    layout = U.make_layout([dict(read_id="R1", elems=[U.bc("cb", 16), U.umi(12)], max_len=28),
                            dict(read_id="R2", elems=[U.cdna(1, 1000)], is_reverse=True, max_len=0)], {"cb": wl})
    n = 30000
    full = U.gen_reads(rng, layout, n, [32, 128], p_sub=0.04, p_n=0.02, p_short=0.03, hot=300, read_len=[28, 91])
    for b in full:
        b.length[:] = np.maximum(b.length, 1)
    td = str(tmp_path)
    f1, f2, fw = os.path.join(td, "R1.fq.gz"), os.path.join(td, "R2.fq"), os.path.join(td, "wl.txt")
    write_fastq(f1, [f"r{i}/1".encode() for i in range(n)], full[0], True)
    write_fastq(f2, [f"r{i}/2".encode() for i in range(n)], full[1], False)
    with open(fw, "wb") as fh:
        fh.write(b"\n".join(wl) + b"\n")
    p = subprocess.run([exe, f1, f2, fw, "0.9", "7000"], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    got = json.loads(p.stdout.strip().splitlines()[-1])

    pipe = BarcodePipeline(layout)
    try:
        rdr = GroupReader([[f1], [f2]], [pipe.stride(0), 0], [False, False], qual_windows=[pipe.qual_window(0), (0, 0)])
        chunks = []
        while True:
            g = rdr.next(7000)
            if g is None:
                break
            ch = pipe.new_chunk(len(g[0].batch.length))
            ch.reset(len(g[0].batch.length))
            for r in range(2):
                ch.upload(r, g[r].batch)
            pipe.count(ch)
            pipe.sync()
            chunks.append(ch)
        rdr.close()
        pipe.allreduce()
        want = dict(records=0, exact=0, corrected=0, no_match=0, low_confidence=0, absent=0, targets=0)
        checksum = 0
        for ch in chunks:
            pipe.correct(ch, 0.9)
            r1, r2 = ch.fetch(0), ch.fetch(1)
            code = (r1.fstatus >> 4) & 7
            want["records"] += len(code)
            for name, c in (("exact", _ffi.BC_EXACT), ("corrected", _ffi.BC_CORRECTED), ("no_match", _ffi.BC_NO_MATCH),
                            ("low_confidence", _ffi.BC_LOW_CONF), ("absent", _ffi.BC_ABSENT)):
                want[name] += int((code == c).sum())
            want["targets"] += int(((r2.fstatus & 4) != 0).sum())
            assigned = (code == _ffi.BC_EXACT) | (code == _ffi.BC_CORRECTED)
            for i in range(len(code)):
                checksum = mix(checksum, int(r1.fstatus[i]))
                checksum = mix(checksum, int(r1.bc_key[0][i]) if assigned[i] else 0)
                checksum = mix(checksum, int(r1.umi_key[0][i]))
                checksum = mix(checksum, int(r2.fstatus[i]))
                checksum = mix(checksum, int(r2.tcoord[i]))
        counts, mismatch, total = pipe.counts(0)
        csum = 0
        for c in counts:
            csum = mix(csum, int(c))
    finally:
        pipe.close()
    for k, v in want.items():
        assert got[k] == v, (k, got[k], v)
    assert (got["whitelist_total"], got["whitelist_mismatch"]) == (total, mismatch)
    assert got["counts_checksum"] == csum and got["results_checksum"] == checksum
    assert got["corrected"] > 0 and got["no_match"] > 0 and got["kernels_launched"] > 0
