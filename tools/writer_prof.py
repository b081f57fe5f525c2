"""One make_fastq run with the device writer on the scATAC bench input (developer tool; run under ncu by
tools/prof/ncu_writer.sh): python tools/writer_prof.py N [runs]"""
import os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools import api_bench as A

n = int(sys.argv[1])
runs = int(sys.argv[2]) if len(sys.argv) > 2 else 1
with tempfile.TemporaryDirectory() as td:
    w, spec, files = A.make_inputs(td, n)
    for k in range(runs):
        t0 = time.perf_counter()
        A.run_ours(spec, files["plain"], os.path.join(td, "o"), writer="device", compression="device")
        print("run", k, round(time.perf_counter() - t0, 3), "s", flush=True)
