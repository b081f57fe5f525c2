"""cProfile of one make_fastq run (needs a GPU): python tools/api_cprof.py N writer compression [fmt]"""
import cProfile, os, pstats, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools import api_bench as A
from precellar_b200.api import make_fastq

n, writer, comp = int(sys.argv[1]), sys.argv[2], sys.argv[3]
fmt = sys.argv[4] if len(sys.argv) > 4 else "plain"
with tempfile.TemporaryDirectory() as td:
    w, spec, files = A.make_inputs(td, n)
    out = os.path.join(td, "o")
    for _ in range(2):
        A.run_ours(spec, files[fmt], out, writer=writer, compression=comp)
    a = A.assay_for(spec, files[fmt])
    pr = cProfile.Profile()
    t0 = time.perf_counter()
    pr.enable()
    make_fastq(a, modality="atac", out_dir=out, correct_barcode=True, writer=writer, compression=comp)
    pr.disable()
    print("make_fastq", time.perf_counter() - t0)
    pstats.Stats(pr).sort_stats("cumulative").print_stats(45)
    pstats.Stats(pr).sort_stats("tottime").print_stats(25)
