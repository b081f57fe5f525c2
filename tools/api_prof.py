"""Where an API-level run spends its wall clock: python tools/api_prof.py N [writer:compression ...] (needs a GPU)."""
import os, sys, tempfile, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["PCL_API_TIMING"] = "1"
from tools import api_bench as A
from precellar_b200 import api
from precellar_b200.api import make_fastq

n = int(sys.argv[1])
modes = [m.split(":") for m in sys.argv[2:]] or [["host", "zstd"], ["device", "zstd"], ["device", "device"]]
with tempfile.TemporaryDirectory() as td:
    w, spec, files = A.make_inputs(td, n)
    digests = {}
    for fmt in ("plain", "gzip", "bgzf"):
        for writer, comp in modes:
            out = os.path.join(td, "o_%s_%s" % (writer, comp))
            A.run_ours(spec, files[fmt], out, writer=writer, compression=comp)
            best = None
            for _ in range(2):
                api.TIMING.clear()
                t0 = time.perf_counter(); a = A.assay_for(spec, files[fmt]); t_assay = time.perf_counter() - t0
                t0 = time.perf_counter()
                make_fastq(a, modality="atac", out_dir=out, correct_barcode=True, writer=writer, compression=comp)
                tot = time.perf_counter() - t0
                if best is None or tot + t_assay < best[0]:
                    best = (tot + t_assay, t_assay, tot, dict(api.TIMING))
            os.environ["PCL_API_KERNEL_MS"] = "1"                     # one more run with every kernel bracketed by events
            api.TIMING.clear()
            make_fastq(A.assay_for(spec, files[fmt]), modality="atac", out_dir=out, correct_barcode=True, writer=writer, compression=comp)
            kernel_ms = api.TIMING.get("kernel_ms")
            del os.environ["PCL_API_KERNEL_MS"]
            sizes = [os.path.getsize(os.path.join(out, f)) for f in ("R1.fq.zst", "R2.fq.zst")]
            d = A._digest(os.path.join(out, "R1.fq.zst")) + A._digest(os.path.join(out, "R2.fq.zst"))
            digests.setdefault(fmt, set()).add(d)
            print(json.dumps({"fmt": fmt, "writer": writer, "compression": comp, "n": n, "total_s": round(best[0], 3),
                              "triples_per_s": round(n / best[0]), "assay_s": round(best[1], 3), "make_fastq_s": round(best[2], 3),
                              "out_bytes": sizes, "kernel_ms": kernel_ms, "stages": {k: round(v, 3) for k, v in best[3].items()}}), flush=True)
    print("identical outputs per format:", {k: len(v) == 1 for k, v in digests.items()})
