import os, sys, tempfile, time, json
sys.path.insert(0, '/root/repo')
os.environ["PCL_API_TIMING"]="1"
from tools import api_bench as A
from precellar_b200 import api
n=int(sys.argv[1])
with tempfile.TemporaryDirectory() as td:
    w, spec, files = A.make_inputs(td, n)
    for fmt in ("plain","gzip"):
        A.run_ours(spec, files[fmt], os.path.join(td,"o"))
        api.TIMING.clear()
        t0=time.perf_counter(); a=A.assay_for(spec, files[fmt]); t_assay=time.perf_counter()-t0
        t0=time.perf_counter()
        from precellar_b200.api import make_fastq
        make_fastq(a, modality="atac", out_dir=os.path.join(td,"o"), correct_barcode=True)
        tot=time.perf_counter()-t0
        print(fmt, n, "assay", round(t_assay,3), "make_fastq", round(tot,3), {k: round(v,3) for k,v in api.TIMING.items()})
