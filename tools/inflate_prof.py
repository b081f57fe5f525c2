"""Compressed input through the API, new decoder against old, alternating in one process (needs a GPU):
python tools/inflate_prof.py N [out.json] [bgzf|gzip]
  bgzf: members inflated on the device (k_bgzf_inflate) against the host threads (PCL_DEVICE_INFLATE=0)
  gzip: one stream per file through csrc/fastinflate.h against zlib (PCL_ZLIB_INFLATE=1)
FASTQ files -> make_fastq(writer=device, compression=device)."""
import json
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["PCL_API_TIMING"] = "1"
from tools import api_bench as A
from precellar_b200 import api, textingest
from precellar_b200.api import make_fastq

n = int(sys.argv[1])
fmt = sys.argv[3] if len(sys.argv) > 3 else "bgzf"
SWITCH = {"bgzf": ("PCL_DEVICE_INFLATE", {"old": "0", "new": "1"}), "gzip": ("PCL_ZLIB_INFLATE", {"old": "1", "new": "0"})}[fmt]
rows = []
with tempfile.TemporaryDirectory() as td:
    w, spec, files = A.make_inputs(td, n)
    comp_bytes = sum(os.path.getsize(p) for p in files[fmt].values())
    text_bytes = sum(os.path.getsize(p) for p in files["plain"].values())
    digests = set()
    for which in ("old", "new", "old", "new"):
        mode = "1" if which == "new" else "0"
        os.environ[SWITCH[0]] = SWITCH[1][which]
        out = os.path.join(td, "o_" + mode)
        best = None
        for rep in range(3):
            for k in textingest.INFLATED_STATS:
                textingest.INFLATED_STATS[k] = 0
            api.TIMING.clear()
            t0 = time.perf_counter()
            a = A.assay_for(spec, files[fmt])
            make_fastq(a, modality="atac", out_dir=out, correct_barcode=True, writer="device", compression="device")
            tot = time.perf_counter() - t0
            if rep and (best is None or tot < best[0]):
                best = (tot, dict(api.TIMING), dict(textingest.INFLATED_STATS))
        digests.add(A._digest(os.path.join(out, "R1.fq.zst")) + A._digest(os.path.join(out, "R2.fq.zst")))
        row = {"format": fmt, "decoder": {"bgzf": ("host threads (zlib)", "k_bgzf_inflate"), "gzip": ("zlib", "fastinflate.h")}[fmt][which == "new"],
               "device_inflate": fmt == "bgzf" and which == "new", "n": n, "total_s": round(best[0], 4), "triples_per_s": round(n / best[0]),
               "pass1_text_parse_s": round(best[1].get("pass1_text_parse", 0.0), 4), "inflate": best[2],
               "text_bytes": text_bytes, "compressed_bytes": comp_bytes}
        if best[2]["kernel_ms"]:
            row["inflate_text_GBps"] = round(best[2]["out_bytes"] / best[2]["kernel_ms"] / 1e6, 2)
        rows.append(row)
        print(json.dumps(row), flush=True)
    ok = len(digests) == 1
    print("outputs identical with both decoders:", ok)
if len(sys.argv) > 2:
    with open(sys.argv[2], "w") as fh:
        for r in rows:
            fh.write(json.dumps(r) + "\n")
        fh.write(json.dumps({"outputs_identical": ok}) + "\n")
sys.exit(0 if ok else 1)
