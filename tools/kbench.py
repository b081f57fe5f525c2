"""Kernel timing only (no assertions): v3 layout, device-resident. Diagnostic tool."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from precellar_b200 import _ffi, configs
from precellar_b200.pipeline import BarcodePipeline
from tools import synth as S
n = int(sys.argv[1]) if len(sys.argv) > 1 else 60_000_000
wl = S.make_whitelist(6_800_000, 16)
syn = S.Synth(S.SynthSpec(6_800_000), wl)
pipe = BarcodePipeline(configs.v3_layout(wl, 91))
ch = pipe.new_chunk(n); ch.reset(n)
s, q, l = ch.device_ptrs(0); _, _, l2 = ch.device_ptrs(1)
syn.fill_device(0, 0, n, s, q, l, ch.stream, qual_window=pipe.qual_window(0)); syn.fill_u16_device(0, l2, 91, n, ch.stream); pipe.sync()
def step():
    pipe.reset_counts(); pipe.count(ch); pipe.allreduce(); pipe.correct(ch, 0.975)
for _ in range(3): step()
pipe.sync(); pipe.profile_enable(True); pipe.profile_reset()
for _ in range(5): step()
out = {k: pipe.profile_get(v)[0] / 5 for k, v in (("count", _ffi.K_COUNT), ("lenonly", _ffi.K_COUNT_LENONLY), ("correct", _ffi.K_CORRECT))}
print(json.dumps(dict(exp=os.environ.get("PCL_EXPERIMENT", "0"), n=n, **out)))
