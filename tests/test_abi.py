"""CPU tier: the C-ABI library builds, loads, exports every symbol the header declares, and refuses to
compute without a device (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import pytest

from precellar_b200 import _ffi, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def libpath():
    return build.build()


def header_symbols():
    text = open(os.path.join(ROOT, "include", "precellar_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pcl_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    assert header_symbols() == sorted(_ffi.SYMBOLS)


def test_library_exports_every_declared_symbol(libpath):
    out = subprocess.check_output(["nm", "-D", "--defined-only", libpath], text=True)
    exported = {ln.split()[-1] for ln in out.splitlines() if " T " in ln}
    missing = [s for s in header_symbols() if s not in exported]
    assert not missing, missing


def test_library_is_sm100a_only(libpath):
    out = subprocess.run(["cuobjdump", "--list-elf", libpath], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_loads_and_struct_sizes(libpath):
    L = _ffi.lib()
    assert L.pcl_abi_version() == 1
    # layout of the POD structs must match the header (checked against a compiled probe)
    probe = os.path.join(ROOT, "tests", "emul", "sizes_probe")
    src = probe + ".c"
    with open(src, "w") as fh:
        fh.write('#include <stdio.h>\n#include "../../include/precellar_b200.h"\n'
                 'int main(){printf("%zu %zu %zu %zu %zu\\n", sizeof(pcl_elem_desc), sizeof(pcl_read_desc), '
                 'sizeof(pcl_read_info), sizeof(pcl_read_results), sizeof(pcl_emit_result));return 0;}\n')
    subprocess.check_call(["gcc", "-o", probe, src])
    sizes = [int(x) for x in subprocess.check_output([probe], text=True).split()]
    assert sizes == [C.sizeof(_ffi.ElemDesc), C.sizeof(_ffi.ReadDesc), C.sizeof(_ffi.ReadInfo), C.sizeof(_ffi.ReadResults),
                     C.sizeof(_ffi.EmitResult)]
    os.remove(probe)
    os.remove(src)


def test_key_decode():
    L = _ffi.lib()
    buf = C.create_string_buffer(40)
    key = (1 << 8) | (0 << 6) | (1 << 4) | (2 << 2) | 3          # "ACGT"
    assert L.pcl_key_decode(key, buf) == 4 and buf.value == b"ACGT"
    assert L.pcl_key_decode(0, buf) == -1
    from precellar_b200.pipeline import decode_key_py
    assert decode_key_py(key) == b"ACGT"


def test_no_cpu_fallback():
    """Without a CUDA device every compute path fails loudly."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    L = _ffi.lib()
    ctx = C.c_void_p()
    rc = L.pcl_ctx_create(C.byref(ctx), 0)
    assert rc == _ffi.PCL_ERR_CUDA and not ctx.value
    assert L.pcl_last_error(None)
    from tests import util as U
    from tests.util import bc, make_layout, umi
    from precellar_b200.pipeline import BarcodePipeline
    layout = make_layout([dict(read_id="R1", elems=[bc("cb", 16), umi(12)], max_len=28)], {"cb": [b"A" * 16]})
    with pytest.raises(_ffi.PclError):
        BarcodePipeline(layout)


def test_product_does_not_reference_the_oracle():
    """Nothing under precellar_b200/ may import, link or call the oracle or the emulation harness."""
    bad = []
    for dp, _, files in os.walk(os.path.join(ROOT, "precellar_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dp, f), errors="replace").read()
                if re.search(r"\boracle\b|pyoracle|liboracle|libemul|emu_run", text) and f != "hd_core.h":
                    bad.append(f)
    assert not bad, bad
