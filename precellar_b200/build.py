"""In-tree build of libprecellar_b200.so (sm_100a only).

    python -m precellar_b200.build [--force]

nvcc cross-compiles without a GPU; the resulting .so is git-ignored but travels with the tree.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libprecellar_b200.so")
SOURCES = [os.path.join(CSRC, "pcl.cu"), os.path.join(CSRC, "ingest.cpp"), os.path.join(CSRC, "gzsrc.cpp")]
DEPS = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh", ".h", ".cpp"))] + \
       [os.path.join(HERE, "..", "include", "precellar_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",   # B200 only: no PTX for other targets, no multi-arch fatbin
    "-lineinfo", "-O3", "-std=c++17", "--shared", "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
]


def nvcc_path() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def is_stale() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(d) > t for d in DEPS if os.path.exists(d))


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return OUT
    cmd = [nvcc_path()] + NVCC_FLAGS + ["-o", OUT] + SOURCES + ["-ldl", "-lz"]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    log = proc.stdout + proc.stderr
    with open(os.path.join(HERE, "build.log"), "w") as fh:
        fh.write(" ".join(cmd) + "\n" + log)
    if proc.returncode != 0:
        sys.stderr.write(log)
        raise RuntimeError("nvcc failed")
    if verbose:
        print(log)
    return OUT


def build_tools(force: bool = False) -> None:
    """Bench / test tooling that also contains sm_100a code (the synthetic read generator)."""
    import importlib
    sys.path.insert(0, os.path.dirname(HERE))
    importlib.import_module("tools.synth").build(force)
    build_examples(force)


def build_examples(force: bool = False) -> str:
    """examples/v3_demo: the hot path driven from C++ through the C ABI only (tests/test_gpu_cabi_demo.py runs it)."""
    root = os.path.dirname(HERE)
    src, exe = os.path.join(root, "examples", "v3_demo.cpp"), os.path.join(root, "examples", "v3_demo")
    deps = [src, os.path.join(root, "include", "precellar_b200.h"), OUT]
    if force or not os.path.exists(exe) or any(os.path.getmtime(d) > os.path.getmtime(exe) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-I", os.path.join(root, "include"), src, "-o", exe, "-L", HERE,
                               "-lprecellar_b200", "-Wl,-rpath,$ORIGIN/../precellar_b200"])
    return exe


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
    build_tools(force="--force" in sys.argv)
