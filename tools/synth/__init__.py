"""Python front-end of the synthetic read generator (bench / test tooling; see synth.cu)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libpcl_synth.so")
SEED = 0x5EED0001


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "synth.cu")
    if force or not os.path.exists(LIB) or os.path.getmtime(src) > os.path.getmtime(LIB):
        nvcc = "/usr/local/cuda/bin/nvcc"
        subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
                               "--shared", "-Xcompiler", "-fPIC", "-o", LIB, src])
    return LIB


class _Params(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("n_whitelist", C.c_uint32), ("bc_len", C.c_uint32), ("umi_len", C.c_uint32),
                ("prefix_len", C.c_uint32), ("stride", C.c_uint32), ("n_cells", C.c_uint32), ("reverse", C.c_uint32),
                ("pad", C.c_uint32)]


class _Params3(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("n_hp", C.c_uint32), ("n_rt", C.c_uint32), ("n_cells", C.c_uint32),
                ("stride", C.c_uint32)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB)
        L.synth_create.restype = C.c_void_p
        L.synth_create.argtypes = [C.POINTER(_Params), C.c_void_p, C.c_void_p, C.c_void_p]
        L.synth_destroy.argtypes = [C.c_void_p]
        L.synth_fill_host.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.synth_fill_device.argtypes = [C.c_void_p, C.c_int, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_uint32, C.c_uint32]
        L.synth_fill_u16_device.argtypes = [C.c_int, C.c_void_p, C.c_uint16, C.c_uint64, C.c_void_p]
        L.synth_memcpy_d2h.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_uint64]
        L.synth3_create.restype = C.c_void_p
        L.synth3_create.argtypes = [C.POINTER(_Params3), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.synth3_destroy.argtypes = [C.c_void_p]
        L.synth3_fill_host.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.synth3_fill_device.argtypes = [C.c_void_p, C.c_int, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def make_whitelist(n: int, k: int, seed: int = SEED) -> np.ndarray:
    """n distinct uniform-random ACGT k-mers as an [n, k] uint8 array (file order = generation order)."""
    assert k <= 31
    rng = np.random.default_rng(seed)
    space = 4 ** k
    if n > space:
        raise ValueError("whitelist larger than the k-mer space")
    keys = np.zeros(0, np.uint64)
    while len(keys) < n:
        draw = rng.integers(0, space, size=int((n - len(keys)) * 1.05) + 16, dtype=np.uint64)
        keys = np.concatenate([keys, draw])
        _, first = np.unique(keys, return_index=True)
        keys = keys[np.sort(first)]                      # drop duplicates, keep generation order
    keys = keys[:n]
    out = np.empty((n, k), np.uint8)
    lut = np.frombuffer(b"ACGT", np.uint8)
    for p in range(k):
        out[:, p] = lut[((keys >> np.uint64(2 * (k - 1 - p))) & np.uint64(3)).astype(np.int64)]
    return out


def zipf_thresholds(n_cells: int) -> np.ndarray:
    """thr[r] = floor(CDF_Zipf(s=1)(r) * 2^64) with the last entry saturated, as uint64."""
    w = 1.0 / np.arange(1, n_cells + 1, dtype=np.float64)
    cdf = np.cumsum(w) / np.sum(w)
    thr = np.minimum(np.floor(cdf * 18446744073709551616.0), 18446744073709549568.0).astype(np.uint64)
    thr[-1] = np.uint64(0xFFFFFFFFFFFFFFFF)
    return thr


@dataclass
class SynthSpec:
    n_whitelist: int
    bc_len: int = 16
    umi_len: int = 12
    prefix_len: int = 0
    stride: int = 32
    n_cells: int = 10000
    reverse: bool = False
    seed: int = SEED


class Synth:
    def __init__(self, spec: SynthSpec, whitelist: np.ndarray):
        assert whitelist.shape == (spec.n_whitelist, spec.bc_len) and whitelist.dtype == np.uint8
        assert spec.prefix_len + spec.bc_len + spec.umi_len <= spec.stride
        self.spec = spec
        self.whitelist = np.ascontiguousarray(whitelist)
        rng = np.random.default_rng(spec.seed ^ 0xC311)
        self.cells = rng.permutation(spec.n_whitelist)[:spec.n_cells].astype(np.uint32)
        self.thr = zipf_thresholds(len(self.cells))
        p = _Params(spec.seed, spec.n_whitelist, spec.bc_len, spec.umi_len, spec.prefix_len, spec.stride,
                    len(self.cells), int(spec.reverse), 0)
        self.h = C.c_void_p(lib().synth_create(C.byref(p), self.whitelist.ctypes.data, self.cells.ctypes.data,
                                               self.thr.ctypes.data))

    @property
    def record_len(self) -> int:
        return self.spec.prefix_len + self.spec.bc_len + self.spec.umi_len

    def fill_host(self, first: int, n: int, seq: np.ndarray, qual: np.ndarray, length: np.ndarray, n_threads: int = 0):
        assert seq.shape == (n, self.spec.stride) and qual.shape == (n, self.spec.stride) and length.shape == (n,)
        assert seq.flags.c_contiguous and qual.flags.c_contiguous and length.dtype == np.uint16
        lib().synth_fill_host(self.h, first, n, seq.ctypes.data, qual.ctypes.data, length.ctypes.data,
                              n_threads or os.cpu_count() or 1)

    def host_arrays(self, first: int, n: int, n_threads: int = 0):
        seq = np.empty((n, self.spec.stride), np.uint8)
        qual = np.empty((n, self.spec.stride), np.uint8)
        length = np.empty(n, np.uint16)
        self.fill_host(first, n, seq, qual, length, n_threads)
        return seq, qual, length

    def fill_device(self, device: int, first: int, n: int, seq_ptr: int, qual_ptr: int, len_ptr: int, stream: int = 0,
                    qual_window=None):
        """qual_window = (offset, stride) of the device quality plane (BarcodePipeline.qual_window)."""
        qoff, qstride = qual_window if qual_window is not None else (0, self.spec.stride)
        rc = lib().synth_fill_device(self.h, device, first, n, seq_ptr, qual_ptr, len_ptr, stream, qoff, qstride)
        if rc:
            raise RuntimeError(f"synth_fill_device: cudaError {rc}")

    @staticmethod
    def fill_u16_device(device: int, ptr: int, value: int, n: int, stream: int = 0):
        rc = lib().synth_fill_u16_device(device, ptr, value, n, stream)
        if rc:
            raise RuntimeError(f"synth_fill_u16_device: cudaError {rc}")

    @staticmethod
    def memcpy_d2h(device: int, dst_ptr: int, src_ptr: int, nbytes: int):
        rc = lib().synth_memcpy_d2h(device, dst_ptr, src_ptr, nbytes)
        if rc:
            raise RuntimeError(f"synth_memcpy_d2h: cudaError {rc}")

    def close(self):
        if self.h:
            lib().synth_destroy(self.h)
            self.h = None


def make_mixed_whitelist(n9: int, n10: int, seed: int):
    """Distinct 9-mers and 10-mers (sci-RNA-seq3 hairpin list: 238 x 9 + 146 x 10 in the reference's fixture),
    shuffled; returns list[bytes]."""
    rng = np.random.default_rng(seed)
    a = [bytes(r) for r in make_whitelist(n9, 9, seed + 1)]
    b = [bytes(r) for r in make_whitelist(n10, 10, seed + 2)]
    items = a + b
    rng.shuffle(items)
    return items


class SynthSci3:
    """sci-RNA-seq3 R1 generator (hairpin 9|10 + CAGAGC + UMI 8 + RT 10, 34 nt)."""

    def __init__(self, hp, rt, n_cells: int = 10000, stride: int = 48, seed: int = SEED):
        self.hp, self.rt, self.stride = list(hp), list(rt), stride
        hp10 = np.full((len(hp), 10), ord("A"), np.uint8)
        hl = np.zeros(len(hp), np.uint8)
        for i, x in enumerate(hp):
            hp10[i, :len(x)] = np.frombuffer(x, np.uint8)
            hl[i] = len(x)
        rt10 = np.stack([np.frombuffer(x, np.uint8) for x in rt])
        rng = np.random.default_rng(seed ^ 0x5C13)
        n_cells = min(n_cells, len(hp) * len(rt))
        combos = rng.permutation(len(hp) * len(rt))[:n_cells]
        self.cells = ((combos // len(rt)).astype(np.uint32) << 16) | (combos % len(rt)).astype(np.uint32)
        self.thr = zipf_thresholds(n_cells)
        self._keep = (hp10, hl, rt10)
        p = _Params3(seed, len(hp), len(rt), n_cells, stride)
        self.h = C.c_void_p(lib().synth3_create(C.byref(p), hp10.ctypes.data, hl.ctypes.data, rt10.ctypes.data,
                                                self.cells.ctypes.data, self.thr.ctypes.data))

    def host_arrays(self, first: int, n: int, n_threads: int = 0):
        seq = np.empty((n, self.stride), np.uint8)
        qual = np.empty((n, self.stride), np.uint8)
        length = np.empty(n, np.uint16)
        lib().synth3_fill_host(self.h, first, n, seq.ctypes.data, qual.ctypes.data, length.ctypes.data,
                               n_threads or os.cpu_count() or 1)
        return seq, qual, length

    def fill_device(self, device: int, first: int, n: int, seq_ptr: int, qual_ptr: int, len_ptr: int, stream: int = 0):
        rc = lib().synth3_fill_device(self.h, device, first, n, seq_ptr, qual_ptr, len_ptr, stream)
        if rc:
            raise RuntimeError(f"synth3_fill_device: cudaError {rc}")

    def close(self):
        if self.h:
            lib().synth3_destroy(self.h)
            self.h = None
