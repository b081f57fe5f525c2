"""Minimal zstd codec over the system libzstd.so.1 (ctypes; the image ships the library without headers).

Used for ``.zst`` inputs (seqspec/src/utils.rs:44-56) and the ``R1.fq.zst`` / ``R2.fq.zst`` outputs of
``make_fastq`` (python/src/lib.rs:136-146).  Only stable public entry points of libzstd are used.
"""
from __future__ import annotations

import ctypes as C
import ctypes.util

_lib = None


class _Buf(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("size", C.c_size_t), ("pos", C.c_size_t)]


def lib():
    global _lib
    if _lib is None:
        name = ctypes.util.find_library("zstd") or "libzstd.so.1"
        try:
            L = C.CDLL(name)
        except OSError as e:
            raise RuntimeError("libzstd is not available on this system") from e
        L.ZSTD_compressBound.restype = C.c_size_t
        L.ZSTD_compressBound.argtypes = [C.c_size_t]
        L.ZSTD_compress.restype = C.c_size_t
        L.ZSTD_compress.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int]
        L.ZSTD_isError.restype = C.c_uint
        L.ZSTD_isError.argtypes = [C.c_size_t]
        L.ZSTD_getErrorName.restype = C.c_char_p
        L.ZSTD_getErrorName.argtypes = [C.c_size_t]
        L.ZSTD_createDStream.restype = C.c_void_p
        L.ZSTD_freeDStream.argtypes = [C.c_void_p]
        L.ZSTD_initDStream.restype = C.c_size_t
        L.ZSTD_initDStream.argtypes = [C.c_void_p]
        L.ZSTD_decompressStream.restype = C.c_size_t
        L.ZSTD_decompressStream.argtypes = [C.c_void_p, C.POINTER(_Buf), C.POINTER(_Buf)]
        _lib = L
    return _lib


def _check(code: int) -> int:
    L = lib()
    if L.ZSTD_isError(code):
        raise RuntimeError("zstd: " + L.ZSTD_getErrorName(code).decode())
    return code


def compress(data, level: int = 9) -> bytes:
    """One frame (the reference's default level for zstd outputs is 9, seqspec/src/utils.rs:33-35).  `data`: bytes, or a
    writable contiguous buffer (memoryview / numpy slice), which is compressed in place without a copy."""
    L = lib()
    n_in = len(data)
    cap = L.ZSTD_compressBound(n_in)
    out = C.create_string_buffer(cap)
    src = data if isinstance(data, (bytes, bytearray)) and not isinstance(data, bytearray) else (C.c_char * n_in).from_buffer(data)
    n = _check(L.ZSTD_compress(out, cap, src, n_in, level))
    return out.raw[:n]


def decompress(data: bytes) -> bytes:
    """All frames of a buffer (streaming API: content sizes need not be recorded)."""
    L = lib()
    ds = L.ZSTD_createDStream()
    try:
        _check(L.ZSTD_initDStream(ds))
        src = C.create_string_buffer(data, len(data))
        inb = _Buf(C.cast(src, C.c_void_p), len(data), 0)
        chunk = 1 << 20
        dst = C.create_string_buffer(chunk)
        parts = []
        while inb.pos < inb.size:
            outb = _Buf(C.cast(dst, C.c_void_p), chunk, 0)
            _check(L.ZSTD_decompressStream(ds, C.byref(outb), C.byref(inb)))
            parts.append(dst.raw[:outb.pos])
        # drain
        while True:
            outb = _Buf(C.cast(dst, C.c_void_p), chunk, 0)
            r = _check(L.ZSTD_decompressStream(ds, C.byref(outb), C.byref(inb)))
            if outb.pos:
                parts.append(dst.raw[:outb.pos])
            if outb.pos < chunk or r == 0:
                break
        return b"".join(parts)
    finally:
        L.ZSTD_freeDStream(ds)


def decompress_file(path: str) -> bytes:
    with open(path, "rb") as fh:
        return decompress(fh.read())


class Reader:
    """The decompressed bytes of a (multi-frame) .zst file as a stream: ``readinto`` fills the caller's buffer straight
    from ZSTD_decompressStream, the file is read 4 MB at a time.  What open_file's zstd::stream::Decoder does
    (seqspec/src/utils.rs:44-56), without holding the whole text in memory."""

    def __init__(self, path: str, piece: int = 4 << 20):
        self.L = lib()
        self.fh = open(path, "rb", buffering=0)
        self.path = path
        self.ds = self.L.ZSTD_createDStream()
        _check(self.L.ZSTD_initDStream(self.ds))
        self.src = C.create_string_buffer(piece)
        self.inb = _Buf(C.cast(self.src, C.c_void_p), 0, 0)
        self.file_eof = False
        self.mid_frame = False

    def readinto(self, mv) -> int:
        n = len(mv)
        if n == 0 or self.ds is None:
            return 0
        dst = (C.c_uint8 * n).from_buffer(mv)
        outb = _Buf(C.cast(dst, C.c_void_p), n, 0)
        while outb.pos < n:
            if self.inb.pos == self.inb.size and not self.file_eof:
                got = self.fh.readinto(self.src)
                self.inb = _Buf(C.cast(self.src, C.c_void_p), got or 0, 0)
                self.file_eof = not got
            before = (self.inb.pos, outb.pos)
            r = _check(self.L.ZSTD_decompressStream(self.ds, C.byref(outb), C.byref(self.inb)))
            if (self.inb.pos, outb.pos) != before:
                self.mid_frame = r != 0                  # 0: a frame was decoded and flushed completely
            elif self.file_eof and self.inb.pos == self.inb.size:
                if self.mid_frame:
                    raise RuntimeError("zstd: truncated frame in " + self.path)
                break
        return outb.pos

    def close(self) -> None:
        if self.ds is not None:
            self.L.ZSTD_freeDStream(self.ds)
            self.ds = None
        self.fh.close()


class Writer:
    """Append-only .zst writer: independent frames of up to `block` bytes (a valid multi-frame zstd file), compressed
    by `threads` workers (ZSTD_compress through ctypes releases the GIL) and written in order.  The reference writes
    its zstd outputs at level 9 with 8 worker threads (seqspec/src/utils.rs:33-36, python/src/lib.rs:138-142)."""

    def __init__(self, path: str, level: int = 9, block: int = 4 << 20, threads: int = 8):
        from concurrent.futures import ThreadPoolExecutor
        self.fh = open(path, "wb")
        self.level, self.block = level, block
        self.buf = bytearray()
        self.pool = ThreadPoolExecutor(max_workers=max(1, threads))
        self.pending = []                       # futures in file order
        self.max_pending = 4 * max(1, threads)

    def _drain(self, keep: int) -> None:
        while len(self.pending) > keep:
            self.fh.write(self.pending.pop(0).result())

    def _submit(self, chunk: bytes) -> None:
        self.pending.append(self.pool.submit(compress, chunk, self.level))
        self._drain(self.max_pending)

    def write(self, data) -> None:
        mv = memoryview(data).cast("B")
        if not self.buf and len(mv) >= self.block:          # large appends: frames straight out of the caller's buffer,
            k = 0                                           # no copy when it is writable (the slices keep it alive)
            while len(mv) - k >= self.block:
                piece = mv[k:k + self.block]
                self._submit(piece if not mv.readonly else bytes(piece))
                k += self.block
            mv = mv[k:]
        self.buf += mv
        while len(self.buf) >= self.block:
            self._submit(bytes(self.buf[:self.block]))
            del self.buf[:self.block]

    def close(self) -> None:
        if self.fh:
            if self.buf:
                self._submit(bytes(self.buf))
                self.buf = bytearray()
            self._drain(0)
            self.pool.shutdown(wait=True)
            self.fh.close()
            self.fh = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()
