"""Device ingest: raw FASTQ text -> chunk planes, parsed on the GPU (csrc/textparse.cuh).

The host side only moves bytes: the (decompressed) text of every read file is streamed into pinned blocks, the
device indexes the lines and fills the SoA planes (``pcl_text_scan`` / ``pcl_text_fill``).  What
``BatchedFqReader::next`` enforces (precellar/src/align/fastq.rs:400-448: the same number of records in every
file, equal names after stripping /1 /2) is enforced here too, the names by a device-side hash compare.

Blocks are double-buffered: while the device parses block k, a worker thread already reads block k+1 behind a
reserved gap, into which the unconsumed tail of block k is copied once it is known.
"""
from __future__ import annotations

import ctypes as C
import io
import os
from concurrent.futures import ThreadPoolExecutor
from typing import List, Optional, Sequence

import numpy as np

from . import _ffi
from ._ffi import lib
from .pipeline import BarcodePipeline, Chunk


# Pinned blocks are expensive to create (the driver pins and maps every page: about a second per GB here), so the blocks
# of a finished reader are parked and handed to the next one of this process instead of being freed.
_PINNED_POOL: List[tuple] = []                   # (nbytes, address)
_PINNED_POOL_MAX_BYTES = 2 << 30


def _pinned(nbytes: int) -> np.ndarray:
    nbytes = int(nbytes)
    best = None
    for k, (sz, _p) in enumerate(_PINNED_POOL):
        if nbytes <= sz <= 2 * nbytes + (1 << 20) and (best is None or sz < _PINNED_POOL[best][0]):
            best = k
    if best is not None:
        sz, p = _PINNED_POOL.pop(best)
    else:
        sz, p = nbytes, lib().pcl_host_alloc(nbytes)
        if not p:
            raise MemoryError(f"pinned allocation of {nbytes} bytes failed")
    return np.frombuffer((C.c_uint8 * sz).from_address(p), dtype=np.uint8), p


def _unpin(arr: np.ndarray, p: int) -> None:
    if not p:
        return
    if sum(sz for sz, _ in _PINNED_POOL) + len(arr) <= _PINNED_POOL_MAX_BYTES:
        _PINNED_POOL.append((len(arr), p))
    else:
        lib().pcl_host_free(p)


def release_pinned_pool() -> None:
    while _PINNED_POOL:
        lib().pcl_host_free(_PINNED_POOL.pop()[1])


class GzSource:
    """Decompressed bytes of one gzip file through pcl_gzsrc_* (csrc/gzsrc.cpp): BGZF blocks are inflated by
    several threads straight into the destination, any other gzip file by one zlib stream."""

    def __init__(self, path: str, n_threads: Optional[int] = None):
        h = C.c_void_p()
        nt = n_threads if n_threads is not None else max(1, min(8, (os.cpu_count() or 2) // 2))
        rc = lib().pcl_gzsrc_open(path.encode(), nt, C.byref(h))
        if rc < 0:
            raise _ffi.PclError(rc, f"cannot open file: {path}")
        self.h = h
        self.bgzf = bool(lib().pcl_gzsrc_is_bgzf(h))

    def readinto(self, mv) -> int:
        n = len(mv)
        if n == 0:
            return 0
        buf = (C.c_uint8 * n).from_buffer(mv)
        got = lib().pcl_gzsrc_read(self.h, C.addressof(buf), n)
        if got < 0:
            raise _ffi.PclError(int(got), lib().pcl_gzsrc_error(self.h).decode())
        return int(got)

    def close(self) -> None:
        if self.h:
            lib().pcl_gzsrc_close(self.h)
            self.h = None


def _open_raw(path: str):
    """open_file (seqspec/src/utils.rs:44-56): gzip by magic, zstd by extension, else plain."""
    with open(path, "rb") as f:
        magic = f.read(2)
    if magic == b"\x1f\x8b":
        return GzSource(path)
    if path.endswith(".zst"):
        from . import _zstd
        return _zstd.Reader(path)
    return open(path, "rb", buffering=0)


_IO_SLICE = 8 << 20
_io_pool: Optional[ThreadPoolExecutor] = None


def _io_workers() -> ThreadPoolExecutor:
    """Shared pool for sliced reads of plain files: one core copies page cache into pinned memory at a few GB/s,
    well below what the PCIe link takes, so a block is read by several threads at once."""
    global _io_pool
    if _io_pool is None:
        _io_pool = ThreadPoolExecutor(max_workers=max(2, min(8, (os.cpu_count() or 2) // 2)))
    return _io_pool


class TextSource:
    """The concatenated, decompressed bytes of the files of one read, as a sliding block in pinned memory."""

    def __init__(self, paths: Sequence[str], block_bytes: int):
        if not paths:
            raise _ffi.PclError(_ffi.PCL_ERR_INPUT, "no FASTQ files")
        self.paths = list(paths)
        self.next_path = 0
        self.fh = None
        self.last_byte = 0x0A
        self.eof = False
        self.cap = int(block_bytes)
        self._alloc(2 * self.cap)
        self.cur = 0
        self.start = self.end = self.reserve
        self.pending = None

    def _alloc(self, size: int) -> None:
        self.bufs, self.ptrs = [], []
        for _ in range(2):
            a, p = _pinned(int(size))
            self.bufs.append(a)
            self.ptrs.append(p)
        self.size = min(len(a) for a in self.bufs)       # a block out of the pool may be larger than asked for
        self.reserve = self.size // 2

    def close(self) -> None:
        if self.pending is not None:
            try:
                self.pending.result()
            except Exception:
                pass
            self.pending = None
        if self.fh is not None:
            self.fh.close()
            self.fh = None
        for a, p in zip(self.bufs, self.ptrs):
            _unpin(a, p)
        self.ptrs, self.bufs = [], []

    def _read_into(self, which: int, offset: int, want: int) -> int:
        """Read up to `want` bytes into bufs[which][offset:]; a file that does not end in a newline gets one, so the
        next file of the read starts on its own line (Read::open chains the files, read.rs:98-108)."""
        mv = memoryview(self.bufs[which])
        got = 0
        while got < want and not self.eof:
            if self.fh is None:
                if self.next_path >= len(self.paths):
                    self.eof = True
                    break
                self.fh = _open_raw(self.paths[self.next_path])
                self.next_path += 1
            if isinstance(self.fh, io.FileIO) and want - got > _IO_SLICE:
                n = self._pread_sliced(mv, offset + got, want - got)
            else:
                n = self.fh.readinto(mv[offset + got:offset + want])
            if not n:
                self.fh.close()
                self.fh = None
                if self.last_byte != 0x0A:
                    mv[offset + got] = 0x0A
                    got += 1
                    self.last_byte = 0x0A
                continue
            got += n
            self.last_byte = mv[offset + got - 1]
        return got

    def _pread_sliced(self, mv: memoryview, offset: int, want: int) -> int:
        fd, pos = self.fh.fileno(), self.fh.tell()
        left = os.fstat(fd).st_size - pos
        want = min(want, left)
        if want <= 0:
            return 0
        futs = [_io_workers().submit(os.preadv, fd, [mv[offset + o:offset + min(o + _IO_SLICE, want)]], pos + o)
                for o in range(0, want, _IO_SLICE)]
        got = 0
        for o, f in zip(range(0, want, _IO_SLICE), futs):
            n = f.result()
            if n != min(_IO_SLICE, want - o):
                raise _ffi.PclError(_ffi.PCL_ERR_INPUT, f"short read from {self.paths[self.next_path - 1]}")
            got += n
        self.fh.seek(pos + got)
        return got

    def block(self) -> np.ndarray:
        return self.bufs[self.cur][self.start:self.end]

    def prefetch(self, pool: ThreadPoolExecutor, want: int) -> None:
        """Start reading the next block into the other buffer, behind the gap kept for the tail of this one."""
        assert self.pending is None
        want = max(0, min(int(want), self.size - self.reserve))
        self.pending = pool.submit(self._read_into, 1 - self.cur, self.reserve, want)

    def advance(self, used: int) -> None:
        """Drop `used` bytes from the front of the block and append what the prefetch read."""
        got = self.pending.result() if self.pending is not None else 0
        self.pending = None
        cur, other = self.bufs[self.cur], self.bufs[1 - self.cur]
        lo = self.start + int(used)
        tail = self.end - lo
        if tail <= self.reserve:
            other[self.reserve - tail:self.reserve] = cur[lo:self.end]
            self.start, self.end = self.reserve - tail, self.reserve + got
            self.cur = 1 - self.cur
            return
        # this read is far ahead of the others: rebuild the block at the front of a (possibly larger) buffer
        need = tail + got
        fresh = np.concatenate([cur[lo:self.end], other[self.reserve:self.reserve + got]])
        if need > self.size - self.size // 2:
            old = list(zip(self.bufs, self.ptrs))
            self._alloc(4 * max(need, self.cap))
            for a, p in old:
                _unpin(a, p)
            self.cur = 0
        else:
            self.cur = 1 - self.cur
        self.bufs[self.cur][self.reserve - 0:self.reserve + need] = fresh
        self.start, self.end = self.reserve, self.reserve + need


class _Done:
    """Stands in for the future of a prefetch that needs no I/O."""
    def result(self):
        return 0


class MmapSource:
    """The files of one read as windows of their memory maps: uncompressed files that end in a newline need no staging
    copy at all - the device copy reads the page cache (pageable memory) directly.  Same interface as TextSource; a
    window never spans two files (the chunk at a file boundary is simply shorter)."""

    def __init__(self, paths: Sequence[str], block_bytes: int):
        import mmap
        self.paths = list(paths)
        self.maps, self.views = [], []
        for p in self.paths:
            fh = open(p, "rb")
            try:
                m = mmap.mmap(fh.fileno(), 0, access=mmap.ACCESS_READ)
            finally:
                fh.close()
            self.maps.append(m)
            self.views.append(np.frombuffer(m, dtype=np.uint8))
        self.k = 0
        self.start = self.end = 0
        self.pending = None
        self._want = 0
        self.cap = int(block_bytes)
        self.eof = False                         # like TextSource: set by the advance that found nothing left to add

    @staticmethod
    def eligible(paths: Sequence[str]) -> bool:
        if not paths:
            return False
        for p in paths:
            try:
                with open(p, "rb") as fh:
                    head = fh.read(2)
                    size = os.fstat(fh.fileno()).st_size
                    if size == 0 or head == b"\x1f\x8b" or p.endswith(".zst"):
                        return False
                    fh.seek(size - 1)
                    if fh.read(1) != b"\n":
                        return False
            except OSError:
                return False
        return True

    def bytes_per_record(self, default: float) -> float:
        """Mean size of the first (up to 64) records of the first file."""
        head = self.views[0][:1 << 16]
        nl = np.flatnonzero(head == 10)
        k = min(len(nl) // 4, 64)
        return float(nl[4 * k - 1] + 1) / k if k else float(default)


    def block(self) -> np.ndarray:
        return self.views[self.k][self.start:self.end]

    def prefetch(self, pool, want: int) -> None:
        assert self.pending is None
        self._want = max(0, int(want))
        self.pending = _Done()

    def advance(self, used: int) -> None:
        want, self._want, self.pending = (self._want if self.pending is not None else 0), 0, None
        self.start += int(used)
        size = len(self.views[self.k])
        if self.k + 1 >= len(self.views) and self.end >= size:
            self.eof = True
        if self.start >= size and self.k + 1 < len(self.views):          # this file is consumed: on to the next one
            self.k += 1
            self.start = self.end = 0
            size = len(self.views[self.k])
        self.end = min(size, max(self.end, self.start) + want)

    def close(self) -> None:
        self.views = []
        for m in self.maps:
            try:
                m.close()
            except BufferError:                  # a view is still referenced somewhere: the map goes with it
                pass
        self.maps = []


class DevBlock:
    """A text block that already sits in device memory (a window of a BgzfDeviceSource buffer)."""

    def __init__(self, backend, slot: int, base: int, start: int, n: int):
        self.backend, self.slot, self.start, self.n = backend, slot, int(start), int(n)
        self.ptr = int(base) + int(start)

    def __len__(self) -> int:
        return self.n

    def __bytes__(self) -> bytes:
        return self.backend.fetch(self.slot, self.start, self.n)


class DeviceInflater:
    """pcl_inflater_* of one ctx: device text buffers ("slots") and the kernel that inflates BGZF members into them."""

    def __init__(self, ctx):
        self.h = C.c_void_p()
        _ffi.check(ctx, lib().pcl_inflater_create(ctx, C.byref(self.h)))

    def _check(self, rc: int) -> None:
        if rc < 0:
            raise _ffi.PclError(rc, lib().pcl_inflater_error(self.h).decode(errors="replace"))

    def buffer(self, slot: int, nbytes: int) -> int:
        p = C.c_void_p()
        self._check(lib().pcl_inflater_buffer(self.h, slot, int(nbytes), C.byref(p)))
        return p.value or 0

    def run(self, comp: np.ndarray, members, n_members: int, slot: int, dst_off: int) -> None:
        self._check(lib().pcl_inflater_run(self.h, comp.ctypes.data, len(comp), members, n_members, slot, int(dst_off)))

    def move(self, src_slot: int, src_off: int, dst_slot: int, dst_off: int, n: int) -> None:
        self._check(lib().pcl_inflater_move(self.h, src_slot, int(src_off), dst_slot, int(dst_off), int(n)))

    def fetch(self, slot: int, off: int, n: int) -> bytes:
        buf = C.create_string_buffer(max(int(n), 1))
        self._check(lib().pcl_inflater_fetch(self.h, slot, int(off), int(n), buf))
        return buf.raw[:int(n)]

    def poke(self, slot: int, off: int, data: bytes) -> None:
        self._check(lib().pcl_inflater_poke(self.h, slot, int(off), data, len(data)))

    def stats(self) -> dict:
        ms, nb, nl = C.c_double(), C.c_uint64(), C.c_uint64()
        lib().pcl_inflater_stats(self.h, C.byref(ms), C.byref(nb), C.byref(nl))
        return {"kernel_ms": ms.value, "out_bytes": nb.value, "launches": nl.value}

    def close(self) -> None:
        if self.h:
            lib().pcl_inflater_destroy(self.h)
            self.h = None


def device_inflate_enabled() -> bool:
    """BGZF input is inflated on the device unless PCL_DEVICE_INFLATE=0 (then: the host threads of csrc/gzsrc.cpp)."""
    return os.environ.get("PCL_DEVICE_INFLATE", "1") != "0"


def bgzf_walk(view: np.ndarray, want_out: int, members, cap: int):
    """pcl_bgzf_walk over the bytes of `view`: (number of members, compressed bytes consumed, inflated bytes)."""
    n, used, out = C.c_uint32(), C.c_uint64(), C.c_uint64()
    rc = lib().pcl_bgzf_walk(view.ctypes.data if len(view) else None, len(view), int(want_out), members, cap,
                             C.byref(n), C.byref(used), C.byref(out))
    return rc, n.value, used.value, out.value


INFLATED_STATS = {"kernel_ms": 0.0, "out_bytes": 0, "launches": 0}      # summed over the closed sources of this process


class BgzfDeviceSource:
    """The files of one read, all BGZF, as a sliding block in DEVICE memory: the compressed bytes are read from the files'
    memory maps, copied to the device as they are and inflated there (csrc/inflate.cuh), one warp per member.  Same
    interface and the same double-buffer discipline as TextSource; block() is a DevBlock.

    A batch always holds whole members, so a read may overshoot what was asked for by less than one member (64 KB):
    the buffers carry that much slack."""

    SLACK = 1 << 17
    MEMBERS_CAP = 1 << 15
    # A launch of the inflate kernel lasts as long as its slowest member takes one warp (about 4 ms for 64 KB), whatever the
    # number of members up to the 4 736 warps a B200 holds: the source inflates up to this much text per launch (a whole
    # file of a few hundred MB at once) and then nothing until fewer than two chunks are buffered (TextGroupReader._want).
    AHEAD_BYTES = 256 << 20

    def __init__(self, paths: Sequence[str], block_bytes: int, backend, ahead_bytes: int = 0):
        import mmap
        self.paths = list(paths)
        self.backend = backend
        self.maps, self.views = [], []
        for p in self.paths:
            with open(p, "rb") as fh:
                m = mmap.mmap(fh.fileno(), 0, access=mmap.ACCESS_READ)
            self.maps.append(m)
            self.views.append(np.frombuffer(m, dtype=np.uint8))
        # text inflated per launch: no more than the files can hold (FASTQ text deflates to a fifth; 8 is generous)
        self.ahead_bytes = min(int(ahead_bytes), 8 * sum(len(v) for v in self.views) + (1 << 16)) if ahead_bytes else 0
        block_bytes = max(int(block_bytes), self.ahead_bytes)
        self.k = 0                               # file being read, and the position in it
        self.fpos = 0
        self.last_byte = 0x0A
        self.eof = False
        self.cap = int(block_bytes)
        self.members = (_ffi.BgzfMember * self.MEMBERS_CAP)()
        self.slots = [0, 1]
        self.base = [0, 0]
        self._alloc(self.slots, 2 * self.cap)
        self.cur = 0
        self.start = self.end = self.reserve
        self.pending = None

    @staticmethod
    def eligible(paths: Sequence[str]) -> bool:
        """Every file starts with a BGZF member (gzip magic, FEXTRA with a 'BC' field)."""
        if not paths:
            return False
        members = (_ffi.BgzfMember * 1)()
        for p in paths:
            try:
                with open(p, "rb") as fh:
                    head = np.frombuffer(fh.read(1 << 16), dtype=np.uint8)
                    size = os.fstat(fh.fileno()).st_size
            except OSError:
                return False
            if size == 0 or len(head) < 18 or head[0] != 0x1F or head[1] != 0x8B:
                return False
            rc, _n, used, _out = bgzf_walk(head, 1, members, 1)
            if rc < 0 or (used == 0 and size <= len(head)):
                return False
        return True

    def _alloc(self, slots, size: int) -> None:
        self.size = int(size)
        self.reserve = self.size // 2
        for k, s in enumerate(slots):
            self.base[k] = self.backend.buffer(s, self.size + self.SLACK + 16)

    def close(self) -> None:
        if self.pending is not None:
            try:
                self.pending.result()
            except Exception:
                pass
            self.pending = None
        if self.backend is not None:
            if hasattr(self.backend, "stats"):
                for key, v in self.backend.stats().items():
                    INFLATED_STATS[key] += v
            self.backend.close()
            self.backend = None
        self.views = []
        for m in self.maps:
            try:
                m.close()
            except BufferError:
                pass
        self.maps = []

    def _read_into(self, which: int, offset: int, want: int) -> int:
        """Inflate whole members into buffer `which` from `offset` on until at least `want` bytes are there (or the input
        ends); a file that does not end in a newline gets one (Read::open chains the files, read.rs:98-108)."""
        slot = self.slots[which]
        got = 0
        while got < want and not self.eof:
            if self.k >= len(self.views):
                self.eof = True
                break
            v = self.views[self.k]
            if self.fpos >= len(v):                                   # this file is done
                if self.last_byte != 0x0A:
                    self.backend.poke(slot, offset + got, b"\n")
                    got += 1
                    self.last_byte = 0x0A
                self.k += 1
                self.fpos = 0
                continue
            rest = v[self.fpos:]
            rc, n, used, out = bgzf_walk(rest, want - got, self.members, self.MEMBERS_CAP)
            if rc < 0 or used == 0:
                raise _ffi.PclError(_ffi.PCL_ERR_INPUT, f"truncated or non-BGZF block in {self.paths[self.k]}")
            if n:
                self.backend.run(rest[:used], self.members, n, slot, offset + got)
                got += out
                # the last byte so far: what is left of the file may be empty members only (the end-of-file marker)
                self.last_byte = self.backend.fetch(slot, offset + got - 1, 1)[0]
            self.fpos += used
        return got

    def block(self, limit: Optional[int] = None) -> DevBlock:
        """The buffered text, or its first `limit` bytes (the reader indexes one chunk's worth, not everything inflated ahead)."""
        n = self.end - self.start
        if limit is not None:
            n = min(n, int(limit))
        return DevBlock(self.backend, self.slots[self.cur], self.base[self.cur], self.start, n)

    def prefetch(self, pool: ThreadPoolExecutor, want: int) -> None:
        assert self.pending is None
        want = max(0, min(int(want), self.size - self.reserve))
        self.pending = pool.submit(self._read_into, 1 - self.cur, self.reserve, want)

    def advance(self, used: int) -> None:
        """Drop `used` bytes from the front of the block and append what the prefetch inflated."""
        got = self.pending.result() if self.pending is not None else 0
        self.pending = None
        if got == 0:                                     # nothing new behind the gap of the other buffer: the block shrinks in place
            self.start += int(used)
            return
        lo = self.start + int(used)
        tail = self.end - lo
        cur, other = self.slots[self.cur], self.slots[1 - self.cur]
        if tail <= self.reserve:
            self.backend.move(cur, lo, other, self.reserve - tail, tail)
            self.start, self.end = self.reserve - tail, self.reserve + got
            self.cur = 1 - self.cur
            return
        # this read is far ahead of the others: rebuild the block at the front half of a fresh (possibly larger) pair
        need = tail + got
        spare = [s for s in range(4) if s not in self.slots]
        old = list(self.slots)
        old_reserve = self.reserve
        size = self.size if need <= self.size - self.size // 2 else 4 * max(need, self.cap)
        self.slots = spare
        self._alloc(spare, size)
        self.backend.move(cur, lo, spare[0], self.reserve, tail)
        self.backend.move(other, old_reserve, spare[0], self.reserve + tail, got)
        for s in old:
            self.backend.buffer(s, 0)
        self.cur = 0
        self.start, self.end = self.reserve, self.reserve + need


class TextGroupReader:
    """Fills chunks of a pipeline from the FASTQ files of a read-group, one record per file per group."""

    def __init__(self, pipe: BarcodePipeline, files_per_read: Sequence[Sequence[str]], records_per_chunk: int,
                 bytes_per_record: Optional[Sequence[int]] = None):
        self.pipe = pipe
        self.n_reads = len(files_per_read)
        self.target = int(records_per_chunk)
        guess = list(bytes_per_record) if bytes_per_record is not None else [400] * self.n_reads
        self.bpr = [float(g) for g in guess]
        self.pool = ThreadPoolExecutor(max_workers=max(1, self.n_reads))
        self.sources = []
        for r, f in enumerate(files_per_read):
            if MmapSource.eligible(f):
                src = MmapSource(f, 0)
                self.bpr[r] = src.bytes_per_record(self.bpr[r])          # measured on the first records, not guessed
            else:
                src = None
                if device_inflate_enabled() and BgzfDeviceSource.eligible(f):
                    inf = DeviceInflater(pipe.ctx)
                    try:
                        # room for the two chunks that may still be buffered when the next batch is inflated behind them
                        src = BgzfDeviceSource(f, int(self.target * self.bpr[r] * 2.2) + (1 << 16), inf, BgzfDeviceSource.AHEAD_BYTES)
                    except _ffi.PclError as exc:                     # no room for the device text buffers: the host threads inflate
                        inf.close()
                        if exc.code != _ffi.PCL_ERR_NOMEM:
                            self.close()
                            raise
                if src is None:
                    src = TextSource(f, int(self.target * self.bpr[r] * 1.25) + (1 << 16))
            self.sources.append(src)
        self.first = True
        self.records = 0

    def close(self) -> None:
        for s in self.sources:
            s.close()
        self.pool.shutdown(wait=True)

    def _want(self, r: int, blocks_ahead: int = 1) -> int:
        """Bytes to read so that, after `blocks_ahead - 1` chunks were cut off the front, a full chunk is buffered."""
        have = self.sources[r].end - self.sources[r].start
        per = self.target * self.bpr[r]
        ahead = getattr(self.sources[r], "ahead_bytes", 0)
        if ahead:                                        # device inflate: big batches, and none while two chunks are buffered
            if blocks_ahead > 1 and have >= int(per * 2.03) + 4096:
                return 0
            return max(0, max(ahead, int(per * (blocks_ahead + 0.03)) + 4096) - have)
        return max(0, int(per * (blocks_ahead + 0.03)) + 4096 - have)

    def next_into(self, chunk: Chunk) -> int:
        """Parse the next records into `chunk` (at most its capacity / the reader's target); 0 at the end of input."""
        cap = min(chunk.capacity, self.target)
        if self.first:                                   # nothing read yet: fetch the first blocks synchronously
            for r, s in enumerate(self.sources):
                s.prefetch(self.pool, self._want(r))
            for s in self.sources:
                s.advance(0)
            self.first = False
        grow = 1
        while True:
            # a source that inflates ahead hands over one chunk's worth of its text (more when no whole group was in it)
            blocks = [s.block(int(self.target * self.bpr[r] * 1.03 * grow) + 4096) if getattr(s, "ahead_bytes", 0) else s.block()
                      for r, s in enumerate(self.sources)]
            for r, s in enumerate(self.sources):
                if s.pending is None:
                    s.prefetch(self.pool, self._want(r, 2) if grow == 1 else grow * (1 << 20))
            for r, b in enumerate(blocks):
                chunk.text_scan(r, b)
            lines = [chunk.text_lines(r) for r in range(self.n_reads)]
            n = min(min(l // 4 for l in lines), cap)
            if n > 0:
                break
            # no complete group in the blocks: either the input ended or a record is larger than the block
            for s in self.sources:
                s.advance(0)
            if any(len(b) < s.end - s.start for b, s in zip(blocks, self.sources)):       # only a window was looked at: widen it
                grow *= 2
                continue
            if all(s.eof for s in self.sources):
                rest = [bytes(s.block()).strip() for s in self.sources]
                if all(not x for x in rest):
                    return 0
                if any(not x for x in rest) or any(l // 4 for l in lines):
                    raise _ffi.PclError(_ffi.PCL_ERR_INPUT, "Unequal number of reads in the chunk")     # fastq.rs:435-436
                raise _ffi.PclError(_ffi.PCL_ERR_INPUT, "truncated FASTQ record at the end of the input")
            grow *= 2
        chunk.reset(n)
        used = [chunk.text_fill(r) for r in range(self.n_reads)]
        bad = chunk.text_names_check()
        if bad >= 0:
            raise _ffi.PclError(_ffi.PCL_ERR_INPUT, f"read names mismatch (group {self.records + bad})")   # fastq.rs:438-442
        for r, (s, u) in enumerate(zip(self.sources, used)):
            self.bpr[r] = 0.5 * self.bpr[r] + 0.5 * (u / n)
            s.advance(u)
        self.records += n
        return n
