"""GPU tier: the block coder of the device writer (k_zh_blocks / k_zh_compact through pcl_zstd_frame) gives the same bytes
as its host emulation (tests/emul, checked against libzstd in the CPU tier) and decodes with libzstd."""
import ctypes as C

import numpy as np
import pytest

from precellar_b200 import _ffi, _zstd
from tests.test_zhuff import fastq_text, illumina_text, zh_frame

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    L = _ffi.lib()
    h = C.c_void_p()
    _ffi.check(None, L.pcl_ctx_create(C.byref(h), 0))
    yield h
    L.pcl_ctx_destroy(h)


def device_frame(ctx, data: bytes) -> bytes:
    L = _ffi.lib()
    cap = 14 + len(data) + 3 * (len(data) // 131072 + 1) + 64
    out = np.zeros(cap, np.uint8)
    src = np.frombuffer(data, np.uint8) if data else np.zeros(1, np.uint8)
    used = C.c_uint64()
    _ffi.check(ctx, L.pcl_zstd_frame(ctx, src.ctypes.data, len(data), out.ctypes.data, cap, C.byref(used)))
    return out[:used.value].tobytes()


def test_frames_equal_the_emulation_and_decode(ctx):
    rng = np.random.default_rng(12)
    text = fastq_text(rng, 9000)
    noise = rng.integers(0, 256, 300000, dtype=np.uint8).tobytes()
    w = 2.0 ** (-0.9 * np.arange(40))
    skew = rng.choice(np.arange(40, 80), 400000, p=w / w.sum()).astype(np.uint8).tobytes()
    ill = illumina_text(rng, 5000)
    cases = [ill, ill[:131072 + 9], b"@same name here\nACGTACGTACGTACGTACGT\n+\nFFFFFFFFFFFFFFFFFFFF\n" * 4000,
             b"".join(b"ab%d\n" % (i % 7) for i in range(60000)), text, text[:131072], text[:131073], text[:1024], text[:1023], text[:5], b"", noise, skew, b"A" * 200000,
             text[:700000] + noise[:1000] + text[:300000]]
    for k, data in enumerate(cases):
        got = device_frame(ctx, data)
        assert _zstd.decompress(got) == data, k
        want, _ = zh_frame(data)
        assert got == want, k
    big = fastq_text(rng, 60000) + illumina_text(rng, 40000)      # ~ 20 MB: a few hundred CTAs
    got = device_frame(ctx, big)
    assert _zstd.decompress(got) == big and len(got) < 0.6 * len(big)
    assert got == zh_frame(big)[0]
