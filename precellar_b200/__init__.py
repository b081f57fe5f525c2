"""precellar_b200 -- B200-native implementation of precellar's barcode hot path.

Segment extraction -> whitelist frequency count -> Phred-weighted 1-Hamming barcode correction, as
hand-written sm_100a CUDA kernels behind a C ABI (include/precellar_b200.h), with a host layer that keeps
precellar's API surface for this path (``Assay``, ``make_fastq``).  There is no CPU fallback.
"""
from .seqspec import Assay  # noqa: F401

__version__ = "0.1.0"


def __getattr__(name):
    # heavy / library-bound modules are imported on first use
    if name in ("BarcodePipeline", "HostBatch"):
        from . import pipeline
        return getattr(pipeline, name)
    if name == "compile_layout":
        from .layout import compile_layout
        return compile_layout
    if name in ("make_fastq", "FastqProcessor"):
        from . import api
        return getattr(api, name)
    raise AttributeError(name)
