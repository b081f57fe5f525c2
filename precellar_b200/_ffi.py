"""ctypes binding of the C ABI in include/precellar_b200.h (libprecellar_b200.so).

The shared library is built in-tree by ``precellar_b200.build`` / ``__graft_entry__.build()``.
There is no fallback: if the library is missing or no CUDA device is present, calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libprecellar_b200.so")

PCL_MAX_READS = 8
PCL_MAX_ELEMS = 16
PCL_MAX_ANCHOR = 64
PCL_MAX_REGIONS = 8
PCL_MAX_BC_PER_READ = 4
PCL_MAX_UMI_PER_READ = 4
PCL_MAX_KEY_LEN = 31

PCL_OK, PCL_ERR_ARG, PCL_ERR_CUDA, PCL_ERR_NCCL, PCL_ERR_UNSUPPORTED, PCL_ERR_INPUT, PCL_ERR_NOMEM = 0, -1, -2, -3, -4, -5, -6
KIND_OTHER, KIND_BARCODE, KIND_UMI, KIND_TARGET = 0, 1, 2, 3
SPLIT_OK, SPLIT_ANCHOR_NOT_FOUND, SPLIT_PATTERN_MISMATCH, SPLIT_MULTI_TARGET = 0, 1, 2, 3
FSTATUS_TARGET = 0x4
BC_EXACT, BC_ABSENT, BC_NO_MATCH, BC_CORRECTED, BC_LOW_CONF, BC_EXCEED_EE = 0, 3, 4, 5, 6, 7
KEY_ABSENT32, KEY_ABSENT64, COORD_ABSENT = 0xFFFFFFFF, 0xFFFFFFFFFFFFFFFF, 0xFFFFFFFF
K_COUNT, K_CORRECT, K_COUNT_LENONLY, K_ENUM, K_QC, K_TEXT_INDEX, K_TEXT_FILL, K_FILTER, K_EMIT, K_ZSTD = 0, 1, 2, 3, 4, 5, 6, 7, 8, 9
F64_MAX = 1.7976931348623157e308


class ElemDesc(C.Structure):
    _fields_ = [("kind", C.c_uint8), ("is_fixed_seq", C.c_uint8), ("min_len", C.c_uint16), ("max_len", C.c_uint16),
                ("region_slot", C.c_int16), ("seq_len", C.c_uint8), ("reserved", C.c_uint8 * 3),
                ("sequence", C.c_char * PCL_MAX_ANCHOR)]


class ReadDesc(C.Structure):
    _fields_ = [("n_elems", C.c_uint16), ("is_reverse", C.c_uint8), ("reserved", C.c_uint8), ("max_len", C.c_uint32),
                ("elems", ElemDesc * PCL_MAX_ELEMS)]


ROUTE_WALK, ROUTE_FAST, ROUTE_FAST_SPEC, ROUTE_PLAN, ROUTE_PLAN_GEO, ROUTE_LENONLY = range(6)     # pcl_read_info.route


class ReadInfo(C.Structure):
    _fields_ = [("n_bc", C.c_uint8), ("n_umi", C.c_uint8), ("has_target", C.c_uint8), ("fixed_layout", C.c_uint8),
                ("needs_payload", C.c_uint8), ("qual_only", C.c_uint8), ("route", C.c_uint8), ("reserved", C.c_uint8),
                ("bc_elem", C.c_uint8 * PCL_MAX_BC_PER_READ), ("bc_key_bytes", C.c_uint8 * PCL_MAX_BC_PER_READ),
                ("umi_elem", C.c_uint8 * PCL_MAX_UMI_PER_READ), ("umi_key_bytes", C.c_uint8 * PCL_MAX_UMI_PER_READ),
                ("static_start", C.c_uint16 * PCL_MAX_ELEMS), ("qual_offset", C.c_uint16), ("qual_stride", C.c_uint16)]


class ReadResults(C.Structure):
    _fields_ = [("fstatus", C.c_void_p), ("tcoord", C.c_void_p),
                ("bc_key", C.c_void_p * PCL_MAX_BC_PER_READ), ("umi_key", C.c_void_p * PCL_MAX_UMI_PER_READ),
                ("bcoord", C.c_void_p * PCL_MAX_BC_PER_READ), ("ucoord", C.c_void_p * PCL_MAX_UMI_PER_READ)]


class HostRead(C.Structure):
    _fields_ = [("text", C.c_void_p), ("off", C.c_void_p), ("len", C.c_void_p), ("res", ReadResults)]


class EmitResult(C.Structure):
    _fields_ = [("r1", C.c_void_p), ("r1_bytes", C.c_uint64), ("r1_raw_bytes", C.c_uint64),
                ("r2", C.c_void_p), ("r2_bytes", C.c_uint64), ("r2_raw_bytes", C.c_uint64), ("n_written", C.c_uint64)]


class GroupInfo(C.Structure):
    _fields_ = [("n_records", C.c_uint64), ("n_assigned", C.c_uint64), ("n_ungroupable", C.c_uint64),
                ("n_barcodes", C.c_uint64), ("n_fragments", C.c_uint64)]


class BgzfMember(C.Structure):
    _fields_ = [("src_off", C.c_uint32), ("clen", C.c_uint32), ("dst_off", C.c_uint32), ("isize", C.c_uint32), ("crc", C.c_uint32)]


class PclError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"precellar_b200 error {code}: {msg}")
        self.code = code


# every symbol include/precellar_b200.h declares (tests check that the library exports all of them)
SYMBOLS = [
    "pcl_abi_version", "pcl_last_error", "pcl_device_count", "pcl_ctx_create", "pcl_ctx_destroy", "pcl_sync",
    "pcl_comm_unique_id", "pcl_comm_init", "pcl_layout_set", "pcl_read_info_get", "pcl_whitelist_load",
    "pcl_whitelist_size", "pcl_chunk_create", "pcl_chunk_destroy", "pcl_chunk_reset", "pcl_read_stride",
    "pcl_chunk_upload", "pcl_chunk_device_ptrs", "pcl_chunk_stream", "pcl_counts_reset", "pcl_count",
    "pcl_counts_allreduce", "pcl_correct", "pcl_chunk_fetch", "pcl_counts_fetch", "pcl_qc_fetch", "pcl_key_decode",
    "pcl_host_alloc", "pcl_host_free", "pcl_profile_enable", "pcl_profile_get", "pcl_profile_reset",
    "pcl_launch_count", "pcl_timer_start", "pcl_timer_stop_ms",
    "pcl_qc_enable", "pcl_qc_accumulate", "pcl_qc_fetch_categories",
    "pcl_fastq_open", "pcl_fastq_close", "pcl_fastq_error", "pcl_fastq_read", "pcl_fastq_text_full", "pcl_chunk_fetch_async",
    "pcl_make_fastq_batch",
    "pcl_text_scan", "pcl_text_lines", "pcl_text_fill", "pcl_text_names_check", "pcl_text_line_ends",
    "pcl_chunk_upload_compact", "pcl_read_payload_bytes",
    "pcl_gzsrc_open", "pcl_gzsrc_read", "pcl_gzsrc_is_bgzf", "pcl_gzsrc_error", "pcl_gzsrc_close",
    "pcl_counts_freeze", "pcl_chunk_bytes", "pcl_mem_info", "pcl_verify_enable", "pcl_chunk_set_uniform_len", "pcl_chunk_uniform_results_async",
    "pcl_group_build", "pcl_group_build_multi", "pcl_group_info_get", "pcl_group_fetch", "pcl_group_key_decode", "pcl_group_destroy",
    "pcl_text_keep", "pcl_emit_fastq", "pcl_zstd_frame",
    "pcl_bgzf_walk", "pcl_inflater_create", "pcl_inflater_destroy", "pcl_inflater_error", "pcl_inflater_buffer", "pcl_inflater_run",
    "pcl_inflater_move", "pcl_inflater_fetch", "pcl_inflater_poke", "pcl_inflater_stats", "pcl_text_scan_device",
]

_lib = None


def lib() -> C.CDLL:
    """Load libprecellar_b200.so (raises if it has not been built: there is no fallback path)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} not found: build it with `python -m precellar_b200.build` "
                          "(the CUDA library is the only implementation; there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, i32, u32, u64, f64 = C.c_void_p, C.c_int, C.c_uint32, C.c_uint64, C.c_double
    sig = {
        "pcl_abi_version": (i32, []),
        "pcl_last_error": (C.c_char_p, [vp]),
        "pcl_device_count": (i32, []),
        "pcl_ctx_create": (i32, [C.POINTER(vp), i32]),
        "pcl_ctx_destroy": (i32, [vp]),
        "pcl_sync": (i32, [vp]),
        "pcl_comm_unique_id": (i32, [vp]),
        "pcl_comm_init": (i32, [vp, i32, i32, vp]),
        "pcl_layout_set": (i32, [vp, C.POINTER(ReadDesc), i32]),
        "pcl_read_info_get": (i32, [vp, i32, C.POINTER(ReadInfo)]),
        "pcl_whitelist_load": (i32, [vp, i32, vp, vp, u64]),
        "pcl_whitelist_size": (u64, [vp, i32]),
        "pcl_chunk_create": (i32, [vp, u64, C.POINTER(vp)]),
        "pcl_chunk_destroy": (i32, [vp]),
        "pcl_chunk_reset": (i32, [vp, u64]),
        "pcl_read_stride": (u32, [vp, i32]),
        "pcl_chunk_upload": (i32, [vp, i32, vp, vp, vp]),
        "pcl_chunk_upload_compact": (i32, [vp, i32, vp, u32, vp, vp]),
        "pcl_read_payload_bytes": (u32, [vp, i32]),
        "pcl_chunk_device_ptrs": (i32, [vp, i32, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]),
        "pcl_chunk_stream": (vp, [vp]),
        "pcl_counts_reset": (i32, [vp]),
        "pcl_count": (i32, [vp, vp]),
        "pcl_counts_allreduce": (i32, [vp]),
        "pcl_correct": (i32, [vp, vp, f64, f64]),
        "pcl_chunk_fetch": (i32, [vp, i32, C.POINTER(ReadResults)]),
        "pcl_chunk_fetch_async": (i32, [vp, i32, C.POINTER(ReadResults)]),
        "pcl_counts_fetch": (i32, [vp, i32, vp, C.POINTER(u64), C.POINTER(u64)]),
        "pcl_qc_fetch": (i32, [vp, vp]),
        "pcl_key_decode": (i32, [u64, C.c_char_p]),
        "pcl_host_alloc": (vp, [u64]),
        "pcl_host_free": (None, [vp]),
        "pcl_profile_enable": (i32, [vp, i32]),
        "pcl_profile_get": (i32, [vp, i32, C.POINTER(f64), C.POINTER(u64)]),
        "pcl_profile_reset": (i32, [vp]),
        "pcl_launch_count": (u64, [vp]),
        "pcl_timer_start": (i32, [vp]),
        "pcl_timer_stop_ms": (i32, [vp, C.POINTER(f64)]),
        "pcl_qc_enable": (i32, [vp, i32]),
        "pcl_qc_accumulate": (i32, [vp, vp]),
        "pcl_qc_fetch_categories": (i32, [vp, vp]),
        "pcl_make_fastq_batch": (i32, [vp, C.POINTER(HostRead), u64, i32, i32, vp, u64, C.POINTER(u64), vp, u64,
                                       C.POINTER(u64), C.POINTER(u64)]),
        "pcl_text_scan": (i32, [vp, i32, vp, u64]),
        "pcl_text_lines": (i32, [vp, i32, C.POINTER(u64)]),
        "pcl_text_fill": (i32, [vp, i32, C.POINTER(u64)]),
        "pcl_text_names_check": (i32, [vp, C.POINTER(C.c_int64)]),
        "pcl_text_line_ends": (i32, [vp, i32, vp, u64]),
        "pcl_text_keep": (i32, [vp, i32]),
        "pcl_emit_fastq": (i32, [vp, i32, i32, i32, C.POINTER(EmitResult)]),
        "pcl_zstd_frame": (i32, [vp, vp, u64, vp, u64, C.POINTER(u64)]),
        "pcl_gzsrc_open": (i32, [C.c_char_p, i32, C.POINTER(vp)]),
        "pcl_gzsrc_read": (C.c_int64, [vp, vp, u64]),
        "pcl_gzsrc_is_bgzf": (i32, [vp]),
        "pcl_gzsrc_error": (C.c_char_p, [vp]),
        "pcl_gzsrc_close": (None, [vp]),
        "pcl_counts_freeze": (i32, [vp, i32]),
        "pcl_chunk_bytes": (u64, [vp, u64]),
        "pcl_mem_info": (i32, [vp, C.POINTER(u64), C.POINTER(u64)]),
        "pcl_verify_enable": (i32, [vp, i32]),
        "pcl_chunk_set_uniform_len": (i32, [vp, i32, C.c_uint16]),
        "pcl_chunk_uniform_results_async": (i32, [vp, i32, vp]),
        "pcl_group_build": (i32, [vp, vp, vp, C.POINTER(vp)]),
        "pcl_group_build_multi": (i32, [vp, C.POINTER(vp), i32, vp, C.POINTER(vp)]),
        "pcl_group_info_get": (i32, [vp, C.POINTER(GroupInfo)]),
        "pcl_group_fetch": (i32, [vp, vp, vp, vp, vp]),
        "pcl_group_key_decode": (i32, [u64, C.c_char_p]),
        "pcl_group_destroy": (i32, [vp]),
        "pcl_fastq_open": (i32, [C.POINTER(C.c_char_p), i32, C.POINTER(vp)]),
        "pcl_fastq_close": (None, [vp]),
        "pcl_fastq_error": (C.c_char_p, [vp]),
        "pcl_fastq_read": (C.c_int64, [vp, u64, u32, u32, u32, vp, vp, vp, vp, vp, u64, vp, C.POINTER(u64)]),
        "pcl_fastq_text_full": (C.c_int, [vp]),
        "pcl_bgzf_walk": (i32, [vp, u64, u64, C.POINTER(BgzfMember), u32, C.POINTER(u32), C.POINTER(u64), C.POINTER(u64)]),
        "pcl_inflater_create": (i32, [vp, C.POINTER(vp)]),
        "pcl_inflater_destroy": (None, [vp]),
        "pcl_inflater_error": (C.c_char_p, [vp]),
        "pcl_inflater_buffer": (i32, [vp, i32, u64, C.POINTER(vp)]),
        "pcl_inflater_run": (i32, [vp, vp, u64, C.POINTER(BgzfMember), u32, i32, u64]),
        "pcl_inflater_move": (i32, [vp, i32, u64, i32, u64, u64]),
        "pcl_inflater_fetch": (i32, [vp, i32, u64, u64, vp]),
        "pcl_inflater_poke": (i32, [vp, i32, u64, vp, u64]),
        "pcl_inflater_stats": (i32, [vp, C.POINTER(f64), C.POINTER(u64), C.POINTER(u64)]),
        "pcl_text_scan_device": (i32, [vp, i32, vp, u64]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    if L.pcl_abi_version() != 1:
        raise ImportError("libprecellar_b200.so ABI version mismatch")
    _lib = L
    return L


def check(ctx, rc: int) -> int:
    if rc < 0:
        msg = lib().pcl_last_error(ctx)
        raise PclError(rc, msg.decode(errors="replace") if msg else "")
    return rc
