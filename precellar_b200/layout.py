"""Region-table compiler: Assay + modality -> the flat per-read element tables the kernels consume.

What the reference derives on the fly for every run (file:line relative to the precellar repo):
  * the reads that take part                 precellar/src/align/fastq.rs:139-148, 460-471
  * their SegmentInfo after truncate_max      seqspec/src/lib.rs:416-440, seqspec/src/read.rs:182-212
  * the whitelist of every barcode region     seqspec/src/lib.rs:443-472, precellar/src/barcode.rs:66-78
is computed here once and handed to ``pcl_layout_set`` / ``pcl_whitelist_load``.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional

from . import _ffi
from .seqspec import Assay, SegmentInfoElem, parse_modality

_KIND = {"barcode": _ffi.KIND_BARCODE, "umi": _ffi.KIND_UMI, "gdna": _ffi.KIND_TARGET, "cdna": _ffi.KIND_TARGET}


@dataclass
class ReadPlan:
    read_id: str
    is_reverse: bool
    max_len: int                      # Read.max_len (FastqReader truncation, 0 = none)
    elems: List[SegmentInfoElem]
    region_slots: List[int]           # per element: whitelist slot for barcodes, else -1
    files: List[str] = field(default_factory=list)

    @property
    def has_barcode(self) -> bool:
        return any(e.is_barcode() for e in self.elems)


@dataclass
class CompiledLayout:
    modality: str
    reads: List[ReadPlan]             # sequence_spec order == join order (fastq.rs:362-379)
    region_ids: List[str]             # whitelist slot -> region id (get_whitelists order)
    whitelists: Dict[str, List[bytes]]
    onlist_type: Dict[str, bool]      # region id -> sequence_type == onlist (barcode.rs:71-76)
    contains_umi: bool

    def descs(self):
        """ctypes array of pcl_read_desc for pcl_layout_set."""
        arr = (_ffi.ReadDesc * len(self.reads))()
        for r, plan in enumerate(self.reads):
            if len(plan.elems) > _ffi.PCL_MAX_ELEMS:
                raise _ffi.PclError(_ffi.PCL_ERR_UNSUPPORTED, f"read {plan.read_id}: more than {_ffi.PCL_MAX_ELEMS} elements")
            d = arr[r]
            d.n_elems = len(plan.elems)
            d.is_reverse = 1 if plan.is_reverse else 0
            d.max_len = plan.max_len
            for e, el in enumerate(plan.elems):
                de = d.elems[e]
                de.kind = _KIND.get(el.region_type, _ffi.KIND_OTHER)
                if el.max_len > 0xFFFF:
                    raise _ffi.PclError(_ffi.PCL_ERR_UNSUPPORTED, f"element {el.region_id}: max_len {el.max_len} > 65535")
                de.min_len = el.min_len
                de.max_len = el.max_len
                de.region_slot = plan.region_slots[e]
                # A Fixed sequence is only ever compared when it is searched for as an anchor, i.e. when a
                # variable-length element precedes it; split() uses linker_tolerance = 1.0 (segment.rs:169-171,325).
                if el.sequence_type == "fixed" and _can_anchor(plan.elems, e):
                    seq = el.sequence.encode()
                    if len(seq) > _ffi.PCL_MAX_ANCHOR:
                        raise _ffi.PclError(_ffi.PCL_ERR_UNSUPPORTED,
                                            f"element {el.region_id}: anchor longer than {_ffi.PCL_MAX_ANCHOR}")
                    de.is_fixed_seq = 1
                    de.seq_len = len(seq)
                    de.sequence = seq
        return arr


def _can_anchor(elems: List[SegmentInfoElem], e: int) -> bool:
    """A Fixed element is searched for only when a variable-length element precedes it (segment.rs:250-254)."""
    return any(not x.is_fixed_len() for x in elems[:e])


def compile_layout(assay: Assay, modality: str, require_files: bool = False) -> CompiledLayout:
    modality = parse_modality(modality)
    wl = assay.get_whitelists(modality)
    region_ids = list(wl.keys())
    slot_of = {rid: i for i, rid in enumerate(region_ids)}
    if len(region_ids) > _ffi.PCL_MAX_REGIONS:
        raise _ffi.PclError(_ffi.PCL_ERR_UNSUPPORTED, f"more than {_ffi.PCL_MAX_REGIONS} barcode regions")
    onlist_type = {rid: assay._region_map[rid].sequence_type == "onlist" for rid in region_ids}
    reads: List[ReadPlan] = []
    contains_umi = False
    for read, info in assay.get_segments_by_modality(modality):
        contains_umi |= any(e.is_umi() for e in info)                  # barcode.rs:61-63
        # FastqAnnotator::new is None unless the read holds a barcode, UMI or target (fastq.rs:460-471)
        if not any(e.is_barcode() or e.is_umi() or e.is_target() for e in info):
            continue
        files = read.fastq_paths()
        if require_files and not files:                                # read.open()? (fastq.rs:145)
            continue
        slots = []
        for e in info:
            if e.is_barcode():
                if e.region_id not in slot_of:
                    raise KeyError(f"Region '{e.region_id}' not found in whitelist builders")   # barcode.rs:106-109
                slots.append(slot_of[e.region_id])
            else:
                slots.append(-1)
        reads.append(ReadPlan(read.read_id, read.is_reverse(), int(read.max_len), list(info), slots, files))
    if len(reads) > _ffi.PCL_MAX_READS:
        raise _ffi.PclError(_ffi.PCL_ERR_UNSUPPORTED, f"more than {_ffi.PCL_MAX_READS} reads")
    return CompiledLayout(modality, reads, region_ids, wl, onlist_type, contains_umi)


def is_paired_end(layout: CompiledLayout) -> bool:
    """AnnotatedFastqReader::is_paired_end (fastq.rs:314-329), including its naming quirk:
    a target on a reverse read sets `has_read1`, one on a forward read `has_read2`."""
    a = any(e.is_target() for p in layout.reads if p.is_reverse for e in p.elems)
    b = any(e.is_target() for p in layout.reads if not p.is_reverse for e in p.elems)
    return a and b


def layout_from_elems(reads, whitelists, modality: str = "rna") -> CompiledLayout:
    """Build a CompiledLayout from explicit element tables (no YAML): ``reads`` is a list of dicts with
    read_id, elems (SegmentInfoElem list, already truncated), is_reverse, max_len; ``whitelists`` maps
    barcode region ids to byte-string lists (or [n, k] uint8 arrays) in whitelist-slot order."""
    region_ids = list(whitelists.keys())
    slot_of = {r: i for i, r in enumerate(region_ids)}
    plans = []
    for r in reads:
        elems = list(r["elems"])
        slots = [slot_of[e.region_id] if e.is_barcode() else -1 for e in elems]
        plans.append(ReadPlan(r["read_id"], bool(r.get("is_reverse", False)), int(r.get("max_len", 0)), elems, slots,
                              list(r.get("files", []))))
    # arrays and seqspec.OnlistItems (deduplicated, with a cached packed form) pass through; anything else is copied
    wl = {k: (v if (hasattr(v, "shape") or getattr(v, "unique", False)) else list(v)) for k, v in whitelists.items()}
    return CompiledLayout(parse_modality(modality), plans, region_ids, wl, {k: True for k in region_ids},
                          any(e.is_umi() for p in plans for e in p.elems))
