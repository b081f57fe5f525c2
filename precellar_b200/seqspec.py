"""seqspec library-structure model: the host-side mirror of the reference's ``seqspec`` crate
for the barcode hot path (Assay, Region, Read, region-table derivation, whitelists).

Mirrors (file:line relative to the precellar repository):
  * ``Assay``            seqspec/src/lib.rs:36-549 and its PyO3 wrapper python/src/pyseqspec.rs:47-396
  * ``Region``/``LibSpec``  seqspec/src/region.rs:18-261, RegionType :276-425, SequenceType :427-444, Onlist :446-469
  * ``Read``             seqspec/src/read.rs:56-213
  * ``SegmentInfo.truncate_max``  seqspec/src/read/segment.rs:144-164

This module is configuration-time host logic (runs once per assay); the per-read work happens in the
CUDA library behind ``include/precellar_b200.h``.
"""
from __future__ import annotations

import copy
import glob as _glob
import gzip
import hashlib
import io
import logging
import os
from dataclasses import dataclass, field
from typing import Dict, Iterator, List, Optional, Tuple

import yaml

log = logging.getLogger("precellar_b200")

MODALITIES = ("dna", "rna", "tag", "protein", "atac", "crispr")       # lib.rs:551-560

# region.rs:378-413 -- canonical names and accepted aliases (case-insensitive)
_REGION_TYPES = {
    "barcode": "barcode", "cdna": "cdna", "custom_primer": "custom_primer", "customprime": "custom_primer",
    "fastq": "fastq", "fastq_link": "fastq_link", "fastqlink": "fastq_link", "gdna": "gdna", "hic": "hic",
    "illumina_p5": "illumina_p5", "illuminap5": "illumina_p5", "illumina_p7": "illumina_p7",
    "illuminap7": "illumina_p7", "index5": "index5", "index7": "index7", "linker": "linker", "me1": "me1",
    "me2": "me2", "methyl": "methyl", "named": "named", "nextera_read1": "nextera_read1",
    "nexteraread1": "nextera_read1", "nextera_read2": "nextera_read2", "nexteraread2": "nextera_read2",
    "poly_a": "poly_a", "polya": "poly_a", "poly_g": "poly_g", "polyg": "poly_g", "poly_t": "poly_t",
    "polyt": "poly_t", "poly_c": "poly_c", "polyc": "poly_c", "s5": "s5", "s7": "s7",
    "truseq_read1": "truseq_read1", "truseqread1": "truseq_read1", "truseq_read2": "truseq_read2",
    "truseqread2": "truseq_read2", "umi": "umi",
}
_SEQUENCING_PRIMERS = {"custom_primer", "truseq_read1", "truseq_read2", "nextera_read1", "nextera_read2",
                       "illumina_p5", "illumina_p7"}                     # region.rs:365-376
_SEQUENCE_TYPES = ("fixed", "random", "onlist", "joined")               # region.rs:427-434


def parse_modality(s: str) -> str:
    """Modality::from_str (lib.rs:591-605): case-insensitive."""
    m = str(s).lower()
    if m not in MODALITIES:
        raise ValueError(f"Invalid modality: {s}")
    return m


def parse_region_type(s: str) -> str:
    """RegionType::from_str_case_insensitive (region.rs:378-413); modalities are region types too."""
    k = str(s).lower()
    if k in _REGION_TYPES:
        return _REGION_TYPES[k]
    if k in MODALITIES:
        return k
    raise ValueError(f"Invalid region type: {s}")


def _as_str(v) -> str:
    # serde reads md5: null | 0 | hex-string into a String in the templates we ship with; be lenient.
    return "" if v is None else str(v)


# ---------------------------------------------------------------------------------------------
# compressed-file helpers (seqspec/src/utils.rs:44-56,110-125)
# ---------------------------------------------------------------------------------------------
def open_file(path: str) -> io.BufferedIOBase:
    """open_file: gzip detected by magic (multi-member ok), zstd by the ``.zst`` extension."""
    with open(path, "rb") as fh:
        magic = fh.read(2)
    if magic == b"\x1f\x8b":
        return gzip.open(path, "rb")
    if path.endswith(".zst"):
        from . import _zstd
        return io.BytesIO(_zstd.decompress_file(path))
    return open(path, "rb")


def read_lines(path: str) -> List[bytes]:
    """``BufRead::lines()`` semantics (region.rs:467-468): split on \\n, strip one trailing \\r."""
    with open_file(path) as fh:
        data = fh.read()
    lines = data.split(b"\n")
    if lines and lines[-1] == b"":
        lines.pop()
    return [ln[:-1] if ln.endswith(b"\r") else ln for ln in lines]


# ---------------------------------------------------------------------------------------------
# Onlist / File / Region / Read
# ---------------------------------------------------------------------------------------------
_ONLIST_CACHE: dict = {}


class OnlistItems(list):
    """The entries of an onlist file (duplicates dropped, file order) with a cached packed form - one byte blob and the
    offsets of the entries - which is what ``pcl_whitelist_load`` takes: a 737 K-entry list is joined once per process,
    not once per pipeline."""
    unique = True

    def packed(self):
        got = getattr(self, "_packed", None)
        if got is None:
            import numpy as np
            offs = np.zeros(len(self) + 1, dtype=np.uint64)
            if self:
                np.cumsum(np.fromiter(map(len, self), dtype=np.uint64, count=len(self)), out=offs[1:])
            blob = np.frombuffer(b"".join(self), dtype=np.uint8) if self else np.zeros(1, np.uint8)
            got = self._packed = (blob, offs)
        return got


@dataclass
class Onlist:                                                            # region.rs:446-456
    file_id: str = ""
    filename: str = ""
    filetype: str = ""
    filesize: int = 0
    url: str = ""
    urltype: str = "local"
    location: Optional[str] = None
    md5: str = ""

    @classmethod
    def from_dict(cls, d: dict) -> "Onlist":
        return cls(_as_str(d.get("file_id")), _as_str(d.get("filename")), _as_str(d.get("filetype")),
                   int(d.get("filesize") or 0), _as_str(d.get("url")), _as_str(d.get("urltype") or "local").lower(),
                   (None if d.get("location") is None else str(d["location"]).lower()), _as_str(d.get("md5")))

    def to_dict(self) -> dict:
        return dict(file_id=self.file_id, filename=self.filename, filetype=self.filetype, filesize=self.filesize,
                    url=self.url, urltype=self.urltype, location=self.location, md5=self.md5)

    def path(self) -> str:
        """Onlist::read's path resolution (region.rs:459-466).  Remote lists are looked up in the
        reference's cache directory (~/.cache/seqspec/<filename>); there is no downloader here."""
        if self.urltype == "local":
            return self.url
        cached = os.path.join(os.path.expanduser("~"), ".cache", "seqspec", self.filename)
        if os.path.exists(cached):
            return cached
        raise FileNotFoundError(
            f"onlist '{self.filename}' is remote ({self.url}) and not in ~/.cache/seqspec; "
            "point the region's onlist at a local file (urltype: local)")

    def read(self) -> List[bytes]:
        """Lines in file order with duplicates dropped (IndexSet, region.rs:459-469).  The parsed list of a file is kept
        per (path, size, mtime): update_read's verification and the run itself would otherwise parse a multi-million-line
        whitelist several times."""
        path = self.path()
        try:
            st = os.stat(path)
            key = (path, st.st_size, st.st_mtime_ns)
        except OSError:
            key = None
        if key is not None and key in _ONLIST_CACHE:
            return _ONLIST_CACHE[key]
        items = OnlistItems(dict.fromkeys(read_lines(path)))
        if key is not None:
            if len(_ONLIST_CACHE) >= 8:
                _ONLIST_CACHE.pop(next(iter(_ONLIST_CACHE)))
            _ONLIST_CACHE[key] = items
        return items


@dataclass
class File:                                                              # read.rs:222-231
    file_id: str = ""
    filename: str = ""
    filetype: str = "fastq"
    filesize: int = 0
    url: str = ""
    urltype: str = "local"
    md5: str = "0"

    @classmethod
    def from_dict(cls, d: dict) -> "File":
        return cls(_as_str(d.get("file_id")), _as_str(d.get("filename")), _as_str(d.get("filetype")),
                   int(d.get("filesize") or 0), _as_str(d.get("url")), _as_str(d.get("urltype") or "local").lower(),
                   _as_str(d.get("md5")))

    @classmethod
    def from_fastq(cls, path: str, compute_md5: bool = False) -> "File":  # read.rs:234-251
        name = os.path.basename(path)
        md5 = "0"
        if compute_md5:
            h = hashlib.md5()
            with open(path, "rb") as fh:
                for blk in iter(lambda: fh.read(1 << 20), b""):
                    h.update(blk)
            md5 = h.hexdigest()
        return cls(name, name, "fastq", os.path.getsize(path), str(path), "local", md5)

    def to_dict(self) -> dict:
        return dict(file_id=self.file_id, filename=self.filename, filetype=self.filetype, filesize=self.filesize,
                    url=self.url, urltype=self.urltype, md5=self.md5)


@dataclass
class Region:                                                            # region.rs:167-179
    region_id: str
    region_type: str
    name: str = ""
    sequence_type: str = "random"
    sequence: str = ""
    min_len: int = 0
    max_len: int = 0
    onlist: Optional[Onlist] = None
    subregions: List["Region"] = field(default_factory=list)

    @classmethod
    def from_dict(cls, d: dict) -> "Region":
        st = str(d.get("sequence_type", "")).lower()
        if st not in _SEQUENCE_TYPES:
            raise ValueError(f"Invalid sequence type: {d.get('sequence_type')}")
        subs = d.get("regions") or []                                    # `regions: null` == empty (region.rs:263-275)
        onl = d.get("onlist")
        return cls(region_id=str(d["region_id"]), region_type=parse_region_type(d["region_type"]),
                   name=_as_str(d.get("name")), sequence_type=st, sequence=_as_str(d.get("sequence")),
                   min_len=int(d["min_len"]), max_len=int(d["max_len"]),
                   onlist=Onlist.from_dict(onl) if onl else None,
                   subregions=[Region.from_dict(x) for x in subs])

    def to_dict(self) -> dict:
        return dict(region_id=self.region_id, region_type=self.region_type, name=self.name,
                    sequence_type=self.sequence_type, sequence=self.sequence, min_len=self.min_len,
                    max_len=self.max_len, onlist=self.onlist.to_dict() if self.onlist else None,
                    regions=[r.to_dict() for r in self.subregions] or None)

    # region.rs:337-376
    def is_barcode(self) -> bool: return self.region_type == "barcode"
    def is_umi(self) -> bool: return self.region_type == "umi"
    def is_target(self) -> bool: return self.region_type in ("gdna", "cdna")
    def is_sequencing_primer(self) -> bool: return self.region_type in _SEQUENCING_PRIMERS
    def is_modality(self) -> bool: return self.region_type in MODALITIES
    def is_fixed_length(self) -> bool: return self.min_len == self.max_len

    def len(self) -> Optional[int]:                                      # region.rs:204-210
        return self.min_len if self.min_len == self.max_len else None

    def get_nesting_depth(self) -> int:                                  # region.rs:219-231
        if not self.subregions:
            return 0
        return 1 + max(r.get_nesting_depth() for r in self.subregions)


@dataclass
class Read:                                                              # read.rs:56-66
    read_id: str = ""
    name: Optional[str] = None
    modality: str = "dna"
    primer_id: str = ""
    min_len: int = 0
    max_len: int = 0
    strand: str = "pos"
    files: Optional[List[File]] = None

    @classmethod
    def from_dict(cls, d: dict) -> "Read":
        strand = str(d.get("strand", "pos")).lower()
        if strand not in ("pos", "neg"):
            raise ValueError(f"Invalid strand: {d.get('strand')}")
        files = d.get("files")
        return cls(read_id=str(d["read_id"]), name=d.get("name"), modality=parse_modality(d["modality"]),
                   primer_id=str(d["primer_id"]), min_len=int(d["min_len"]), max_len=int(d["max_len"]),
                   strand=strand, files=None if files is None else [File.from_dict(f) for f in files])

    def to_dict(self) -> dict:
        return dict(read_id=self.read_id, name=self.name, modality=self.modality, primer_id=self.primer_id,
                    min_len=self.min_len, max_len=self.max_len, strand=self.strand,
                    files=None if self.files is None else [f.to_dict() for f in self.files])

    def is_reverse(self) -> bool:                                        # read.rs:175-180
        return self.strand == "neg"

    def fastq_paths(self) -> List[str]:
        """Files that Read::open would concatenate (read.rs:137-155); empty -> open() is None."""
        return [f.url for f in (self.files or []) if f.filetype == "fastq"]


@dataclass
class SegmentInfoElem:                                                   # segment.rs:371-378
    region_id: str
    region_type: str
    sequence_type: str
    sequence: str
    min_len: int
    max_len: int

    def is_fixed_len(self) -> bool: return self.min_len == self.max_len
    def is_barcode(self) -> bool: return self.region_type == "barcode"
    def is_umi(self) -> bool: return self.region_type == "umi"
    def is_target(self) -> bool: return self.region_type in ("gdna", "cdna")


@dataclass
class SegmentInfo:                                                       # segment.rs:93-97
    elems: List[SegmentInfoElem]
    is_reverse: bool

    def truncate_max(self, length: int) -> "SegmentInfo":               # segment.rs:144-164
        total, out = 0, []
        for e in self.elems:
            if total + e.max_len > length:
                if total + e.min_len <= length:
                    out.append(e)
                break
            total += e.max_len
            out.append(e)
        return SegmentInfo(out, self.is_reverse)

    def __len__(self) -> int:
        return len(self.elems)

    def __iter__(self) -> Iterator[SegmentInfoElem]:
        return iter(self.elems)


# ---------------------------------------------------------------------------------------------
# YAML loading: every mapping in the templates carries a !Assay/!Region/!Read/!Onlist/!File tag
# ---------------------------------------------------------------------------------------------
class _Loader(yaml.SafeLoader):
    pass


def _tagged(loader, suffix, node):
    if isinstance(node, yaml.MappingNode):
        return loader.construct_mapping(node, deep=True)
    if isinstance(node, yaml.SequenceNode):
        return loader.construct_sequence(node, deep=True)
    return loader.construct_scalar(node)


_Loader.add_multi_constructor("!", _tagged)


class Assay:
    """A seqspec assay.  Same constructor and methods as ``precellar.Assay`` (python/src/pyseqspec.rs:47-396).

    Parameters
    ----------
    path : str | os.PathLike
        Local path of the seqspec YAML (URLs are not supported: there is no network layer here).
    """

    def __init__(self, path):
        path = os.fspath(path)
        if "://" in path:
            raise ValueError("Assay.from_url is not available in this build (no network layer); pass a local path")
        with open(path, "r") as fh:
            doc = yaml.load(fh, Loader=_Loader)
        if not isinstance(doc, dict):
            raise ValueError("Failed to parse YAML")
        self.file: Optional[str] = path
        self._load(doc)
        self.normalize_all_paths()
        self.validate_structure()

    @classmethod
    def from_path(cls, path) -> "Assay":                                # lib.rs:57-65
        return cls(path)

    def _load(self, doc: dict) -> None:
        self.seqspec_version = _as_str(doc.get("seqspec_version"))
        self.assay_id = _as_str(doc.get("assay_id"))
        self.name = _as_str(doc.get("name"))
        self.doi = _as_str(doc.get("doi"))
        self.date = _as_str(doc.get("date"))
        self.description = _as_str(doc.get("description"))
        self._modalities = [parse_modality(m) for m in (doc.get("modalities") or [])]
        self.lib_struct = _as_str(doc.get("lib_struct"))
        self.library_protocol = doc.get("library_protocol")
        self.library_kit = doc.get("library_kit")
        self._sequence_protocol = doc.get("sequence_protocol")
        self._sequence_kit = doc.get("sequence_kit")
        cs = doc.get("chemistry_strandedness")
        self._chemistry_strandedness = None if cs is None else str(cs).lower()
        self.sequence_spec: Dict[str, Read] = {}
        for r in doc.get("sequence_spec") or []:
            read = Read.from_dict(r)
            if read.read_id in self.sequence_spec:
                raise ValueError(f"Duplicate read id: {read.read_id}")   # read.rs:45-49
            self.sequence_spec[read.read_id] = read
        self._set_library([Region.from_dict(r) for r in (doc.get("library_spec") or [])])

    def _set_library(self, regions: List[Region]) -> None:
        """LibSpec::new (region.rs:52-85)."""
        self.library_spec: Dict[str, Region] = {}      # modality -> region
        self._region_map: Dict[str, Region] = {}
        self._parent_map: Dict[str, Region] = {}
        for region in regions:
            if not region.is_modality():
                raise ValueError("Top-level regions must be modalities")
            if region.region_type in self.library_spec:
                raise ValueError(f"Duplicate modality: {region.region_type}")
            self.library_spec[region.region_type] = region
            if region.region_id in self._region_map:
                raise ValueError(f"Duplicate region id: {region.region_id}")
            self._region_map[region.region_id] = region
            for sub in region.subregions:
                if sub.region_id in self._region_map:
                    raise ValueError(f"Duplicate region id: {sub.region_id}")
                self._region_map[sub.region_id] = sub
                self._parent_map[sub.region_id] = region

    # -- structure ------------------------------------------------------------------------------
    def validate_structure(self) -> None:                               # lib.rs:75-90
        for m in self._modalities:
            region = self.library_spec.get(m)
            if region is not None:
                depth = region.get_nesting_depth()
                if depth != 1:
                    raise ValueError(f"Invalid assay structure for modality {m}: nesting depth must be exactly "
                                     f"2 levels (found {depth} levels)")

    def normalize_all_paths(self) -> None:                              # lib.rs:110-130
        if not self.file:
            return
        base = os.path.dirname(os.path.abspath(self.file))

        def norm(url: str) -> str:
            p = url if os.path.isabs(url) else os.path.join(base, url)
            rp = os.path.realpath(p)
            if not os.path.exists(rp):
                log.warning("Failed to normalize path: %r", p)
                return url
            return rp

        for read in self.sequence_spec.values():
            for f in read.files or []:
                if f.urltype == "local":
                    f.url = norm(f.url)
        for region in self._region_map.values():
            if region.onlist is not None and region.onlist.urltype == "local":
                region.onlist.url = norm(region.onlist.url)

    # -- simple properties (pyseqspec.rs:69-178) -----------------------------------------------------
    @property
    def id(self) -> str: return self.assay_id
    @id.setter
    def id(self, v: str) -> None: self.assay_id = str(v)

    @property
    def chemistry_strandedness(self) -> Optional[str]: return self._chemistry_strandedness
    @chemistry_strandedness.setter
    def chemistry_strandedness(self, v: str) -> None:
        if v not in ("forward", "reverse", "unstranded"):
            raise ValueError(f"Invalid chemistry_strandedness: {v}. Must be one of 'forward', 'reverse', or 'unstranded'.")
        self._chemistry_strandedness = v

    @property
    def sequence_protocol(self) -> str:
        p = self._sequence_protocol
        return p if isinstance(p, str) else f"custom: {p!r}"
    @sequence_protocol.setter
    def sequence_protocol(self, v: str) -> None: self._sequence_protocol = str(v)

    @property
    def sequence_kit(self) -> str:
        p = self._sequence_kit
        return p if isinstance(p, str) else f"custom: {p!r}"
    @sequence_kit.setter
    def sequence_kit(self, v: str) -> None: self._sequence_kit = str(v)

    @property
    def filename(self) -> Optional[str]: return self.file

    def modalities(self) -> List[str]:
        return list(self._modalities)

    def modality(self) -> str:                                          # lib.rs:93-106
        if len(self._modalities) == 1:
            return self._modalities[0]
        raise ValueError("Multiple modalities found: " + ", ".join(self._modalities))

    # -- reads -----------------------------------------------------------------------------------------
    def iter_reads(self, modality: str) -> Iterator[Read]:              # lib.rs:474-478
        m = parse_modality(modality)
        return (r for r in self.sequence_spec.values() if r.modality == m)

    def delete_read(self, read_id: str) -> Optional[Read]:              # lib.rs:396-398
        return self.sequence_spec.pop(read_id, None)

    def delete_all_reads(self, modality: str) -> None:                  # lib.rs:400-408
        for rid in [r.read_id for r in self.iter_reads(modality)]:
            self.delete_read(rid)

    def delete_region(self, region_id: str) -> None:                    # lib.rs:410-413, region.rs:152-162
        tops = []
        for region in self.library_spec.values():
            if region.region_id == region_id:
                continue
            r = copy.copy(region)
            r.subregions = [s for s in region.subregions if s.region_id != region_id]
            tops.append(r)
        self._set_library(tops)

    def add_illumina_reads(self, modality: str, *, length: int = 150, forward_strand_workflow: bool = False) -> None:
        """Derive R1/R2/I1/I2 from the P5/P7/read-primer layout (lib.rs:154-307)."""
        modality = parse_modality(modality)

        def get_length(regions: List[Region], reverse: bool) -> int:    # lib.rs:176-195
            seq = regions if reverse else list(reversed(regions))
            i = 0
            while i < len(seq) and seq[i].sequence_type == "fixed":
                i += 1
            total = 0
            for r in seq[i:]:
                if r.len() is None:
                    raise ValueError(f"region {r.region_id} has no fixed length")   # .len().unwrap()
                total += r.len()
            return total

        self.delete_all_reads(modality)
        top = self.library_spec.get(modality)
        if top is None:
            raise ValueError(f"Modality not found: {modality}")
        it = iter(top.subregions)

        def advance_until(pred) -> Optional[Tuple[Region, List[Region]]]:   # lib.rs:160-174
            acc = []
            for nxt in it:
                if pred(nxt):
                    return nxt, acc
                acc.append(nxt)
            return None

        is_read1 = lambda r: r.region_type in ("nextera_read1", "truseq_read1")
        is_read2 = lambda r: r.region_type in ("nextera_read2", "truseq_read2")
        for cur in it:
            if cur.region_type == "illumina_p5":
                got = advance_until(is_read1)
                if got is not None:
                    nxt, acc = got
                    self.update_read(f"{modality}-R1", modality=modality, primer_id=nxt.region_id, is_reverse=False,
                                     min_len=length, max_len=length)
                    if forward_strand_workflow:
                        acc_len = get_length(acc, False)
                        if acc_len > 0:
                            self.update_read(f"{modality}-I2", modality=modality, primer_id=cur.region_id,
                                             is_reverse=False, min_len=acc_len, max_len=acc_len)
                    else:
                        acc_len = get_length(acc, True)
                        if acc_len > 0:
                            self.update_read(f"{modality}-I2", modality=modality, primer_id=nxt.region_id,
                                             is_reverse=True, min_len=acc_len, max_len=acc_len)
            elif is_read2(cur):
                got = advance_until(lambda r: r.region_type == "illumina_p7")
                if got is not None:
                    _, acc = got
                    acc_len = get_length(acc, False)
                    self.update_read(f"{modality}-R2", modality=modality, primer_id=cur.region_id, is_reverse=True,
                                     min_len=length, max_len=length)
                    if acc_len > 0:
                        self.update_read(f"{modality}-I1", modality=modality, primer_id=cur.region_id,
                                         is_reverse=False, min_len=acc_len, max_len=acc_len)

    def update_read(self, read_id: str, *, modality: Optional[str] = None, primer_id: Optional[str] = None,
                    is_reverse: Optional[bool] = None, fastq=None, min_len: Optional[int] = None,
                    max_len: Optional[int] = None, compute_md5: bool = False, infer_read_length: bool = True,
                    infer_read_length_sample: int = 10000, verify: bool = True) -> None:
        """Create or update a read (lib.rs:310-394; argument handling of pyseqspec.rs:246-291)."""
        if read_id in self.sequence_spec:
            read = copy.deepcopy(self.sequence_spec[read_id])
        else:
            assert modality is not None, "modality must be provided for a new read"
            assert primer_id is not None, "primer_id must be provided for a new read"
            assert is_reverse is not None, "is_reverse must be provided for a new read"
            read = Read(read_id=read_id)
        if is_reverse is not None:
            read.strand = "neg" if is_reverse else "pos"
        if modality is not None:
            read.modality = parse_modality(modality)
        if primer_id is not None:
            read.primer_id = primer_id
        if fastq is not None:
            if isinstance(fastq, (list, tuple)):
                paths = [os.fspath(p) for p in fastq]
            elif isinstance(fastq, str):
                paths = sorted(_glob.glob(fastq))                        # a str is a glob pattern (pyseqspec.rs:265-270)
            else:
                paths = [os.fspath(fastq)]
            read.files = [File.from_fastq(p, compute_md5) for p in paths]

        if (min_len is None or max_len is None) and infer_read_length:  # lib.rs:363-379
            from .fastq import infer_length
            lo, hi = infer_length(read.fastq_paths(), infer_read_length_sample)
            log.info("The read length of %s was inferred to be %s.", read_id, lo if lo == hi else f"{lo}-{hi}")
            read.min_len = lo if min_len is None else min_len
            read.max_len = hi if max_len is None else max_len
        else:
            read.min_len = min_len or 0
            read.max_len = max_len or 0

        if read.primer_id not in self._parent_map:                      # verify (lib.rs:482-485)
            raise ValueError(f"Primer not found: {read.primer_id}")
        if verify:
            self._verify(read, infer_read_length_sample)
        region = self._region_map.get(read.primer_id)
        if region is not None and not region.is_sequencing_primer():
            log.warning("primer_id '%s' is not a sequencing primer (type: %s). This will cause errors. If you are "
                        "sure this is what you want, please change region_type to custom_primer.",
                        read.primer_id, region.region_type)
        self.sequence_spec[read_id] = read

    # -- verification of a read against a sample of its records -----------------------------------------------------------
    def _verify(self, read: "Read", sample_size: int) -> Optional[dict]:
        """Assay::verify (lib.rs:481-548): split the first ``sample_size`` records with tolerances (0.2, 0.2) and warn when
        fewer than half are valid or fewer than half of a barcode region's segments are on its onlist.  The sample goes
        through the CUDA path (pcl_verify_enable: the same walk, fixed sequences compared too); there is no CPU route, so
        without a usable device the check is skipped with a log line.  Returns the percentages (for tests)."""
        info = self._segments_of(read)
        paths = read.fastq_paths()
        if info is None or not paths:
            return None
        try:
            import numpy as np
            from . import _ffi
            from .ingest import FastqReader
            from .layout import layout_from_elems
            from .pipeline import BarcodePipeline
            elems, wl = [], {}
            for e in info:
                if e.is_barcode():
                    region = self._region_map.get(e.region_id)
                    onlist = region.onlist.read() if (region is not None and region.onlist is not None) else None
                    if onlist:
                        wl[e.region_id] = onlist
                        elems.append(e)
                    else:                                               # no onlist: only the split matters (lib.rs:498-506)
                        elems.append(SegmentInfoElem(e.region_id, "linker", e.sequence_type, e.sequence, e.min_len, e.max_len))
                else:
                    elems.append(e)
            layout = layout_from_elems([dict(read_id=read.read_id, elems=elems, is_reverse=read.is_reverse(), max_len=int(read.max_len))],
                                       wl, read.modality)
            pipe = BarcodePipeline(layout, verify=True)
        except (ImportError, OSError) as exc:
            log.info("verification of read '%s' skipped: %s", read.read_id, exc)
            return None
        except Exception as exc:                                        # no device, or a layout the kernels do not take
            code = getattr(exc, "code", None)
            if code in (-2, -4):                                        # PCL_ERR_CUDA, PCL_ERR_UNSUPPORTED
                log.info("verification of read '%s' skipped: %s", read.read_id, exc)
                return None
            raise
        try:
            rdr = FastqReader(paths)
            try:
                b = rdr.read(sample_size, pipe.stride(0), qual_only=bool(pipe.infos[0].qual_only), qual_window=pipe.qual_window(0))
            finally:
                rdr.close()
            if b is None:
                return None
            n = len(b.batch.length)
            ch = pipe.new_chunk(n)
            ch.reset(n)
            ch.upload(0, b.batch)
            pipe.count(ch)
            res = ch.fetch(0)
            fs = res.fstatus.astype(np.uint32)
            valid = (fs & 3) == 0
            out = {"total": n, "percent_valid": float(valid.sum()) / n * 100.0, "percent_matched": {}}
            if out["percent_valid"] < 50.0:
                log.warning("Read '%s' has low percentage of valid records. Percentage of valid records: %.2f%%",
                            read.read_id, out["percent_valid"])
            j = 0
            per_region: Dict[str, int] = {}
            for e in elems:
                if e.is_barcode():
                    code = (fs >> (4 + 3 * j)) & 7
                    per_region[e.region_id] = per_region.get(e.region_id, 0) + int((valid & (code == _ffi.BC_EXACT)).sum())
                    j += 1
            for rid, n_matched in per_region.items():
                pm = n_matched / n * 100.0
                out["percent_matched"][rid] = pm
                if pm < 50.0:
                    log.warning("Read '%s' has low percentage of matched records for region '%s'. Percentage of matched records: %.2f%%",
                                read.read_id, rid, pm)
            return out
        finally:
            pipe.close()

    # -- region table -----------------------------------------------------------------------------------
    def get_segments(self, read_id: str) -> Optional[SegmentInfo]:
        """Atomic regions covered by a read (lib.rs:433-440 -> read.rs:182-212 -> segment.rs:144-164)."""
        read = self.sequence_spec.get(read_id)
        if read is None:
            return None
        return self._segments_of(read)

    def _segments_of(self, read: "Read") -> Optional[SegmentInfo]:
        parent = self._parent_map.get(read.primer_id)
        if parent is None or parent.sequence_type != "joined":
            return None
        subs = list(reversed(parent.subregions)) if read.is_reverse() else list(parent.subregions)
        i = 0
        while i < len(subs) and not (subs[i].is_sequencing_primer() and subs[i].region_id == read.primer_id):
            i += 1                                                       # skip_while(!found)
        elems = [SegmentInfoElem(r.region_id, r.region_type, r.sequence_type, r.sequence, r.min_len, r.max_len)
                 for r in subs[i + 1:]]                                  # .skip(1)
        info = SegmentInfo(elems, read.is_reverse()).truncate_max(read.max_len)
        return info if len(info) > 0 else None

    def get_segments_by_modality(self, modality: str) -> List[Tuple[Read, SegmentInfo]]:   # lib.rs:416-430
        m = parse_modality(modality)
        out = []
        for read in self.sequence_spec.values():
            if read.modality == m:
                info = self.get_segments(read.read_id)
                if info is None:
                    raise RuntimeError(f"Cannot find index for Read: {read.read_id}")
                out.append((read, info))
        return out

    def get_whitelists(self, modality: str) -> Dict[str, List[bytes]]:  # lib.rs:443-472
        m = parse_modality(modality)
        top = self.library_spec.get(m)
        if top is None:
            raise KeyError(f"modality not found: {m}")
        out: Dict[str, List[bytes]] = {}
        for r in top.subregions:
            if r.is_barcode():
                if r.sequence_type == "onlist" and r.onlist is not None:
                    out[r.region_id] = r.onlist.read()
                else:
                    if r.sequence_type == "onlist":
                        log.info("Barcode region '%s' does not have a whitelist", r.region_id)
                    out[r.region_id] = []
        return out

    def whitelist(self, modality: str) -> Optional[List[str]]:
        """Cartesian product of all barcode onlists (region.rs:109-141, pyseqspec.rs:307-317)."""
        m = parse_modality(modality)
        top = self.library_spec.get(m)
        if top is None:
            return None
        lists = [r.onlist.read() for r in top.subregions if r.is_barcode() and r.onlist is not None]
        if not lists:
            return None
        acc = lists[0]
        for nxt in lists[1:]:
            acc = [a + b for a in acc for b in nxt]
        return [x.decode() for x in acc]

    # -- serialisation ------------------------------------------------------------------------------------
    def _to_doc(self) -> dict:
        return dict(seqspec_version=self.seqspec_version, assay_id=self.assay_id, name=self.name, doi=self.doi,
                    date=self.date, description=self.description, modalities=list(self._modalities),
                    lib_struct=self.lib_struct, library_protocol=self.library_protocol,
                    library_kit=self.library_kit, sequence_protocol=self._sequence_protocol,
                    sequence_kit=self._sequence_kit, chemistry_strandedness=self._chemistry_strandedness,
                    sequence_spec=[r.to_dict() for r in self.sequence_spec.values()],
                    library_spec=[r.to_dict() for r in self.library_spec.values()])

    def to_yaml(self) -> str:
        return yaml.safe_dump(self._to_doc(), sort_keys=False)

    def save(self, out) -> None:
        """Write the YAML with paths made relative to the output location (pyseqspec.rs:358-368)."""
        out = os.fspath(out)
        base = os.path.dirname(os.path.abspath(out))
        doc = copy.deepcopy(self)

        def rel(url: str) -> str:
            try:
                return os.path.relpath(url, base)
            except ValueError:
                return url

        for read in doc.sequence_spec.values():
            for f in read.files or []:
                if f.urltype == "local":
                    f.url = rel(f.url)
        for region in doc._region_map.values():
            if region.onlist is not None and region.onlist.urltype == "local":
                region.onlist.url = rel(region.onlist.url)
        with open(out, "w") as fh:
            fh.write(doc.to_yaml())

    def __repr__(self) -> str:                                          # pyseqspec.rs:370-396 (tree rendering)
        by_primer: Dict[str, List[Read]] = {}
        for r in self.sequence_spec.values():
            by_primer.setdefault(r.primer_id, []).append(r)
        lines = [self.assay_id]
        tops = list(self.library_spec.values())
        for ti, top in enumerate(tops):
            last_top = ti == len(tops) - 1
            lines.append(("└── " if last_top else "├── ") + f"{top.region_id}({top.min_len}-{top.max_len})")
            pad = "    " if last_top else "│   "
            for si, sub in enumerate(top.subregions):
                ln = f"{sub.min_len}" if sub.min_len == sub.max_len else f"{sub.min_len}-{sub.max_len}"
                label = f"{sub.region_id}({ln})"
                reads = by_primer.get(sub.region_id, [])
                if reads:
                    label += " [" + ", ".join(("↑" if r.is_reverse() else "↓") + r.read_id +
                                              f"({r.min_len}-{r.max_len})" for r in reads) + "]"
                lines.append(pad + ("└── " if si == len(top.subregions) - 1 else "├── ") + label)
        return "\n".join(lines)

    __str__ = __repr__
