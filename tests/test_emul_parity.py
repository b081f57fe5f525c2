"""CPU tier: the __host__ __device__ logic the kernels wrap (hd_core.h / host_build.h, run on the host by
tests/emul) against the oracle, bit-exact, on randomised layouts and reads."""
import numpy as np
import pytest

from oracle import pyoracle as O
from precellar_b200 import _ffi
from tests import util as U
from tests.util import bc, cdna, gdna, linker, make_layout, poly, umi, var


def _layouts(rng):
    wl16 = U.gen_whitelist(rng, 300, 16)
    wl_hp = U.gen_whitelist(rng, 120, [9, 10])
    wl_rt = U.gen_whitelist(rng, 96, 10)
    wl8a, wl8b, wl8c = (U.gen_whitelist(rng, 48, 8) for _ in range(3))
    wl24 = U.gen_whitelist(rng, 200, 24)
    wl7 = U.gen_whitelist(rng, 64, 7)
    out = {}
    # 10x v3: R1 [B16 U12], R2 [cDNA 1-1000] neg
    out["v3"] = make_layout([
        dict(read_id="R1", elems=[bc("cb", 16), umi(12)], max_len=28),
        dict(read_id="R2", elems=[cdna(1, 1000)], is_reverse=True, max_len=91)], {"cb": wl16})
    # fast-path shapes: 12-base barcode (marker form) + UMI + trailing target in a 32-byte record; 10x v2 (UMI 10)
    wl12 = U.gen_whitelist(rng, 300, 12)
    out["fast12_tail"] = make_layout([
        dict(read_id="R1", elems=[linker("ACG"), bc("cb", 12), umi(9), cdna(1, 500)], max_len=0),
        dict(read_id="R2", elems=[bc("cb", 12), umi(15, "u15")], is_reverse=True, max_len=27)], {"cb": wl12})
    out["v2"] = make_layout([
        dict(read_id="R1", elems=[bc("cb", 16), umi(10)], max_len=26),
        dict(read_id="R2", elems=[gdna(3, 7, "g0"), cdna(1, 1000)], is_reverse=True, max_len=98)], {"cb": wl16})
    # 10x ATAC: I2 [B16] neg, R1/R2 gdna
    out["atac"] = make_layout([
        dict(read_id="R1", elems=[gdna(1, 1000)], max_len=50),
        dict(read_id="I2", elems=[bc("cb", 16)], is_reverse=True, max_len=16),
        dict(read_id="R2", elems=[gdna(1, 1000)], is_reverse=True, max_len=50)], {"cb": wl16}, "atac")
    # multiome ATAC: I2 [O8 fixed, B16] neg
    out["multiome_atac"] = make_layout([
        dict(read_id="R1", elems=[gdna(1, 1000)], max_len=50),
        dict(read_id="I2", elems=[linker("CGCGTCTG"), bc("cb", 16)], is_reverse=True, max_len=24),
        dict(read_id="R2", elems=[gdna(1, 1000)], is_reverse=True, max_len=50)], {"cb": wl16}, "atac")
    # sci-RNA-seq3: R1 [B 9-10, O6 "CAGAGC", U8, B10], R2 [cdna]
    out["sci3"] = make_layout([
        dict(read_id="R1", elems=[bc("hp", 9, 10), linker("CAGAGC"), umi(8), bc("rt", 10)], max_len=34),
        dict(read_id="R2", elems=[cdna(1, 1000)], is_reverse=True, max_len=100)], {"hp": wl_hp, "rt": wl_rt})
    # SHARE-seq-like: three 8-mers separated by fixed linkers, one read; forward
    out["share"] = make_layout([
        dict(read_id="R1", elems=[gdna(1, 1000)], max_len=40),
        dict(read_id="I1", elems=[bc("b1", 8), linker("ACGTACGTACGTACG"), bc("b2", 8), linker("TTGGCCAATTGGCCA"), bc("b3", 8)], max_len=54)],
        {"b1": wl8a, "b2": wl8b, "b3": wl8c}, "atac")
    # dscATAC-like: phase block 0-4 nt, 7-mers around anchors, then the insert in the same read
    out["dsc"] = make_layout([
        dict(read_id="R1", elems=[var(0, 4, "phase"), bc("b1", 7), linker("TATGCATGAC"), bc("b2", 7), linker("AGTCACTGAG"),
                                  bc("b3", 7), linker("TCGTCGGCAGCGTC"), gdna(1, 1000)], max_len=100),
        dict(read_id="R2", elems=[gdna(1, 1000)], is_reverse=True, max_len=60)],
        {"b1": wl7, "b2": wl7[:40], "b3": wl7[10:]}, "atac")
    # long barcode (64-bit keys), UMI of 16 (64-bit), reverse strand with a variable run and an anchor
    out["k64_rev"] = make_layout([
        dict(read_id="R1", elems=[bc("cb", 24), umi(16)], max_len=40),
        dict(read_id="R2", elems=[var(1, 3), linker("GATTACA"), umi(5, "umi2"), cdna(1, 200)], is_reverse=True, max_len=80)],
        {"cb": wl24})
    # two UMI elements in one read (the later one wins inside a file), UMI in a second file too, targets both strands
    out["multi_umi"] = make_layout([
        dict(read_id="R1", elems=[umi(4, "u1"), bc("cb", 16), umi(6, "u2"), cdna(2, 30)], max_len=60),
        dict(read_id="R2", elems=[umi(3, "u3"), gdna(1, 100)], is_reverse=True, max_len=40)], {"cb": wl16})
    # the reference's own vector layouts (segment.rs:687-705) with a barcode whitelist
    wl4, wl2 = U.gen_whitelist(rng, 40, 4), U.gen_whitelist(rng, 10, 2)
    out["vec11"] = make_layout([
        dict(read_id="R1", elems=[bc("b4", 4), var(1, 3), umi(2), linker("ACTGG"), var(2, 4, "v2"), bc("b2a", 2), linker("GGGG", "l2"),
                                  var(2, 4, "v3"), bc("b2b", 2), linker("GGGG", "l3"), cdna(1, 100)], max_len=0)],
        {"b4": wl4, "b2a": wl2, "b2b": wl2[:7]})
    # joined remainder [TO], the same whitelist used by two elements, and a trailing variable barcode
    out["joined"] = make_layout([
        dict(read_id="R1", elems=[bc("b4", 4), umi(2), cdna(2, 10), poly(2, 10), bc("b4", 4), umi(2, "u2")], max_len=0),
        dict(read_id="R2", elems=[umi(3, "u3"), bc("b4", 3, 5)], max_len=0)],
        {"b4": wl4})
    return out


@pytest.fixture(scope="module")
def layouts():
    return _layouts(np.random.default_rng(20240607))


NAMES = ["v3", "fast12_tail", "v2", "atac", "multiome_atac", "sci3", "share", "dsc", "k64_rev", "multi_umi", "vec11", "joined"]


@pytest.mark.parametrize("name", NAMES)
@pytest.mark.parametrize("mode", ["correct", "nocorrect", "ee", "lower", "correct-generic", "correct-direct", "lower-direct", "correct-nobits",
                                  "lower-generic"])
def test_emul_matches_oracle(layouts, name, mode):
    # default: specialised kernels + neighbourhood tables; "-generic": interpreter kernels only;
    # "-direct": specialised kernels with the 3*L direct probes instead of the neighbourhood tables;
    # "-nobits": neighbourhood tables without the partial-key bitmaps in front of them
    use_fast = {"generic": 0, "direct": 2, "nobits": 3}.get(mode.split("-")[-1], 1)
    mode = mode.split("-")[0]
    layout = layouts[name]
    rng = np.random.default_rng(U.seed_of(name, mode))
    infos, strides = U.emul_infos(layout)
    n = 700
    # the reference asserts equal totals across whitelists (barcode.rs:126-129): with several barcode
    # regions only reads too short for every barcode may be truncated
    multi = len(layout.region_ids) > 1
    full = U.gen_reads(rng, layout, n, strides, p_lower=0.02 if mode == "lower" else 0.0,
                       p_short=0.0 if multi else 0.03, p_tiny=0.03 if multi else 0.0)
    dev = U.narrow(full, strides, infos)
    kw = dict(correct=mode != "nocorrect", threshold=0.975 if mode != "lower" else 0.6,
              max_expected_errors=1.5 if mode == "ee" else _ffi.F64_MAX)
    infos, results, counts, defects = U.run_emul(layout, dev, use_fast=use_fast, **kw)
    ores, ocounts = U.run_oracle(layout, full, **kw)
    g = U.compare_with_oracle(layout, infos, full, results, counts, defects, ores, ocounts, correct=kw["correct"])
    # the test data must exercise the interesting branches
    codes = np.concatenate([b["code"] for b in g.bc])
    assert (codes == _ffi.BC_EXACT).any() and (codes == _ffi.BC_ABSENT).any()
    if mode == "correct":
        assert (codes == _ffi.BC_CORRECTED).any() and (codes == _ffi.BC_NO_MATCH).any()
    if mode == "ee":
        assert (codes == _ffi.BC_EXCEED_EE).any()


def test_reference_vectors_through_device_split():
    """The 11 known-answer vectors of segment.rs:617-705 through pcl_split (element coordinates)."""
    from tests.test_oracle_golden import GOLDEN
    for read, oelems, expected in GOLDEN:
        seq = "".join(read.split()).encode()
        elems, wls = [], {}
        for k, e in enumerate(oelems):
            rid = f"{e.region_id}{k}"
            if e.rtype == O.BARCODE:                 # one shared whitelist: totals stay equal when a barcode is swallowed
                elems.append(bc("wl", e.min_len, e.max_len)); wls["wl"] = [b"AAAA", b"AATT", b"CA", b"AT"]
            elif e.rtype == O.UMI:
                elems.append(umi(e.min_len, rid))
            elif e.rtype == O.TARGET:
                elems.append(cdna(e.min_len, e.max_len, rid))
            elif e.seq_fixed:
                elems.append(linker(e.sequence, rid))
            else:
                elems.append(var(e.min_len, e.max_len, rid))
        layout = make_layout([dict(read_id="R", elems=elems, max_len=0)], wls)
        infos, strides = U.emul_infos(layout)
        w = max(strides[0], len(seq))
        sq = np.full((1, w), ord("N"), np.uint8); sq[0, :len(seq)] = np.frombuffer(seq, np.uint8)
        ql = np.full((1, w), ord("@"), np.uint8)
        full = [U.HostBatch(sq, ql, np.array([len(seq)], np.uint16))]
        infos, results, counts, defects = U.run_emul(layout, U.narrow(full, strides, infos))
        ores, ocounts = U.run_oracle(layout, full)
        U.compare_with_oracle(layout, infos, full, results, counts, defects, ores, ocounts)
        # and the segments named in the expected string are where the device says they are
        res = results[0]
        want = dict()
        for part in expected.split(","):
            tag, s = part[1:part.index("]")], part[part.index("]") + 1:]
            want.setdefault(tag, []).append(s)
        got_b = []
        for j in range(infos[0].n_bc):
            c = int(res.bcoord[j][0]) if res.bcoord[j] is not None else None
            if c is None:
                code = (int(res.fstatus[0]) >> (4 + 3 * j)) & 7
                if code != _ffi.BC_ABSENT:
                    st = infos[0].static_start[infos[0].bc_elem[j]]
                    got_b.append(seq[st:st + elems[infos[0].bc_elem[j]].max_len].decode())
            elif c != _ffi.COORD_ABSENT:
                got_b.append(seq[c >> 16:(c >> 16) + (c & 0xFFFF)].decode())
        assert got_b == want.get("B", []), (read, got_b, want)
        if res.tcoord is not None and int(res.fstatus[0]) & _ffi.FSTATUS_TARGET:
            c = int(res.tcoord[0])
            t = seq[c >> 16:(c >> 16) + (c & 0xFFFF)].decode()
            assert [t] == (want.get("T") or want.get("TO")), (read, t, want)
        else:
            assert "T" not in want and "TO" not in want


def test_fast_path_selection(layouts):
    """Which reads take the register-resident / length-only kernels (1 / 2), the AnchorPlan kernel (3), and which
    the interpreter (0)."""
    L = U.emul_lib()
    def sel(name):
        d = layouts[name].descs()
        return [L.emu_fast_enabled(C.byref(d[r])) for r in range(len(layouts[name].reads))]
    import ctypes as C
    assert sel("v3") == [1, 2]
    assert sel("atac") == [2, 1, 2]
    assert sel("multiome_atac") == [2, 1, 2]
    assert sel("fast12_tail") == [1, 1]
    assert sel("sci3") == [3, 2]                 # [variable barcode][anchor][UMI][barcode]
    assert sel("k64_rev") == [0, 3]              # 64-bit keys -> interpreter; neg strand [variable][anchor][UMI][target] -> plan
    assert sel("dsc") == [0, 2]                  # a fixed barcode between the variable run and the anchor: interpreter
    assert sel("share") == [2, 0]
    # pcl_read_info.route: the compile-time instances (10x v3 R1 field geometry, sci-RNA-seq3 R1 plan geometry)
    from precellar_b200 import _ffi
    route = lambda name: [i.route for i in U.emul_infos(layouts[name])[0]]
    assert _ffi.ROUTE_PLAN_GEO == 4
    assert route("sci3") == [_ffi.ROUTE_PLAN_GEO, _ffi.ROUTE_LENONLY]
    assert route("v3") == [_ffi.ROUTE_FAST_SPEC, _ffi.ROUTE_LENONLY]
    assert route("k64_rev")[1] == _ffi.ROUTE_PLAN and route("dsc")[0] == _ffi.ROUTE_WALK


def _dense_layouts(rng):
    """Whitelists that stress the tables: every 6-mer (all buckets overflow, neighbourhood chains are long), 15- and
    16-mers (marker vs marker-less 32-bit keys), a 15/16 mix (64-bit table), 31-mers (longest packable barcode)."""
    import itertools
    out = {}
    all6 = [bytes(p) for p in itertools.product(b"ACGT", repeat=6)]
    rng.shuffle(all6)
    out["all6"] = make_layout([dict(read_id="R1", elems=[bc("cb", 6), umi(4)], max_len=10)], {"cb": all6[:4000]})
    out["len15"] = make_layout([dict(read_id="R1", elems=[bc("cb", 15), umi(5)], max_len=20)], {"cb": U.gen_whitelist(rng, 500, 15)})
    out["mix15_16"] = make_layout([dict(read_id="R1", elems=[bc("cb", 15, 16)], max_len=16)],
                                  {"cb": U.gen_whitelist(rng, 300, 15) + U.gen_whitelist(rng, 300, 16)})
    out["len31"] = make_layout([dict(read_id="R1", elems=[bc("cb", 31), umi(9)], max_len=40)], {"cb": U.gen_whitelist(rng, 300, 31)})
    # near-duplicate whitelist: many entries within Hamming distance 1 of each other -> several candidates per barcode
    base = U.gen_whitelist(rng, 40, 16)
    near = []
    for b in base:
        near.append(b)
        for pos in (0, 7, 15):
            for c in b"ACGT":
                v = bytearray(b); v[pos] = c
                near.append(bytes(v))
    out["near_dups"] = make_layout([dict(read_id="R1", elems=[bc("cb", 16), umi(12)], max_len=28)], {"cb": near})
    return out


@pytest.mark.parametrize("name", ["all6", "len15", "mix15_16", "len31", "near_dups"])
@pytest.mark.parametrize("use_fast", [0, 1, 2, 3])
def test_emul_dense_and_boundary_whitelists(name, use_fast):
    layout = _dense_layouts(np.random.default_rng(77))[name]
    rng = np.random.default_rng(U.seed_of(name, "dense"))
    infos, strides = U.emul_infos(layout)
    full = U.gen_reads(rng, layout, 1500, strides, p_sub=0.05, p_n=0.02, hot=10 ** 6)
    dev = U.narrow(full, strides, infos)
    infos, results, counts, defects = U.run_emul(layout, dev, threshold=0.6, use_fast=use_fast)
    ores, ocounts = U.run_oracle(layout, full, threshold=0.6)
    g = U.compare_with_oracle(layout, infos, full, results, counts, defects, ores, ocounts)
    codes = g.bc[0]["code"]
    assert (codes == _ffi.BC_CORRECTED).any() and (codes == _ffi.BC_EXACT).any()
    if name in ("near_dups", "all6"):
        assert (codes == _ffi.BC_LOW_CONF).any()          # several neighbours share the posterior


def _random_plan_layout(rng):
    """[fixed ...][one variable element][anchor][fixed ...][optional trailing variable element], random kinds."""
    wls, elems = {}, []
    n_bc = n_umi = n_t = 0

    def fixed_elem(k):
        nonlocal n_bc, n_umi, n_t
        u = rng.random()
        if u < 0.35 and n_bc < 3:
            L = int(rng.integers(4, 13))
            rid = f"b{n_bc}_{k}"
            wls[rid] = U.gen_whitelist(rng, int(rng.integers(20, 120)), L)
            n_bc += 1
            return bc(rid, L)
        if u < 0.55 and n_umi < 2:
            n_umi += 1
            return umi(int(rng.integers(3, 13)), f"u{k}")
        if u < 0.7:
            return linker(bytes(rng.choice(np.frombuffer(b"ACGT", np.uint8), int(rng.integers(2, 9)))).decode(), f"l{k}")   # taken at face value
        return U.E(f"o{k}", "linker", int(rng.integers(1, 7)))
    for k in range(int(rng.integers(0, 3))):
        elems.append(fixed_elem(k))
    a = int(rng.integers(0, 9))
    b = a + int(rng.integers(1, 7))
    if rng.random() < 0.5 and n_bc < 3 and b <= 15 and a >= 4:
        rid = f"vb{n_bc}"
        wls[rid] = U.gen_whitelist(rng, 80, list(range(a, b + 1)))
        n_bc += 1
        elems.append(bc(rid, a, b))
    else:
        elems.append(var(a, b, "v"))
    elems.append(linker(bytes(rng.choice(np.frombuffer(b"ACGT", np.uint8), int(rng.integers(1, 13)))).decode(), "anchor"))
    for k in range(int(rng.integers(0, 4))):
        elems.append(fixed_elem(10 + k))
    if rng.random() < 0.5:
        elems.append(cdna(1, 200) if rng.random() < 0.7 else var(0, 9, "tailv"))
    if n_bc == 0:
        wls["bx"] = U.gen_whitelist(rng, 50, 6)
        elems.insert(0, bc("bx", 6))
    total = sum(e.max_len for e in elems if e.max_len < 100)
    if total > 60:
        return None
    return make_layout([dict(read_id="R1", elems=elems, is_reverse=bool(rng.random() < 0.4), max_len=int(rng.choice([0, 0, total, total - 3])))], wls)


def test_emul_random_anchor_plan_layouts():
    """Random layouts of the AnchorPlan shape (both strands, variable barcodes, trailing targets, fixed sequences behind the
    anchor, Read.max_len cuts): the collapsed split + uniform-offset extraction must equal the oracle, and the
    interpreter path (use_fast=2) on the same data too."""
    import ctypes as C
    rng = np.random.default_rng(U.seed_of("random-plans"))
    L = U.emul_lib()
    n_plan = 0
    tries = 0
    while n_plan < 40 and tries < 400:
        tries += 1
        layout = _random_plan_layout(rng)
        if layout is None:
            continue
        d = layout.descs()
        try:
            infos, strides = U.emul_infos(layout)
        except _ffi.PclError:
            continue
        if L.emu_fast_enabled(C.byref(d[0])) != 3:
            continue
        n_plan += 1
        multi = len(layout.region_ids) > 1
        full = U.gen_reads(rng, layout, 250, strides, p_sub=0.04, p_n=0.02, p_lower=0.02 if layout.reads[0].is_reverse else 0.0,
                           p_short=0.0 if multi else 0.15, p_tiny=0.05 if multi else 0.0, p_anchor_mut=0.3)
        dev = U.narrow(full, strides, infos)
        try:
            ores, ocounts = U.run_oracle(layout, full, threshold=0.8)
        except Exception:
            n_plan -= 1                                   # the reference asserts equal totals across whitelists: not a split question
            continue
        for use_fast in (1, 2):
            infos2, results, counts, defects = U.run_emul(layout, dev, threshold=0.8, use_fast=use_fast)
            U.compare_with_oracle(layout, infos2, full, results, counts, defects, ores, ocounts)
    assert n_plan >= 25, (n_plan, tries)


def test_emul_whitelist_with_duplicates_keeps_file_order():
    """WhitelistBuilder::new collects the barcodes into a map and Onlist::read into an IndexSet (file order, duplicates
    dropped, region.rs:459-469): the table builder must number the entries the same way when it is handed duplicates."""
    rng = np.random.default_rng(U.seed_of("dups"))
    base = U.gen_whitelist(rng, 300, 16)
    wl = []
    for b in base:
        wl.append(b)
        if rng.random() < 0.3:
            wl.append(base[int(rng.integers(len(base)))])          # an earlier (or later) entry again
    layout = make_layout([dict(read_id="R1", elems=[bc("cb", 16), umi(10)], max_len=26)], {"cb": wl})
    infos, strides = U.emul_infos(layout)
    full = U.gen_reads(rng, layout, 3000, strides, p_sub=0.04, p_n=0.01, hot=10 ** 6)
    dev = U.narrow(full, strides, infos)
    ores, ocounts = U.run_oracle(layout, full, threshold=0.9)
    uniq = list(dict.fromkeys(wl))
    assert len(uniq) < len(wl) and len(ocounts[0][0]) == len(uniq)
    for use_fast in (0, 1):
        infos2, results, counts, defects = U.run_emul(layout, dev, threshold=0.9, use_fast=use_fast)
        assert len(counts[0][0]) == len(uniq)
        U.compare_with_oracle(layout, infos2, full, results, counts, defects, ores, ocounts)
