"""Warp-stall samples of one kernel per source line (developer tool): joins `ncu --page source --csv` (SASS addresses) with the
line table of the cubin inside the built library.

    python tools/ncu_src_lines.py gpurun_out/src_writer.csv k_zh_blocks [top]
"""
import collections, csv, glob, os, re, subprocess, sys, tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    path, pat = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
    td = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "precellar_b200", "libprecellar_b200.so")], cwd=td, capture_output=True)
    cubin = [c for c in glob.glob(os.path.join(td, "*.cubin")) if os.path.basename(c).startswith("pcl.")][0]
    out = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
    cur_fn = cur_line = None
    addr2line = {}
    for ln in out.splitlines():
        m = re.match(r"\s*\.text\.(\S+):", ln)
        if m:
            cur_fn = m.group(1)
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur_line = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+\S", ln)
        if m and cur_fn and pat in cur_fn and cur_fn.startswith("_Z") and "$" not in cur_fn:
            addr2line[int(m.group(1), 16)] = cur_line
    rows = list(csv.reader(open(path)))
    hdr = next(r for r in rows if r and r[0] == "Address")
    ia, isamp, iinst = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    base = None
    agg, agg_i, tot = collections.Counter(), collections.Counter(), 0
    for r in rows:
        if not r or not r[0].startswith("0x"):
            continue
        a = int(r[ia], 16)
        base = a if base is None else base
        s = int(r[isamp] or 0)
        tot += s
        l = addr2line.get(a - base)
        agg[l] += s
        agg_i[l] += int(r[iinst] or 0)
    print("samples", tot)
    cache = {}
    for l, c in agg.most_common(top):
        src = ""
        if l:
            f = os.path.join(ROOT, "precellar_b200", "csrc", l[0])
            if os.path.exists(f):
                cache.setdefault(f, open(f).read().splitlines())
                src = cache[f][l[1] - 1].strip()[:105]
        print(f"{c:8d} {100 * c / max(tot, 1):5.1f}%  inst={agg_i[l]:10d}  {l}  {src}")


if __name__ == "__main__":
    main()
