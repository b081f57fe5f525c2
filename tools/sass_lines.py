"""Static SASS instruction count per source line of one kernel (developer tool; needs a cubin built with -lineinfo).

    python tools/sass_lines.py /tmp/t.cubin k_count_planILb0E11PlanGeoSci3 [min_count]
"""
import collections
import re
import subprocess
import sys


def main():
    cubin, pat = sys.argv[1], sys.argv[2]
    floor = int(sys.argv[3]) if len(sys.argv) > 3 else 4
    out = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
    cur_fn, cur_line = None, None
    counts = collections.Counter()
    inl = collections.Counter()
    for ln in out.splitlines():
        m = re.match(r"\s*\.text\.(\S+):", ln)
        if m:
            cur_fn = m.group(1)
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
        if m:
            cur_line = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        if cur_fn and pat in cur_fn and re.match(r"\s+/\*[0-9a-f]{4}\*/\s+\S", ln):
            counts[cur_line] += 1
    total = sum(counts.values())
    print(f"{total} instructions with line info in *{pat}*")
    cache = {}
    for (f, l), c in sorted(counts.items(), key=lambda kv: -kv[1]):
        if c < floor:
            break
        src = ""
        for d in ("precellar_b200/csrc/", "include/"):
            try:
                cache.setdefault(d + f, open(d + f).read().splitlines())
                src = cache[d + f][l - 1].strip()[:110]
                break
            except OSError:
                pass
        print(f"{c:5d}  {f}:{l}  {src}")


if __name__ == "__main__":
    main()
