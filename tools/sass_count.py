"""Static SASS statistics of one kernel of libprecellar_b200.so (developer tool): instruction count, the opcode histogram,
and the instructions between the first backward branch target and that branch (the main loop).

    python tools/sass_count.py k_count_fastILi14ELi1 [path/to/lib.so]
"""
import collections
import re
import subprocess
import sys


def functions(lib):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    cur, body = None, collections.OrderedDict()
    for ln in out.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            body[cur] = []
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m and cur:
            body[cur].append((int(m.group(1), 16), m.group(2).strip()))
    return body


def main():
    pat = sys.argv[1]
    lib = sys.argv[2] if len(sys.argv) > 2 else "precellar_b200/libprecellar_b200.so"
    for name, ins in functions(lib).items():
        if pat not in name:
            continue
        ops = collections.Counter()
        for _, t in ins:
            t = re.sub(r"^@!?U?P\d+\s+", "", t)
            ops[t.split()[0].split(".")[0]] += 1
        # largest backward branch span = the main loop
        best = (0, 0, 0)
        for addr, t in ins:
            m = re.search(r"BRA\S*\s+(?:.*?)(0x[0-9a-f]+)", t)
            if m:
                tgt = int(m.group(1), 16)
                if tgt < addr and addr - tgt > best[0]:
                    best = (addr - tgt, tgt, addr)
        loop = [t for a, t in ins if best[1] <= a <= best[2]]
        lops = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", t).split()[0].split(".")[0] for t in loop)
        print(f"{name}: {len(ins)} instructions, main loop {len(loop)}")
        print("  loop ops:", dict(lops.most_common(14)))


if __name__ == "__main__":
    main()
