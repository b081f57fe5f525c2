"""Summarise an .ncu-rep (developer tool): one row per profiled launch with the metrics the roofline discussion uses.

    python tools/ncu_summary.py gpurun_out/x.ncu-rep [out.csv]
"""
import csv
import io
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "time_us"),
    ("dram__bytes_read.sum", "dram_read_MB"), ("dram__bytes_write.sum", "dram_write_MB"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
    ("lts__t_sectors.sum", "l2_sectors_M"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1_ld_sectors_M"),
    ("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1_ld_requests_M"),
    ("sm__inst_executed.sum", "warp_inst_M"),
    ("sm__inst_executed.avg.per_cycle_active", "ipc_per_sm"),
    ("smsp__issue_active.avg.pct", "issue_active_pct"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu_pct"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma_pct"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy_pct"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"),
]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    out = [["kernel"] + [n for _, n in WANT]]
    for r in data:
        line = [r[col["Kernel Name"]].split("(")[0]]
        for m, n in WANT:
            if m not in col:
                line.append("")
                continue
            v = r[col[m]].replace(",", "")
            u = units[col[m]]
            try:
                f = float(v)
                if n.endswith("_MB"):
                    f = f / 1e6 if u == "byte" else f * {"Kbyte": 1e-3, "Mbyte": 1, "Gbyte": 1e3}.get(u, 1)
                elif n.endswith("_M"):
                    f = f / 1e6
                elif n == "time_us":
                    f = f * {"ns": 1e-3, "us": 1, "ms": 1e3, "second": 1e6}.get(u, 1)
                line.append(f"{f:.4g}")
            except ValueError:
                line.append(v)
        out.append(line)
    w = csv.writer(open(sys.argv[2], "w", newline="") if len(sys.argv) > 2 else sys.stdout)
    w.writerows(out)


if __name__ == "__main__":
    main()
