#!/usr/bin/env python3
"""Condense an `ncu --set full` report (or a launch-list csv) into the few numbers DESIGN.md and
bench.py quote.  Runs here (no GPU needed): `ncu -i <rep> --page raw --csv` does the decoding.

    python profiles/summarize_ncu.py gpurun_out/x.ncu-rep   > profiles/rNN_x.summary.txt
    python profiles/summarize_ncu.py gpurun_out/launches.csv > profiles/rNN_launches.summary.txt
"""
import csv
import subprocess
import sys
from collections import OrderedDict

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("sm__cycles_elapsed.max", "cycles"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
    ("launch__occupancy_limit_registers", "occ limit regs (blocks)"),
    ("launch__occupancy_limit_shared_mem", "occ limit smem (blocks)"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sm__inst_executed.avg.per_cycle_elapsed", "IPC per SM"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
    ("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "FMA-heavy pipe % (IMAD)"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wavefronts"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_pipe_throttle"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall not_selected"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall mio_throttle"),
]


def short(name):
    name = name.replace("<unnamed>::", "").replace("void ", "")
    return name.split("(")[0][:60]


def report(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True)
    rows = list(csv.reader(raw.stdout.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    print("# %s: %d profiled launches (ncu --set full --clock-control none)" % (path.split("/")[-1], len(data)))
    for r in data:
        print("\n## %s  grid %s block %s" % (short(r[col["Kernel Name"]]), r[col["Grid Size"]], r[col["Block Size"]]))
        for key, label in KEYS:
            if key in col:
                print("  %-28s %s %s" % (label, r[col[key]], units[col[key]]))


def launches(path):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    ki, vi, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
    agg = OrderedDict()
    total = 0.0
    for r in rows[1:]:
        key = (short(r[ki]), r[gi])
        v = float(r[vi].replace(",", "")) / 1e3
        a = agg.setdefault(key, [0, 0.0])
        a[0] += 1
        a[1] += v
        total += v
    print("# %s: %d launches, %.1f us total (gpu__time_duration.sum, cold-cache, serialised)" %
          (path.split("/")[-1], len(rows) - 1, total))
    print("%-62s %-16s %6s %10s %7s" % ("kernel", "grid", "count", "avg us", "share"))
    for (k, g), (c, t) in agg.items():
        print("%-62s %-16s %6d %10.2f %6.1f%%" % (k, g, c, t / c, 100 * t / total))


if __name__ == "__main__":
    p = sys.argv[1]
    (launches if p.endswith(".csv") else report)(p)
