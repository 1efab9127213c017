#!/usr/bin/env python3
"""Key metrics per profiled launch from an `ncu --set full` report (read here, no GPU needed).
usage: summarise_ncu.py <report.ncu-rep>"""
import csv
import io
import subprocess
import sys

WANT = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__thread_inst_executed_pred_on_per_inst_executed.ratio", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "l1tex__t_bytes.sum", "lts__t_bytes.sum"]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for n, r in enumerate(rows[2:]):
        print(f"--- launch {n}")
        for w in WANT:
            if w in idx:
                print(f"{w} = {r[idx[w]]} {units[idx[w]]}")
        stalls = [(float(r[i]), h) for h, i in idx.items() if h.startswith("smsp__average_warp") and "issue_stalled" in h and h.endswith("_per_issue_active.ratio") and r[i]]
        if not stalls:
            stalls = [(float(r[i]), h) for h, i in idx.items() if "warp_issue_stalled" in h and h.endswith(".pct") and r[i]]
        for v, h in sorted(stalls, reverse=True)[:6]:
            print(f"stall {h} = {v:.3f}")


if __name__ == "__main__":
    main(sys.argv[1])
