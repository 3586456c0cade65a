#!/usr/bin/env python
"""Text summary of an ncu report for profiles/: one column per captured launch, the metrics the design discussion uses.
    python tools/ncu_summary.py report.ncu-rep "header comment" > profiles/rNN_xxx_ncu_summary.txt"""
import csv
import subprocess
import sys

METRICS = [
    "Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "sm__cycles_elapsed.avg",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor", "smsp__inst_executed.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
]


def main():
    rep, note = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    print("# " + note)
    print("# source: %s (ncu --set full --clock-control none; one column per captured launch; serialised, cold-ish caches:" % rep)
    print("# use for shares and utilisation, not absolute time)")
    for m in METRICS:
        if m in hdr:
            i = hdr.index(m)
            print("%s [%s]: %s" % (m, units[i], " | ".join(r[i][:44] for r in data)))


if __name__ == "__main__":
    main()
