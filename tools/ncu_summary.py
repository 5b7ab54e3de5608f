#!/usr/bin/env python
"""Prints the per-kernel metrics this repo quotes from an .ncu-rep (run where ncu is installed; no GPU needed):
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [more metric substrings ...]"""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers",
    "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_fma.sum",
    "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_xu.sum",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "l1tex__t_bytes.sum", "lts__t_bytes.sum",
]


def main():
    rep = sys.argv[1]
    extra = sys.argv[2:]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("==", r[hdr.index("Kernel Name")][:100])
        for k in KEYS:
            if k in hdr:
                print(f"  {k:75s} {r[hdr.index(k)]:>18s} {units[hdr.index(k)]}")
        for i, h in enumerate(hdr):
            if ("issue_stalled" in h and h.endswith("per_warp_active.pct")) or any(e in h for e in extra):
                try:
                    v = float(r[i])
                except ValueError:
                    continue
                if v >= 1.0 or any(e in h for e in extra):
                    print(f"  {h:75s} {r[i]:>18s} {units[i]}")


if __name__ == "__main__":
    main()
