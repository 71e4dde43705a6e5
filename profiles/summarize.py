#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed) into a small text file for profiles/.
usage: python profiles/summarize.py gpurun_out/prof.ncu-rep > profiles/<name>.txt"""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_active.avg.per_cycle_active",
        "smsp__warps_eligible.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum",
        "l1tex__t_sector_hit_rate.pct", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sm__cycles_elapsed.max", "smsp__cycles_active.avg"]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(l for l in raw.splitlines() if l.startswith('"')))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("kernel:", r[hdr.index("Kernel Name")])
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"  {k:75s} {r[i]:>18s} {units[i]}")
        stalls = []
        for i, h in enumerate(hdr):
            if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h:
                try:
                    stalls.append((float(r[i]), h.replace("smsp__pcsamp_warps_issue_stalled_", "")))
                except ValueError:
                    pass
        tot = sum(v for v, _ in stalls) or 1.0
        print("  stall reasons (pc sampling): " + ", ".join(f"{h} {100 * v / tot:.1f}%" for v, h in sorted(stalls, reverse=True)[:8]))
        print()


if __name__ == "__main__":
    main(sys.argv[1])
