#!/usr/bin/env python
"""profiles/traffic.json from the ncu summaries of a measurement pass (made by profiles/summarize.py in tools/gpu_measure.sh).
  python tools/update_traffic.py <tag> [<commit>]     e.g. r2y; reads profiles/<tag>_k_*_full.txt
The table holds the per-launch counters bench.py cannot measure itself (DRAM bytes, L1TEX sectors, L2 hit rate, pipe
utilisation); bench.py combines them with the kernel time it measures live."""
import json, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MAP = {  # summary file -> (table key, note)
    "k_align_full.txt": ("k_align/scan2map/scans=65536/res=0.25/K=1", "the default bench configuration (room world)"),
    "k_align_dense_full.txt": ("k_align/dense/scans=16384/res=0.25/K=1", "dense world (bench.py --world dense --scans 16384; the `dense` leg of the default line)"),
    "k_eval_poses_sweep_full.txt": ("k_eval_poses/sweep/hyps=1000000/res=0.25/K=1", "configs[3] on one GPU"),
    "k_pairs_fused_full.txt": ("k_pairs_fused/odometry/scans=16384/res=0.5/K=1", "bench.py --workload odometry: 16 383 consecutive pairs, fused shared-memory path"),
}
FIELDS = {"dram_read": "dram__bytes_read.sum", "dram_write": "dram__bytes_write.sum", "l1tex_sectors": "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
          "l1tex_requests": "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l2_hit_pct": "lts__t_sector_hit_rate.pct",
          "inst_executed": "smsp__inst_executed.sum", "kernel_ms_under_ncu": "gpu__time_duration.sum",
          "issue_active_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
          "fma_pipe_cycles_active_pct": "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
          "l1tex_data_pipe_pct": "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"}
SCALE = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0, "ms": 1.0, "us": 1e-3, "s": 1e3}


def parse(path):
    out = {}
    for line in open(path):
        m = re.match(r"\s+(\S+)\s+([-0-9.e+]+)\s*(\S*)", line)
        if m:
            out[m.group(1)] = float(m.group(2)) * SCALE.get(m.group(3), 1.0)
    return out


def main():
    tag = sys.argv[1]
    commit = sys.argv[2] if len(sys.argv) > 2 else subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    path = os.path.join(ROOT, "profiles", "traffic.json")
    tab = json.load(open(path))
    for suffix, (key, note) in MAP.items():
        f = os.path.join(ROOT, "profiles", f"{tag}_{suffix}")
        if not os.path.exists(f):
            print("missing", f); continue
        v = parse(f)
        e = {k: v[m] for k, m in FIELDS.items() if m in v}
        e["dram_bytes"] = e.get("dram_read", 0.0) + e.get("dram_write", 0.0)
        e.update(source=f"profiles/{tag}_{suffix}", commit=commit, spec="v4", note=note)
        for keep in ("mix_floor_cycles", "loop_sample_share", "mix_note"):   # from tools/mix_probe.py / tools/srcstall.py, not from a summary
            if keep in tab.get(key, {}):
                e[keep] = tab[key][keep]
        tab[key] = e
        print(key, {k: e[k] for k in ("dram_bytes", "l1tex_sectors", "kernel_ms_under_ncu") if k in e})
    json.dump(tab, open(path, "w"), indent=1)


if __name__ == "__main__":
    main()
