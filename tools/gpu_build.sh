#!/bin/bash
set -u
mkdir -p gpurun_out
python bench.py --workload build > gpurun_out/r1h_build.json 2> gpurun_out/r1h_build.err; echo "build rc=$?"; cat gpurun_out/r1h_build.json | cut -c1-1500
python bench.py --workload build --res 2.0 1.0 0.5 2>/dev/null | cut -c1-300
python bench.py --workload build --overlap 1 2>/dev/null | cut -c1-300
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_accumulate|k_finalize" -s 4 -c 2 -f -o gpurun_out/r1h_prof_build \
    python bench.py --workload build --steps 2 --warmup 3 > gpurun_out/r1h_ncu_build.log 2>&1; echo "ncu rc=$?"
