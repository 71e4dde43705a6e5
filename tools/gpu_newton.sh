#!/bin/bash
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/nw_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/nw_pytest.log
python bench.py --workload newton > gpurun_out/r1i_newton.json 2> gpurun_out/r1i_newton.err; echo "newton rc=$?"; cut -c1-1800 gpurun_out/r1i_newton.json
python bench.py --workload newton --overlap 1 2>/dev/null | cut -c1-260
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_eval_poses -s 1 -c 1 -f -o gpurun_out/r1i_prof_newton \
    python bench.py --workload newton --steps 1 --warmup 1 > gpurun_out/r1i_ncu_newton.log 2>&1; echo "ncu rc=$?"
