#!/bin/bash
# round-1 measurement pass "r1f": parity, bench lines, ncu launch list + full captures
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r1f_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r1f_pytest.log
python bench.py > gpurun_out/r1f_bench.json 2> gpurun_out/r1f_bench.err; echo "bench rc=$?"
python bench.py --workload sweep > gpurun_out/r1f_sweep.json 2> gpurun_out/r1f_sweep.err; echo "sweep rc=$?"
python bench.py --workload sweep --overlap 1 --no-cpu-baseline > gpurun_out/r1f_sweep_k4.json 2>> gpurun_out/r1f_sweep.err
python bench.py --workload pyramid --no-cpu-baseline > gpurun_out/r1f_pyramid.json 2> gpurun_out/r1f_pyramid.err; echo "pyramid rc=$?"
python bench.py --overlap 1 --no-cpu-baseline > gpurun_out/r1f_k4.json 2> gpurun_out/r1f_k4.err; echo "k4 rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1f_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r1f_ncu_launch.log 2>&1; echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_align -s 1 -c 1 -f -o gpurun_out/r1f_prof_align \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r1f_ncu_align.log 2>&1; echo "ncu align rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_eval_poses -s 1 -c 1 -f -o gpurun_out/r1f_prof_sweep \
    python bench.py --workload sweep --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r1f_ncu_sweep.log 2>&1; echo "ncu sweep rc=$?"
head -c 600 gpurun_out/r1f_bench.json; echo; head -c 400 gpurun_out/r1f_sweep.json; echo
