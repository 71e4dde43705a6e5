#!/bin/bash
# round-1 measurement pass "${TAG}": parity, bench lines, ncu launch list + full captures
set -u
TAG=${1:-r1g}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/${TAG}_pytest.log
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python bench.py --workload sweep > gpurun_out/${TAG}_sweep.json 2> gpurun_out/${TAG}_sweep.err; echo "sweep rc=$?"
python bench.py --workload sweep --overlap 1 --no-cpu-baseline > gpurun_out/${TAG}_sweep_k4.json 2>> gpurun_out/${TAG}_sweep.err
python bench.py --workload pyramid --no-cpu-baseline > gpurun_out/${TAG}_pyramid.json 2> gpurun_out/${TAG}_pyramid.err; echo "pyramid rc=$?"
python bench.py --overlap 1 --no-cpu-baseline > gpurun_out/${TAG}_k4.json 2> gpurun_out/${TAG}_k4.err; echo "k4 rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_launch.log 2>&1; echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_align -s 1 -c 1 -f -o gpurun_out/${TAG}_prof_align \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_align.log 2>&1; echo "ncu align rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_eval_poses -s 1 -c 1 -f -o gpurun_out/${TAG}_prof_sweep \
    python bench.py --workload sweep --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_sweep.log 2>&1; echo "ncu sweep rc=$?"
head -c 600 gpurun_out/${TAG}_bench.json; echo; head -c 400 gpurun_out/${TAG}_sweep.json; echo
