#!/bin/bash
# Measurement pass "${TAG}" on one B200 (round 2): parity tests, both arms of the default bench line (every leg), the
# other workloads, the ncu launch list and one --set full capture per dominant kernel (room and dense world, sweep, fused
# pairs). Summaries are made here (ncu is on the box as well) so that the .ncu-rep files need not travel.
set -u
TAG=${1:-r2z}
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -q --timeout 900 > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/${TAG}_pytest.log
python bench.py --impl reference --steps 5 --warmup 1 > $O/${TAG}_reference.json 2> $O/${TAG}_reference.err; echo "reference rc=$?"
python bench.py --steps 20 --warmup 5 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench rc=$?"; tail -3 $O/${TAG}_bench.err
python bench.py --workload odometry > $O/${TAG}_odometry.json 2> $O/${TAG}_odometry.err; echo "odometry rc=$?"
python bench.py --workload odometry --res 2.0 1.0 0.5 --perturb 0.1 1.0 > $O/${TAG}_odometry_pyramid.json 2>> $O/${TAG}_odometry.err; echo "odometry pyramid rc=$?"
python bench.py --workload sweep > $O/${TAG}_sweep.json 2> $O/${TAG}_sweep.err; echo "sweep rc=$?"
python bench.py --workload sweep --overlap 1 > $O/${TAG}_sweep_k4.json 2>> $O/${TAG}_sweep.err; echo "sweep k4 rc=$?"
python bench.py --overlap 1 --no-cpu-baseline > $O/${TAG}_k4.json 2> $O/${TAG}_k4.err; echo "k4 rc=$?"
python bench.py --workload newton > $O/${TAG}_newton.json 2> $O/${TAG}_newton.err; echo "newton rc=$?"
python bench.py --workload build > $O/${TAG}_build.json 2> $O/${TAG}_build.err; echo "build rc=$?"
python bench.py --world dense --scans 16384 --no-cpu-baseline > $O/${TAG}_dense.json 2> $O/${TAG}_dense.err; echo "dense rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --legs sweep,pyramid,odometry > $O/${TAG}_ncu_launch.log 2>&1; echo "ncu launches rc=$?"
cap() {  # name, kernel regex, skip, bench args...
    local name=$1 rx=$2 skip=$3; shift 3
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -f -o $O/${TAG}_prof_$name \
        python bench.py "$@" > $O/${TAG}_ncu_$name.log 2>&1; echo "ncu $name rc=$?"
    python profiles/summarize.py $O/${TAG}_prof_$name.ncu-rep > $O/${TAG}_${name}_full.txt 2>/dev/null
}
cap k_align k_align 1 --steps 1 --warmup 1 --no-cpu-baseline --legs none
cap k_align_dense k_align 1 --world dense --scans 16384 --steps 1 --warmup 1 --no-cpu-baseline
cap k_eval_poses_sweep k_eval_poses 1 --workload sweep --steps 1 --warmup 1
cap k_pairs_fused k_pairs_fused 2 --workload odometry --steps 2 --warmup 1
for f in reference bench odometry odometry_pyramid sweep sweep_k4 k4 newton build dense; do python - <<PY
import json
try:
    d = json.loads(open('$O/${TAG}_$f.json').read().strip().splitlines()[-1])
    print('$f', round(d['value'] / 1e6, 3), 'M', d['unit'], 'ms/step', round(d['ms_per_step'], 4), 'e2e', round(d['e2e']['value'] / 1e6, 3), 'iters', d.get('mean_iterations'))
except Exception as e:
    print('$f FAILED', e)
PY
done
