#!/bin/bash
set -u
TAG=${1:-r2d}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -k "pairs or cpp_host or vs_f64" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/${TAG}_pytest.log
timeout 300 python bench.py --workload odometry > gpurun_out/${TAG}_odometry.json 2> gpurun_out/${TAG}_odometry.err; echo "odometry rc=$?"; tail -3 gpurun_out/${TAG}_odometry.err
timeout 300 python bench.py --workload odometry --res 2.0 1.0 0.5 --perturb 0.1 1.0 > gpurun_out/${TAG}_odometry_pyramid.json 2>> gpurun_out/${TAG}_odometry.err; echo "odometry pyramid rc=$?"
for f in odometry odometry_pyramid; do python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/${TAG}_$f.json').read().strip().splitlines()[-1])
    print('$f', round(d['value'] / 1e6, 3), 'M', d['unit'], 'ms/step', round(d['ms_per_step'], 4), 'e2e', round(d['e2e']['value'] / 1e6, 3), 'iters', d.get('mean_iterations'), 'status', d.get('status_counts'), d.get('sequential_set_target_plus_align'))
except Exception as e:
    print('$f FAILED', e)
PY
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pairs_fused -s 2 -c 1 -f -o gpurun_out/${TAG}_prof_pairs \
    python bench.py --workload odometry --steps 2 --warmup 1 > gpurun_out/${TAG}_ncu_pairs.log 2>&1; echo "ncu pairs rc=$?"
