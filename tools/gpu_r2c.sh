#!/bin/bash
set -u
TAG=${1:-r2c}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/${TAG}_pytest.log
timeout 300 python bench.py --workload odometry > gpurun_out/${TAG}_odometry.json 2> gpurun_out/${TAG}_odometry.err; echo "odometry rc=$?"; tail -3 gpurun_out/${TAG}_odometry.err
NDT2D_PAIRS_FUSED=0 timeout 300 python bench.py --workload odometry > gpurun_out/${TAG}_odometry_general.json 2> gpurun_out/${TAG}_odometry_general.err; echo "odometry general rc=$?"
timeout 300 python bench.py --workload odometry --res 2.0 1.0 0.5 --perturb 0.1 1.0 > gpurun_out/${TAG}_odometry_pyramid.json 2>> gpurun_out/${TAG}_odometry.err; echo "odometry pyramid rc=$?"
for f in odometry odometry_general odometry_pyramid; do python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/${TAG}_$f.json').read().strip().splitlines()[-1])
    print('$f', round(d['value'] / 1e6, 3), 'M', d['unit'], 'ms/step', round(d['ms_per_step'], 4), 'e2e', round(d['e2e']['value'] / 1e6, 3), 'iters', d.get('mean_iterations'), 'status', d.get('status_counts'), d.get('sequential_set_target_plus_align'))
except Exception as e:
    print('$f FAILED', e)
PY
done
