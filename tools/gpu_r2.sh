#!/bin/bash
# round-2 measurement pass "${TAG}": parity tests, the bench lines of every workload (no ncu)
set -u
TAG=${1:-r2a}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/${TAG}_pytest.log
python bench.py --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python bench.py --workload sweep > gpurun_out/${TAG}_sweep.json 2> gpurun_out/${TAG}_sweep.err; echo "sweep rc=$?"
python bench.py --workload pyramid --no-cpu-baseline > gpurun_out/${TAG}_pyramid.json 2> gpurun_out/${TAG}_pyramid.err; echo "pyramid rc=$?"
python bench.py --overlap 1 --no-cpu-baseline > gpurun_out/${TAG}_k4.json 2> gpurun_out/${TAG}_k4.err; echo "k4 rc=$?"
python bench.py --workload newton > gpurun_out/${TAG}_newton.json 2> gpurun_out/${TAG}_newton.err; echo "newton rc=$?"
python bench.py --workload odometry > gpurun_out/${TAG}_odometry.json 2> gpurun_out/${TAG}_odometry.err; echo "odometry rc=$?"
python bench.py --workload build > gpurun_out/${TAG}_build.json 2> gpurun_out/${TAG}_build.err; echo "build rc=$?"
for f in bench sweep pyramid k4 newton odometry build; do python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/${TAG}_$f.json').read().strip().splitlines()[-1])
    print('$f', round(d['value'] / 1e6, 3), 'M', d['unit'], 'ms/step', round(d['ms_per_step'], 4), 'e2e', round(d['e2e']['value'] / 1e6, 3), 'iters', d.get('mean_iterations'), 'status', d.get('status_counts'))
except Exception as e:
    print('$f FAILED', e); print(open('gpurun_out/${TAG}_$f.err').read()[-1500:])
PY
done
