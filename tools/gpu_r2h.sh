#!/bin/bash
set -u
TAG=${1:-r2h}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/${TAG}_pytest.log
timeout 300 python bench.py --workload odometry > gpurun_out/${TAG}_odometry.json 2> gpurun_out/${TAG}_odometry.err; echo "odometry rc=$?"; tail -3 gpurun_out/${TAG}_odometry.err
python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; tail -5 gpurun_out/${TAG}_bench.err
python - <<PY
import json
for f in ('odometry', 'bench'):
    try:
        d = json.loads(open('gpurun_out/${TAG}_%s.json' % f).read().strip().splitlines()[-1])
        print(f, round(d['value']/1e6,3), 'e2e', round(d['e2e']['value']/1e6,3), 'iters', d.get('mean_iterations'))
        if f == 'bench':
            for k in ('prior2','pyramid','dense','odometry'):
                print(' ', k, round(d[k]['value']/1e6,3), d[k].get('mean_iterations'))
            s = d['sweep']; print('  sweep', round(s['value']/1e6,1), s['ms_per_query'], s['ok'], s['relocalize']['ms_per_query'])
            print('  single align', d['single_align_latency_us'], 'config0', d['config0']['gpu_set_target_plus_align_ms'], d['config0']['cpu_oracle_1_thread_ms'])
    except Exception as e:
        print(f, 'FAILED', e)
PY
