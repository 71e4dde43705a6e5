#!/bin/bash
# round-2 pass "${TAG}": parity tests, the default bench line with every leg, reference arm, variants, ncu captures
set -u
TAG=${1:-r2b}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/${TAG}_pytest.log
python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; tail -5 gpurun_out/${TAG}_bench.err
python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/${TAG}_bench.json').read().strip().splitlines()[-1])
    print('value', round(d['value']/1e6,3), 'e2e', round(d['e2e']['value']/1e6,3), 'iters', d['mean_iterations'], d['e2e']['numa'], d['e2e']['per_rank'])
    for k in ('prior2','pyramid','dense'):
        print(k, round(d[k]['value']/1e6,3), d[k].get('mean_iterations'), d[k].get('frac_within_5cm_of_truth'), d[k].get('status_counts'))
    s = d['sweep']; print('sweep', round(s['value']/1e6,1), s['ms_per_query'], s['combine_equals_host_api_result'], s['exchange_check'], s['relocalize'])
    print('config0', d['config0']); print('precision', d['precision'])
    print('cpu', d['cpu_baseline'])
except Exception as e:
    print('bench line FAILED', e)
PY
python tools/exp.py gen > /dev/null 2>&1
python tools/exp.py run base jr1 mb2 2>/dev/null | tee gpurun_out/${TAG}_variants.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --legs none > gpurun_out/${TAG}_ncu_launch.log 2>&1; echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_align -s 1 -c 1 -f -o gpurun_out/${TAG}_prof_align \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --legs none > gpurun_out/${TAG}_ncu_align.log 2>&1; echo "ncu align rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_eval_poses -s 1 -c 1 -f -o gpurun_out/${TAG}_prof_sweep \
    python bench.py --workload sweep --steps 1 --warmup 1 > gpurun_out/${TAG}_ncu_sweep.log 2>&1; echo "ncu sweep rc=$?"
