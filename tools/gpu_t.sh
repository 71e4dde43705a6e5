#!/bin/bash
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "align_pairs" 2>&1 | tail -2
python bench.py --workload odometry 2>&1 | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']/1e6,3),'M/s', round(d['ms_per_step'],3),'ms e2e',round(d['e2e']['value']/1e6,3), d['sequential_set_target_plus_align'], d['status_counts'])"
python bench.py --workload odometry --res 2.0 1.0 0.5 --perturb 0.1 1.0 2>&1 | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']/1e6,3),'M/s', round(d['ms_per_step'],3),'ms e2e',round(d['e2e']['value']/1e6,3), d['mean_iterations'], d['status_counts'])"
ncu --metrics gpu__time_duration.sum --clock-control none --csv -k regex:"k_pairs_build|k_align" -c 4 python bench.py --workload odometry --steps 1 --warmup 1 2>/dev/null | grep -E "k_pairs|k_align" | awk -F'","' '{print $5, $NF}' | cut -c1-120
