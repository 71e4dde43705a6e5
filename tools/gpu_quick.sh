#!/bin/bash
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/q_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/q_pytest.log
examples/slam_frontend /tmp/g.g2o; echo "slam rc=$?"
python bench.py --no-cpu-baseline 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value',round(d['value']/1e6,2),'e2e',{k:round(v['value']/1e6,2) for k,v in d['e2e']['by_input'].items()}, d['single_align_latency_us'])"
