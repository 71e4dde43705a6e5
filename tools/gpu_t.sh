#!/bin/bash
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
examples/slam_frontend /tmp/g.g2o | grep -E "timing|batched|odometry:"
python bench.py --no-cpu-baseline 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value',round(d['value']/1e6,2),'e2e',round(d['e2e']['value']/1e6,2), d['single_align_latency_us'])"
