#!/bin/bash
python bench.py --no-cpu-baseline 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']/1e6,2), d['roofline']['frac'], d['roofline']['gather_pipe'], round(d['e2e']['value']/1e6,2))"
python bench.py --workload sweep --no-cpu-baseline 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']/1e6,2), d['roofline']['frac'], d['roofline']['gather_pipe'])"
python bench.py --impl reference --steps 2 --warmup 1 | cut -c1-200
