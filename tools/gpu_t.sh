#!/bin/bash
python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -2
python -m pytest tests -x -q -m gpu 2>&1 | tail -2
python bench.py > gpurun_out/final_bench.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/final_bench.json').read().strip().splitlines()[-1]); print('value',round(d['value']/1e6,2),'e2e',round(d['e2e']['value']/1e6,2),d['e2e']['input'],'frac',round(d['roofline']['frac'],2),'traffic',d['roofline']['traffic'],'cpu',round(d['cpu_baseline']['value']),d['cpu_baseline']['cores'],'launches',d['gpu_launches'],d['clocks'])"
python bench.py --impl reference | cut -c1-330
