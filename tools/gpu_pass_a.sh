#!/bin/bash
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pa_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pa_pytest.log
for c in 8192 4096 2048; do
  NDT2D_CHUNK_SCANS=$c python bench.py --no-cpu-baseline > gpurun_out/pa_chunk_$c.json 2>/dev/null
  python -c "
import json;d=json.load(open('gpurun_out/pa_chunk_$c.json'));print('chunk $c value',round(d['value']/1e6,2),'e2e',{k:round(v['value']/1e6,2) for k,v in d['e2e']['by_input'].items()})"
done
