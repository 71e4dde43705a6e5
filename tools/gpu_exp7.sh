#!/bin/bash
set -u
mkdir -p gpurun_out
python tools/exp.py gen --scans 65536 2>&1 | tail -1
{ python tools/exp.py run base ld0 ld5 ld6 ld7 b2 pipe2 u2; python tools/exp.py run base ld5 ld6 ld7 --overlap 1 --scans 32768; } > gpurun_out/x7_variants.jsonl 2> gpurun_out/x7_variants.err; cat gpurun_out/x7_variants.jsonl
python bench.py --workload sweep --no-cpu-baseline > gpurun_out/x7_sw_base.json 2>/dev/null
for v in ld5 ld6 ld7 sw_b6 sw_b8 sw_p4 sw_p6; do
  NDT2D_LIB=build/variants/libndt2d_$v.so python bench.py --workload sweep --no-cpu-baseline > gpurun_out/x7_sw_$v.json 2>/dev/null
done
for v in base ld5 ld6 ld7 sw_b6 sw_b8 sw_p4 sw_p6; do python -c "
import json;d=json.load(open('gpurun_out/x7_sw_$v.json'));print('sweep $v',round(d['value']/1e6,1),'Mhyp/s',round(d['ms_per_step'],3),'ms')"; done
