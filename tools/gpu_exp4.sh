#!/bin/bash
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/x4_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/x4_pytest.log
python tools/exp.py gen --scans 65536 2>&1 | tail -1
python tools/exp.py run base > gpurun_out/x4_variants.jsonl 2> gpurun_out/x4_variants.err; cat gpurun_out/x4_variants.jsonl
python bench.py --workload sweep --no-cpu-baseline > gpurun_out/x4_sw_base.json 2>/dev/null
for v in sw_b5 sw_b6 sw_b8 sw_p3 sw_p4 sw_p5 sw_p6; do
  NDT2D_LIB=build/variants/libndt2d_$v.so python bench.py --workload sweep --no-cpu-baseline > gpurun_out/x4_$v.json 2>/dev/null
done
for v in sw_base sw_b5 sw_b6 sw_b8 sw_p3 sw_p4 sw_p5 sw_p6; do python -c "
import json;d=json.load(open('gpurun_out/x4_$v.json'));print('$v',round(d['value']/1e6,1),'Mhyp/s',round(d['ms_per_step'],3),'ms')"; done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_align -c 1 -f -o gpurun_out/x4_prof_base \
     python tools/exp.py run base --scans 16384 --steps 1 --warmup 0 > gpurun_out/x4_ncu_base.log 2>&1
