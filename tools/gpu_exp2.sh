#!/bin/bash
set -u
mkdir -p gpurun_out
python tools/exp.py gen --scans 65536 2>&1 | tail -1
python tools/exp.py run base async2 async2cg > gpurun_out/x2_variants.jsonl 2> gpurun_out/x2_variants.err
cat gpurun_out/x2_variants.jsonl; tail -3 gpurun_out/x2_variants.err
for v in async2; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_align -c 1 -f -o gpurun_out/x2_prof_$v \
     python tools/exp.py run $v --scans 16384 --steps 1 --warmup 0 > gpurun_out/x2_ncu_$v.log 2>&1
  echo "ncu $v rc=$?"
done
