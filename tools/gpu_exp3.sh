#!/bin/bash
set -u
mkdir -p gpurun_out
python tools/exp.py gen --scans 65536 2>&1 | tail -1
python tools/exp.py run base pipe2 pipe5x4 pipe6x4 pipe3x6 t128b6 t128b7 > gpurun_out/x3_variants.jsonl 2> gpurun_out/x3_variants.err
cat gpurun_out/x3_variants.jsonl; tail -3 gpurun_out/x3_variants.err
for v in pipe2 pipe5x4; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_align -c 1 -f -o gpurun_out/x3_prof_$v \
     python tools/exp.py run $v --scans 16384 --steps 1 --warmup 0 > gpurun_out/x3_ncu_$v.log 2>&1
  echo "ncu $v rc=$?"
done
