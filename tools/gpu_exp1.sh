#!/bin/bash
# round-1 experiment batch 1: parity, queue/occupancy/pipelining variants, ncu captures
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/x1_pytest.log 2>&1; echo "pytest rc=$?" 
tail -3 gpurun_out/x1_pytest.log
python tools/exp.py gen --scans 65536 2>&1 | tail -1
python tools/exp.py run base q1 q2 q2w64 b2 pipe2 u2 > gpurun_out/x1_variants.jsonl 2> gpurun_out/x1_variants.err
python tools/exp.py run base q2 --shuffle 1 >> gpurun_out/x1_variants.jsonl 2>> gpurun_out/x1_variants.err
cat gpurun_out/x1_variants.jsonl
for v in base q2 pipe2; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_align -c 1 -f -o gpurun_out/x1_prof_$v \
     python tools/exp.py run $v --scans 16384 --steps 1 --warmup 0 > gpurun_out/x1_ncu_$v.log 2>&1
  echo "ncu $v rc=$?"
done
