#!/bin/bash
set -u
mkdir -p gpurun_out
python tools/exp.py gen --scans 65536 2>&1 | tail -1
python tools/exp.py run base b2 t128b5 t128b6 t128b7 pipe2 pipe5x4 u2 ld4 ld1 > gpurun_out/x5_variants.jsonl 2> gpurun_out/x5_variants.err; cat gpurun_out/x5_variants.jsonl
