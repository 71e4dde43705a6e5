#!/bin/bash
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/x6_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/x6_pytest.log
python tools/exp.py gen --scans 65536 2>&1 | tail -1
{ python tools/exp.py run base ld4 --overlap 1 --scans 32768; python tools/exp.py run base ld4 --res 2.0 1.0 0.5 --scans 32768; python tools/exp.py run base ld4 --shuffle 1; } > gpurun_out/x6_variants.jsonl 2> gpurun_out/x6_variants.err; cat gpurun_out/x6_variants.jsonl
python bench.py --workload sweep --no-cpu-baseline > gpurun_out/x6_sweep.json 2> gpurun_out/x6_sweep.err; head -c 300 gpurun_out/x6_sweep.json; echo
NDT2D_LIB=build/variants/libndt2d_ld4.so python bench.py --workload sweep --no-cpu-baseline 2>/dev/null | head -c 200; echo
