#!/bin/bash
set -u
python tools/exp.py gen --scans 65536 2>&1 | tail -1
python tools/exp.py run base po base po 2>/dev/null | cut -c1-120
python bench.py --workload sweep --no-cpu-baseline 2>/dev/null | cut -c1-140
NDT2D_LIB=build/variants/libndt2d_po.so python bench.py --workload sweep --no-cpu-baseline 2>/dev/null | cut -c1-140
