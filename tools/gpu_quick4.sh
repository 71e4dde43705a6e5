#!/bin/bash
python bench.py --workload newton 2>/dev/null | cut -c1-170
NDT2D_LIB=build/variants/libndt2d_ld0.so python bench.py --workload newton 2>/dev/null | cut -c1-170
