#!/bin/bash
for v in "" build/variants/libndt2d_pb2.so build/variants/libndt2d_pb1.so; do
NDT2D_LIB=$v python bench.py --workload odometry 2>&1 | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('$v', round(d['value']/1e6,3),'M/s', round(d['ms_per_step'],3),'ms')"
done
python bench.py --no-cpu-baseline 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('scan2map', round(d['value']/1e6,2))"
