#!/bin/bash
python bench.py --workload odometry > gpurun_out/r1j_odometry.json 2>/dev/null; cut -c1-300 gpurun_out/r1j_odometry.json
python bench.py --workload odometry --res 2.0 1.0 0.5 --perturb 0.1 1.0 > gpurun_out/r1j_odometry_pyramid.json 2>/dev/null; cut -c1-200 gpurun_out/r1j_odometry_pyramid.json
python -m pytest tests -x -q -m gpu 2>&1 | tail -2
