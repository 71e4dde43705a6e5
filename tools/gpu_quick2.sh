#!/bin/bash
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/q2_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/q2_pytest.log
python tools/exp.py gen --scans 65536 2>&1 | tail -1
{ python tools/exp.py run base nojr base nojr; python tools/exp.py run base nojr --overlap 1 --scans 32768; python tools/exp.py run base nojr --res 2.0 1.0 0.5 --scans 32768; } 2>/dev/null | cut -c1-200
