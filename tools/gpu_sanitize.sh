#!/bin/bash
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/san_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/san_pytest.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 3 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/san_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/san_memcheck.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 3 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/san_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -3 gpurun_out/san_racecheck.log
timeout 600 compute-sanitizer --tool synccheck --error-exitcode 3 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/san_synccheck.log 2>&1; echo "synccheck rc=$?"; tail -3 gpurun_out/san_synccheck.log
