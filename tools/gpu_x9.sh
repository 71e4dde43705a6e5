#!/bin/bash
# experiment pass x9: parity, k_align with j-from-r at K = 1, fused pairs with shared-window sort buffers
set -u
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q --timeout 900 > $O/x9_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/x9_pytest.log
python tools/exp.py gen > /dev/null 2>&1
python tools/exp.py run head base head base --steps 10 > $O/x9_variants.jsonl 2> $O/x9_variants.err
cat $O/x9_variants.jsonl
python bench.py --workload odometry > $O/x9_odometry.json 2> $O/x9_odometry.err; echo "odometry rc=$?"
NDT2D_LIB=build/variants/libndt2d_head.so python bench.py --workload odometry > $O/x9_odometry_head.json 2>> $O/x9_odometry.err; echo "odometry head rc=$?"
python bench.py --workload sweep > $O/x9_sweep.json 2> $O/x9_sweep.err; echo "sweep rc=$?"
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/x9_bench.json 2> $O/x9_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
for f in ("odometry", "odometry_head", "sweep", "bench"):
    try:
        d = json.loads(open(f"gpurun_out/x9_{f}.json").read().strip().splitlines()[-1])
        print(f, d["value"], d["unit"], d["ms_per_step"], {k: (v.get("value") if isinstance(v, dict) else v) for k, v in d.items() if k in ("pyramid", "sweep", "odometry", "dense", "prior2")})
    except Exception as e:
        print(f, "FAILED", e)
PY
