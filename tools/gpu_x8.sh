#!/bin/bash
# experiment pass x8: parity of the restructured LM loop, k_align variants (result hashes must agree), L1-hit probe
set -u
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q --timeout 900 > $O/x8_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/x8_pytest.log
python tools/exp.py gen > /dev/null 2>&1
python tools/exp.py run head base jr1 one jr1one esel0 b2 head base --steps 10 > $O/x8_variants.jsonl 2> $O/x8_variants.err
cat $O/x8_variants.jsonl
python bench.py --workload newton > $O/x8_newton_base.json 2> $O/x8_newton.err; echo "newton rc=$?"
NDT2D_LIB=build/variants/libndt2d_probe.so python bench.py --workload newton > $O/x8_newton_probe.json 2>> $O/x8_newton.err; echo "newton probe rc=$?"
python - <<'PY'
import json
for f in ("base", "probe"):
    try:
        d = json.loads(open(f"gpurun_out/x8_newton_{f}.json").read().strip().splitlines()[-1])
        print(f, d["value"], d["unit"], d["ms_per_step"])
    except Exception as e:
        print(f, "FAILED", e)
PY
