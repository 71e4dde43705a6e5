for h in 1000000 500000 250000 125000; do python bench.py --workload sweep --hyps $h --no-cpu-baseline --steps 20 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print($h, round(d['ms_per_step'],4), round(d['value']/1e6,1))"; done
