#!/bin/bash
# multi-GPU pass: the default bench line (every leg) under torchrun on N ranks, plus the exchange check tool
set -u
N=${1:-2}; TAG=${2:-r2n}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/${TAG}_n${N}_topo.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/${TAG}_n${N}_bench.json 2> gpurun_out/${TAG}_n${N}_bench.err; echo "bench rc=$?"; tail -5 gpurun_out/${TAG}_n${N}_bench.err
python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/${TAG}_n${N}_bench.json').read().strip().splitlines()[-1])
    print('N=$N value', round(d['value']/1e6,2), 'e2e', round(d['e2e']['value']/1e6,2), d['e2e']['per_rank'], d['e2e']['numa'])
    print('relay', d['e2e'].get('relay'), 'without', d['e2e'].get('without_relay'), 'by_input', {k: round(v['value']/1e6,2) for k, v in d['e2e']['by_input'].items()})
    for k in ('prior2','pyramid'):
        print(k, round(d[k]['value']/1e6,3), d[k].get('mean_iterations'))
    s = d['sweep']; print('sweep', round(s['value']/1e6,1), s['ms_per_query'], s['combine_equals_host_api_result'], s.get('nccl'), s['exchange_check'], s['relocalize'], s['ok'])
except Exception as e:
    print('bench line FAILED', e)
PY
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/${TAG}_n${N}_pytest.log 2>&1; echo "pytest multi rc=$?"; tail -3 gpurun_out/${TAG}_n${N}_pytest.log
