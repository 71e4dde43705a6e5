#!/bin/bash
set -u
N=${1:-8}; TAG=${2:-r1h}; STEPS=${3:-50}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29512 bench.py --gpus $N --workload sweep --steps $STEPS --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_n${N}_sweep_s$STEPS.json 2> gpurun_out/${TAG}_n${N}_sweep_s$STEPS.err; echo "sweep p2p rc=$?"
timeout 400 $TR --master-port 29513 bench.py --gpus $N --workload sweep --combine nccl --steps $STEPS --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_n${N}_sweep_nccl_s$STEPS.json 2> gpurun_out/${TAG}_n${N}_sweep_nccl_s$STEPS.err; echo "sweep nccl rc=$?"
for f in sweep_s$STEPS sweep_nccl_s$STEPS; do python -c "
import json;d=json.loads(open('gpurun_out/${TAG}_n${N}_$f.json').read().strip().splitlines()[-1]);print('$f N=$N',round(d['value']/1e6,2),d['unit'],round(d['ms_per_step'],4),'ms/step', d.get('combine_equals_host_api_result'), d['clocks'])"; done
