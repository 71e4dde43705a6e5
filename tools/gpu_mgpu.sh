#!/bin/bash
# multi-GPU measurement pass: N ranks on one node (run under gpurun --gpus N)
set -u
N=${1:-2}; TAG=${2:-r1h}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29511 tools/mgpu_exchange_check.py > gpurun_out/${TAG}_n${N}_exchange.log 2>&1; echo "exchange rc=$?"; grep "exchange check" gpurun_out/${TAG}_n${N}_exchange.log
timeout 400 $TR --master-port 29512 bench.py --gpus $N --workload sweep --no-cpu-baseline > gpurun_out/${TAG}_n${N}_sweep.json 2> gpurun_out/${TAG}_n${N}_sweep.err; echo "sweep p2p rc=$?"
timeout 400 $TR --master-port 29513 bench.py --gpus $N --workload sweep --combine nccl --no-cpu-baseline > gpurun_out/${TAG}_n${N}_sweep_nccl.json 2> gpurun_out/${TAG}_n${N}_sweep_nccl.err; echo "sweep nccl rc=$?"
timeout 400 $TR --master-port 29514 bench.py --gpus $N --no-cpu-baseline > gpurun_out/${TAG}_n${N}_bench.json 2> gpurun_out/${TAG}_n${N}_bench.err; echo "scan2map rc=$?"
for f in sweep sweep_nccl bench; do python -c "
import json;d=json.loads(open('gpurun_out/${TAG}_n${N}_$f.json').read().strip().splitlines()[-1]);print('$f N=$N',round(d['value']/1e6,2),d['unit'],round(d['ms_per_step'],4),'ms/step e2e',round(d['e2e']['value']/1e6,2), d.get('combine_equals_host_api_result'))"; done
