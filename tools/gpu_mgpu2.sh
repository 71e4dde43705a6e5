#!/bin/bash
set -u
N=${1:-2}
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/mgpu_exchange_check.py > gpurun_out/mg${N}_exchange.log 2>&1; echo "exchange rc=$?"; tail -5 gpurun_out/mg${N}_exchange.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --workload sweep --no-cpu-baseline > gpurun_out/mg${N}_sweep.json 2> gpurun_out/mg${N}_sweep.err; echo "sweep rc=$?"; head -c 330 gpurun_out/mg${N}_sweep.json; echo
