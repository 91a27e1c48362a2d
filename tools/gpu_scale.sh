#!/bin/bash
# Scaling pass (gpurun --gpus N): parity of the blocked loop with the peer-store exchange at N ranks, then the C4 bench.
set -u
N=${1:-8}
TAG=${2:-r02}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29511 tools/check_sharded_nccl.py --rows 1001 --seed 3 --exchange p2p --block 16 2>&1 | grep "sharded check" | tee gpurun_out/${TAG}_check_n$N.log
timeout 300 $TR --master-port 29512 tools/check_sharded_nccl.py --rows 2100 --seed 1 --exchange p2p 2>&1 | grep "sharded check" | tee -a gpurun_out/${TAG}_check_n$N.log
timeout 600 $TR --master-port 29513 bench.py --gpus $N --e2e-steps 2 > gpurun_out/${TAG}_bench_n${N}.json 2> gpurun_out/${TAG}_bench_n${N}.err
tail -2 gpurun_out/${TAG}_bench_n${N}.err | cut -c1-300; cat gpurun_out/${TAG}_bench_n${N}.json
