#!/bin/bash
# Scaling pass (gpurun --gpus N): one parity check of the peer-store exchange, then the C4 bench at N (p2p, then nccl).
set -u
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29511 tools/check_sharded_nccl.py --rows 1000 --seed 3 --exchange p2p 2>&1 | grep "sharded check" | tee gpurun_out/check_p2p_n$N.log
for X in p2p nccl; do
  timeout 600 $TR --master-port 29513 bench.py --gpus $N --exchange $X --e2e-steps 2 > gpurun_out/bench_n${N}_$X.json 2> gpurun_out/bench_n${N}_$X.err
  tail -2 gpurun_out/bench_n${N}_$X.err | cut -c1-300; cat gpurun_out/bench_n${N}_$X.json
done
