#!/bin/bash
# Multi-GPU pass (gpurun --gpus N): parity of the row-sharded solver with both exchanges, then the C4 bench at N
# with the peer-store exchange and with the NCCL all-gather.  Everything lands in gpurun_out/.
set -u
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_f64_engine_gpu.py tests/test_api_fp64_gpu.py tests/test_sharded_gpu.py -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_f64_multi.log
for X in p2p nccl; do
  timeout 300 $TR --master-port 29511 tools/check_sharded_nccl.py --rows 512 --exchange $X 2>&1 | grep -v "^W\|^\*\*\*" | tail -4 | tee gpurun_out/check_${X}_n$N.log
  timeout 300 $TR --master-port 29512 tools/check_sharded_nccl.py --rows 1000 --seed 3 --exchange $X 2>&1 | grep -v "^W\|^\*\*\*" | tail -4 | tee -a gpurun_out/check_${X}_n$N.log
done
for X in p2p nccl; do
  timeout 900 $TR --master-port 29513 bench.py --gpus $N --exchange $X > gpurun_out/bench_n${N}_$X.json 2> gpurun_out/bench_n${N}_$X.err
  tail -2 gpurun_out/bench_n${N}_$X.err; cat gpurun_out/bench_n${N}_$X.json
done
