#!/bin/bash
# Multi-GPU pass (gpurun --gpus N): parity of the row-sharded solver (blocked loop with peer stores, rank-1 loop with
# peer stores, rank-1 loop with the NCCL all-gather), then the C4 bench at N.  Everything lands in gpurun_out/.
set -u
N=${1:-2}
TAG=${2:-r02}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
for ARGS in "--exchange p2p --block 16" "--exchange p2p --block 0" "--exchange nccl" "--exchange p2p --block 5"; do
  timeout 300 $TR --master-port 29511 tools/check_sharded_nccl.py --rows 512 $ARGS 2>&1 | grep "sharded check" | tee -a gpurun_out/${TAG}_check_n$N.log
  timeout 300 $TR --master-port 29512 tools/check_sharded_nccl.py --rows 1001 --seed 3 $ARGS 2>&1 | grep "sharded check" | tee -a gpurun_out/${TAG}_check_n$N.log
done
timeout 900 $TR --master-port 29513 bench.py --gpus $N --e2e-steps 2 > gpurun_out/${TAG}_bench_n${N}.json 2> gpurun_out/${TAG}_bench_n${N}.err
tail -2 gpurun_out/${TAG}_bench_n${N}.err | cut -c1-300; cat gpurun_out/${TAG}_bench_n${N}.json
if [ "${3:-}" = "legacy" ]; then
  timeout 900 $TR --master-port 29514 bench.py --gpus $N --block 0 --pivots-per-step 256 --e2e-steps 1 > gpurun_out/${TAG}_bench_n${N}_rank1.json 2> gpurun_out/${TAG}_bench_n${N}_rank1.err
  cat gpurun_out/${TAG}_bench_n${N}_rank1.json
fi
