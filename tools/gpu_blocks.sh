#!/bin/bash
# Blocked fp64 loop: GPU test suite, block-size sweep at C4 / C2 / the 8-GPU shard shape, ncu capture of the fused update.
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/r02_pytest_gpu.log
for K in 8 12 16; do
  LPS_F64_BLOCK=$K python tools/bench_update.py --m 32768 --pivots 256 --warmup 16 2>&1 | tail -1
done | tee gpurun_out/r02_sweep_c4.txt
for K in 0 4 8 16; do
  LPS_F64_BLOCK=$K python tools/bench_update.py --m 2000 --pivots 2000 --warmup 64 --no-prof 2>&1 | tail -1
done | tee gpurun_out/r02_sweep_c2.txt
for K in 16; do
  LPS_F64_BLOCK=$K python tools/bench_update.py --m 4096 --n 65536 --pivots 256 --warmup 16 2>&1 | tail -1
done | tee gpurun_out/r02_sweep_shard8.txt
if [ "${1:-}" = "ncu" ]; then
  CMD="python tools/bench_update.py --m 32768 --pivots 64 --warmup 16"
  ncu --set full --clock-control none --import-source on -k regex:k_blk_update -s 1 -c 2 -o gpurun_out/r02_prof_blk_update $CMD > gpurun_out/ncu_full.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:k_blk_panel -s 1 -c 1 -o gpurun_out/r02_prof_blk_panel $CMD > gpurun_out/ncu_full2.log 2>&1
  ls -la gpurun_out
fi
