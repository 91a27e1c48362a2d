#!/bin/bash
# Blocked fp64 loop: block-size sweep at C4 / C2 / the 8-GPU shard shape, panel shapes on small and medium tableaux
# (LPS_F64_BLOCK: pivots per block, 0 = rank-1 loop; LPS_F64_PANEL_CTAS: grid-wide panel shape; LPS_F64_PANEL_CLUSTER:
# 0 / 8 / 16 CTAs of the cluster panel).  Everything lands in gpurun_out/.
set -u
TAG=${1:-r02}
mkdir -p gpurun_out
for K in 0 8 12 16; do
  LPS_F64_BLOCK=$K python tools/bench_update.py --m 32768 --pivots 256 --warmup 16 2>&1 | tail -1
done | tee gpurun_out/${TAG}_sweep_c4.txt
LPS_F64_BLOCK=16 python tools/bench_update.py --m 4096 --n 65536 --pivots 256 --warmup 16 2>&1 | tail -1 | tee gpurun_out/${TAG}_sweep_shard8.txt
for M in 50 100 500 1000 2000 4000 7000; do
  for CFG in "LPS_F64_BLOCK=0" "LPS_F64_BLOCK=16 LPS_F64_PANEL_CLUSTER=0" "LPS_F64_BLOCK=16"; do
    env $CFG python tools/bench_update.py --m $M --pivots 1500 --warmup 32 --no-prof 2>&1 | tail -1 | cut -c1-160 | sed "s/^/$CFG  /"
  done
done | tee gpurun_out/${TAG}_panel_shapes.txt
