#!/bin/bash
# One-GPU pass: all GPU parity tests, smoke, both bench arms, the C2 / C3 / shard-shaped loops, C5 batch bench.
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
tail -3 gpurun_out/bench_n1.err; cat gpurun_out/bench_ref.json gpurun_out/bench_n1.json
python tools/bench_update.py --m 2000 --pivots 2000 --warmup 50 2>&1 | tail -1 | tee gpurun_out/loop_c2.json
python tools/bench_update.py --m 4096 --n 65536 --pivots 400 --warmup 20 2>&1 | tail -1 | tee gpurun_out/loop_shard8.json
python tools/bench_exact.py --m 1000 --hi 127 --pivots 40 2>&1 | tail -1 | tee gpurun_out/loop_c3.json
python tools/bench_exact.py --m 50 --hi 9 --pivots 1000 2>&1 | tail -1 | tee gpurun_out/loop_c1.json
if [ "${1:-}" = "ncu" ]; then
  CMD="python tools/bench_update.py --m 4096 --n 65536 --pivots 12 --warmup 4"
  ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_shard8.csv $CMD > gpurun_out/ncu_shard8.log 2>&1
  CMD="python tools/bench_update.py --m 2000 --pivots 12 --warmup 4"
  ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_c2.csv $CMD > gpurun_out/ncu_c2.log 2>&1
  CMD="python tools/bench_exact.py --m 1000 --hi 127 --pivots 4 --warmup 2"
  ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_c3.csv $CMD > gpurun_out/ncu_c3.log 2>&1
fi
