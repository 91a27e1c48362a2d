#!/bin/bash
# One GPU-box pass: parity tests, smoke, bench (both arms); with "ncu": launch list of the bench command and full
# captures of the two kernels of the blocked loop.  Run under gpurun from the repo root; everything lands in gpurun_out/.
set -u
TAG=${2:-r02}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/host.txt; grep MemTotal /proc/meminfo >> gpurun_out/host.txt
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/${TAG}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/${TAG}_smoke.log
python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err
python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err
tail -3 gpurun_out/${TAG}_bench_n1.err
cat gpurun_out/${TAG}_bench_ref.json gpurun_out/${TAG}_bench_n1.json
if [ "${1:-}" = "ncu" ]; then
  CMD="python bench.py --steps 2 --warmup 3 --pivots-per-step 64 --rank1-pivots 4 --e2e-steps 1 --no-cpu-baseline"
  $CMD > gpurun_out/plain.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_bench_c4.csv $CMD > gpurun_out/ncu_launches.log 2>&1
  CMD2="python tools/bench_update.py --m 32768 --pivots 64 --warmup 16"
  $CMD2 > gpurun_out/plain2.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:k_blk_update -s 1 -c 1 -o gpurun_out/${TAG}_prof_blk_update $CMD2 > gpurun_out/ncu_full.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:k_blk_panel -s 1 -c 1 -o gpurun_out/${TAG}_prof_blk_panel $CMD2 > gpurun_out/ncu_full2.log 2>&1
  ls -la gpurun_out | tail -8
fi
