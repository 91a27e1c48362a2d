#!/bin/bash
# One GPU-box pass: parity tests, smoke, bench (both arms), C5 batch bench; with "ncu": launch lists and full
# captures of the dominant kernels.  Run under gpurun from the repo root; everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/host.txt; grep MemTotal /proc/meminfo >> gpurun_out/host.txt
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke.log
python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
tail -3 gpurun_out/bench_n1.err
cat gpurun_out/bench_ref.json gpurun_out/bench_n1.json
timeout 600 python tools/bench_batch.py --cluster 0 --reps 5 --cpu-seconds 10 > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err
cat gpurun_out/bench_c5.json | cut -c1-1200
if [ "${1:-}" = "ncu" ]; then
  CMD="python bench.py --steps 2 --warmup 3 --pivots-per-step 4 --e2e-steps 1 --no-cpu-baseline"
  $CMD > gpurun_out/plain.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
  CMD2="python tools/bench_update.py --m 32768 --pivots 6"
  $CMD2 > gpurun_out/plain2.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:k_update_f64 -s 4 -c 2 -o gpurun_out/prof_update $CMD2 > gpurun_out/ncu_full.log 2>&1
  ls -la gpurun_out
fi
