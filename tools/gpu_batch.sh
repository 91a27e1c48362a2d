#!/bin/bash
# GPU pass for the batched node-LP kernel: parity tests first (bounded by timeout), then the C5 bench.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_batch_gpu.py -x -q 2>&1 | tail -40 | tee gpurun_out/pytest_batch.log
if grep -q "failed\|error" gpurun_out/pytest_batch.log; then exit 1; fi
timeout 600 python tools/bench_batch.py --cluster ${1:-0,4,8} --reps 3 --cpu-seconds 0 2>&1 | tee gpurun_out/bench_c5.log | python -c "
import sys, json
for line in sys.stdin:
    try: d = json.loads(line)
    except Exception: print(line.rstrip()[:300]); continue
    print({k: d[k] for k in ('cluster','primes','rows_per_cta','smem_bytes_per_cta','resident_clusters','ms','node_lps_per_s','pivots_per_s','optimal')}, d['e2e']['node_lps_per_s'])
"
