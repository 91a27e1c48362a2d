#!/bin/bash
# GPU pass for the batched node-LP kernel: parity tests first (bounded by timeout), then the C5 bench.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_batch_gpu.py -x -q 2>&1 | tail -40 | tee gpurun_out/pytest_batch.log
