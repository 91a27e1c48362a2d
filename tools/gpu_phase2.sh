#!/bin/bash
# GPU pass for the device-side phase switch: the parity tests that go through Simplex's constructor, then the
# equality-form time-to-optimal cases.
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_exact_gpu.py tests/test_api_fp64_gpu.py tests/test_api_golden.py tests/test_linprog_gpu.py -x -q -m gpu 2>&1 | tail -40 | tee gpurun_out/pytest_phase2.log
timeout 600 python tools/bench_tto.py --only-phase1 2>&1 | tail -8 | tee gpurun_out/tto_phase1.jsonl
