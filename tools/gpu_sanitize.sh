#!/bin/bash
# compute-sanitizer passes over small cases of every kernel family (memcheck everywhere, racecheck on the kernels
# that synchronise through shared memory or in-kernel flags).  Each step is bounded by its own timeout.
set -u
mkdir -p gpurun_out
S=/usr/local/cuda/bin/compute-sanitizer
run() { name=$1; shift; timeout 600 "$@" > gpurun_out/san_$name.log 2>&1; echo "$name rc=$? $(grep -c 'ERROR SUMMARY' gpurun_out/san_$name.log) $(grep 'ERROR SUMMARY\|passed\|failed' gpurun_out/san_$name.log | tail -2 | tr '\n' ' ')"; }
run mem_f64 $S --tool memcheck --error-exitcode 9 python -m pytest tests/test_f64_engine_gpu.py -x -q -k "textbook or zero_pivot or beale or find_pivot or max_pivots or is_canonical or (solve_bit and 50)"
run mem_exact $S --tool memcheck --error-exitcode 9 python -m pytest tests/test_exact_gpu.py -x -q -k "known_answer or special or (pivots_match and 6) or rational_rows or find_pivot"
run mem_batch $S --tool memcheck --error-exitcode 9 python -m pytest tests/test_batch_gpu.py -x -q -k "special or rule_switch"
run race_f64 $S --tool racecheck --error-exitcode 9 python -m pytest tests/test_f64_engine_gpu.py -x -q -k "beale or max_pivots"
run race_batch $S --tool racecheck --error-exitcode 9 python -m pytest tests/test_batch_gpu.py -x -q -k "special"
run sync_batch $S --tool synccheck --error-exitcode 9 python -m pytest tests/test_batch_gpu.py -x -q -k "special"
