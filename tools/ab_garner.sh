#!/bin/bash
# A/B of the Garner kernels inside the exact pivot loop: warp-per-number (default) against block-per-8-numbers.
for REP in 1 2 3; do
  for V in warp block; do
    if [ $V = block ]; then export LPS_GARNER_BLOCK=1; else unset LPS_GARNER_BLOCK; fi
    for CASE in "1000 127 1300 40" "200 9 1000 3" "50 9 1000 3"; do
      set -- $CASE
      python tools/bench_exact.py --m $1 --hi $2 --pivots $3 --warmup $4 --no-prof 2>&1 | tail -1 | \
        python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$V', 'm', d['m'], 'pivots', d['pivots'], 'pivots/s', round(d['pivots_per_s']))"
    done
  done
done
