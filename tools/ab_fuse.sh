#!/bin/bash
# A/B of the fused small-tableau decision kernels (three launches per pivot) against the six-launch pivot.
for REP in 1 2 3; do
  for V in fused plain; do
    if [ $V = plain ]; then export LPS_NO_FUSE=1; else unset LPS_NO_FUSE; fi
    for CASE in "50 9 1000 3" "100 9 1000 3" "200 9 1000 3"; do
      set -- $CASE
      python tools/bench_exact.py --m $1 --hi $2 --pivots $3 --warmup $4 --no-prof 2>&1 | tail -1 | \
        python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$V', 'm', d['m'], 'pivots', d['pivots'], 'pivots/s', round(d['pivots_per_s']))"
    done
  done
done
