#!/bin/bash
# Launch-shape sweep of k_update_f64 (column chunk x rows per CTA) at a few sizes; prints update time per launch.
for M in ${SIZES:-500 1000 2000 4000}; do
  for C in 512 1024; do
    for R in 8 16 32 64; do
      LPS_F64_CHUNK=$C LPS_F64_ROWS=$R python tools/bench_update.py --m $M --pivots 300 --warmup 20 2>&1 | tail -1 | \
        python -c "import sys,json; d=json.loads(sys.stdin.read()); print('m',d['m'],'chunk',$C,'rows',$R,'update_us',round(1e3*d['update_ms_avg'],2),'GB/s',round(d['update_GBps']))"
    done
  done
done
