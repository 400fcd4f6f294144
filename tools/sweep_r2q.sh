#!/bin/bash
# tile size / ring depth of k_bsr3 with the final kernel (config 3, jittered cloud)
cd /root/repo
OUT=gpurun_out/r2_q.txt
: > $OUT
for mt in 5 6 7; do for d in 1 2 3; do
  for w in config3 config3apply; do
    SBP_BSR_MT=$mt SBP_BSR_DEPTH=$d SBP_BSR_VERBOSE=1 timeout -s KILL 60 python tools/time_config.py --workload $w --steps 30 2>gpurun_out/r2_q.err | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('mt=$mt d=$d', d['workload'], d['ms_median'], round(d['hbm_gbs']))" >> $OUT
    grep -m1 "k_bsr3:" gpurun_out/r2_q.err | sed 's/.*MT=/   MT=/' | cut -c1-60 >> $OUT
  done
done; done
cat $OUT
