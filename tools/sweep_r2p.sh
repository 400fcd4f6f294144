#!/bin/bash
# L2 prefetch distance with the greedy box producer
cd /root/repo
OUT=gpurun_out/r2_p.txt
: > $OUT
for pf in -1 0 2 4 6 8 12; do
  for w in "config3" "config3apply" "config1b" "config1b --n 104" "banded" "config3 --n 30" "config3 --n 44"; do
    SBP_BSR_PF=$pf timeout -s KILL 60 python tools/time_config.py --workload $w --steps 30 2>>gpurun_out/r2_p.err | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('pf=$pf', d['workload'], d['N'], d['ms_median'], round(d['hbm_gbs']))" >> $OUT
  done
done
cat $OUT
