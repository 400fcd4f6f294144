#!/bin/bash
# staggered chains (SBP_BSR_SEM = k: gate open after k/8 of a chain), variants s0 (full kernel) and s32 (no memory traffic)
cd /root/repo
OUT=gpurun_out/r2_m.txt
: > $OUT
for t in s32 s0; do
  L=$PWD/summationbypartsoperatorsextra.jl_b200/libsbp_b200_$t.so
  for k in 0 3 4 5 6 7 8; do
  for w in config3 config3apply; do
    SBP_BSR_SEM=$k SBP_B200_LIB=$L timeout -s KILL 60 python tools/time_config.py --workload $w --steps 30 2>>gpurun_out/r2_m.err | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$t sem=$k', d['workload'], d['N'], d['ms_median'], round(d['hbm_gbs']))" >> $OUT
  done
  done
done
cat $OUT
