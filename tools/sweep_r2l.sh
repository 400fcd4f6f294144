#!/bin/bash
# what-if variants of k_bsr3 (tools/mkvar.sh), config 3; usage: sweep_r2l.sh tag1 tag2 ...
cd /root/repo
OUT=gpurun_out/r2_l.txt
: > $OUT
for t in "$@"; do
  L=$PWD/summationbypartsoperatorsextra.jl_b200/libsbp_b200_$t.so
  for w in config3 config3apply; do
    SBP_B200_LIB=$L timeout -s KILL 60 python tools/time_config.py --workload $w --steps 30 2>>gpurun_out/r2_l.err | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$t', d['workload'], d['N'], d['ms_median'], round(d['hbm_gbs']))" >> $OUT
  done
done
cat $OUT
