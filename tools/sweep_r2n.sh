#!/bin/bash
# three consumer warps per scheduler without memory traffic (what-if 32): is the consumer loop faster at all?
cd /root/repo
OUT=gpurun_out/r2_n.txt
: > $OUT
run() { # tag lib env...
  t=$1; L=$PWD/summationbypartsoperatorsextra.jl_b200/libsbp_b200_$2.so; shift 2
  for w in config3 config3apply; do
    env "$@" SBP_BSR_VERBOSE=1 SBP_B200_LIB=$L timeout -s KILL 60 python tools/time_config.py --workload $w --steps 30 2>gpurun_out/r2_n.err | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$t', d['workload'], d['ms_median'], round(d['hbm_gbs']))" >> $OUT
    grep -m1 "k_bsr3:" gpurun_out/r2_n.err | cut -c1-160 >> $OUT
  done
}
for lib in f2 f3; do
run "$lib cw8 help" $lib SBP_BSR_CW=8 SBP_BSR_HELP=1
run "$lib cw8 nohelp" $lib SBP_BSR_CW=8 SBP_BSR_HELP=0
for mt in 5 6 7; do
run "$lib cw12 mt$mt" $lib SBP_BSR_CW=12 SBP_BSR_HELP=0 SBP_BSR_MT=$mt
done
done
cat $OUT
