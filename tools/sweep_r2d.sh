#!/bin/bash
# r2 experiment 4: with L2 prefetch on, do the warp organisations matter now?  12 consumer warps + chain semaphore, pipe token (PING)
cd /root/repo
OUT=gpurun_out/r2_pf2.txt
: > $OUT
run() { # lib pf mt depth sem out cw help
  SBP_B200_LIB=$PWD/summationbypartsoperatorsextra.jl_b200/$1 SBP_BSR_V3=2 SBP_BSR_PF=$2 SBP_BSR_MT=$3 SBP_BSR_DEPTH=$4 SBP_BSR_SEM=$5 SBP_BSR_OUT=$6 SBP_BSR_CW=$7 SBP_BSR_HELP=$8 \
    timeout -s KILL 120 python tools/time_config.py --workload config3 --steps 30 --tag "$1 pf$2 mt$3 d$4 sem$5 out$6 cw$7 help$8" 2>>gpurun_out/r2_pf2.err | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['tag'], d['ms_median'], round(d['hbm_gbs']))" >> $OUT
}
for mt in 6 7; do for d in 1 2; do for sem in 0 2; do for out in 1 2; do run libsbp_b200_f3.so 6 $mt $d $sem $out 12 0; done; done; done; done
for pf in 0 6; do run libsbp_b200_ping.so $pf 7 1 0 2 8 1; run libsbp_b200_ping.so $pf 6 2 0 2 8 1; run libsbp_b200_ping.so $pf 7 2 0 1 8 0; done
cat $OUT
