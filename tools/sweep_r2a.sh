#!/bin/bash
# r2 experiment 1: 12 consumer warps + chain semaphore (at most `sem` of the 3 warps of a scheduler multiply at a time)
cd /root/repo
OUT=gpurun_out/r2_sem.txt
: > $OUT
run() { # lib mt depth sem out cw help
  SBP_B200_LIB=$PWD/summationbypartsoperatorsextra.jl_b200/$1 SBP_BSR_V3=2 SBP_BSR_MT=$2 SBP_BSR_DEPTH=$3 SBP_BSR_SEM=$4 SBP_BSR_OUT=$5 SBP_BSR_CW=$6 SBP_BSR_HELP=$7 SBP_BSR_VERBOSE=1 \
    timeout -s KILL 120 python tools/time_config.py --workload config3 --steps 30 --tag "$1 mt$2 d$3 sem$4 out$5 cw$6 help$7" 2>gpurun_out/r2_sem.err | tail -1 >> $OUT
  grep -m1 "k_bsr" gpurun_out/r2_sem.err >> $OUT
}
run libsbp_b200.so 7 1 0 2 8 1      # shipped default
for lib in libsbp_b200_f3.so libsbp_b200_f4.so; do
 for mt in 5 6 7; do
  for d in 1 2 3; do
   for sem in 0 2 1; do
     run $lib $mt $d $sem 1 12 0
   done
  done
 done
done
for mt in 5 6; do for d in 2 3; do for sem in 0 2; do run libsbp_b200_f3.so $mt $d $sem 2 12 0; done; done; done
cat $OUT
