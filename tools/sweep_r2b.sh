#!/bin/bash
# r2 experiment 2: what-if wide output stores (EXP 11: two {32 x TI} boxes per item, 12: four {16 x TI}), no output (4), helper fence
cd /root/repo
OUT=gpurun_out/r2_wide.txt
: > $OUT
run() { # lib tag
  SBP_B200_LIB=$PWD/summationbypartsoperatorsextra.jl_b200/$1 \
    timeout -s KILL 120 python tools/time_config.py --workload config3 --steps 30 --tag "$1" 2>>gpurun_out/r2_wide.err | tail -1 | cut -c1-200 >> $OUT
}
for lib in libsbp_b200.so libsbp_b200_e11.so libsbp_b200_e12.so libsbp_b200_e4.so libsbp_b200_hf.so; do run $lib; done
cat $OUT
