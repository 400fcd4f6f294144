#!/bin/bash
# r2 experiment 3: L2 prefetch distance of the box producer (k_bsr3 + helpers, config 3), MT / depth variants
cd /root/repo
OUT=gpurun_out/r2_pf.txt
: > $OUT
run() { # lib pf mt depth
  SBP_B200_LIB=$PWD/summationbypartsoperatorsextra.jl_b200/$1 SBP_BSR_PF=$2 SBP_BSR_MT=$3 SBP_BSR_DEPTH=$4 \
    timeout -s KILL 120 python tools/time_config.py --workload config3 --steps 30 --tag "$1 pf$2 mt$3 d$4" 2>>gpurun_out/r2_pf.err | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['tag'], d['ms_median'], round(d['hbm_gbs']))" >> $OUT
}
for pf in 0 2 4 6 8 12 16; do run libsbp_b200.so $pf 7 1; done
for pf in 4 8; do run libsbp_b200_hf.so $pf 7 1; done
for pf in 0 4 8; do run libsbp_b200.so $pf 6 2; run libsbp_b200.so $pf 6 1; run libsbp_b200.so $pf 5 2; done
cat $OUT
