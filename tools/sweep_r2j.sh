#!/bin/bash
cd /root/repo
OUT=gpurun_out/r2_j.txt
: > $OUT
export SBP_B200_LIB=$PWD/summationbypartsoperatorsextra.jl_b200/libsbp_b200_ch12.so
run() { env "$@" timeout -s KILL 120 python tools/time_config.py --workload config3 --steps 30 2>gpurun_out/r2_j.err | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$*', d['ms_median'], round(d['hbm_gbs']))" >> $OUT; }
for mt in 4 5; do for sem in 0 2 1; do run SBP_BSR_MT=$mt SBP_BSR_SEM=$sem; done; done
run SBP_BSR_MT=4 SBP_BSR_SEM=2 SBP_BSR_DEPTH=1
run SBP_BSR_MT=5 SBP_BSR_SEM=2 SBP_BSR_PF=8
cat $OUT
