#!/bin/bash
cd /root/repo
OUT=gpurun_out/r2_i.txt
: > $OUT
export SBP_B200_LIB=$PWD/summationbypartsoperatorsextra.jl_b200/libsbp_b200_ch12.so
run() { env "$@" SBP_BSR_VERBOSE=1 timeout -s KILL 120 python tools/time_config.py --workload config3 --steps 30 2>gpurun_out/r2_i.err | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$*', d['ms_median'], round(d['hbm_gbs']))" >> $OUT; grep -m1 "k_bsr3" gpurun_out/r2_i.err | cut -c1-140 >> $OUT; }
run SBP_BSR_MT=0
run SBP_BSR_MT=4
run SBP_BSR_MT=5
run SBP_BSR_MT=6
run SBP_BSR_MT=5 SBP_BSR_PF=0
run SBP_BSR_MT=5 SBP_BSR_PF=8
timeout 300 python -m pytest tests -m gpu -x -q -k "rhs_advection or apply_sparse or stage_update or deterministic" 2>&1 | tail -2 >> $OUT
cat $OUT
