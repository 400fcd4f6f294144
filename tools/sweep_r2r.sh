#!/bin/bash
# box width 12 instead of 20 nodes (smaller ring -> MT = 8 fits for config 3?)
cd /root/repo
OUT=gpurun_out/r2_r.txt
: > $OUT
L=$PWD/summationbypartsoperatorsextra.jl_b200/libsbp_b200_bx12.so
SBP_B200_LIB=$L timeout -s KILL 300 python -m pytest tests -x -q -m gpu -k "bsr or sparse or rhs or stage or odd or advection or determin" 2>&1 | tail -2 >> $OUT
for w in "config3" "config3apply" "config1b" "config1b --n 104" "banded" "banded --n 104" "config3 --n 30" "config3 --n 44"; do
  SBP_BSR_VERBOSE=1 SBP_B200_LIB=$L timeout -s KILL 60 python tools/time_config.py --workload $w --steps 30 2>gpurun_out/r2_r.err | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('bx12', d['workload'], d['N'], d['ms_median'], round(d['hbm_gbs']))" >> $OUT
  grep -m1 "k_bsr" gpurun_out/r2_r.err | sed 's/.*MT=/   MT=/' | cut -c1-70 >> $OUT
done
for mt in 7 8; do
  SBP_BSR_MT=$mt SBP_BSR_VERBOSE=1 SBP_B200_LIB=$L timeout -s KILL 60 python tools/time_config.py --workload config3 --steps 30 2>gpurun_out/r2_r.err | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('bx12 mt=$mt', d['workload'], d['N'], d['ms_median'], round(d['hbm_gbs']))" >> $OUT
  grep -m1 "k_bsr" gpurun_out/r2_r.err | sed 's/.*MT=/   MT=/' | cut -c1-70 >> $OUT
done
SBP_B200_LIB=$L timeout -s KILL 100 python tools/time_stage.py 2>/dev/null | tail -1 | cut -c1-330 >> $OUT
cat $OUT
