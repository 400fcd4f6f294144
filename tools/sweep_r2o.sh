#!/bin/bash
# variants (tools/mkvar.sh) on config 3: RHS, plain apply and the stage kernels; usage: sweep_r2o.sh tag1 tag2 ...
cd /root/repo
OUT=gpurun_out/r2_o.txt
: > $OUT
for t in "$@"; do
  L=$PWD/summationbypartsoperatorsextra.jl_b200/libsbp_b200_$t.so
  for w in config3 config3apply; do
    SBP_B200_LIB=$L timeout -s KILL 60 python tools/time_config.py --workload $w --steps 30 2>>gpurun_out/r2_o.err | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$t', d['workload'], d['N'], d['ms_median'], round(d['hbm_gbs']))" >> $OUT
  done
  SBP_B200_LIB=$L timeout -s KILL 100 python tools/time_stage.py 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$t stage', d['rhs_ms'], d['stage_own_ms'], d['stage_extra_ms'], d['ssprk53_step_fused_ms'])" >> $OUT
done
cat $OUT
