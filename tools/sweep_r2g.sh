#!/bin/bash
cd /root/repo
OUT=gpurun_out/r2_g.txt
: > $OUT
run() { pf=$1; shift; SBP_BSR_PF=$pf timeout -s KILL 120 python tools/time_config.py "$@" 2>>gpurun_out/r2_g.err | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('pf$pf', d['workload'], d['N'], d['ms_median'], round(d['hbm_gbs']))" >> $OUT; }
for pf in 0 2 3 4 6; do
run $pf --workload config1b --steps 30
run $pf --workload config1b --n 104 --steps 30
run $pf --workload config3 --n 44 --steps 30
run $pf --workload config3 --n 30 --steps 30
done
run 3 --workload config3 --steps 30
run 5 --workload config3 --steps 30
cat $OUT
