#!/bin/bash
cd /root/repo
OUT=gpurun_out/r2_h.txt
: > $OUT
run() { lib=$1; shift; SBP_B200_LIB=$PWD/summationbypartsoperatorsextra.jl_b200/$lib timeout -s KILL 120 python tools/time_config.py "$@" 2>>gpurun_out/r2_h.err | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$lib', d['workload'], d['N'], d['ms_median'], round(d['hbm_gbs']))" >> $OUT; }
for lib in libsbp_b200.so libsbp_b200_ping.so; do
run $lib --workload config3 --steps 30
run $lib --workload config3apply --steps 30
run $lib --workload config1b --steps 30
run $lib --workload config1b --n 104 --steps 30
run $lib --workload config3 --n 44 --steps 30
done
cat $OUT
