#!/bin/bash
cd /root/repo
OUT=gpurun_out/r2_k.txt
: > $OUT
run() { timeout -s KILL 120 python tools/time_config.py "$@" 2>>gpurun_out/r2_k.err | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['workload'], d['N'], d['ms_median'], round(d['hbm_gbs']))" >> $OUT; }
run --workload config3 --steps 30
run --workload config3apply --steps 30
run --workload config1b --steps 30
run --workload config1b --n 104 --steps 30
run --workload banded --steps 30
run --workload banded --n 104 --steps 30
run --workload config3 --n 44 --steps 30
run --workload config3 --n 30 --steps 30
python tools/time_stage.py 2>/dev/null | cut -c1-330 >> $OUT
cat $OUT
