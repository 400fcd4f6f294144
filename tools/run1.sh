P=$PWD/summationbypartsoperatorsextra.jl_b200
for cfg in "2 1" "1 2"; do set -- $cfg
SBP_B200_LIB=$P/libsbp_b200_ping3.so SBP_BSR_V3=1 SBP_BSR_MT=7 SBP_BSR_OUT=$1 SBP_BSR_DEPTH=$2 timeout -s KILL 40 python tools/time_config.py --workload config3 --steps 40 --tag "ping3_out$1_d$2" 2>&1 | cut -c1-150
done
SBP_B200_LIB=$P/libsbp_b200_ping3.so timeout -s KILL 120 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "advection2d" 2>&1 | tail -2
