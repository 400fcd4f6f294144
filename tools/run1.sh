P=$PWD/summationbypartsoperatorsextra.jl_b200
for v in solo1a tracea; do echo == $v; SBP_B200_LIB=$P/libsbp_b200_$v.so python tools/time_config.py --workload config3 --steps 1 2>&1 | grep "warp [0-7] \|ms_median" | sort | uniq | awk "NR%8==1" | cut -c1-200 | head -4; done
