for cfg in "5 2" "6 1" "6 2" "4 3" "5 3"; do set -- $cfg
  SBP_BSR_VERBOSE=1 SBP_BSR_MT=$1 SBP_BSR_DEPTH=$2 python tools/time_config.py --workload config3 --steps 30 --tag mt$1_d$2 2>&1 | sort -u | cut -c1-170 | grep -v "^k_bsr<"
done
