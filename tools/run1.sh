for n1 in 30 32 34 36 38 40 44; do
SBP_BSR_VERBOSE=1 timeout 120 python tools/time_config.py --workload config3 --n $n1 --steps 30 --tag "n1_$n1" 2>&1 | sort -u | grep -v "k_bsr<" | cut -c1-230
done
