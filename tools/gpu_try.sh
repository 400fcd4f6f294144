for wl in config1b "config1b --n 104" "banded --n 104" "banded --n 101" "config3 --n 30" "config3 --n 44"; do
for v in "0 0" "1 1"; do set -- $v
SBP_BSR_V3=$1 SBP_BSR_HELP=$2 timeout -s KILL 30 python tools/time_config.py --workload $wl --steps 40 --tag "v3$1_help$2" 2>&1 | cut -c1-175
done; done
