P=$PWD/summationbypartsoperatorsextra.jl_b200
python -m pytest tests -m gpu -x -q 2>&1 | tail -8
for wl in config3 config1b "banded --n 104" "banded --n 101" "config1b --n 104"; do SBP_BSR_VERBOSE=1 python tools/time_config.py --workload $wl --steps 30 --tag new 2>&1 | sort -u; done 2>&1 | cut -c1-220
