for wl in config3 config1b "config1b --n 104" "banded --n 104" "banded --n 101"; do
SBP_BSR_VERBOSE=1 SBP_BSR_V3=1 timeout -s KILL 40 python tools/time_config.py --workload $wl --steps 40 --tag "defer" 2>&1 | sort -u | grep -v "k_bsr<" | cut -c1-210
done
SBP_BSR_V3=1 SBP_BSR_MT=8 timeout -s KILL 40 python tools/time_config.py --workload config3 --steps 40 --tag "defer_mt8" 2>&1 | cut -c1-210
timeout -s KILL 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "sparse or advection or mfsbp or ssprk or odd or container" 2>&1 | tail -3
