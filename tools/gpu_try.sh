timeout -s KILL 200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "stage or ssprk or full_size" 2>&1 | tail -2
for wl in config3 config1b; do timeout -s KILL 60 python tools/time_stage.py --workload $wl; done
