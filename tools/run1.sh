timeout -s KILL 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
for wl in config3 config1b; do timeout -s KILL 100 python tools/time_stage.py --workload $wl; done
