TAG=${1:-latest}
set -x
timeout -s KILL 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_${TAG}.log 2>&1; tail -2 gpurun_out/pytest_gpu_${TAG}.log
timeout -s KILL 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout -s KILL 600 python bench.py > gpurun_out/bench_${TAG}.log 2> gpurun_out/bench_${TAG}.err; tail -c 400 gpurun_out/bench_${TAG}.log; tail -3 gpurun_out/bench_${TAG}.err
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_bench_${TAG}.csv python bench.py --steps 20 --warmup 3 > gpurun_out/ncu_bench_${TAG}.log 2>&1; tail -1 gpurun_out/ncu_bench_${TAG}.log | cut -c1-120
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:k_bsr -c 1 -s 2 -o gpurun_out/prof_config3stage_${TAG} -f python tools/prof_run.py --workload config3stage --steps 4 > gpurun_out/ncu_c3stage_${TAG}.log 2>&1; tail -1 gpurun_out/ncu_c3stage_${TAG}.log
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:k_bsr -c 1 -s 2 -o gpurun_out/prof_config3_${TAG} -f python tools/prof_run.py --workload config3 --steps 4 > gpurun_out/ncu_c3_${TAG}.log 2>&1; tail -1 gpurun_out/ncu_c3_${TAG}.log
