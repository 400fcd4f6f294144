set -x
timeout -s KILL 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r1k.log 2>&1; tail -3 gpurun_out/pytest_gpu_r1k.log
timeout -s KILL 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout -s KILL 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_r1k.log 2> gpurun_out/bench_ref_r1k.err; tail -c 600 gpurun_out/bench_ref_r1k.log
timeout -s KILL 600 python bench.py > gpurun_out/bench_r1k.log 2> gpurun_out/bench_r1k.err; tail -c 1500 gpurun_out/bench_r1k.log
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_bench_r1k.csv python bench.py --steps 20 --warmup 3 > gpurun_out/ncu_bench_r1k.log 2>&1; tail -2 gpurun_out/ncu_bench_r1k.log | cut -c1-200
