# final evidence of a round: GPU tests, smoke, bench (native + reference arm), ncu launch list of the bench command, ncu --set full
# captures of the dominant kernels.  usage (on the GPU box): bash tools/run_final.sh r2z
TAG=${1:-latest}
set -x
timeout -s KILL 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_${TAG}.log 2>&1; tail -2 gpurun_out/pytest_gpu_${TAG}.log
timeout -s KILL 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout -s KILL 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref_${TAG}.log 2> gpurun_out/bench_ref_${TAG}.err; cut -c1-200 gpurun_out/bench_ref_${TAG}.log
timeout -s KILL 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_${TAG}.log 2> gpurun_out/bench_${TAG}.err; tail -c 300 gpurun_out/bench_${TAG}.log; tail -3 gpurun_out/bench_${TAG}.err
timeout -s KILL 600 python bench.py > gpurun_out/bench_default_${TAG}.log 2> gpurun_out/bench_default_${TAG}.err; tail -c 200 gpurun_out/bench_default_${TAG}.log
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_bench_${TAG}.csv python bench.py --steps 20 --warmup 3 > gpurun_out/ncu_bench_${TAG}.log 2>&1; tail -1 gpurun_out/ncu_bench_${TAG}.log | cut -c1-120
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:k_dense_small_sym -c 1 -s 2 -o gpurun_out/prof_config2_${TAG} -f python tools/prof_run.py --workload config2 --steps 4 > gpurun_out/ncu_c2_${TAG}.log 2>&1; tail -1 gpurun_out/ncu_c2_${TAG}.log
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:k_bsr -c 1 -s 2 -o gpurun_out/prof_config3_${TAG} -f python tools/prof_run.py --workload config3 --steps 4 > gpurun_out/ncu_c3_${TAG}.log 2>&1; tail -1 gpurun_out/ncu_c3_${TAG}.log
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:k_bsr -c 1 -s 2 -o gpurun_out/prof_config3stage_${TAG} -f python tools/prof_run.py --workload config3stage --steps 4 > gpurun_out/ncu_c3stage_${TAG}.log 2>&1; tail -1 gpurun_out/ncu_c3stage_${TAG}.log
