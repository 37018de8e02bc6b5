cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/w5_gpu_tests.log 2>&1; tail -5 gpurun_out/w5_gpu_tests.log
timeout 600 python bench.py --config c4_mcar --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/w5_bench_c4_mcar.json 2> gpurun_out/w5_bench_c4_mcar.err; tail -c 1500 gpurun_out/w5_bench_c4_mcar.json; tail -3 gpurun_out/w5_bench_c4_mcar.err
