# Scratch command list for one `gpurun -- 'bash scripts/_dev_run.sh'` call during development (edited per experiment; the GPU box has no shell).
# The commands the round's evidence came from are kept in profiles/*.md; the multi-GPU ones are scripts/scale_configs.sh.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/dev_gpu_tests.log 2>&1; tail -3 gpurun_out/dev_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/dev_bench_c2.json 2> gpurun_out/dev_bench_c2.err; tail -c 600 gpurun_out/dev_bench_c2.json
