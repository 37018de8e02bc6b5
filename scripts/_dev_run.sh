cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2g_gpu_tests.log 2>&1; tail -3 gpurun_out/r2g_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 2>/dev/null | cut -c1-400
GS_DEV_ITERS=1 timeout 600 ncu --set full --clock-control none -k regex:"collect_f16|gae_kernel" -c 2 -o gpurun_out/r2g_collect_gae -f python tests/dev_step_profile.py > gpurun_out/r2g_ncu_collect.log 2>&1; tail -1 gpurun_out/r2g_ncu_collect.log
