cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python tests/dev_wide_check.py 19072 1048576 > gpurun_out/w2_check.log 2>&1; grep "rel L2" gpurun_out/w2_check.log
GS_DEV_PROFILE=1 GS_DEV_HIDDEN=256 timeout 300 python tests/dev_update_time.py --child > gpurun_out/w1_time256.log 2>&1; grep "gs_ppo_step\|wide" gpurun_out/w1_time256.log | cut -c1-70,150-230
timeout 600 python -m pytest tests/test_gpu_update.py -x -q 2>&1 | tail -1
