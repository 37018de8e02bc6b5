cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_update.py -x -q > gpurun_out/w7_update_tests.log 2>&1; tail -2 gpurun_out/w7_update_tests.log
timeout 600 python tests/dev_wide_check.py 19072 1048576 > gpurun_out/w2_check.log 2>&1; grep "rel L2" gpurun_out/w2_check.log
GS_DEV_HIDDEN=256 timeout 300 python tests/dev_update_time.py --child > gpurun_out/w7_plain_update256.log 2>&1 && tail -1 gpurun_out/w7_plain_update256.log &&
GS_DEV_HIDDEN=256 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"update_wide|wgrad_wide|stage_w2" -s 15 -c 3 -o gpurun_out/r2_update_wide -f python tests/dev_update_time.py --child > gpurun_out/w7_ncu_update256.log 2>&1
tail -2 gpurun_out/w7_ncu_update256.log
