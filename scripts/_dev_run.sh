cd $GRAFT_REPO_ROOT
python tests/dev_tc_fullsize_error.py > gpurun_out/r2_err_f16.log 2>&1
cat gpurun_out/r2_err_f16.log
python -m pytest tests/test_gpu_update.py -x -q 2>&1 | tail -15
