cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python tests/dev_wide_trace.py > gpurun_out/w3_trace.log 2>&1; tail -4 gpurun_out/w3_trace.log
timeout 600 python tests/dev_wide_check.py 19072 1048576 > gpurun_out/w2_check.log 2>&1; grep "rel L2" gpurun_out/w2_check.log
