cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python tests/dev_wide_check.py 128 18944 19072 37888 100000 1048576 > gpurun_out/w2_check.log 2>&1; tail -80 gpurun_out/w2_check.log
