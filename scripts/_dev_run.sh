cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_update.py -x -q 2>&1 | tail -5
GS_DEV_TRACK=0 python tests/dev_update_time.py --child 2>&1 | tail -1
cp _exp/trace.so gymnasium_solver_b200/csrc/libgs_engine.so
python tests/dev_f16_trace.py 2>&1 | tail -11
