cd $GRAFT_REPO_ROOT
for f in 1 32; do cp _exp/flush$f.so gymnasium_solver_b200/csrc/libgs_engine.so; GS_DEV_SAVE=gpurun_out/g_flush$f.npy python tests/dev_tc_fullsize_error.py 2>&1 | grep "impl 0"; GS_DEV_TRACK=0 python tests/dev_update_time.py --child 2>&1 | tail -1; done
python - <<'PY'
import numpy as np
a=np.load("gpurun_out/g_flush1.npy"); b=np.load("gpurun_out/g_flush32.npy")
print("flush 32 vs flush 1: L2 rel diff", np.linalg.norm(a-b)/np.linalg.norm(a), "max abs", np.abs(a-b).max(), "max|g|", np.abs(a).max())
PY
