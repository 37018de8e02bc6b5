cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
cp gymnasium_solver_b200/csrc/libgs_engine.so /tmp/orig.so
for v in u8 u16; do cp _exp/$v.so gymnasium_solver_b200/csrc/libgs_engine.so; echo "== $v"; timeout 600 python scripts/microbench_sweep.py --min-log2 12 2>&1 | cut -d'|' -f1-10; done > gpurun_out/w9_sweep.log 2>&1
cp /tmp/orig.so gymnasium_solver_b200/csrc/libgs_engine.so
timeout 600 python -m pytest tests/test_gpu_returns.py -x -q 2>&1 | tail -1
cat gpurun_out/w9_sweep.log
