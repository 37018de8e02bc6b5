cd $GRAFT_REPO_ROOT
for v in 0 1 2; do
  cp _exp/cp$v.so gymnasium_solver_b200/csrc/libgs_engine.so
  GS_DEV_TRACK=0 python tests/dev_update_time.py --child 2>&1 | tail -1
done
for v in 0 1 2; do
  cp _exp/cp$v.so gymnasium_solver_b200/csrc/libgs_engine.so
  GS_DEV_TRACK=0 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:update_f16 -s 6 -c 1 python tests/dev_update_time.py --child 2>&1 | grep -E "dram__bytes|gpu__time" | sed "s/^/variant $v: /"
done
