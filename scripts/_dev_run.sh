cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2_plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_launch.log 2>&1
python tests/dev_update_time.py --child > gpurun_out/r2_plain_update.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:update_f16 -s 6 -c 1 -o gpurun_out/r2_update_f16 -f python tests/dev_update_time.py --child > gpurun_out/r2_ncu_update.log 2>&1
GS_DEV_ITERS=1 python tests/dev_step_profile.py > gpurun_out/r2_plain_step.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"collect_f16|gae_kernel|update_finish|gather_offsets|rollout_pack" -s 40 -c 8 -o gpurun_out/r2_step_kernels -f python tests/dev_step_profile.py > gpurun_out/r2_ncu_step.log 2>&1
tail -2 gpurun_out/r2_ncu_launch.log gpurun_out/r2_ncu_update.log gpurun_out/r2_ncu_step.log
ls -la gpurun_out/*.ncu-rep gpurun_out/r2_launches.csv
