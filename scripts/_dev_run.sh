cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2f_gpu_tests.log 2>&1; tail -2 gpurun_out/r2f_gpu_tests.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2f_bench_c2.json 2> gpurun_out/r2f_bench_c2.err; python - <<'PY'
import json
d=[json.loads(l) for l in open("gpurun_out/r2f_bench_c2.json") if l.startswith("{")][-1]
print("c2", d["value"]/1e6, d["ms_per_step"], "e2e", d["e2e"]["value"]/1e6, "cpu", d["cpu_baseline"], "upd", d["roofline"]["avg_launch_s"], "gae", d["roofline_gae"]["frac"], "collect", d["roofline_collect"]["avg_call_s"])
PY
timeout 300 python tests/dev_step_profile.py 2>&1 | cut -c1-72,150-230 | grep -E "gs::|Memset" > gpurun_out/r2f_step_profile.log; cat gpurun_out/r2f_step_profile.log
# ncu: launch list of one bench step, then full captures of the dominant kernels (each after its plain run above exited 0)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2f_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2f_ncu_launch.log 2>&1
GS_DEV_HIDDEN=256 timeout 300 python tests/dev_update_time.py --child > gpurun_out/r2f_plain_update256.log 2>&1 && tail -1 gpurun_out/r2f_plain_update256.log &&
GS_DEV_HIDDEN=256 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"update_wide|wgrad_wide" -s 10 -c 2 -o gpurun_out/r2f_update_wide -f python tests/dev_update_time.py --child > gpurun_out/r2f_ncu_update256.log 2>&1
timeout 300 python tests/dev_update_time.py --child > gpurun_out/r2f_plain_update64.log 2>&1 && tail -1 gpurun_out/r2f_plain_update64.log &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:update_f16 -s 6 -c 1 -o gpurun_out/r2f_update_f16 -f python tests/dev_update_time.py --child > gpurun_out/r2f_ncu_update64.log 2>&1
GS_DEV_ITERS=1 timeout 900 ncu --set full --clock-control none -k regex:"collect_f16|gae_kernel|update_finish_kernel|gather_offsets|rollout_pack" -s 4 -c 12 -o gpurun_out/r2f_step_kernels -f python tests/dev_step_profile.py > gpurun_out/r2f_ncu_step.log 2>&1
ls -la gpurun_out/r2f_*.ncu-rep gpurun_out/r2f_launches.csv
