cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/w14_gpu_tests.log 2>&1; tail -2 gpurun_out/w14_gpu_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/w14_bench_c2.json 2> gpurun_out/w14_bench_c2.err; python - <<'PY'
import json
d=[json.loads(l) for l in open("gpurun_out/w14_bench_c2.json") if l.startswith("{")][-1]
print("c2", d["value"]/1e6, d["ms_per_step"], "e2e", d["e2e"]["value"]/1e6, "ep_rew", d["config"]["last_ep_rew_mean"])
PY
