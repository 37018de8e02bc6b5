cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_returns.py tests/test_gpu_fullsize.py -x -q > gpurun_out/w8_returns_tests.log 2>&1; tail -2 gpurun_out/w8_returns_tests.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/w8_bench_c2.json 2> gpurun_out/w8_bench_c2.err; python - <<'PY'
import json
d=[json.loads(l) for l in open("gpurun_out/w8_bench_c2.json") if l.startswith("{")][-1]
print(d["value"]/1e6, d["ms_per_step"], d["roofline_gae"], d["roofline"]["avg_launch_s"], d["roofline_collect"]["avg_call_s"])
PY
