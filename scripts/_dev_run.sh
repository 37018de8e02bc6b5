cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/w11_gpu_tests.log 2>&1; tail -3 gpurun_out/w11_gpu_tests.log
for c in c4_mcar c4_acrobot; do
timeout 600 python bench.py --config $c --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/w11_bench_$c.json 2> gpurun_out/w11_bench_$c.err; python - <<PY
import json
d=[json.loads(l) for l in open("gpurun_out/w11_bench_$c.json") if l.startswith("{")][-1]
print("$c", d["value"]/1e6, d["ms_per_step"], "e2e", d["e2e"]["value"]/1e6, d["clocks"])
PY
done
