#!/bin/bash
# BASELINE.json configs[1..4] at N GPUs of one box (bench.py's own JSON lines, one file per config under gpurun_out/).
#   usage (on the GPU box): [CONFIGS="c2 c3 ..."] bash scripts/scale_configs.sh N [tag]
# c2 weak-scales (65,536 envs per GPU), c3 / c4 are strong-scaled (total env count fixed, sharded over the ranks), c5 rows are the full
# collect + GAE + update path at 1M and 4M envs TOTAL (64x64 MLP) -- the update-step microbench of configs[4] through the sharded path.
N=${1:-1}; TAG=${2:-r2}
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
run() {  # name, then bench.py arguments
  local name=$1; shift
  if [ "$N" = 1 ]; then
    timeout 900 python bench.py --gpus 1 "$@" --no-cpu-baseline > gpurun_out/${TAG}_scale_${name}_n${N}.json 2> gpurun_out/${TAG}_scale_${name}_n${N}.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N "$@" --no-cpu-baseline \
      > gpurun_out/${TAG}_scale_${name}_n${N}.json 2> gpurun_out/${TAG}_scale_${name}_n${N}.err
  fi
  echo "$name n=$N rc=$? $(python - <<PY
import json
try:
    d=[json.loads(l) for l in open("gpurun_out/${TAG}_scale_${name}_n${N}.json") if l.startswith("{")][-1]
    print(f"{d['value']/1e6:.1f} M env-steps/s, {d['ms_per_step']:.2f} ms/step, clocks {d['clocks']['sm_mhz']} {d['clocks']['reasons']}")
except Exception as e: print("no line:", e)
PY
)"
}
CONFIGS=${CONFIGS:-"c2 c3 c4_acrobot c4_mcar c5_1m c5_4m"}
for c in $CONFIGS; do
  case $c in
    c2) run c2 --steps 10 --warmup 3 ;;
    c3) run c3 --config c3 --steps 20 --warmup 3 ;;
    c4_acrobot) run c4_acrobot --config c4_acrobot --steps 4 --warmup 3 ;;
    c4_mcar) run c4_mcar --config c4_mcar --steps 3 --warmup 3 ;;
    c5_1m) run c5_1m --n-envs $((1048576 / N)) --batch-size 1048576 --steps 4 --warmup 3 ;;
    c5_4m) run c5_4m --n-envs $((4194304 / N)) --batch-size 1048576 --steps 3 --warmup 3 ;;
  esac
done
