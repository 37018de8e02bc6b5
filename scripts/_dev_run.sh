cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
cp gymnasium_solver_b200/csrc/libgs_engine.so /tmp/orig.so
for v in old new old new; do
  cp _exp/$v.so gymnasium_solver_b200/csrc/libgs_engine.so
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err
  python -c "
import json
d=[json.loads(l) for l in open('gpurun_out/ab_$v.json') if l.startswith('{')][-1]; print('$v', round(d['value']/1e6,1), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']/1e6,1))"
done
cp /tmp/orig.so gymnasium_solver_b200/csrc/libgs_engine.so
