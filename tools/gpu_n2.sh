# usage: bash tools/gpu_n2.sh TAG  (run with gpurun --gpus 2): 2-rank GPU tests, then bench at N=1 and N=2
cd $GRAFT_REPO_ROOT
TAG=${1:-n2}
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_ddp_gpu.py -m gpu -q -s --timeout 500 2>&1 | grep -E "^E  |passed|failed|relative|Error" | cut -c1-300 | tail -20
for N in 1 2; do
  if [ $N = 1 ]; then CMD="python bench.py"; else CMD="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py"; fi
  timeout -s KILL 600 $CMD --gpus $N --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_bench_$N.json 2> gpurun_out/${TAG}_bench_$N.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${TAG}_bench_$N.json').read().strip().splitlines()[-1])
    print($N, {k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d.get('comm'), d['config']['workload'][-40:])
except Exception as e:
    print('bench failed', e); print(open('gpurun_out/${TAG}_bench_$N.err').read()[-3000:])
PY
done
