# usage: bash tools/gpu_n8.sh N TAG  (run with gpurun --gpus N): bench at N GPUs exactly as the driver launches it
cd $GRAFT_REPO_ROOT
N=${1:-8}; TAG=${2:-n8}
mkdir -p gpurun_out
timeout -s KILL 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_$N.json 2> gpurun_out/${TAG}_bench_$N.err
echo rc $?
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${TAG}_bench_$N.json').read().strip().splitlines()[-1])
    print($N, {k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d.get('comm'), d['clocks'])
except Exception as e:
    print('bench failed', e); print(open('gpurun_out/${TAG}_bench_$N.err').read()[-2500:])
PY
