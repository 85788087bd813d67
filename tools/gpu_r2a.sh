# round 2, first GPU pass: full suite (new parity tests print their error tables), bench (with cpu + cuda-eager baselines), step trace
cd $GRAFT_REPO_ROOT
TAG=${1:-r2a}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader | head -2
nproc; free -g | head -2
timeout -s KILL 1500 python -m pytest tests -m gpu -q --timeout 900 -s 2>&1 | grep -E "^(E  |FAILED|ERROR|[0-9]+ (passed|failed)|worst|\[|    nerf_|.*passed|.*failed)" | cut -c1-400 | head -150 > gpurun_out/${TAG}_pytest.log
tail -80 gpurun_out/${TAG}_pytest.log
timeout -s KILL 900 python bench.py --steps 30 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${TAG}_bench.json').read())
    print({k:d[k] for k in ('value','ms_per_step')}, d['e2e'], d['roofline']['frac'], d['roofline_mlp']['fused_trunk']['frac'], d['roofline_mlp']['families_ms_per_step'])
    print(d.get('cpu_baseline')); print(d.get('cuda_eager_baseline'))
except Exception as e:
    print('bench failed', e); print(open('gpurun_out/${TAG}_bench.err').read()[-3000:])
PY
timeout -s KILL 600 python bench.py --workload render --steps 3 --warmup 3 > gpurun_out/${TAG}_render.json 2> gpurun_out/${TAG}_render.err
cut -c1-1500 gpurun_out/${TAG}_render.json; tail -5 gpurun_out/${TAG}_render.err
timeout -s KILL 300 python tools/step_trace.py --steps 6 --top 45 > gpurun_out/${TAG}_step_trace.txt 2>&1
sed -n 2,30p gpurun_out/${TAG}_step_trace.txt | cut -c1-150
