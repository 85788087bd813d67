# usage: bash tools/gpu_iter.sh [tag] -- GPU suite + short bench + CUPTI step trace (iteration loop)
cd $GRAFT_REPO_ROOT
TAG=${1:-iter}
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -q --timeout 600 -x 2>&1 | grep -E "^(E  |FAILED|ERROR|[0-9]+ (passed|failed)|worst)" | cut -c1-300 | head -40 > gpurun_out/${TAG}_pytest.log
tail -15 gpurun_out/${TAG}_pytest.log
timeout -s KILL 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${TAG}_bench.json').read())
    print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['frac'], d['roofline_mlp']['fused_trunk']['frac'], d['roofline_mlp']['families_ms_per_step'])
except Exception as e:
    print('bench failed', e); print(open('gpurun_out/${TAG}_bench.err').read()[-2000:])
PY
timeout -s KILL 300 python tools/step_trace.py --steps 6 --top 45 > gpurun_out/${TAG}_step_trace.txt 2>&1
sed -n 2,40p gpurun_out/${TAG}_step_trace.txt | cut -c1-150
