# usage: bash tools/gpu_quick.sh TAG "pytest args" [bench env...] -- a subset of GPU tests, then a short bench
cd $GRAFT_REPO_ROOT
TAG=$1; shift
PYT=$1; shift
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest $PYT -m gpu -q --timeout 600 -x 2>&1 | tail -40 | cut -c1-400 > gpurun_out/${TAG}_pytest.log
tail -25 gpurun_out/${TAG}_pytest.log
env "$@" timeout -s KILL 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${TAG}_bench.json').read())
    print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['e2e']['ms_per_step_runs'], d['roofline']['frac'], d['roofline_mlp']['families_ms_per_step'])
except Exception as e:
    print('bench failed', e); print(open('gpurun_out/${TAG}_bench.err').read()[-3000:])
PY
