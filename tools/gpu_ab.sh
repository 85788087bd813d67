# usage: bash tools/gpu_ab.sh TAG "ENV1=.. ENV2=.." "ENV.." ... -- GPU suite once, then one short bench per environment setting
cd $GRAFT_REPO_ROOT
TAG=$1; shift
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -q --timeout 600 -x 2>&1 | grep -E "^(E  |FAILED|ERROR|[0-9]+ (passed|failed)|worst)" | cut -c1-300 | head -40 > gpurun_out/${TAG}_pytest.log
tail -15 gpurun_out/${TAG}_pytest.log
i=0
for ENVS in "$@"; do
  i=$((i+1))
  env $ENVS timeout -s KILL 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_bench_$i.json 2> gpurun_out/${TAG}_bench_$i.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${TAG}_bench_$i.json').read())
    print('$ENVS', {k:round(d[k],3) for k in ('value','ms_per_step')}, round(d['e2e']['value']), d['roofline_mlp']['families_ms_per_step'])
except Exception as e:
    print('$ENVS bench failed', e); print(open('gpurun_out/${TAG}_bench_$i.err').read()[-2000:])
PY
done
timeout -s KILL 300 python tools/step_trace.py --steps 6 --top 45 > gpurun_out/${TAG}_step_trace.txt 2>&1
sed -n 2,12p gpurun_out/${TAG}_step_trace.txt | cut -c1-150
