# usage: bash tools/gpu_check.sh  -- full GPU test suite + bench line (run under gpurun)
cd $GRAFT_REPO_ROOT
timeout -s KILL 900 python -m pytest tests -m gpu -q --timeout 600 2>&1 | grep -E "^(E  |FAILED|ERROR|[0-9]+ (passed|failed)|worst)" | cut -c1-300 | head -40
timeout -s KILL 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_last.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_last.json').read())
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['frac'], d['roofline_mlp']['fused_trunk']['frac'], d['roofline_mlp']['families_ms_per_step'])
PY
