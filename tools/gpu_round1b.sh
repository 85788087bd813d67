cd $GRAFT_REPO_ROOT
timeout -s KILL 900 python -m pytest tests -m gpu -q --timeout 600 2>&1 | grep -E "^(E  |FAILED|ERROR|[0-9]+ (passed|failed)|worst)" | cut -c1-300 | head -40
timeout -s KILL 300 python tools/bench_hbm.py 2>&1 | tee gpurun_out/r1_hbm_kernels.jsonl | cut -c1-250
timeout -s KILL 600 python tools/bench_render.py 2>&1 | tee gpurun_out/r1_render_config5.jsonl | cut -c1-400
timeout -s KILL 600 python bench.py --steps 50 --warmup 5 2>&1 | tail -1 > gpurun_out/bench_last.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_last.json').read())
print({k:d[k] for k in ('value','ms_per_step','clocks','cpu_baseline')}, d['e2e']['value'])
print(json.dumps(d['roofline']))
print(json.dumps(d['roofline_mlp']['fused_trunk']), d['roofline_mlp']['all_tcgen05']['frac'])
print(d['roofline_mlp']['families_ms_per_step'])
PY
