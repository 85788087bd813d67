# ncu of the fused trunk kernels (forward with masks + backward) and the full GPU suite
cd $GRAFT_REPO_ROOT
TAG=${1:-r2c}
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mlp_trunk_fwd -s 14 -c 1 -o gpurun_out/${TAG}_trunk_fwd python tools/bench_gemm.py trunk > gpurun_out/${TAG}_ncu_fwd.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mlp_trunk_bwd -s 2 -c 1 -o gpurun_out/${TAG}_trunk_bwd python tools/bench_gemm.py trunk > gpurun_out/${TAG}_ncu_bwd.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
timeout -s KILL 1500 python -m pytest tests -m gpu -q --timeout 900 -s 2>&1 | grep -E "^(E  |FAILED|ERROR|[0-9]+ (passed|failed)|worst|\[|.*passed|.*failed)" | cut -c1-300 | head -100 > gpurun_out/${TAG}_pytest.log
tail -50 gpurun_out/${TAG}_pytest.log
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_bench.json").read())
    print(d["value"], d["ms_per_step"], d["e2e"]["ms_per_step_runs"], d["roofline"]["frac"], d["roofline_mlp"]["families_ms_per_step"])
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/${TAG}_bench.err").read()[-2500:])
PY
