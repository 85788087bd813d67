# round 2, second GPU pass: tnet + fixed tests, dual-tile trunk A/B, bench, trace
cd $GRAFT_REPO_ROOT
TAG=${1:-r2b}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tnet_gpu.py tests/test_train_step_gpu.py tests/test_eval_utils.py tests/test_tto_gpu.py tests/test_baseline_size_gpu.py tests/test_gemm_gpu.py -m gpu -q -s --timeout 600 2>&1 | grep -E "^(E  |FAILED|ERROR|[0-9]+ (passed|failed)|worst|\[|    nerf_|.*passed|.*failed)" | cut -c1-300 | head -80 > gpurun_out/${TAG}_pytest.log
tail -60 gpurun_out/${TAG}_pytest.log
echo "--- trunk A/B (single-tile multicast vs dual-tile cta_group::2 with copy-out warps)"
for v in 0 1; do echo "UPNERF_TRUNK_DUAL=$v"; UPNERF_TRUNK_DUAL=$v timeout 120 python tools/bench_gemm.py trunk 2>&1 | grep mlp_trunk; done | tee gpurun_out/${TAG}_trunk_ab.txt
echo "UPNERF_TRUNK_DUAL=1 UPNERF_TRUNK_LSU_STORE=0 (TMA stores)"; UPNERF_TRUNK_DUAL=1 UPNERF_TRUNK_LSU_STORE=0 timeout 120 python tools/bench_gemm.py trunk 2>&1 | grep mlp_trunk_fwd | tee -a gpurun_out/${TAG}_trunk_ab.txt
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_bench.json").read())
    print(d["value"], d["ms_per_step"], d["e2e"]["ms_per_step_runs"], d["roofline"]["frac"], d["roofline_mlp"]["families_ms_per_step"])
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/${TAG}_bench.err").read()[-2500:])
PY
UPNERF_TRUNK_DUAL=1 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('DUAL=1 step', d['value'], d['ms_per_step'], d['roofline_mlp']['families_ms_per_step'])"
timeout 300 python bench.py --workload render --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_render.json 2> gpurun_out/${TAG}_render.err
cut -c1-1000 gpurun_out/${TAG}_render.json; tail -3 gpurun_out/${TAG}_render.err
timeout -s KILL 300 python tools/step_trace.py --steps 6 --top 30 > gpurun_out/${TAG}_step_trace.txt 2>&1
sed -n 2,34p gpurun_out/${TAG}_step_trace.txt | cut -c1-150
