# usage: bash tools/gpu_trunk.sh TAG [ncu] -- fused-trunk iteration: trunk parity tests, trunk microbench, short bench,
# optionally an ncu --set full capture of the four trunk launches of one step
cd $GRAFT_REPO_ROOT
T=${1:-trunk}
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_gemm_gpu.py -m gpu -q --timeout 120 -x -k "trunk" 2>&1 | tail -15 | cut -c1-300 > gpurun_out/${T}_pytest.log
tail -5 gpurun_out/${T}_pytest.log
timeout -s KILL 200 python tools/bench_gemm.py trunk > gpurun_out/${T}_gemm.txt 2>&1
cat gpurun_out/${T}_gemm.txt | tail -5
timeout -s KILL 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${T}_bench.json').read())
    print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['frac'], d['roofline_mlp']['families_ms_per_step'])
except Exception as e:
    print('bench failed', e); print(open('gpurun_out/${T}_bench.err').read()[-2000:])
PY
if [ "$2" = "ncu" ]; then
timeout -s KILL 400 ncu --set full --import-source on --clock-control none -k regex:"mlp_trunk" --launch-skip 4 --launch-count 4 -o gpurun_out/${T}_trunk -f python tools/one_step.py > /dev/null 2>&1
ls -la gpurun_out/${T}_trunk.ncu-rep
fi
