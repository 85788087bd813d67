# usage: bash tools/gpu_final.sh TAG -- full GPU suite, smoke, the bench line (with CPU baseline) and the render line
cd $GRAFT_REPO_ROOT
T=${1:-final}
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -q --timeout 600 2>&1 | grep -E "^(E  |FAILED|ERROR|[0-9]+ (passed|failed)|worst)" | cut -c1-300 | head -40 > gpurun_out/${T}_pytest.log
tail -5 gpurun_out/${T}_pytest.log
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout -s KILL 600 python bench.py --steps 100 --warmup 5 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
cut -c1-260 gpurun_out/${T}_bench.json
timeout -s KILL 600 python bench.py --workload render --steps 3 --warmup 1 > gpurun_out/${T}_render.json 2> gpurun_out/${T}_render.err
cut -c1-200 gpurun_out/${T}_render.json
timeout -s KILL 300 python tools/step_trace.py --steps 6 --top 30 > gpurun_out/${T}_step_trace.txt 2>&1
sed -n 3,16p gpurun_out/${T}_step_trace.txt | cut -c1-150
