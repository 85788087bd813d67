# usage: bash tools/gpu_round1e.sh -- end-of-round evidence: GPU suite, bench line (with CPU baseline), reference arm,
# CUPTI step trace, ncu launch list, HBM-kernel bench, config-5 render bench, ray batcher bench, write ceilings
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
T=r1e
timeout -s KILL 900 python -m pytest tests -m gpu -q --timeout 600 2>&1 | grep -E "^(E  |FAILED|ERROR|[0-9]+ (passed|failed)|worst)" | cut -c1-300 | head -60 > gpurun_out/${T}_pytest.log
tail -5 gpurun_out/${T}_pytest.log
timeout -s KILL 600 python bench.py --steps 100 --warmup 5 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
cut -c1-300 gpurun_out/${T}_bench.json
timeout -s KILL 300 python tools/step_trace.py --steps 6 --top 70 > gpurun_out/${T}_step_trace.txt 2>&1
timeout -s KILL 300 python tools/bench_hbm.py > gpurun_out/${T}_hbm.jsonl 2> gpurun_out/${T}_hbm.err
timeout -s KILL 300 python tools/bench_render.py > gpurun_out/${T}_render.jsonl 2> gpurun_out/${T}_render.err
timeout -s KILL 300 python tools/bench_batcher.py > gpurun_out/${T}_batcher.jsonl 2> gpurun_out/${T}_batcher.err
timeout -s KILL 200 python tools/write_bw.py > gpurun_out/${T}_write_bw.txt 2>&1
timeout -s KILL 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_ncu_launch.log 2>&1
python tools/launch_summary.py gpurun_out/${T}_launches.csv 50 > gpurun_out/${T}_launches_summary.txt 2>&1
head -12 gpurun_out/${T}_step_trace.txt | cut -c1-160
cat gpurun_out/${T}_hbm.jsonl gpurun_out/${T}_render.jsonl gpurun_out/${T}_batcher.jsonl | cut -c1-260
