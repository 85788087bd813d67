# usage: bash tools/gpu_round1d.sh -- GPU suite + bench line + CUPTI step trace + ncu launch list + HBM kernel bench
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -q --timeout 600 2>&1 | grep -E "^(E  |FAILED|ERROR|[0-9]+ (passed|failed)|worst)" | cut -c1-300 | head -60 > gpurun_out/r1d_pytest.log
tail -5 gpurun_out/r1d_pytest.log
timeout -s KILL 600 python bench.py --steps 50 --warmup 5 > gpurun_out/r1d_bench.json 2> gpurun_out/r1d_bench.err
cut -c1-400 gpurun_out/r1d_bench.json
timeout -s KILL 300 python tools/step_trace.py --steps 6 --top 70 > gpurun_out/r1d_step_trace.txt 2>&1
timeout -s KILL 300 python tools/bench_hbm.py > gpurun_out/r1d_hbm.jsonl 2> gpurun_out/r1d_hbm.err
timeout -s KILL 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r1d_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r1d_ncu_launch.log 2>&1
python tools/launch_summary.py gpurun_out/r1d_launches.csv 50 > gpurun_out/r1d_launches_summary.txt 2>&1
head -30 gpurun_out/r1d_step_trace.txt | cut -c1-160
cat gpurun_out/r1d_hbm.jsonl | cut -c1-300
