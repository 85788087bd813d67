# usage: bash tools/gpu_round2e.sh [TAG] -- end-of-round evidence (1 GPU): GPU suite, bench line (with CPU baseline),
# render bench line, CUPTI step trace + one-step timeline, ncu launch list, ncu --set full of the dominant kernels,
# HBM-kernel bench, dense-primitive microbenches
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
T=${1:-r2e}
timeout -s KILL 900 python -m pytest tests -m gpu -q --timeout 600 2>&1 | grep -E "^(E  |FAILED|ERROR|[0-9]+ (passed|failed)|worst)" | cut -c1-300 | head -60 > gpurun_out/${T}_pytest.log
tail -5 gpurun_out/${T}_pytest.log
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout -s KILL 600 python bench.py --steps 100 --warmup 5 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
cut -c1-300 gpurun_out/${T}_bench.json
timeout -s KILL 600 python bench.py --workload render --steps 3 --warmup 1 > gpurun_out/${T}_render.json 2> gpurun_out/${T}_render.err
cut -c1-300 gpurun_out/${T}_render.json
timeout -s KILL 300 python tools/step_trace.py --steps 6 --top 70 > gpurun_out/${T}_step_trace.txt 2>&1
timeout -s KILL 300 python tools/step_timeline.py > gpurun_out/${T}_timeline.txt 2>/dev/null
timeout -s KILL 300 python tools/bench_hbm.py > gpurun_out/${T}_hbm.jsonl 2> gpurun_out/${T}_hbm.err
timeout -s KILL 300 python tools/bench_gemm.py > gpurun_out/${T}_gemm.txt 2>&1
timeout -s KILL 120 python tools/bench_gemm.py epi > gpurun_out/${T}_gemm_epi.txt 2>&1
timeout -s KILL 120 python tools/bench_gemm.py tf32 > gpurun_out/${T}_gemm_tf32.txt 2>&1
timeout -s KILL 200 python tools/trunk_variants.py > gpurun_out/${T}_trunk_variants.txt 2>&1
# launch list of the bench command (eager warm-up steps + graph replays; per-launch times are cold-cache and serialised)
timeout -s KILL 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_ncu_launch.log 2>&1
python tools/launch_summary.py gpurun_out/${T}_launches.csv 50 > gpurun_out/${T}_launches_summary.txt 2>&1
# full captures: second eager step -- 3 grouped weight-gradient launches (tnet, fine pass, coarse pass), trunk fwd/bwd x2
timeout -s KILL 400 ncu --set full --import-source on --clock-control none -k regex:"wgrad_tc_kernel" --launch-skip 3 --launch-count 3 -o gpurun_out/${T}_wgrad_tc -f python tools/one_step.py > /dev/null 2>&1
timeout -s KILL 400 ncu --set full --import-source on --clock-control none -k regex:"mlp_trunk" --launch-skip 4 --launch-count 4 -o gpurun_out/${T}_trunk -f python tools/one_step.py > /dev/null 2>&1
timeout -s KILL 400 ncu --set full --import-source on --clock-control none -k regex:"gemm_tc_kernel" --launch-skip 30 --launch-count 12 -o gpurun_out/${T}_gemm_tc -f python tools/one_step.py > /dev/null 2>&1
ls -la gpurun_out/${T}_*.ncu-rep
head -12 gpurun_out/${T}_step_trace.txt | cut -c1-160
cat gpurun_out/${T}_hbm.jsonl | cut -c1-260
