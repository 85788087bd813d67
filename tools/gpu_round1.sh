set -x
cd $GRAFT_REPO_ROOT
timeout -s KILL 900 python -m pytest tests -m gpu -q --timeout 600 2>&1 | grep -E "^(E  |FAILED|ERROR|[0-9]+ (passed|failed)|worst)" | cut -c1-300 | head -60 > gpurun_out/r1_pytest.log
timeout -s KILL 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r1_bench.json 2> gpurun_out/r1_bench.err
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 60 -c 2 -o gpurun_out/r1_gemm_tc python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full1.log 2>&1
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:wgrad_tc_kernel -s 30 -c 2 -o gpurun_out/r1_wgrad_tc python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full2.log 2>&1
tail -3 gpurun_out/r1_pytest.log; cut -c1-600 gpurun_out/r1_bench.json; ls -la gpurun_out
