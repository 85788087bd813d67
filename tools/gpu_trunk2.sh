# usage: bash tools/gpu_trunk2.sh TAG -- trunk parity tests + kernel-variant timing A/B
cd $GRAFT_REPO_ROOT
T=${1:-trunk2}
mkdir -p gpurun_out
timeout -s KILL 240 python -m pytest tests/test_gemm_gpu.py -m gpu -q --timeout 100 -x -k "trunk" 2>&1 | tail -15 | cut -c1-400 > gpurun_out/${T}_pytest.log
tail -8 gpurun_out/${T}_pytest.log
timeout -s KILL 200 python tools/trunk_variants.py > gpurun_out/${T}_variants.txt 2>&1
tail -8 gpurun_out/${T}_variants.txt
