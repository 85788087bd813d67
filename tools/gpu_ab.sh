# quick A/B of the fused trunk variants (isolated kernels) + their parity tests
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TAG=${1:-ab}
timeout 600 python -m pytest tests/test_gemm_gpu.py -m gpu -q -x --timeout 300 -k "trunk" 2>&1 | tail -3
for v in 0 1; do echo "UPNERF_TRUNK_DUAL=$v"; UPNERF_TRUNK_DUAL=$v timeout 120 python tools/bench_gemm.py trunk 2>&1 | grep mlp_trunk; done | tee gpurun_out/${TAG}_trunk_ab.txt
