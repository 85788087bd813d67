# usage: bash tools/gpu_ncu_trunk.sh TAG -- ncu --set full of the four trunk launches of one eager step
cd $GRAFT_REPO_ROOT
T=${1:-ncu}
mkdir -p gpurun_out
timeout -s KILL 400 ncu --set full --import-source on --clock-control none -k regex:"mlp_trunk" --launch-skip 4 --launch-count 4 -o gpurun_out/${T}_trunk -f python tools/one_step.py > /dev/null 2>&1
ls -la gpurun_out/${T}_trunk.ncu-rep
nvidia-smi -q -d POWER | grep -i -E "power limit|power draw|cap" | head -12
