cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python scripts/gpu_ab.py C2:256,C5:64 base lightvec lightvec2 base > gpurun_out/r02c_light_loads_ab.log 2>&1
cat gpurun_out/r02c_light_loads_ab.log
