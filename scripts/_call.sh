cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 80 python scripts/gpu_adaptive.py 0 0 C3 C4 2>&1 | grep "aBatch" | tee gpurun_out/r02h_adaptive_c3_c4.log | cut -c1-200
