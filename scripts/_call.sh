cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python scripts/gpu_ab.py C4:8,C3:64 base trial base trial 2>&1 | tee gpurun_out/r02c_trial_refill.log
