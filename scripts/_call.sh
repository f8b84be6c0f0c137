cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 120 python scripts/gpu_ab.py C3:64 base disk base disk 2>&1 | tee gpurun_out/r02c_disk_refill.log
