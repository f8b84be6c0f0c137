cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python scripts/gpu_ab.py C3:64,C1:16 base sahlv0 sahlv0p3 base sahlv0 2>&1 | tee gpurun_out/r02c_pool_sah_layout.log
