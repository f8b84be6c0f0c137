cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python scripts/gpu_ab.py C4:8,C4:16:1920 old base > gpurun_out/r02c_trav_fetch_after_shade.log 2>&1
cat gpurun_out/r02c_trav_fetch_after_shade.log
for m in 4 8 12 16; do echo "trav_min $m"; RT_B200_TRAV_MIN=$m timeout 100 python scripts/prof_render.py C4 8 2; done 2>&1 | tee -a gpurun_out/r02c_trav_fetch_after_shade.log
