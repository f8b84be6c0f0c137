cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(echo "== k_render_pool<SAH> (RT_B200_NO_TRAV=1)"; RT_B200_NO_TRAV=1 timeout 200 python scripts/gpu_trav_threshold.py 500 1000 2000 4000 8000 16000
echo "== k_render_trav (RT_B200_TRAV_ALWAYS=1)"; RT_B200_TRAV_ALWAYS=1 timeout 200 python scripts/gpu_trav_threshold.py 500 1000 2000 4000 8000 16000
echo "== C3 weekend-final @64"; timeout 100 python scripts/prof_render.py C3 64 3; RT_B200_TRAV_ALWAYS=1 timeout 100 python scripts/prof_render.py C3 64 3) > gpurun_out/r02c_trav_threshold.log 2>&1
cat gpurun_out/r02c_trav_threshold.log
