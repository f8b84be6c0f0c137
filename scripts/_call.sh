cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_render_trav -c 1 -o gpurun_out/r02d_c4_trav -f python scripts/prof_render.py C4 8 1 > gpurun_out/ncu_trav.log 2>&1
tail -1 gpurun_out/ncu_trav.log
