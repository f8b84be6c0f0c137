cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(for b in 3 4 5 6 8; do for m in 8 12; do echo -n "burst $b min $m: "; RT_B200_TRAV_BURST=$b RT_B200_TRAV_MIN=$m timeout 100 python scripts/prof_render.py C4 8 3; done; done) 2>&1 | tee gpurun_out/r02c_trav_burst_sweep.log
