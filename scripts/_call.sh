cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(for q in 0 1; do
  if [ $q = 1 ]; then export RT_B200_NO_QNODES=1; echo "== FP32 wide nodes (RT_B200_NO_QNODES=1)"; else echo "== quantised nodes"; fi
  timeout 120 python scripts/prof_render.py C4 8 3
  timeout 120 python scripts/prof_render.py C4 8 3 integrator=wavefront
  timeout 120 python scripts/prof_render.py C4 8 3 aTolerance=0.05
done) > gpurun_out/r02c_qnodes_ab.log 2>&1
unset RT_B200_NO_QNODES
cat gpurun_out/r02c_qnodes_ab.log
timeout 300 python scripts/gpu_sah_vs_reference.py C4:3840:8 C4:960:32 2>&1 | tail -3
timeout 900 python -m pytest tests -x -q -m gpu -k "rain or deep or trav or C4 or degenerate or primary" 2>&1 | tail -3
