cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 300 python scripts/gpu_sah_vs_reference.py C3:1920:64 C4:3840:8 C1:400:16) > gpurun_out/r02c_sah_vs_ref.log 2>&1
cat gpurun_out/r02c_sah_vs_ref.log
timeout 600 python scripts/gpu_ab.py C3:64,C4:8,C1:16 old base > gpurun_out/r02c_node_ch_ab.log 2>&1
cat gpurun_out/r02c_node_ch_ab.log
timeout 600 python scripts/gpu_ab.py C4:8,C3:64 old base -- integrator=wavefront 2>&1
timeout 600 python scripts/gpu_ab.py C4:16:1920,C3:64 old base -- aTolerance=0.05 2>&1
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
