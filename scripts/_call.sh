set -x
cd $GRAFT_REPO_ROOT
python scripts/gpu_ab.py C2:256,C5:64,C1:1024 base > gpurun_out/r02_ab_c2_4.log 2>&1; cat gpurun_out/r02_ab_c2_4.log
python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_full_parity.py::test_converged_image_at_baseline_size > gpurun_out/r02_pytest_4.log 2>&1; tail -8 gpurun_out/r02_pytest_4.log
python scripts/gpu_adaptive.py 1024 1024 C2 2>&1 | tail -5
ncu --set full --import-source on --clock-control none -k regex:k_render_stream -c 1 -o gpurun_out/r02_c2_adaptive python scripts/prof_render.py C2 256 1 aTolerance=0.05 > gpurun_out/ncu_c2a.log 2>&1; tail -3 gpurun_out/ncu_c2a.log
