cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "adaptive" > gpurun_out/r02c_pytest_adaptive.log 2>&1; tail -5 gpurun_out/r02c_pytest_adaptive.log
(echo "== k_render_adaptive"; timeout 200 python scripts/gpu_adaptive.py 0 0 C2 C3 C5 C1; echo "== pixel stream (RT_B200_NO_ADAPTIVE_POOL=1)"; RT_B200_NO_ADAPTIVE_POOL=1 timeout 200 python scripts/gpu_adaptive.py 0 0 C2 C3 C5 C1) > gpurun_out/r02c_adaptive_ab.log 2>&1
cat gpurun_out/r02c_adaptive_ab.log
