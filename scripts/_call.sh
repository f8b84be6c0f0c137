set -x
cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
nproc
python -m pytest tests -m gpu -x -q --ignore=tests/test_gpu_full_parity.py > gpurun_out/r02_pytest_1.log 2>&1; tail -15 gpurun_out/r02_pytest_1.log
python scripts/gpu_ab.py C2:256 base r1 lean rec lr t128b5 t128b4 t128b6 lr_t128b5 lr_t128b4 > gpurun_out/r02_ab_c2_1.log 2>&1; cat gpurun_out/r02_ab_c2_1.log
python scripts/gpu_ab.py C4:16,C3:64,C5:64 base r1 trav128b5 trav128b6 trav128b8 > gpurun_out/r02_ab_c4_1.log 2>&1; cat gpurun_out/r02_ab_c4_1.log
python scripts/gpu_full_parity.py r02a > gpurun_out/r02_full_parity_1.log 2>&1; tail -12 gpurun_out/r02_full_parity_1.log
