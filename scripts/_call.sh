# last check of the committed tree: full GPU suite + smoke
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/r02h_pytest_gpu_full.log 2>&1; tail -3 gpurun_out/r02h_pytest_gpu_full.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
