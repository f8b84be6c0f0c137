cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 800 python scripts/gpu_ab.py C3:64,C2:256 base pad1 pad2 pad3 pad4 pad5 pad6 pad7 base 2>&1 | tee gpurun_out/r02c_code_pad_sweep.log
