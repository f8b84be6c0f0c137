cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python scripts/gpu_ab.py C4:8,C5:64 base pad1 pad2 pad3 pad4 pad5 pad6 pad7 base 2>&1 | tee gpurun_out/r02c_code_pad_sweep2.log
timeout 900 python scripts/gpu_ab.py C2:256,C3:64 base pad1 pad2 pad3 pad4 pad5 pad6 pad7 base -- aTolerance=0.05 2>&1 | tee -a gpurun_out/r02c_code_pad_sweep2.log
