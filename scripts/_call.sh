cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python scripts/gpu_ab.py C2:512 base pad2 pad3 pad4 pad5 base 2>&1 | tee -a gpurun_out/r02c_list_kernel_layout.log
timeout 900 python scripts/gpu_ab.py C2:512,C5:64 base adlv0 base adlv0 -- aTolerance=0.05 2>&1 | tee -a gpurun_out/r02c_list_kernel_layout.log
