cd $GRAFT_REPO_ROOT
python scripts/gpu_ab.py C2:1024,C3:64,C4:16,C5:64 base ph7 base ph7 2>&1 | tail -6
python scripts/gpu_ab.py C2:1024 base ph7 -- aTolerance=0.05 2>&1 | tail -3
