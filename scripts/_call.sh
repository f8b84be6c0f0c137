cd $GRAFT_REPO_ROOT
python scripts/gpu_ab.py C4:32 base 2>&1 | tail -1
RT_B200_TRAV2=1 python scripts/gpu_ab.py C4:32 base 2>&1 | tail -1
for m in 4 8 12 16 20; do echo -n "trav2 min=$m "; RT_B200_TRAV2=1 RT_B200_TRAV_MIN=$m python scripts/gpu_ab.py C4:32 base 2>&1 | tail -1; done
for b in 2 6 8; do echo -n "trav2 burst=$b "; RT_B200_TRAV2=1 RT_B200_TRAV_BURST=$b python scripts/gpu_ab.py C4:32 base 2>&1 | tail -1; done
RT_B200_TRAV2=1 python -m pytest tests -m gpu -q -x -k "deep_tree or C4 or rain or progressive" 2>&1 | tail -3
