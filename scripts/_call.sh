cd $GRAFT_REPO_ROOT
for c in 4 8 12 18 32 64; do echo -n "chunks=$c "; RT_B200_CHUNKS=$c python scripts/prof_render.py C2 1024 4 partIndex=3 partCount=8 2>&1 | tail -1; done
for c in 4 8 18 32 64; do echo -n "POOL chunks=$c "; RT_B200_POOL_LIST=1 RT_B200_CHUNKS=$c python scripts/prof_render.py C2 1024 4 partIndex=3 partCount=8 2>&1 | tail -1; done
for c in 2 3 4; do echo -n "POOL whole chunks=$c "; RT_B200_POOL_LIST=1 RT_B200_CHUNKS=$c python scripts/prof_render.py C2 1024 4 2>&1 | tail -1; done
