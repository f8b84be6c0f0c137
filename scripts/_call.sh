cd $GRAFT_REPO_ROOT
for v in noteam t2 t4 t8 t16; do
echo "== $v"
export RT_B200_LIB=$PWD/mcp_raytracer_b200/csrc/ab_$v.so
python scripts/prof_render.py C2 1024 2 aTolerance=0.05 2>&1 | tail -1
python scripts/prof_render.py C2 1024 2 aTolerance=0.05 width=512 2>&1 | tail -1
python scripts/prof_render.py C1 100 3 aTolerance=0.05 2>&1 | tail -1
python scripts/prof_render.py C3 64 2 aTolerance=0.05 2>&1 | tail -1
done
