cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_shadow_rays.py -q -m gpu -s 2>&1 | tail -40
python scripts/prof_render.py C2 1024 2 aTolerance=0 lightSampling=shadowRays 2>&1 | tail -1
python scripts/prof_render.py C1 100 3 lightSampling=shadowRays 2>&1 | tail -1
