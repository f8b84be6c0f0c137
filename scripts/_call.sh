cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python scripts/gpu_full_parity.py r02b > gpurun_out/r02b_full_parity.log 2>&1; tail -3 gpurun_out/r02b_full_parity.log | cut -c1-300; python - <<'PY'
import json
for r in json.load(open("gpurun_out/r02b_full_parity.json")):
    print(r.get("case"), r.get("bars_met"), r.get("frac_within_3sigma"), r.get("rel_rmse_raw"), r.get("rel_rmse_noise_floor"), r.get("z2_mean"), r.get("ids_differ"), r.get("gpu_nonfinite_pixels"), r.get("oracle_nonfinite_pixels"))
PY
