"""Converged-image parity at BASELINE sizes, all five configs (the cases of tests/test_gpu_full_parity.py), with the
numbers written out: gpurun_out/<tag>_full_parity.json + one JSON line per case on stdout.
usage: gpu_full_parity.py [tag] [cases...]     (cases: names of tests/test_gpu_full_parity.py CASES, c4-primary, c4-same-seed, c4-converged)"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import parity_stats as ps
import test_gpu_full_parity as T

tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
want = sys.argv[2:] or list(T.CASES) + ["c4-primary", "c4-same-seed", "c4-converged"]
rows = []
for name in want:
    t0 = time.perf_counter()
    if name == "c4-primary": r = T.run_c4_primary()
    elif name == "c4-same-seed": r = T.run_c4_same_seed()
    else:
        r = T.run_c4_converged() if name == "c4-converged" else T.run_converged_case(name)
        try:
            ps.check_converged(r, firefly_allowance=0.004 if name == "c4-converged" else 0.002, one_percent_bar=(name != "c4-converged")); r["bars_met"] = True
        except AssertionError:
            r["bars_met"] = False
    r["wall_s"] = time.perf_counter() - t0
    rows.append(r)
    print(json.dumps(r), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", f"{tag}_full_parity.json"), "w"), indent=1)
