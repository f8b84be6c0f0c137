cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 200 python -m pytest tests -q -m gpu -x -k "same_seed or material or scatter or stats or weekend or C3 or C1" 2>&1 | tail -2
timeout 300 python bench.py --workload C3,C4 --no-cpu-baseline > gpurun_out/r02h_c3_c4.jsonl 2>/dev/null; python -c "
import json
for l in open('gpurun_out/r02h_c3_c4.jsonl'):
    j=json.loads(l); print(j['config']['workload'][:24], round(j['ms_per_step'],1), round(j['value']), round(j['grays_per_s'],2), j['fb_sha1'][:12])"
