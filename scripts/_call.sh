cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=$1; WL=$2
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --workload $WL > gpurun_out/r02g_scale_n$N.jsonl 2> gpurun_out/r02g_scale_n$N.err
python - <<PY
import json
for l in open("gpurun_out/r02g_scale_n$N.jsonl"):
    j=json.loads(l); print(j["config"]["workload"][:24], j["n_gpus"], round(j["value"]), round(j["ms_per_step"],2), round(j["e2e"]["value"]), j["rank_ms"]["min"], j["rank_ms"]["max"], j["fb_sha1"][:12])
PY
