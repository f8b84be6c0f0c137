cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "reference_topology_walk or deep_tree" 2>&1 | tail -3
timeout 200 python bench.py --steps 2 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; j=json.loads(sys.stdin.read()); print(j['value'], j['reference_defaults_adaptive'])"
