# last sanity check of the committed tree: smoke + a slice of the GPU suite
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 60 python -m pytest tests -q -m gpu -x -k "reference_topology_walk or metal or c_abi or zero" 2>&1 | tail -2
