set -x
cd $GRAFT_REPO_ROOT
for c in 1 2 3 4 6 8 12; do RT_B200_CHUNKS=$c python scripts/prof_render.py C2 1024 3 | sed "s/^/chunks=$c /"; done 2>&1 | tee gpurun_out/r02_c2_chunks.log
python bench.py --workload C4,C3 --steps 2 --warmup 3 --cpu-seconds 4 > gpurun_out/r02_bench_c4c3.json 2> gpurun_out/r02_bench_c4c3.err; tail -3 gpurun_out/r02_bench_c4c3.err; cut -c1-2500 gpurun_out/r02_bench_c4c3.json
ncu --set full --import-source on --clock-control none -k regex:k_render_pool -c 1 -o gpurun_out/r02_c2_v3 python scripts/prof_render.py C2 256 1 > gpurun_out/ncu_c2v3.log 2>&1; tail -2 gpurun_out/ncu_c2v3.log
