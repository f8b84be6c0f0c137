#!/bin/bash
# Builds A/B variants of the library: scripts/build_variants.sh name1="-DFLAG ..." name2="..." -> csrc/ab_<name>.so
# (git-ignored, travels to the GPU box; select one with RT_B200_LIB=<path>; scripts/gpu_ab.py times them side by side)
set -e
cd "$(dirname "$0")/../mcp_raytracer_b200/csrc"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
for spec in "$@"; do
  name="${spec%%=*}"; flags="${spec#*=}"
  [ "$name" = "$spec" ] && flags=""
  ( $NVCC -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-Wall,-Wno-unused-function -Xptxas -v \
      $flags -shared -o ab_$name.so rt_api.cu rt_megakernel.cu rt_scene.cpp 2> ab_$name.log && echo "built ab_$name.so [$flags]" ) &
done
wait
