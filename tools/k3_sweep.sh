#!/bin/bash
# Rebuild the library with different -D settings on the GPU box and time the fused kernels.
# usage: bash tools/k3_sweep.sh "TAG1:-DFOO=1 -DBAR=2" "TAG2:..."
mkdir -p gpurun_out
for spec in "$@"; do
  tag="${spec%%:*}"; defs="${spec#*:}"
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC -DML_TMA_FAST_BUILD $defs \
     -o momlevel_b200/libmomlevel_b200.so momlevel_b200/csrc/ml_api.cu momlevel_b200/csrc/ml_tma.cu momlevel_b200/csrc/ml_hostpath.cu momlevel_b200/csrc/ml_strat.cu 2> gpurun_out/sweep_build_$tag.log || { echo "build failed $tag"; tail -5 gpurun_out/sweep_build_$tag.log; continue; }
  MOMLEVEL_B200_LIB=$PWD/momlevel_b200/libmomlevel_b200.so timeout 120 python tools/k3_bench.py "$tag" 2>&1 | tail -1 | tee -a gpurun_out/sweep.log
done
