#!/bin/bash
# Dev tool: build an A/B variant of libplen_b200.so into scripts/ab/ (git-ignored, shipped to the GPU box).
# Usage: scripts/build_variant.sh <name> [-DPLEN_...=... ...]
N=$1; shift
mkdir -p scripts/ab
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --shared -Xcompiler -fPIC -Xptxas -v \
  -cudart static -ccbin /usr/bin/g++ -I . "$@" -o scripts/ab/libplen_$N.so plen_ml_walk_b200/csrc/plen_b200.cu plen_ml_walk_b200/csrc/plen_td3.cu 2>&1 \
  | grep -A3 "Compiling.*k_solve" | grep -i "spill\|Used" | tr '\n' ' '; echo " <- $N"
