#!/bin/bash
# A/B at small batch sizes: bash scripts/ab_small.sh <tag>
O=gpurun_out/$1; mkdir -p $O
for E in ${SIZES:-2048 4096 8192 16384}; do
  for L in scripts/ab/*.so plen_ml_walk_b200/libplen_b200.so; do
    timeout 200 python scripts/ab_time.py $L $E 300 2>&1 | tail -1 | sed "s/^/E=$E /" | tee -a $O/ab_small.txt
  done
done
