#!/bin/bash
# A/B pass on one GPU box: time plen_step of every library under scripts/ab/ and of the in-tree build.
# Usage (under gpurun): bash scripts/ab_round.sh <tag> [envs]
TAG=${1:-ab}; E=${2:-131072}
O=gpurun_out/$TAG
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1
for rep in 1 2; do
  for L in scripts/ab/*.so plen_ml_walk_b200/libplen_b200.so; do
    timeout 300 python scripts/ab_time.py $L $E 30 2>&1 | tail -1 | tee -a $O/ab.txt
  done
done
