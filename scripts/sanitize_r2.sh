#!/bin/bash
# compute-sanitizer pass over the round-2 kernels: bash scripts/sanitize_r2.sh <tag>   (under gpurun)
O=gpurun_out/${1:-r2_san}; mkdir -p $O
S=/usr/local/cuda/bin/compute-sanitizer
# steady-state step with box contacts: separate launches (k_solve<false> + k_solve_x, EXT = 2) and the merged kernel (EXT = 1)
PLEN_MERGE_MAX=0 timeout 900 $S --tool memcheck python scripts/profile_steady.py 2500 45 2 > $O/memcheck_sep.log 2>&1; tail -2 $O/memcheck_sep.log
timeout 900 $S --tool memcheck python scripts/profile_steady.py 2500 45 2 > $O/memcheck_merged.log 2>&1; tail -2 $O/memcheck_merged.log
PLEN_MERGE_MAX=0 timeout 900 $S --tool racecheck python scripts/profile_steady.py 1100 40 1 > $O/racecheck_sep.log 2>&1; tail -2 $O/racecheck_sep.log
timeout 900 $S --tool racecheck python scripts/profile_steady.py 1100 40 1 > $O/racecheck_merged.log 2>&1; tail -2 $O/racecheck_merged.log
timeout 900 $S --tool initcheck python scripts/profile_steady.py 1100 40 1 > $O/initcheck.log 2>&1; tail -2 $O/initcheck.log
timeout 900 $S --tool synccheck python scripts/profile_steady.py 1100 40 1 > $O/synccheck.log 2>&1; tail -2 $O/synccheck.log
timeout 1200 $S --tool memcheck python -m pytest tests/test_td3_gpu.py -m gpu -q -k "tensor_core" > $O/memcheck_td3_tc.log 2>&1; tail -3 $O/memcheck_td3_tc.log
