#!/bin/bash
# ncu part of scripts/gpu_round.sh alone (launch list + --set full of the step kernels): bash scripts/ncu_round.sh <tag>
O=gpurun_out/${1:-cur}; mkdir -p $O
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 60 --csv --log-file $O/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_solve|k_dyn|k_rank|k_post" -s 40 -c 6 \
    -o $O/full -f python scripts/profile_step.py 32768 5 > $O/full.log 2>&1
ncu -i $O/full.ncu-rep --page raw --csv > $O/full_raw.csv 2>/dev/null
ncu -i $O/full.ncu-rep --page source --csv -k regex:k_solve > $O/solve_sass.csv 2>/dev/null
ncu -i $O/full.ncu-rep --page source --csv -k regex:k_dyn > $O/dyn_sass.csv 2>/dev/null
rm -f $O/full.ncu-rep
ls -la $O
