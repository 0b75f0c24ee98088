#!/bin/bash
# One GPU-box pass: GPU tests, bench (both arms), ncu launch list, ncu --set full of the two hot kernels.
# Usage (from the repo root, under gpurun): bash scripts/gpu_round.sh <tag>
TAG=${1:-cur}
O=gpurun_out/$TAG
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1
timeout 600 python bench.py > $O/bench.json 2> $O/bench.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > $O/bench_ref.json 2> $O/bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 60 --csv --log-file $O/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_solve|k_dyn|k_rank|k_post" -s 40 -c 6 \
    -o $O/full -f python scripts/profile_step.py 32768 5 > $O/full.log 2>&1
ncu -i $O/full.ncu-rep --page raw --csv > $O/full_raw.csv 2>/dev/null
ncu -i $O/full.ncu-rep --page source --csv -k regex:k_solve > $O/solve_sass.csv 2>/dev/null
ncu -i $O/full.ncu-rep --page source --csv -k regex:k_dyn > $O/dyn_sass.csv 2>/dev/null
timeout 300 python scripts/td3_bench.py 1000 > $O/td3_bench.json 2> $O/td3_bench.err
timeout 200 python scripts/actor_bench.py > $O/actor_bench.json 2> $O/actor_bench.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_actor_forward_tc" -s 30 -c 1 -o $O/actor_tc -f \
    python scripts/actor_bench.py > $O/actor_tc.log 2>&1
ncu -i $O/actor_tc.ncu-rep --page raw --csv > $O/actor_tc_raw.csv 2>/dev/null
timeout 600 python scripts/plen_td3_batched.py --envs 16384 --env-steps 1048576 --updates-per-step 8 > $O/config4_u8.json 2> $O/config4.err
timeout 600 python scripts/plen_td3_batched.py --envs 16384 --env-steps 1048576 --updates-per-step 8 --actor-precision fp16 --batch-size 1024 > $O/config4_u8_b1024_fp16.json 2>> $O/config4.err
timeout 600 python scripts/plen_td3_batched.py --envs 16384 --env-steps 1048576 --no-learner > $O/config4_nolearner.json 2>> $O/config4.err
timeout 600 python scripts/trajectory_eval_batched.py > $O/config3.json 2> $O/config3.err
timeout 300 python scripts/walk_eval_batched.py > $O/walk_eval_fp32.json 2> $O/walk_eval.err
timeout 300 python scripts/walk_eval_batched.py --precision fp16 > $O/walk_eval_fp16.json 2>> $O/walk_eval.err
python scripts/ab_time.py plen_ml_walk_b200/libplen_b200.so 4096 200 > $O/config2_4096.txt 2>&1
python scripts/ab_time.py plen_ml_walk_b200/libplen_b200.so 1048576 10 > $O/config5_1m.txt 2>&1
tail -3 $O/pytest_gpu.log; cat $O/smoke.log | tail -2; cat $O/bench.json; cat $O/bench_ref.json
