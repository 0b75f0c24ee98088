O=gpurun_out/${1:-r2_g2}; mkdir -p $O
for rep in 1 2; do
  for L in scripts/ab/libplen_cur.so plen_ml_walk_b200/libplen_b200.so; do
    python scripts/ab_time.py $L 131072 50 2>&1 | tail -1 | tee -a $O/ab.txt
  done
done
PLEN_MERGE_MAX=0 timeout 900 /usr/local/cuda/bin/compute-sanitizer --tool racecheck python scripts/profile_steady.py 1100 40 1 > $O/racecheck_sep.log 2>&1; tail -1 $O/racecheck_sep.log
timeout 900 /usr/local/cuda/bin/compute-sanitizer --tool racecheck python scripts/profile_steady.py 1100 40 1 > $O/racecheck_merged.log 2>&1; tail -1 $O/racecheck_merged.log
timeout 600 python scripts/trajectory_eval_batched.py > $O/config3.json 2> $O/config3.err; cat $O/config3.json
timeout 300 python scripts/walk_eval_batched.py > $O/walk_eval_fp32.json 2> $O/walk_eval.err; cat $O/walk_eval_fp32.json
timeout 300 python scripts/walk_eval_batched.py --precision fp16 > $O/walk_eval_fp16.json 2>> $O/walk_eval.err; cat $O/walk_eval_fp16.json
timeout 200 python scripts/actor_bench.py > $O/actor_bench.json 2> $O/actor_bench.err; cat $O/actor_bench.json
timeout 600 python -m pytest tests/test_gpu_properties.py tests/test_gpu_parity.py -m gpu -x -q > $O/pytest.log 2>&1; tail -2 $O/pytest.log
