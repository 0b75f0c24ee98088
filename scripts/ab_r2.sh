O=gpurun_out/${1:-r2_man2}; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_properties.py -m gpu -x -q -s -k "manifold" > $O/pytest_man.log 2>&1; grep "manifold mode\|passed\|failed\|Error\|assert" $O/pytest_man.log | tail -12
for m in 0 1; do
  timeout 600 python scripts/walk_eval_batched.py --envs 4096 --sigma 0.1 --sole-manifold $m > $O/walk_eval_sigma0.1_man$m.json 2>> $O/err.txt; cat $O/walk_eval_sigma0.1_man$m.json
done
for L in scripts/ab/libplen_cur.so plen_ml_walk_b200/libplen_b200.so; do
  python scripts/ab_time.py $L 131072 50 2>&1 | tail -1 | tee -a $O/ab.txt
  python scripts/ab_time.py $L 4096 200 2>&1 | tail -1 | tee -a $O/ab.txt
done
tail -3 $O/err.txt
