O=gpurun_out/${1:-r2_ab6}; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; tail -3 $O/pytest_gpu.log
for rep in 1 2; do
  for L in scripts/ab/libplen_head.so plen_ml_walk_b200/libplen_b200.so; do
    python scripts/ab_time.py $L 131072 50 2>&1 | tail -1 | tee -a $O/ab.txt
    PLEN_AB_WARM=150 python scripts/ab_time.py $L 131072 50 2>&1 | tail -1 | sed 's/^/warm150 /' | tee -a $O/ab.txt
    python scripts/ab_time.py $L 4096 200 2>&1 | tail -1 | tee -a $O/ab.txt
    python scripts/ab_time.py $L 65536 50 2>&1 | tail -1 | tee -a $O/ab.txt
    python scripts/ab_time.py $L 1048576 10 2>&1 | tail -1 | tee -a $O/ab.txt
  done
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_dyn|k_solve|k_rank|k_post" -s 2040 -c 68 --csv --log-file $O/steady_launches.csv python scripts/profile_steady.py 131072 60 2 > $O/steady_launches.log 2>&1
