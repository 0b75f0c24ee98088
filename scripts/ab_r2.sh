O=gpurun_out/${1:-r2_ab7}; mkdir -p $O
for rep in 1 2; do
  for L in scripts/ab/libplen_cur.so scripts/ab/libplen_rs1.so scripts/ab/libplen_rs2.so scripts/ab/libplen_rs3.so; do
    python scripts/ab_time.py $L 131072 50 2>&1 | tail -1 | tee -a $O/ab.txt
    PLEN_AB_NOLINKS=1 python scripts/ab_time.py $L 131072 50 2>&1 | tail -1 | sed 's/^/nolinks /' | tee -a $O/ab.txt
  done
done
