O=gpurun_out/${1:-r2_tc5}; mkdir -p $O
timeout 600 python -m pytest tests/test_td3_gpu.py -m gpu -x -q > $O/pytest_td3.log 2>&1; tail -4 $O/pytest_td3.log
timeout 600 python scripts/td3_bench.py 300 > $O/td3_bench.json 2> $O/td3_bench.err; cat $O/td3_bench.json; tail -3 $O/td3_bench.err
timeout 600 ncu --set full --clock-control none -k regex:"k_gemm_tc" -c 24 -o $O/gemm_tc -f python scripts/tc_debug.py 4096 > $O/ncu.log 2>&1
ncu -i $O/gemm_tc.ncu-rep --page raw --csv > $O/gemm_tc_raw.csv 2>/dev/null
rm -f $O/gemm_tc.ncu-rep
