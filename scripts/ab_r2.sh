O=gpurun_out/${1:-r2_tc9}; mkdir -p $O
timeout 600 python -m pytest tests/test_td3_gpu.py -m gpu -x -q > $O/pytest_td3.log 2>&1; tail -3 $O/pytest_td3.log
timeout 600 python scripts/td3_bench.py 300 > $O/td3_bench.json 2> $O/td3_bench.err; python -c "
import json; d=json.load(open('$O/td3_bench.json'))
for k,v in d.items(): print(k, 'fp32 %.1f us  tf32 %.1f us  launches %.1f' % (v['cuda_device_us_per_update'], v['tf32_device_us_per_update'], v['launches_per_update']))"; tail -3 $O/td3_bench.err
