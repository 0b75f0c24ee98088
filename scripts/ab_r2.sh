O=gpurun_out/${1:-r2_g1}; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; tail -3 $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -1 $O/smoke.log
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err; cat $O/bench.json; tail -3 $O/bench.err
timeout 600 python scripts/plen_td3_batched.py --envs 16384 --env-steps 1048576 --updates-per-step 4 --batch-size 4096 --learner-precision tf32 --actor-precision fp16 > $O/config4_u4_b4096_tf32.json 2> $O/config4.err; cat $O/config4_u4_b4096_tf32.json
timeout 600 python scripts/plen_td3_batched.py --envs 16384 --env-steps 1048576 --updates-per-step 4 --batch-size 4096 --actor-precision fp16 > $O/config4_u4_b4096_fp32.json 2>> $O/config4.err; cat $O/config4_u4_b4096_fp32.json
timeout 600 python scripts/plen_td3_batched.py --envs 16384 --env-steps 524288 --updates-per-step 25 --batch-size 16384 --learner-precision tf32 --actor-precision fp16 > $O/config4_u25_b16384_tf32.json 2>> $O/config4.err; cat $O/config4_u25_b16384_tf32.json
tail -3 $O/config4.err
