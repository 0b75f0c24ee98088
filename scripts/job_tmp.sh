O=gpurun_out/r2_man3; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_properties.py -m gpu -x -q -s -k "manifold" > $O/pytest_man.log 2>&1; grep "manifold mode\|passed\|failed\|Error\|assert" $O/pytest_man.log | tail -12
timeout 600 python scripts/walk_eval_batched.py --envs 4096 --sigma 0.1 --sole-manifold 1 > $O/walk_eval_sigma0.1_man1.json 2>> $O/err.txt; cat $O/walk_eval_sigma0.1_man1.json
bash scripts/flops_r2.sh
