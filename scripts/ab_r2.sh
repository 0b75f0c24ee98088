O=gpurun_out/${1:-r2_c4}; mkdir -p $O
R="python scripts/plen_td3_batched.py --envs 16384 --actor-precision fp16"
timeout 600 $R --env-steps 4194304 --no-learner > $O/config4_nolearner.json 2> $O/err.txt; cat $O/config4_nolearner.json
timeout 600 $R --env-steps 4194304 --updates-per-step 8 --batch-size 100 > $O/config4_u8_b100_fp32.json 2>> $O/err.txt; cat $O/config4_u8_b100_fp32.json
timeout 600 $R --env-steps 4194304 --updates-per-step 4 --batch-size 4096 --learner-precision tf32 > $O/config4_u4_b4096_tf32.json 2>> $O/err.txt; cat $O/config4_u4_b4096_tf32.json
timeout 600 $R --env-steps 4194304 --updates-per-step 4 --batch-size 4096 > $O/config4_u4_b4096_fp32.json 2>> $O/err.txt; cat $O/config4_u4_b4096_fp32.json
timeout 600 $R --env-steps 2097152 --updates-per-step 8 --batch-size 16384 --learner-precision tf32 > $O/config4_u8_b16384_tf32.json 2>> $O/err.txt; cat $O/config4_u8_b16384_tf32.json
timeout 600 $R --env-steps 524288 --updates-per-step 100 --batch-size 16384 --learner-precision tf32 > $O/config4_u100_b16384_tf32.json 2>> $O/err.txt; cat $O/config4_u100_b16384_tf32.json
tail -3 $O/err.txt
