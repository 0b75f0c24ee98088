O=gpurun_out/${1:-r2_flops}; mkdir -p $O
M="smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,smsp__sass_thread_inst_executed_op_ffma_pred_on.sum,smsp__inst_executed.sum,gpu__time_duration.sum"
# one env step of 131,072 robots (2 ranges x 4 ticks x (k_dyn, k_rank, k_solve, k_solve_x) + 2 k_post = 34 launches) at step 30 and at step 60
timeout 600 ncu --metrics $M --clock-control none -k regex:"k_dyn|k_solve|k_rank|k_post" -s 1020 -c 34 --csv --log-file $O/flops_step30.csv python scripts/profile_steady.py 131072 30 2 > $O/flops30.log 2>&1
timeout 600 ncu --metrics $M --clock-control none -k regex:"k_dyn|k_solve|k_rank|k_post" -s 2040 -c 34 --csv --log-file $O/flops_step60.csv python scripts/profile_steady.py 131072 60 2 > $O/flops60.log 2>&1
tail -1 $O/flops30.log $O/flops60.log
