# Executed FP32 work of one env step of the bench workload (131,072 robots, 30 warm-up steps), counted from the SASS:
#   bash scripts/flops_r2.sh [outdir]
O=gpurun_out/${1:-r2_flops_sass}; mkdir -p $O
timeout 900 ncu --profile-from-start off --clock-control none --section SourceCounters --import-source on -f -o $O/step \
    python scripts/profile_steady.py 131072 30 1 > $O/ncu_src.log 2>&1
ncu -i $O/step.ncu-rep --page source --csv > /tmp/step_sass.csv 2>> $O/ncu_src.log
python scripts/flops_sass.py /tmp/step_sass.csv 131072 > $O/flops_sass.json; tail -12 $O/flops_sass.json
ls -la $O/step.ncu-rep; rm -f /tmp/step_sass.csv
[ $(stat -c %s $O/step.ncu-rep) -gt 30000000 ] && rm -f $O/step.ncu-rep
timeout 900 ncu --profile-from-start off --clock-control none --csv --log-file $O/metrics.csv \
    --metrics gpu__time_duration.sum,smsp__inst_executed.sum,sm__pipe_fma_cycles_active.sum,sm__cycles_elapsed.sum,smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,smsp__sass_thread_inst_executed_op_ffma_pred_on.sum,dram__bytes_read.sum,dram__bytes_write.sum \
    python scripts/profile_steady.py 131072 30 1 > $O/ncu_met.log 2>&1
tail -2 $O/ncu_met.log
