#!/bin/bash
# FP32 thread-instruction counts + device time of every kernel of two control steps, per BASELINE config (-> profiles/):
#   bash tools/ncu_flops.sh <tag>      (under gpurun, one GPU)
tag=${1:-r02}
M=smsp__sass_thread_inst_executed_op_ffma_pred_on.sum,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,smsp__thread_inst_executed.sum,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
for cfg in "FSTR_OVERRIDES 1048576" "SHELF_OVERRIDES 16384" "PIPE_DR_OVERRIDES 8192" "SHELF_OVERRIDES 1048576" "PIPE_DR_OVERRIDES 1048576"; do
  set -- $cfg
  python tools/one_step.py $1 $2 > gpurun_out/plain_flops_$1_$2.log 2>&1 &&
  ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file gpurun_out/flops_$1_$2_$tag.csv python tools/one_step.py $1 $2 > gpurun_out/ncu_flops_$1_$2.log 2>&1
  tail -1 gpurun_out/plain_flops_$1_$2.log
done
