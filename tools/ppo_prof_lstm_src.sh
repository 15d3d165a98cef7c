# source-level ncu captures of the three heaviest LSTM-path kernels (training-size launches of the second iteration)
FSTR=$(python -c "from vine_robot_isaacgymenvs_b200 import config as c; print(' '.join(c.FSTR_OVERRIDES))")
CMD="python -m vine_robot_isaacgymenvs_b200.train $FSTR num_envs=4096 headless=True max_iterations=3 use_graphs=False"
$CMD > gpurun_out/ppo_lstm_plain.log 2>&1 || exit 1
for k in vine_lstm_step_kernel vine_lstm_head_train_kernel vine_lstm_cell_bwd_tiles_kernel vine_lstm_bwd_gemm_kernel; do
  ncu --set full --import-source on --clock-control none -k regex:$k -s 10 -c 1 -o gpurun_out/src_$k $CMD > gpurun_out/ncu_src_$k.log 2>&1
done
ls gpurun_out | grep src_
