FSTR=$(python -c "from vine_robot_isaacgymenvs_b200 import config as c; print(' '.join(c.FSTR_OVERRIDES))")
ARGS="$FSTR num_envs=65536 headless=True max_iterations=2 use_graphs=False train.params.config.minibatch_size=131072 train.params.config.mini_epochs=1"
python -m vine_robot_isaacgymenvs_b200.train $ARGS > gpurun_out/ppo_lstm_big_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 500 --csv --log-file gpurun_out/launches_ppo_lstm_big_r01.csv python -m vine_robot_isaacgymenvs_b200.train $ARGS > gpurun_out/ncu_ppo_lstm_big.log 2>&1
