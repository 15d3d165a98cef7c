FSTR=$(python -c "from vine_robot_isaacgymenvs_b200 import config as c; print(' '.join(c.FSTR_OVERRIDES))")
python -m vine_robot_isaacgymenvs_b200.train $FSTR num_envs=4096 headless=True max_iterations=3 use_graphs=False > gpurun_out/ppo_lstm_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 800 --csv --log-file gpurun_out/launches_ppo_lstm_r01.csv python -m vine_robot_isaacgymenvs_b200.train $FSTR num_envs=4096 headless=True max_iterations=3 use_graphs=False > gpurun_out/ncu_ppo_lstm.log 2>&1
