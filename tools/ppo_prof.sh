set -x
FSTR=$(python -c "from vine_robot_isaacgymenvs_b200 import config as c; print(' '.join(c.FSTR_OVERRIDES))")
python -m vine_robot_isaacgymenvs_b200.train $FSTR num_envs=4096 headless=True max_iterations=4 train.params.network.rnn=null use_graphs=False > gpurun_out/ppo_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_ppo_r01.csv python -m vine_robot_isaacgymenvs_b200.train $FSTR num_envs=4096 headless=True max_iterations=4 train.params.network.rnn=null use_graphs=False > gpurun_out/ncu_ppo.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:vine_ppo_minibatch_kernel -s 4 -c 1 -o gpurun_out/prof_ppo_minibatch_r01 python -m vine_robot_isaacgymenvs_b200.train $FSTR num_envs=4096 headless=True max_iterations=4 train.params.network.rnn=null use_graphs=False > gpurun_out/ncu_ppo2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:vine_policy_act_kernel -s 20 -c 1 -o gpurun_out/prof_policy_act_r01 python -m vine_robot_isaacgymenvs_b200.train $FSTR num_envs=4096 headless=True max_iterations=4 train.params.network.rnn=null use_graphs=False > gpurun_out/ncu_ppo3.log 2>&1
ls -la gpurun_out | tail -5
