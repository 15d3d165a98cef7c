#!/bin/bash
# FSTR learning curves of the two kernel-only PPO paths (400 iterations) and the shelf config (300), for profiles/
tag=${1:-r02}
F="task=Vine5LinkMovingBase wandb_activate=False task.env.RAIL_SOFT_LIMIT=0.25 RAIL_P_GAIN=30 RAIL_ACCELERATION=6 RAIL_VELOCITY_SCALE=1 task.env.CREATE_SHELF=False task.env.CREATE_PIPE=False vine_randomize=True OBSERVATION_TYPE=TIP_AND_CART_AND_OBJ_INFO task.env.ACTION_DELAY=1 task.env.maxEpisodeLength=100 task.env.SUCCESS_DIST=0.04 task.env.MIN_TARGET_Y=-0.4 task.env.MAX_TARGET_Y=0.4 task.env.MIN_TARGET_Z=0.55 task.env.MAX_TARGET_Z=0.7 task.env.MIN_TARGET_DEPTH_IN_OBSTACLE=0.0 task.env.MAX_TARGET_DEPTH_IN_OBSTACLE=0.0 task.env.CONTACT_FORCE_REWARD_WEIGHT=0.0 task.task.randomization_parameters.DYNAMICS_SCALING_MIN=0.999999 task.task.randomization_parameters.DYNAMICS_SCALING_MAX=1.000001 task.task.randomization_parameters.ACTION_NOISE_STD=0.001 task.task.randomization_parameters.OBSERVATION_NOISE_STD=0.0 +task.task.randomization_parameters.ACCEL_TARGET_SCALING_MIN=0.99 +task.task.randomization_parameters.ACCEL_TARGET_SCALING_MAX=1.05"
export PYTHONPATH=$GRAFT_REPO_ROOT:$PYTHONPATH
cd /tmp && rm -rf runs
python -m vine_robot_isaacgymenvs_b200.train $F num_envs=4096 headless=True max_iterations=400 > $GRAFT_REPO_ROOT/gpurun_out/ppo_fstr_${tag}_training_lstm_native.log 2>&1
python -m vine_robot_isaacgymenvs_b200.train $F num_envs=4096 headless=True max_iterations=400 train.params.network.rnn=null > $GRAFT_REPO_ROOT/gpurun_out/ppo_fstr_${tag}_training_mlp_native.log 2>&1
ls /tmp/runs/Vine5LinkMovingBase /tmp/runs/Vine5LinkMovingBase/nn /tmp/runs/Vine5LinkMovingBase/summaries > $GRAFT_REPO_ROOT/gpurun_out/ppo_run_dir_${tag}.txt 2>&1
python -m vine_robot_isaacgymenvs_b200.train $F num_envs=4096 headless=True test=True checkpoint=/tmp/runs/Vine5LinkMovingBase/nn/Vine5LinkMovingBase.pth train.params.network.rnn=null play_steps=800 >> $GRAFT_REPO_ROOT/gpurun_out/ppo_fstr_${tag}_training_mlp_native.log 2>&1
tail -3 $GRAFT_REPO_ROOT/gpurun_out/ppo_fstr_${tag}_training_lstm_native.log; tail -4 $GRAFT_REPO_ROOT/gpurun_out/ppo_fstr_${tag}_training_mlp_native.log; cat $GRAFT_REPO_ROOT/gpurun_out/ppo_run_dir_${tag}.txt
