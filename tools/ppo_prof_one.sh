#!/bin/bash
# ncu --set full + source of ONE kernel of the recurrent update (second iteration): tools/ppo_prof_one.sh KERNEL_REGEX [skip]
FSTR=$(python -c "from vine_robot_isaacgymenvs_b200 import config as c; print(' '.join(c.FSTR_OVERRIDES))")
CMD="python -m vine_robot_isaacgymenvs_b200.train $FSTR num_envs=4096 headless=True max_iterations=3 use_graphs=False"
$CMD > gpurun_out/ppo_lstm_plain.log 2>&1 || { tail -5 gpurun_out/ppo_lstm_plain.log; exit 1; }
ncu --set full --import-source on --clock-control none -k regex:$1 -s ${2:-10} -c 1 -f -o gpurun_out/src_$1 $CMD > gpurun_out/ncu_src_$1.log 2>&1
ls -la gpurun_out | grep src_
