#!/bin/bash
# launch list (ncu, per-launch durations: cold-cache, serialised) of 3 training iterations of the reference network, no graphs
FSTR=$(python -c "from vine_robot_isaacgymenvs_b200 import config as c; print(' '.join(c.FSTR_OVERRIDES))")
CMD="python -m vine_robot_isaacgymenvs_b200.train $FSTR num_envs=4096 headless=True max_iterations=3 use_graphs=False"
$CMD > gpurun_out/ppo_lstm_plain.log 2>&1 || { tail -5 gpurun_out/ppo_lstm_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none ${NCU_EXTRA} -s 300 -c 800 --csv --log-file gpurun_out/launches_ppo_lstm_${1:-r02}.csv $CMD > gpurun_out/ncu_ppo_lstm.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_ppo_lstm_${1:-r02}.csv | head -30
