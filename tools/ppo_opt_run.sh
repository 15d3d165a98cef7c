#!/bin/bash
# PPO kernel work: the LSTM / PPO GPU tests, then iteration timings of the reference network and of the MLP actor-critic
timeout 900 python -m pytest tests/test_gpu_lstm_net.py tests/test_gpu_ppo.py tests/test_gpu_ppo_fused.py -x -q 2>&1 | tail -3
timeout 300 python tools/ppo_time.py 4096 2>&1 | tail -1
timeout 300 python tools/ppo_time.py 4096 2>&1 | tail -1
