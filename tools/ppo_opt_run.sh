#!/bin/bash
# PPO kernel work: the LSTM / PPO GPU tests, then iteration timings of the reference network and of the MLP actor-critic
timeout 900 python -m pytest tests/test_gpu_lstm_net.py tests/test_gpu_ppo.py tests/test_gpu_ppo_fused.py tests/test_gpu_mlp.py tests/test_gpu_env_api.py -x -q 2>&1 | tail -3
for a in "" "--no-pdl"; do
timeout 300 python tools/ppo_time.py 4096 $a 2>&1 | tail -1
timeout 300 python tools/ppo_time.py 4096 --mlp $a 2>&1 | tail -1
timeout 300 python tools/ppo_time.py 16384 $a 2>&1 | tail -1
done
timeout 300 python tools/ppo_time.py 65536 2>&1 | tail -1
