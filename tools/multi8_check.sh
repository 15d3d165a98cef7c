#!/bin/bash
# 8-GPU record of the final code: the default bench line under torchrun, then the gradient-exchange comparison
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --steps 50 --warmup 5 --ppo-iters 10 --no-sweep > gpurun_out/bench8g.json 2> gpurun_out/bench8g.err; echo "bench8 rc=$?"; tail -2 gpurun_out/bench8g.err | cut -c1-200
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29552 tools/p2p_check.py > gpurun_out/p2p8g.log 2>&1; tail -1 gpurun_out/p2p8g.log | cut -c1-200
