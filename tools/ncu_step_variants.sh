#!/bin/bash
# ncu --set full of the free-space step kernel, both variants (one env per thread / two envs packed), 1 M envs.
# Usage (under gpurun): bash tools/ncu_step_variants.sh <tag>
set -e
tag=${1:-r02}
for v in one_env_per_thread two_envs_packed; do
  cmd="python tools/step_time.py FSTR_OVERRIDES 1048576 +task.sim.vine_step_kernel=$v"
  $cmd > gpurun_out/plain_$v.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:vine_step -s 60 -c 1 -f -o gpurun_out/step_${v}_$tag $cmd > gpurun_out/ncu_$v.log 2>&1
  cat gpurun_out/plain_$v.log
done
