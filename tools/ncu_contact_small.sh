#!/bin/bash
# ncu --set full + source of ONE single-launch obstacle step at BASELINE configs[2] size (shelf, 16,384 envs)
tag=${1:-r02}
cmd="python tools/one_step.py SHELF_OVERRIDES 16384 1"
$cmd > gpurun_out/plain_contact_small.log 2>&1 &&
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:vine_step -c 1 -f -o gpurun_out/contact_small_$tag $cmd > gpurun_out/ncu_contact_small.log 2>&1
tail -2 gpurun_out/plain_contact_small.log; tail -3 gpurun_out/ncu_contact_small.log
for n in 4096 16384 65536; do python tools/step_time.py FSTR_OVERRIDES $n --graph 2>&1 | tail -1; done
python tools/step_time.py SHELF_OVERRIDES 16384 --graph 2>&1 | head -1
python tools/step_time.py PIPE_DR_OVERRIDES 8192 --graph 2>&1 | head -1
