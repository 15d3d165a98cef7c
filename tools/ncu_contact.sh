#!/bin/bash
# ncu --set full of the kernels of ONE routed obstacle step (shelf preset, 1 M envs): near pass, far pass, redo pass
tag=${1:-r02}
cmd="python tools/one_step.py SHELF_OVERRIDES 1048576 1"
$cmd > gpurun_out/plain_contact.log 2>&1 &&
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:vine_step -c 3 -f -o gpurun_out/contact_step_$tag $cmd > gpurun_out/ncu_contact.log 2>&1
tail -2 gpurun_out/plain_contact.log; tail -3 gpurun_out/ncu_contact.log
