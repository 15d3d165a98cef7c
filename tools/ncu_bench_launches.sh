#!/bin/bash
# launch list (per-launch durations, cold-cache and serialised) of the default bench command's env-step part
CMD="python bench.py --steps 2 --warmup 3 --no-ppo --no-sweep --no-cpu-baseline --e2e-steps 3 --no-graph"
$CMD > gpurun_out/bench_plain_r02.json 2> gpurun_out/bench_plain_r02.err || { tail -3 gpurun_out/bench_plain_r02.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:vine_ -c 400 --csv --log-file gpurun_out/launches_bench_r02.csv $CMD > gpurun_out/ncu_bench.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_bench_r02.csv | head -12
