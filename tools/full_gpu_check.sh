#!/bin/bash
# full single-GPU verification: the whole -m gpu suite, smoke(), and the default bench line
python -m pytest tests -x -q -m gpu 2>&1 | tail -6
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 50 --warmup 5 --ppo-iters 10 > gpurun_out/bench_r02.json 2> gpurun_out/bench_r02.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_r02.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r02.json 2>/dev/null; echo "ref rc=$?"
