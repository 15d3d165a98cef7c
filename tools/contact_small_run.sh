#!/bin/bash
# small single-launch obstacle steps: parity/invariance tests, then step times at the BASELINE sizes
timeout 900 python -m pytest tests/test_gpu_env_api.py tests/test_gpu_parity.py tests/test_gpu_presets_parity.py -x -q 2>&1 | tail -3
for n in 4096 8192 16384 28416 32768 65536; do python tools/step_time.py SHELF_OVERRIDES $n --graph 2>&1 | head -1; done
for n in 4096 8192 16384 32768; do python tools/step_time.py PIPE_DR_OVERRIDES $n --graph 2>&1 | head -1; done
