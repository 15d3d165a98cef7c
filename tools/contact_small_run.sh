#!/bin/bash
# obstacle steps: parity/invariance tests, then step times at the BASELINE sizes and at 1 M envs
timeout 900 python -m pytest tests/test_gpu_env_api.py tests/test_gpu_parity.py tests/test_gpu_presets_parity.py -x -q 2>&1 | tail -2
python tools/step_time.py SHELF_OVERRIDES 16384 --graph 2>&1 | head -1
python tools/step_time.py PIPE_DR_OVERRIDES 8192 --graph 2>&1 | head -1
python tools/step_time.py SHELF_OVERRIDES 1048576 --graph 2>&1 | head -1
python tools/step_time.py PIPE_DR_OVERRIDES 1048576 --graph 2>&1 | head -1
