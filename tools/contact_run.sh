python -m pytest tests/test_gpu_presets_parity.py tests/test_gpu_env_api.py tests/test_gpu_parity.py -x -q -m gpu -s 2>&1 | grep -E "wandb_dict|parity |passed|failed|Error|assert" | head -20
for cfg in "SHELF_OVERRIDES 1048576" "PIPE_DR_OVERRIDES 1048576" "SHELF_OVERRIDES 16384" "PIPE_DR_OVERRIDES 8192" "FSTR_OVERRIDES 1048576"; do
python tools/step_time.py $cfg
done
