"""Run a preset for a few control steps with a CUDA-profiler range around the last ones:
    ncu --profile-from-start off ... python tools/one_step.py PRESET N [profiled_steps] [override ...]
Warm-up: 130 steps (episodes of every preset desynchronised, resets and -- with obstacles -- contacts present)."""
import sys

import torch

sys.path.insert(0, ".")
import vine_robot_isaacgymenvs_b200 as vine  # noqa: E402
from vine_robot_isaacgymenvs_b200 import config as vcfg  # noqa: E402

preset, n = getattr(vcfg, sys.argv[1]), int(sys.argv[2])
k = int(sys.argv[3]) if len(sys.argv) > 3 and sys.argv[3].isdigit() else 2
extra = [a for a in sys.argv[3:] if not a.isdigit()]
env = vine.make(cfg=vcfg.compose(preset + [f"num_envs={n}", "headless=True"] + extra))
g = torch.Generator(device="cuda").manual_seed(0)
acts = [torch.rand(n, 2, device="cuda", generator=g) * 2 - 1 for _ in range(8)]
for t in range(130):
    env.actions.copy_(acts[t % 8]); env.step_device()
torch.cuda.synchronize()
torch.cuda.profiler.start()
for t in range(k):
    env.actions.copy_(acts[t % 8]); env.step_device()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok", sys.argv[1], n, float(env.rew_buf.mean()))
