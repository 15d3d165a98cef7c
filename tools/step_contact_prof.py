"""Drive the CONTACT=true step kernel (shelf preset) for profiling: python tools/step_contact_prof.py [preset] [num_envs]"""
import sys

import torch

sys.path.insert(0, ".")
import vine_robot_isaacgymenvs_b200 as vine  # noqa: E402
from vine_robot_isaacgymenvs_b200 import config as vcfg  # noqa: E402

preset = getattr(vcfg, sys.argv[1] if len(sys.argv) > 1 else "SHELF_OVERRIDES")
n = int(sys.argv[2]) if len(sys.argv) > 2 else 262144
env = vine.make(cfg=vcfg.compose(preset + [f"num_envs={n}", "headless=True"]))
g = torch.Generator(device="cuda").manual_seed(0)
for t in range(40):
    a = torch.rand(n, 2, device="cuda", generator=g) * 2 - 1
    env.step(a)
torch.cuda.synchronize()
print("ok", float(env.rew_buf.mean()))
