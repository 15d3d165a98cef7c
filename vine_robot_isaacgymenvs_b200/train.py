"""``python -m vine_robot_isaacgymenvs_b200.train task=Vine5LinkMovingBase num_envs=4096 ...``

Same command-line surface as the reference's ``isaacgymenvs/train.py`` (Hydra overrides; keys of
cfg/config.yaml): ``task=``, ``num_envs=``, ``seed=``, ``max_iterations=``, ``checkpoint=``,
``test=True``, ``multi_gpu=True`` (under torchrun), any ``task.env.*`` / ``train.params.*``.
Checkpoints go to ``runs/<name>/nn/<name>.pth`` like rl_games' (train.py:148-163, YP:69-70).
"""
import json
import os
import sys

import torch

from . import config as vcfg, distributed as vd, make
from .ppo.ppo import PPOAgent


def launch(overrides):
    cfg = vcfg.compose(overrides)
    rank, world, local_rank = vd.rank_world()
    multi = bool(cfg.get("multi_gpu", False)) and world > 1
    dev = f"cuda:{local_rank}" if multi else cfg["rl_device"]
    torch.cuda.set_device(dev)
    if multi:
        vd.init_from_env(device=dev)
        cfg["sim_device"] = cfg["rl_device"] = dev
    seed = int(cfg["seed"]) + (rank if multi else 0)          # train.py:78 offsets the seed by rank
    n = int(cfg["task"]["env"]["numEnvs"])
    env = make(cfg=cfg, seed=int(cfg["seed"]), multi_gpu=multi, global_env_offset=rank * n if multi else 0)
    agent = PPOAgent(env, cfg["train"], device=dev, seed=seed, use_graphs=bool(cfg.get("use_graphs", True)),
                     use_fused_policy=bool(cfg.get("use_fused_policy", True)),
                     use_fused_update=bool(cfg.get("use_fused_update", True)))
    name = cfg["train"]["params"]["config"]["name"]
    out_dir = os.path.join("runs", name, "nn")
    if cfg.get("checkpoint"):
        agent.load_state_dict(torch.load(cfg["checkpoint"], map_location=dev))
    if cfg.get("test"):
        steps = int(cfg.get("play_steps", 400))
        with torch.no_grad():
            for _ in range(steps // agent.T):
                agent.play_steps()
        stats = agent.pop_stats()
        if rank == 0:
            print(json.dumps({"mode": "play", **stats}))
        return stats
    hist = agent.train(int(cfg["train"]["params"]["config"]["max_epochs"]),
                       log=print if rank == 0 else None)
    if rank == 0:
        os.makedirs(out_dir, exist_ok=True)
        torch.save(agent.state_dict(), os.path.join(out_dir, name + ".pth"))
        with open(os.path.join("runs", name, "history.json"), "w") as f:
            json.dump(hist, f)
    return hist


if __name__ == "__main__":
    launch(sys.argv[1:])
