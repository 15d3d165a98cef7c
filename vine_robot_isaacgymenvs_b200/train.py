"""``python -m vine_robot_isaacgymenvs_b200.train task=Vine5LinkMovingBase num_envs=4096 ...``

Same command-line surface as the reference's ``isaacgymenvs/train.py`` (Hydra overrides; keys of
cfg/config.yaml): ``task=``, ``num_envs=``, ``seed=``, ``max_iterations=``, ``checkpoint=``,
``test=True``, ``multi_gpu=True`` (under torchrun), any ``task.env.*`` / ``train.params.*``.
Checkpoints go to ``runs/<name>/nn/<name>.pth`` like rl_games' (train.py:148-163, YP:69-70).
"""
import json
import os
import sys
import time

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
    pc = cfg["train"]["params"]["config"]
    name = pc["name"]
    exp_dir = os.path.join("runs", name)
    out_dir = os.path.join(exp_dir, "nn")
    if cfg.get("checkpoint"):
        agent.load_state_dict(torch.load(cfg["checkpoint"], map_location=dev))
    if cfg.get("test"):   # rl_games player.run: deterministic actions unless the player config says otherwise
        det = bool(cfg["train"]["params"].get("config", {}).get("player", {}).get("deterministic", True))
        stats = agent.evaluate(int(cfg.get("play_steps", 400)), deterministic=det)
        if rank == 0:
            print(json.dumps({"mode": "play", "deterministic": det, **stats}))
        return stats
    if rank == 0:
        write_run_config(cfg, exp_dir)
    save_freq, save_best_after = int(pc.get("save_frequency", 0) or 0), int(pc.get("save_best_after", 0) or 0)
    best = {"reward": -float("inf")}

    def on_epoch(a):
        """rl_games A2CBase.train: `last_<name>_ep_<epoch>_rew_<r>.pth` every save_frequency epochs and `<name>.pth` whenever the
        mean reward improves after save_best_after epochs (YP:69-70)."""
        if rank != 0 or not (save_freq or save_best_after):
            return
        due = save_freq and a.epoch % save_freq == 0
        if not (due or a.epoch >= save_best_after):
            return
        e = a.ep_stats.tolist()                      # device -> host once per epoch, only when checkpointing is configured
        mean_rew = e[2] / e[0] if e[0] > 0 else -float("inf")
        os.makedirs(out_dir, exist_ok=True)
        if due:
            torch.save(a.state_dict(), os.path.join(out_dir, f"last_{name}_ep_{a.epoch}_rew_{mean_rew:.4g}.pth"))
        if a.epoch >= save_best_after and mean_rew > best["reward"]:
            best["reward"] = mean_rew
            sd = a.state_dict()
            sd["last_mean_rewards"] = mean_rew
            torch.save(sd, os.path.join(out_dir, name + ".pth"))

    # what rl_games' A2CBase writes per logged epoch (a2c_common.write_stats) + the reference's observer (train.py:145)
    from .utils.rlgames_utils import RLGPUAlgoObserver, ScalarWriter
    writer = ScalarWriter(os.path.join(exp_dir, "summaries")) if rank == 0 else ScalarWriter(None)
    agent.writer, agent.games_to_track = writer, int(pc.get("games_to_track", 100))
    observer = RLGPUAlgoObserver()
    observer.after_init(agent)
    t_start = time.perf_counter()

    def log(line, st=None):
        if rank != 0:
            return
        print(line)
        if st is not None:
            frame, ep, total = st["frames"], st["epoch"], time.perf_counter() - t_start
            for tag, key in (("losses/a_loss", "a_loss"), ("losses/c_loss", "c_loss"), ("info/kl", "kl"), ("info/last_lr", "lr"),
                             ("performance/step_inference_rl_update_fps", "fps_total")):
                writer.add_scalar(tag, st[key], frame)
            if st["episodes"] > 0:
                for suffix, step in (("frame", frame), ("iter", ep), ("time", total)):
                    writer.add_scalar(f"rewards/{suffix}", st["mean_return"], step)
                    writer.add_scalar(f"episode_lengths/{suffix}", st["mean_length"], step)
                writer.add_scalar("info/success_rate", st["success_rate"], frame)
            writer.add_scalar("info/epochs", ep, frame)
            observer.process_infos(env.extras if isinstance(env.extras, dict) else {}, None)
            observer.after_print_stats(frame, ep, total)

    hist = agent.train(int(pc["max_epochs"]), log=log,
                       on_epoch=on_epoch if (save_freq or save_best_after) else None)
    if rank == 0:
        os.makedirs(out_dir, exist_ok=True)
        torch.save(agent.state_dict(), os.path.join(out_dir, f"last_{name}_ep_{agent.epoch}.pth"))
        if not os.path.exists(os.path.join(out_dir, name + ".pth")):
            torch.save(agent.state_dict(), os.path.join(out_dir, name + ".pth"))
        with open(os.path.join(exp_dir, "history.json"), "w") as f:
            json.dump(hist, f)
    return hist


def write_run_config(cfg, exp_dir, time_str=None):
    """What the reference's train.py leaves beside the checkpoints (train.py:148-163): `config.yaml` (the composed config)
    and `<time>_rlg_config_dict.pkl` = the plain-dict `train` section with `params.config.features` removed -- the file
    isaacgymenvs/vine_robot_test_model.py:112-139 unpickles to rebuild the player."""
    import copy
    import pickle
    from datetime import datetime

    import yaml
    os.makedirs(exp_dir, exist_ok=True)
    with open(os.path.join(exp_dir, "config.yaml"), "w") as f:
        yaml.safe_dump(cfg, f, default_flow_style=False, sort_keys=False)
    rlg = copy.deepcopy(cfg["train"])
    if "params" in rlg and "config" in rlg["params"]:
        rlg["params"]["config"].pop("features", None)
    time_str = time_str or datetime.now().strftime("%Y-%m-%d_%H-%M-%S")   # train.py:68
    path = os.path.join(exp_dir, f"{time_str}_rlg_config_dict.pkl")
    with open(path, "wb") as f:
        pickle.dump(rlg, f)
    return path


if __name__ == "__main__":
    launch(sys.argv[1:])
