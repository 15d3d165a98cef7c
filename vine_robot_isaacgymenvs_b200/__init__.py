"""B200-native Vine5LinkMovingBase hot path.

``make`` mirrors ``isaacgymenvs.make`` (reference: isaacgymenvs/__init__.py:15-56).
"""
from . import abi, config  # noqa: F401


def make(seed=42, task="Vine5LinkMovingBase", num_envs=4096, sim_device="cuda:0", rl_device="cuda:0",
         graphics_device_id=-1, headless=True, multi_gpu=False, virtual_screen_capture=False,
         force_render=False, cfg=None, overrides=None, global_env_offset=0):
    """Create the env.  ``cfg``: a composed root dict (config.compose); else built from ``overrides``."""
    from .utils.rlgames_utils import get_rlgames_env_creator
    if cfg is None:
        ov = [f"task={task}", f"num_envs={num_envs}", f"sim_device={sim_device}", f"rl_device={rl_device}",
              f"seed={seed}"] + list(overrides or [])
        cfg = config.compose(ov)
    create = get_rlgames_env_creator(
        seed=cfg.get("seed", seed), task_config=cfg["task"], task_name=cfg["task"]["name"],
        sim_device=cfg.get("sim_device", sim_device), rl_device=cfg.get("rl_device", rl_device),
        graphics_device_id=graphics_device_id, headless=headless, multi_gpu=multi_gpu,
        virtual_screen_capture=virtual_screen_capture, force_render=force_render,
        global_env_offset=global_env_offset)
    return create()
