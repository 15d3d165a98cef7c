"""The boundary rl_games talks to (reference: isaacgymenvs/utils/rlgames_utils.py:72-92, 151-180).

``get_rlgames_env_creator`` builds the task exactly like the reference does; ``RLGPUEnv`` is the
``IVecEnv`` adapter (step / reset / reset_done / get_env_info).  rl_games itself is optional: when it
is importable the adapter subclasses ``vecenv.IVecEnv`` so ``vecenv.register('RLGPU', ...)`` works
unchanged (train.py:122-127); otherwise it is a plain object with the same methods.
"""
from ..tasks import isaacgym_task_map

try:  # pragma: no cover - rl_games is not installed in the build image
    from rl_games.common import vecenv as _vecenv
    _Base = _vecenv.IVecEnv
except Exception:
    _vecenv = None
    _Base = object


def get_rlgames_env_creator(seed, task_config, task_name, sim_device, rl_device, graphics_device_id, headless,
                            multi_gpu=False, post_create_hook=None, virtual_screen_capture=False,
                            force_render=False, global_env_offset=0):
    """Same parameters and meaning as rlgames_utils.py:36-71; multi-GPU picks cuda:{LOCAL_RANK}
    (the reference does this through Horovod, rlgames_utils.py:58-70)."""
    def create_rlgpu_env(_sim_device=sim_device, _rl_device=rl_device, **kwargs):
        if multi_gpu:
            import os
            rank = int(os.getenv("LOCAL_RANK", "0"))
            _sim_device = _rl_device = f"cuda:{rank}"
            task_config["rank"] = rank
            task_config["rl_device"] = _rl_device
        task_config["seed"] = seed
        env = isaacgym_task_map[task_name](
            cfg=task_config, rl_device=_rl_device, sim_device=_sim_device,
            graphics_device_id=graphics_device_id, headless=headless,
            virtual_screen_capture=virtual_screen_capture, force_render=force_render,
            global_env_offset=global_env_offset)
        if post_create_hook is not None:
            post_create_hook()
        return env
    return create_rlgpu_env


class RLGPUEnv(_Base):
    """rlgames_utils.py:151-180."""

    def __init__(self, config_name, num_actors, env_creator=None, **kwargs):
        if env_creator is None:  # reference path: env_configurations.configurations[config_name]['env_creator']
            from rl_games.common import env_configurations
            env_creator = env_configurations.configurations[config_name]["env_creator"]
        self.env = env_creator(**kwargs)

    def step(self, actions):
        return self.env.step(actions)

    def reset(self):
        return self.env.reset()

    def reset_done(self):
        return self.env.reset_done()

    def get_number_of_agents(self):
        return getattr(self.env, "num_agents", 1)

    def get_env_info(self):
        info = {"action_space": self.env.action_space, "observation_space": self.env.observation_space}
        if hasattr(self.env, "amp_observation_space"):
            info["amp_observation_space"] = self.env.amp_observation_space
        if self.env.num_states > 0:
            info["state_space"] = self.env.state_space
        return info
