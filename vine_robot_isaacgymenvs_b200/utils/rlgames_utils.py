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


class ScalarWriter:
    """Stand-in for the tensorboardX ``SummaryWriter`` rl_games hands to the observer (``algo.writer``): same
    ``add_scalar(tag, value, step)`` call, rows appended to ``<dir>/scalars.jsonl`` (tensorboardX is not installed here;
    pass a real SummaryWriter instead where it is)."""

    def __init__(self, log_dir=None):
        import os
        self.rows, self._f = [], None
        if log_dir is not None:
            os.makedirs(log_dir, exist_ok=True)
            self._f = open(os.path.join(log_dir, "scalars.jsonl"), "a")

    def add_scalar(self, tag, value, step):
        import json
        row = {"tag": tag, "value": float(value), "step": float(step)}
        self.rows.append(row)
        if self._f is not None:
            self._f.write(json.dumps(row) + "\n")
            self._f.flush()


class RLGPUAlgoObserver:
    """rlgames_utils.py:93-148: logs env-provided stats beside the algorithm's.  Same methods and same tags
    (``Episode/<key>``, ``<key>/frame|iter|time``, ``scores/mean|iter|time``); ``algo`` needs ``writer``, ``device`` and
    ``games_to_track``.  For the vine task the env's ``extras`` carry only the ``time_outs`` vector (VT:372), so like in the
    reference nothing but the trainer's own scalars ends up in the log; an env that adds scalar infos gets them logged."""

    def __init__(self):
        pass

    def after_init(self, algo):
        self.algo = algo
        self.mean_scores = []                  # torch_ext.AverageMeter(1, games_to_track): a bounded running window
        self.ep_infos = []
        self.direct_info = {}
        self.writer = self.algo.writer

    def process_infos(self, infos, done_indices):
        import torch
        assert isinstance(infos, dict), "RLGPUAlgoObserver expects dict info"
        if "episode" in infos:
            self.ep_infos.append(infos["episode"])
        if len(infos) > 0:                     # allow direct logging from env: only scalars
            self.direct_info = {k: v for k, v in infos.items()
                                if isinstance(v, (float, int)) or (isinstance(v, torch.Tensor) and v.dim() == 0)}

    def after_clear_stats(self):
        self.mean_scores.clear()

    def after_print_stats(self, frame, epoch_num, total_time):
        import torch
        if self.ep_infos:
            for key in self.ep_infos[0]:
                vals = [torch.as_tensor(ep[key], dtype=torch.float32).reshape(-1).to(self.algo.device) for ep in self.ep_infos]
                self.writer.add_scalar("Episode/" + key, torch.mean(torch.cat(vals)), epoch_num)
            self.ep_infos.clear()
        for k, v in self.direct_info.items():
            self.writer.add_scalar(f"{k}/frame", v, frame)
            self.writer.add_scalar(f"{k}/iter", v, epoch_num)
            self.writer.add_scalar(f"{k}/time", v, total_time)
        if self.mean_scores:
            window = self.mean_scores[-int(getattr(self.algo, "games_to_track", 100)):]
            mean_scores = sum(window) / len(window)
            self.writer.add_scalar("scores/mean", mean_scores, frame)
            self.writer.add_scalar("scores/iter", mean_scores, epoch_num)
            self.writer.add_scalar("scores/time", mean_scores, total_time)
