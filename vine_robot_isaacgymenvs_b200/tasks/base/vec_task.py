"""``Env`` / ``VecTask`` — the env contract of the reference (isaacgymenvs/tasks/base/vec_task.py)
re-stated without Isaac Gym: same constructor arguments, attribute names, dtypes and
``step / reset / reset_done`` semantics (VT:61-108, 169-223, 260-283, 319-427), so rl_games'
``RLGPUEnv`` adapter (utils/rlgames_utils.py:151-180) drives it unchanged.

What differs by construction: there is no ``gym``/``sim`` object; ``create_sim`` builds the native
``VineEnv`` handle and the physics + task logic of one control step is ONE fused CUDA launch.
"""
import abc
from typing import Any, Dict, Tuple

import numpy as np
import torch

try:  # gym is optional (not installed in the build image); rl_games only reads .shape/.low/.high
    from gym import spaces as _spaces  # type: ignore
    Box = _spaces.Box
except Exception:  # pragma: no cover - exercised when gym is absent
    class Box:  # minimal stand-in for gym.spaces.Box
        def __init__(self, low, high, shape=None, dtype=np.float32):
            low = np.asarray(low, dtype=dtype)
            high = np.asarray(high, dtype=dtype)
            if shape is not None:
                low = np.broadcast_to(low, shape).copy()
                high = np.broadcast_to(high, shape).copy()
            self.low, self.high, self.shape, self.dtype = low, high, low.shape, np.dtype(dtype)

        def __repr__(self):
            return f"Box({self.low.min()}, {self.high.max()}, {self.shape}, {self.dtype})"


class Env(abc.ABC):
    def __init__(self, config: Dict[str, Any], rl_device: str, sim_device: str, graphics_device_id: int,
                 headless: bool):
        """VT:61-108."""
        split_device = sim_device.split(":")
        self.device_type = split_device[0]
        self.device_id = int(split_device[1]) if len(split_device) > 1 else 0
        if self.device_type.lower() not in ("cuda", "gpu"):
            raise RuntimeError(
                f"sim_device={sim_device!r}: the B200-native Vine5LinkMovingBase path is CUDA-only and has "
                "no CPU pipeline (the reference's sim_device=cpu/pipeline=cpu is only a measured baseline).")
        config["sim"]["use_gpu_pipeline"] = True
        self.device = "cuda:" + str(self.device_id)
        self.rl_device = rl_device
        self.headless = headless
        self.graphics_device_id = graphics_device_id
        if not config.get("enableCameraSensors", False) and self.headless:
            self.graphics_device_id = -1
        self.num_environments = config["env"]["numEnvs"]
        self.num_agents = config["env"].get("numAgents", 1)
        self.num_observations = config["env"]["numObservations"]
        self.num_states = config["env"].get("numStates", 0)
        self.num_actions = config["env"]["numActions"]
        self.control_freq_inv = config["env"].get("controlFrequencyInv", 1)
        self.obs_space = Box(np.ones(self.num_obs) * -np.inf, np.ones(self.num_obs) * np.inf)
        self.state_space = Box(np.ones(self.num_states) * -np.inf, np.ones(self.num_states) * np.inf)
        self.act_space = Box(np.ones(self.num_actions) * -1., np.ones(self.num_actions) * 1.)
        self.clip_obs = config["env"].get("clipObservations", np.inf)
        self.clip_actions = config["env"].get("clipActions", np.inf)

    @abc.abstractmethod
    def allocate_buffers(self):
        ...

    @abc.abstractmethod
    def step(self, actions: torch.Tensor) -> Tuple[Dict[str, torch.Tensor], torch.Tensor, torch.Tensor, Dict[str, Any]]:
        ...

    @abc.abstractmethod
    def reset(self) -> Dict[str, torch.Tensor]:
        ...

    @abc.abstractmethod
    def reset_idx(self, env_ids: torch.Tensor):
        ...

    @property
    def observation_space(self):
        return self.obs_space

    @property
    def action_space(self):
        return self.act_space

    @property
    def num_envs(self) -> int:
        return self.num_environments

    @property
    def num_acts(self) -> int:
        return self.num_actions

    @property
    def num_obs(self) -> int:
        return self.num_observations


class VecTask(Env):
    metadata = {"render.modes": ["human", "rgb_array"], "video.frames_per_second": 24}

    def __init__(self, config, rl_device, sim_device, graphics_device_id, headless,
                 virtual_screen_capture: bool = False, force_render: bool = False):
        """VT:169-223 (viewer, virtual display and the generic DR engine are out of scope)."""
        super().__init__(config, rl_device, sim_device, graphics_device_id, headless)
        self.virtual_screen_capture = virtual_screen_capture
        self.virtual_display = None
        self.force_render = force_render
        if self.cfg["physics_engine"] not in ("physx", "flex"):
            raise ValueError(f"Invalid physics engine backend: {self.cfg['physics_engine']}")
        self.physics_engine = self.cfg["physics_engine"]
        self.first_randomization = True
        self.dr_randomizations = {}
        self.sim_initialized = False
        self.viewer = None
        self.enable_viewer_sync = self.cfg["sim"].get("enable_viewer_sync_at_start", True)
        self.allocate_buffers()
        self.create_sim()
        self.sim_initialized = True
        self.obs_dict = {}

    def allocate_buffers(self):
        """VT:260-283: same names, shapes and dtypes."""
        n, dev = self.num_envs, self.device
        self.obs_buf = torch.zeros((n, self.num_obs), device=dev, dtype=torch.float)
        self.states_buf = torch.zeros((n, self.num_states), device=dev, dtype=torch.float)
        self.rew_buf = torch.zeros(n, device=dev, dtype=torch.float)
        self.reset_buf = torch.ones(n, device=dev, dtype=torch.long)
        self.timeout_buf = torch.zeros(n, device=dev, dtype=torch.bool)  # bool from the first step on (VT:366)
        self.progress_buf = torch.zeros(n, device=dev, dtype=torch.long)
        self.randomize_buf = torch.zeros(n, device=dev, dtype=torch.long)
        self.extras = {}

    @abc.abstractmethod
    def create_sim(self):
        ...

    def get_state(self):
        """VT:303-305."""
        return torch.clamp(self.states_buf, -self.clip_obs, self.clip_obs).to(self.rl_device)

    @abc.abstractmethod
    def pre_physics_step(self, actions: torch.Tensor):
        ...

    @abc.abstractmethod
    def post_physics_step(self):
        ...

    def zero_actions(self) -> torch.Tensor:
        """VT:382-390."""
        return torch.zeros([self.num_envs, self.num_actions], dtype=torch.float32, device=self.rl_device)

    def reset(self):
        """VT:398-410: called once; returns the still-zero observation buffer, clamped."""
        self.obs_dict["obs"] = torch.clamp(self.obs_buf, -self.clip_obs, self.clip_obs).to(self.rl_device)
        if self.num_states > 0:
            self.obs_dict["states"] = self.get_state()
        return self.obs_dict

    def reset_done(self):
        """VT:412-427."""
        done_env_ids = self.reset_buf.nonzero(as_tuple=False).flatten()
        if len(done_env_ids) > 0:
            self.reset_idx(done_env_ids)
        self.obs_dict["obs"] = torch.clamp(self.obs_buf, -self.clip_obs, self.clip_obs).to(self.rl_device)
        if self.num_states > 0:
            self.obs_dict["states"] = self.get_state()
        return self.obs_dict, done_env_ids

    def render(self, mode="rgb_array"):
        """Headless only: the viewer (VT:429-466) is out of scope."""
        return None
