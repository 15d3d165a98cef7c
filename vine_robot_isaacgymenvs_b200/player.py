"""Deployment-side inference (SURVEY §8 "next" f4): what the reference's ``isaacgymenvs/vine_robot_test_model.py``
does with rl_games' ``PpoPlayerContinuous`` -- load a checkpoint, feed one observation, get one action, keep the
LSTM state between calls -- without rl_games.

``PolicyPlayer.get_action(obs, is_deterministic)``   == rl_games BasePlayer.get_action (vine_robot_test_model.py:112-131)
``VineRobotControlModel.get_action(q, qd, tip_pos, tip_vel, target_pos)`` == vine_robot_test_model.py:143-177

Checkpoints: the rl_games layout ``PPOAgent.state_dict`` writes (``model`` -> ``a2c_network.*``, ``running_mean_std.*``).
Runs wherever torch runs (the robot-side host may have no GPU); on a B200 the batched MLP policy can instead go through
``vine_policy_act``.
"""
import torch

from .ppo.ppo import ActorCritic, RunningMeanStd

REFERENCE_RNN = {"name": "lstm", "units": 256, "layers": 1, "before_mlp": False, "concat_input": True, "layer_norm": True}


class PolicyPlayer:
    def __init__(self, num_obs, num_actions=2, units=(256, 128, 64), rnn="auto", device="cpu", clip_actions=True):
        self.device, self.clip_actions, self._rnn_arg = torch.device(device), clip_actions, rnn
        self.num_obs, self.num_actions, self.units = num_obs, num_actions, tuple(units)
        self.model = self.obs_rms = self.states = None

    def _build(self, has_rnn):
        rnn = REFERENCE_RNN if has_rnn else None
        self.model = ActorCritic(self.num_obs, self.num_actions, self.units, rnn=rnn).to(self.device).eval()
        self.obs_rms = RunningMeanStd((self.num_obs,)).to(self.device)
        self.reset()

    def restore(self, fn_or_state):
        ck = torch.load(fn_or_state, map_location=self.device) if isinstance(fn_or_state, (str, bytes)) or hasattr(fn_or_state, "__fspath__") else fn_or_state
        model = ck["model"]
        has_rnn = any(k.startswith("a2c_network.rnn.") for k in model) if self._rnn_arg == "auto" else bool(self._rnn_arg)
        self._build(has_rnn)
        self.model.load_state_dict({k[len("a2c_network."):]: v for k, v in model.items() if k.startswith("a2c_network.")})
        rms = {k[len("running_mean_std."):]: v for k, v in model.items() if k.startswith("running_mean_std.")}
        if not rms and "running_mean_std" in ck:          # vine_robot_test_model.py:137-139
            rms = ck["running_mean_std"]
        if rms:
            self.obs_rms.load_state_dict({k: v.to(torch.float64) for k, v in rms.items()})
        return self

    def reset(self, batch=1):
        """Zero LSTM state (rl_games player.reset / init_rnn)."""
        if self.model is not None and self.model.has_rnn:
            H = self.model.rnn_units
            self.states = (torch.zeros(batch, H, device=self.device), torch.zeros(batch, H, device=self.device))
        else:
            self.states = None

    @torch.no_grad()
    def get_action(self, obs, is_deterministic=False):
        obs = torch.as_tensor(obs, dtype=torch.float32, device=self.device)
        squeeze = obs.dim() == 1
        obs = obs.reshape(-1, self.num_obs)
        if self.states is not None and self.states[0].shape[0] != obs.shape[0]:
            self.reset(obs.shape[0])
        mu, logstd, _, self.states = self.model(self.obs_rms(obs), self.states)
        action = mu if is_deterministic else mu + torch.exp(logstd) * torch.randn_like(mu)
        if self.clip_actions:
            action = torch.clamp(action, -1.0, 1.0)
        return action[0] if squeeze else action


class VineRobotControlModel(torch.nn.Module):
    """vine_robot_test_model.py:143-177: actions rescaled from [-1, 1] to the hardware ranges."""

    def __init__(self, checkpoint_path, x_range, u_range, num_obs, device="cpu", deterministic=False, **player_kw):
        """``deterministic=False`` is what the reference script does: its ``forward`` calls ``get_action(obs)`` whose
        ``is_determenistic`` defaults to False, i.e. it SAMPLES (vine_robot_test_model.py:112-131,170-171).  Pass True to act
        with the mean."""
        super().__init__()
        self.deterministic = bool(deterministic)
        self.rail_force_min, self.rail_force_max = x_range
        self.u_min, self.u_max = u_range
        self.player = PolicyPlayer(num_obs, device=device, **player_kw).restore(checkpoint_path)

    @staticmethod
    def rescale(x, low, high):
        return (x + 1) * (high - low) / 2 + low

    def forward(self, obs):
        return self.player.get_action(obs, is_deterministic=self.deterministic)

    def get_action(self, *parts):
        """``parts``: the observation pieces in the order the policy was trained on (the reference's script passes
        q, qd, tip_pos, tip_vel, target_pos); returns (rail command, pressure command) in hardware units."""
        obs = torch.cat([torch.as_tensor(p, dtype=torch.float32).reshape(-1) for p in parts])[None, ...]
        action = self.forward(obs)[0].clone()
        if action.numel() == 1:
            return self.rescale(action, self.u_min, self.u_max)
        action[0] = self.rescale(action[0], self.rail_force_min, self.rail_force_max)
        action[1] = self.rescale(action[1], self.u_min, self.u_max)
        return action
