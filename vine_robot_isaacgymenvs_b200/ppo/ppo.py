"""PPO for the vine task without rl_games (SURVEY §8 row a14 and "next" f1).

Restates what rl-games 1.5.2's ``A2CAgent`` does for ``cfg/train/Vine5LinkMovingBasePPO.yaml``
(``a2c_continuous`` / ``continuous_a2c_logstd`` / ``actor_critic``): horizon-16 rollouts, GAE
(``vine_gae`` CUDA kernel), running mean/std of observations and values, advantage normalisation,
clipped actor loss, clipped critic loss x critic_coef, bound loss, Adam with the ``legacy``
adaptive-KL learning-rate schedule, value bootstrap on ``time_outs``, reward shaper.
In-repo analogue of the same math: isaacgymenvs/learning/common_agent.py:257-314 (rollout),
:319-435 (update), :482-517 (losses).  rl_games itself is not vendored in the reference
(setup.py:22) and not installed here: PARITY UNPINNED, checked by learning curves only.

Multi-GPU: one process per GPU, envs sharded; per minibatch ONE all-reduce carrying the flattened
gradients plus the KL scalar; per epoch one all-reduce of the running-statistics moments.
"""
import ctypes as C
import math
import time

import torch
import torch.nn as nn

from .. import abi, distributed as vd


class RunningMeanStd(nn.Module):
    """rl_games RunningMeanStd (per-feature, eps 1e-5, clamp +-5 on normalised output)."""

    def __init__(self, shape, eps=1e-5):
        super().__init__()
        self.eps = eps
        self.register_buffer("running_mean", torch.zeros(shape, dtype=torch.float64))
        self.register_buffer("running_var", torch.ones(shape, dtype=torch.float64))
        self.register_buffer("count", torch.ones((), dtype=torch.float64))

    @torch.no_grad()
    def update(self, x):
        x = x.reshape(-1, *self.running_mean.shape) if self.running_mean.dim() else x.reshape(-1)
        n = x.shape[0]
        mean = x.mean(0)
        m2 = ((x - mean) ** 2).sum(0)
        n, mean, m2 = vd.merge_moments(n, mean, m2)
        n = float(n)
        var = m2.double() / max(n - 1.0, 1.0)
        mean = mean.double()
        delta = mean - self.running_mean
        tot = self.count + n
        new_mean = self.running_mean + delta * n / tot
        m_a = self.running_var * self.count
        m_b = var * n
        self.running_var.copy_((m_a + m_b + delta ** 2 * self.count * n / tot) / tot)
        self.running_mean.copy_(new_mean)
        self.count.copy_(tot)

    def forward(self, x, unnorm=False):
        mean = self.running_mean.float()
        std = torch.sqrt(self.running_var.float() + self.eps)
        if unnorm:
            return torch.clamp(x, -5.0, 5.0) * std + mean
        return torch.clamp((x - mean) / std, -5.0, 5.0)


class ActorCritic(nn.Module):
    """``actor_critic`` network of Vine5LinkMovingBasePPO.yaml:10-30: shared MLP [256,128,64] ELU,
    mu head, value head, state-independent learnable log-std initialised to 0 (fixed_sigma: True)."""

    def __init__(self, num_obs, num_actions, units=(256, 128, 64)):
        super().__init__()
        layers, d = [], num_obs
        for u in units:
            layers += [nn.Linear(d, u), nn.ELU()]
            d = u
        self.mlp = nn.Sequential(*layers)
        self.mu = nn.Linear(d, num_actions)
        self.value = nn.Linear(d, 1)
        self.sigma = nn.Parameter(torch.zeros(num_actions))

    def forward(self, obs):
        h = self.mlp(obs)
        return self.mu(h), self.sigma.expand(obs.shape[0], -1), self.value(h)


def neglogp(x, mean, std, logstd):
    return 0.5 * (((x - mean) / std) ** 2).sum(-1) + 0.5 * math.log(2.0 * math.pi) * x.shape[-1] + logstd.sum(-1)


def policy_kl(p0_mu, p0_sigma, p1_mu, p1_sigma):
    c1 = torch.log(p1_sigma / p0_sigma + 1e-5)
    c2 = (p0_sigma ** 2 + (p1_mu - p0_mu) ** 2) / (2.0 * (p1_sigma ** 2 + 1e-5))
    return (c1 + c2 - 0.5).sum(-1).mean()


class PPOAgent:
    def __init__(self, env, train_cfg, device=None, seed=42):
        c = train_cfg["params"]["config"]
        self.env, self.c = env, c
        self.device = device or env.device
        self.n, self.T = env.num_envs, int(c["horizon_length"])
        self.O, self.A = env.num_obs, env.num_acts
        self.gamma, self.tau = float(c["gamma"]), float(c["tau"])
        self.e_clip, self.critic_coef = float(c["e_clip"]), float(c["critic_coef"])
        self.entropy_coef, self.bounds_coef = float(c["entropy_coef"]), float(c["bounds_loss_coef"])
        self.mini_epochs = int(c["mini_epochs"])
        self.batch = self.n * self.T
        self.minibatch = min(int(c["minibatch_size"]), self.batch)
        assert self.batch % self.minibatch == 0, "batch must be a multiple of minibatch_size"
        self.lr, self.kl_threshold = float(c["learning_rate"]), float(c["kl_threshold"])
        self.adaptive = c.get("lr_schedule") == "adaptive"
        self.reward_scale = float(c.get("reward_shaper", {}).get("scale_value", 1.0))
        self.value_bootstrap = bool(c.get("value_bootstrap", False))
        self.normalize_input, self.normalize_value = bool(c["normalize_input"]), bool(c["normalize_value"])
        self.normalize_advantage = bool(c["normalize_advantage"])
        self.truncate_grads, self.grad_norm = bool(c.get("truncate_grads", False)), float(c.get("grad_norm", 1.0))
        self.bf16 = bool(c.get("mixed_precision", False))
        units = train_cfg["params"]["network"]["mlp"]["units"]
        torch.manual_seed(seed)
        self.model = ActorCritic(self.O, self.A, units).to(self.device)
        self.obs_rms = RunningMeanStd((self.O,)).to(self.device)
        self.val_rms = RunningMeanStd(()).to(self.device)
        self.world = vd.rank_world()[1]
        if self.world > 1:  # hvd.setup_algo equivalent: identical parameters on every rank
            for p in self.model.parameters():
                torch.distributed.broadcast(p.data, 0)
        self.opt = torch.optim.Adam(self.model.parameters(), lr=self.lr, eps=1e-8)
        self._lib = abi.load_library()
        T, n, dev = self.T, self.n, self.device
        f = lambda *s: torch.zeros(*s, device=dev)  # noqa: E731
        self.b_obs, self.b_act, self.b_mu = f(T, n, self.O), f(T, n, self.A), f(T, n, self.A)
        self.b_nlp, self.b_val, self.b_rew, self.b_done = f(T, n), f(T, n), f(T, n), f(T, n)
        self.b_adv, self.b_ret = f(T, n), f(T, n)
        self.obs = env.reset()["obs"].clone()
        self.dones = torch.ones(n, device=dev)
        self.epoch = 0
        self.frames = 0
        self.stats = {"episodes": 0, "successes": 0, "return_sum": 0.0, "len_sum": 0}
        self.ep_ret = torch.zeros(n, device=dev)
        self.ep_len = torch.zeros(n, device=dev)

    # ------------------------------------------------------------------ acting
    @torch.no_grad()
    def _policy(self, obs):
        x = self.obs_rms(obs) if self.normalize_input else obs
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=self.bf16):
            mu, logstd, value = self.model(x)
        mu, logstd, value = mu.float(), logstd.float(), value.float().squeeze(-1)
        if self.normalize_value:
            value = self.val_rms(value, unnorm=True)
        return mu, logstd, value

    @torch.no_grad()
    def play_steps(self):
        env = self.env
        for t in range(self.T):
            mu, logstd, value = self._policy(self.obs)
            sigma = torch.exp(logstd)
            action = mu + sigma * torch.randn_like(mu)
            self.b_obs[t], self.b_act[t], self.b_mu[t] = self.obs, action, mu
            self.b_nlp[t], self.b_val[t], self.b_done[t] = neglogp(action, mu, sigma, logstd), value, self.dones
            od, rew, dones, infos = env.step(torch.clamp(action, -1.0, 1.0))   # rl_games preprocess_actions
            shaped = rew * self.reward_scale
            if self.value_bootstrap:                                           # Vine5LinkMovingBasePPO.yaml:56
                shaped = shaped + self.gamma * value * infos["time_outs"].float()
            self.b_rew[t] = shaped
            # episode statistics: success == the 1000-point "Position Success" term fired (V5:1507)
            self.ep_ret += rew
            self.ep_len += 1
            d = dones.bool()
            if bool(d.any()):
                self.stats["episodes"] += int(d.sum())
                self.stats["successes"] += int((d & (rew > 500.0)).sum())
                self.stats["return_sum"] += float(self.ep_ret[d].sum())
                self.stats["len_sum"] += int(self.ep_len[d].sum())
                self.ep_ret[d] = 0
                self.ep_len[d] = 0
            self.obs = od["obs"].clone()
            self.dones = dones.float()
        _, _, last_value = self._policy(self.obs)
        p = lambda x: C.c_void_p(x.data_ptr())  # noqa: E731
        rc = self._lib.vine_gae(p(self.b_rew), p(self.b_val), p(self.b_done), p(last_value.contiguous()),
                                p(self.dones), self.T, self.n, self.gamma, self.tau, p(self.b_adv), p(self.b_ret),
                                C.c_void_p(torch.cuda.current_stream().cuda_stream))
        assert rc == 0
        self.frames += self.T * self.n * self.world

    # ------------------------------------------------------------------ learning
    def _update_lr(self, kl):
        if not self.adaptive:
            return
        if kl > 2.0 * self.kl_threshold:
            self.lr = max(self.lr / 1.5, 1e-6)
        if kl < 0.5 * self.kl_threshold:
            self.lr = min(self.lr * 1.5, 1e-2)
        for g in self.opt.param_groups:
            g["lr"] = self.lr

    def train_epoch(self):
        self.play_steps()
        B = self.batch
        obs, act, mu_old = self.b_obs.view(B, self.O), self.b_act.view(B, self.A), self.b_mu.view(B, self.A)
        nlp_old, val_old, ret = self.b_nlp.view(B), self.b_val.view(B), self.b_ret.view(B)
        adv = ret - val_old
        if self.normalize_input:
            self.obs_rms.update(obs)
            obs = self.obs_rms(obs)
        if self.normalize_value:
            self.val_rms.update(torch.cat([val_old, ret]))
            val_old, ret = self.val_rms(val_old), self.val_rms(ret)
        if self.normalize_advantage:
            s = torch.stack([adv.sum(), (adv * adv).sum(), torch.tensor(float(B), device=adv.device)]).double()
            if self.world > 1:
                torch.distributed.all_reduce(s)
            mean = s[0] / s[2]
            std = torch.sqrt(torch.clamp((s[1] - s[2] * mean * mean) / (s[2] - 1.0), min=0.0))
            adv = (adv - mean.float()) / (std.float() + 1e-8)
        sigma_old = torch.exp(self.model.sigma.detach()).expand(B, -1).clone()
        params = [p for p in self.model.parameters()]
        info = {"a_loss": 0.0, "c_loss": 0.0, "kl": 0.0}
        nmb = 0
        for _ in range(self.mini_epochs):
            for i in range(0, B, self.minibatch):
                sl = slice(i, i + self.minibatch)
                with torch.autocast("cuda", dtype=torch.bfloat16, enabled=self.bf16):
                    mu, logstd, value = self.model(obs[sl])
                mu, logstd, value = mu.float(), logstd.float(), value.float().squeeze(-1)
                sigma = torch.exp(logstd)
                nlp = neglogp(act[sl], mu, sigma, logstd)
                ratio = torch.exp(nlp_old[sl] - nlp)
                a = adv[sl]
                a_loss = torch.max(-a * ratio, -a * torch.clamp(ratio, 1.0 - self.e_clip, 1.0 + self.e_clip)).mean()
                v_clip = val_old[sl] + (value - val_old[sl]).clamp(-self.e_clip, self.e_clip)
                c_loss = torch.max((value - ret[sl]) ** 2, (v_clip - ret[sl]) ** 2).mean()
                b_loss = (torch.clamp_min(mu - 1.1, 0.0) ** 2 + torch.clamp_max(mu + 1.1, 0.0) ** 2).sum(-1).mean()
                entropy = (0.5 + 0.5 * math.log(2 * math.pi) + logstd).sum(-1).mean()
                loss = a_loss + 0.5 * c_loss * self.critic_coef - self.entropy_coef * entropy + b_loss * self.bounds_coef
                self.opt.zero_grad(set_to_none=True)
                loss.backward()
                with torch.no_grad():
                    kl = policy_kl(mu.detach(), sigma.detach(), mu_old[sl], sigma_old[sl])
                    if self.world > 1:  # ONE collective per minibatch: gradients + KL
                        flat = torch.cat([p.grad.reshape(-1) for p in params])
                        flat, extra = vd.allreduce_mean_(flat, kl.reshape(1))
                        kl = extra[0]
                        o = 0
                        for p in params:
                            p.grad.copy_(flat[o:o + p.numel()].view_as(p))
                            o += p.numel()
                    if self.truncate_grads:
                        nn.utils.clip_grad_norm_(params, self.grad_norm)
                self.opt.step()
                klv = float(kl)
                self._update_lr(klv)
                info["a_loss"] += float(a_loss); info["c_loss"] += float(c_loss); info["kl"] += klv
                nmb += 1
        self.epoch += 1
        return {k: v / nmb for k, v in info.items()}

    def pop_stats(self):
        s = self.stats
        eps = max(s["episodes"], 1)
        out = {"episodes": s["episodes"], "success_rate": s["successes"] / eps, "mean_return": s["return_sum"] / eps,
               "mean_length": s["len_sum"] / eps}
        self.stats = {"episodes": 0, "successes": 0, "return_sum": 0.0, "len_sum": 0}
        return out

    def train(self, max_epochs, log_every=25, log=print):
        t0 = time.perf_counter()
        hist = []
        for ep in range(max_epochs):
            info = self.train_epoch()
            if (ep + 1) % log_every == 0 or ep == max_epochs - 1:
                torch.cuda.synchronize()
                st = self.pop_stats()
                st.update(info)
                st.update({"epoch": ep + 1, "frames": self.frames, "lr": self.lr,
                           "fps_total": self.frames / (time.perf_counter() - t0)})
                hist.append(st)
                if log:
                    log(" ".join(f"{k}={v:.4g}" if isinstance(v, float) else f"{k}={v}" for k, v in st.items()))
        return hist

    # checkpoint layout follows what the reference's deployment loader reads
    # (isaacgymenvs/vine_robot_test_model.py:135-139: keys 'model' and 'running_mean_std')
    def state_dict(self):
        return {"model": self.model.state_dict(), "running_mean_std": self.obs_rms.state_dict(),
                "reward_mean_std": self.val_rms.state_dict(), "optimizer": self.opt.state_dict(),
                "epoch": self.epoch, "frame": self.frames, "last_lr": self.lr}

    def load_state_dict(self, sd):
        self.model.load_state_dict(sd["model"])
        self.obs_rms.load_state_dict(sd["running_mean_std"])
        self.val_rms.load_state_dict(sd["reward_mean_std"])
        if "optimizer" in sd:
            self.opt.load_state_dict(sd["optimizer"])
        self.epoch, self.frames, self.lr = sd.get("epoch", 0), sd.get("frame", 0), sd.get("last_lr", self.lr)
