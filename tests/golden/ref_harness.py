"""Reference-driven harness (runs ONLY in the build container, where /root/reference exists).

Loads the reference's *unmodified* ``Vine5LinkMovingBase`` and ``VecTask`` from
``/root/reference`` under stub ``isaacgym`` / ``gym`` modules and drives them with a fake
``gym`` object:

  * the reference's own ``__init__``, ``step``, ``pre_physics_step``,
    ``compute_and_set_dof_actuation_force_tensor``, ``post_physics_step``, ``reset_idx``,
    ``compute_observations``, ``compute_reward`` (+ ``compute_reward_jit`` / ``compute_reset_jit``)
    run as they are;
  * ``gym.simulate`` (closed PhysX in the reference) is replaced by the oracle's dynamics, with
    Isaac Gym's tensor-API semantics: rigid-body / contact tensors change only on ``refresh_*``
    and reflect the state after the last ``simulate`` (so they are stale after ``reset_idx``);
  * the reference's RNG call sites (``torch.FloatTensor(..).uniform_``, ``torch.randn_like``)
    are fed from the same Philox streams the product uses, so whole steps are comparable.

Nothing from the reference is copied; this file only imports it.  Used by generate_golden.py.
"""
import ctypes as C
import importlib.util
import logging
import math
import os
import sys
import types
from unittest.mock import MagicMock

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

from oracle import oracle as O  # noqa: E402
from vine_robot_isaacgymenvs_b200 import abi, config as vcfg  # noqa: E402

REF_ROOT = "/root/reference"
REF_TASKS = os.path.join(REF_ROOT, "isaacgymenvs", "tasks")

VINE_BODIES = ["slider", "cart", "link_0", "link_1", "link_2", "link_3", "link_4", "tip"]
DOF_NAMES = ["slider_to_cart", "cart_to_link_0", "link_0_to_link_1", "link_1_to_link_2",
             "link_2_to_link_3", "link_3_to_link_4"]

_loaded = None


def load_reference():
    """Import V5 and VT from their real paths under stubbed third-party modules."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not os.path.isdir(REF_ROOT):
        raise RuntimeError("/root/reference is not available (the harness only runs in the build container)")
    if not hasattr(np, "Inf"):
        np.Inf = np.inf  # VT:102 uses the alias numpy 2 removed

    gymapi = MagicMock(name="gymapi")
    gymapi.DOF_MODE_EFFORT = 3
    gymtorch = MagicMock(name="gymtorch")
    gymtorch.wrap_tensor = lambda t: t
    gymtorch.unwrap_tensor = lambda t: t
    torch_utils = types.ModuleType("isaacgym.torch_utils")

    def to_torch(x, dtype=torch.float, device="cpu", requires_grad=False):
        return torch.tensor(np.asarray(x), dtype=dtype, device=device, requires_grad=requires_grad)

    def quat_from_angle_axis(angle, axis):
        theta = (angle / 2).unsqueeze(-1)
        xyz = axis / axis.norm(p=2, dim=-1, keepdim=True) * theta.sin()
        return torch.cat([xyz, theta.cos()], dim=-1)

    torch_utils.to_torch = to_torch
    torch_utils.quat_from_angle_axis = quat_from_angle_axis
    isaacgym = MagicMock(name="isaacgym")
    isaacgym.gymapi, isaacgym.gymtorch, isaacgym.torch_utils = gymapi, gymtorch, torch_utils
    stubs = {
        "isaacgym": isaacgym, "isaacgym.gymapi": gymapi, "isaacgym.gymtorch": gymtorch,
        "isaacgym.gymutil": MagicMock(name="gymutil"), "isaacgym.torch_utils": torch_utils,
        "gym": MagicMock(name="gym"), "gym.spaces": MagicMock(name="gym.spaces"),
        "isaacgymenvs": MagicMock(name="isaacgymenvs"), "isaacgymenvs.utils": MagicMock(),
        "isaacgymenvs.utils.dr_utils": MagicMock(), "wandb": MagicMock(name="wandb"),
    }
    sys.modules.update(stubs)
    # fake packages so that `from .base.vec_task import VecTask` (V5:40) resolves to the real file
    pkg = types.ModuleType("reftasks")
    pkg.__path__ = [REF_TASKS]
    base = types.ModuleType("reftasks.base")
    base.__path__ = [os.path.join(REF_TASKS, "base")]
    sys.modules["reftasks"], sys.modules["reftasks.base"] = pkg, base

    def _load(name, path):
        spec = importlib.util.spec_from_file_location(name, path)
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
        return mod

    vt = _load("reftasks.base.vec_task", os.path.join(REF_TASKS, "base", "vec_task.py"))
    v5 = _load("reftasks.Vine5LinkMovingBase", os.path.join(REF_TASKS, "Vine5LinkMovingBase.py"))
    logging.disable(logging.INFO)
    _loaded = (v5, vt, gymapi)
    return _loaded


class PhiloxFeed:
    """Feeds the reference's RNG call sites from the product's Philox streams."""

    def __init__(self, seed, num_envs, global_env_offset=0):
        self.seed, self.n, self.off = seed, num_envs, global_env_offset
        self.step = 0            # control step index (all envs in lockstep)
        self.sim_i = 0           # sim step within the control step
        self.reset_ids = None    # env ids of the reset_idx call in flight
        self.reset_k = 0
        self.reset_step_flag = 0
        self.last_scale = None   # [N,5,4] multipliers of the last dynamics draw

    def _uniform_ab(self, gid, site, step, block, a, b):
        u = O.uniform4(self.seed, gid, site, step, block)
        a32, b32 = np.float32(a), np.float32(b)
        rng = np.float32(b32 - a32)
        # fmaf(u, rng, a) evaluated exactly: the product of two f32 is exact in f64
        return np.array([np.float32(np.float64(x) * np.float64(rng) + np.float64(a32)) for x in u], np.float32)

    def float_tensor(self, *shape):
        feed = self

        class _FT:
            def uniform_(self_inner, a, b):
                if len(shape) == 3:      # V5:1054 dynamics scaling [N,5,20]
                    out = torch.ones(*shape, dtype=torch.float32)
                    sc = np.ones((feed.n, 5, 4), np.float32)
                    a32 = np.float32(a)
                    rng = np.float32(np.float32(b) - a32)
                    for e in range(feed.n):
                        u = O.dynamics_uniforms(feed.seed, feed.off + e, feed.step, feed.sim_i)[:20]
                        # fmaf(u, rng, a) evaluated exactly: a 16-bit u times an f32 is exact in f64
                        vals = np.array([np.float32(np.float64(x) * np.float64(rng) + np.float64(a32)) for x in u], np.float32)
                        sc[e] = vals.reshape(5, 4)
                    for j in range(5):
                        for term in range(4):
                            out[:, j, 5 * term + j] = torch.from_numpy(sc[:, j, term])
                    feed.last_scale = sc
                    feed.sim_i += 1
                    return out
                n = shape[0]             # reset_idx draws, V5:780-909
                assert feed.reset_ids is not None and n == len(feed.reset_ids)
                k = feed.reset_k
                feed.reset_k += 1
                vals = [feed._uniform_ab(feed.off + int(e), O.SITE_RESET, feed.step | feed.reset_step_flag,
                                         k // 4, a, b)[k % 4] for e in feed.reset_ids]
                return torch.tensor(np.array(vals, np.float32))
        return _FT()

    def randn_like(self, t):
        n, w = t.shape
        site = O.SITE_ACTION_NOISE if w == 2 else O.SITE_OBS_NOISE
        out = np.zeros((n, 4 * ((w + 3) // 4)), np.float32)
        for e in range(n):
            for blk in range((w + 3) // 4):
                out[e, 4 * blk:4 * blk + 4] = O.normal4(self.seed, self.off + e, site, self.step, blk)
        return torch.from_numpy(out[:, :w].copy())


class FakeGym:
    """Minimal Isaac Gym stand-in; dynamics = oracle_simulate (f64 by default)."""

    def __init__(self, gymapi, vine_cfg, num_envs, feed, use_f64=True):
        self.gymapi, self.vc, self.n, self.feed, self.use_f64 = gymapi, vine_cfg, num_envs, feed, use_f64
        self.task = None
        self.bodies = (["shelf", "shelf_link"] if vine_cfg.create_shelf else []) + \
                      (["base_link"] if vine_cfg.create_pipe else []) + VINE_BODIES
        self.nb = len(self.bodies)
        self.n_actors = 1 + int(vine_cfg.create_shelf) + int(vine_cfg.create_pipe)
        self._actor_counter = 0
        self._env_counter = 0
        n = num_envs
        self.dof_state = torch.zeros(n * 6, 2)
        self.root_state = torch.zeros(n * self.n_actors, 13)
        self.rb_state = torch.zeros(n * self.nb, 13)
        self.contact = torch.zeros(n * self.nb, 3)
        # caches = state after the last simulate; initial pose q = 0
        self.c_tip = np.zeros((n, 3), np.float32)
        self.c_tip[:, 1] = np.float32(-5 * 0.0885 * np.sin(3.1415))
        self.c_tip[:, 2] = np.float32(0.965 + 5 * 0.0885 * np.cos(3.1415))
        self.c_tipvel = np.zeros((n, 3), np.float32)
        self.c_cart_y = np.zeros(n, np.float32)
        self.c_cart_vy = np.zeros(n, np.float32)
        self.c_lip = np.zeros(n, np.float32)
        self.efforts = torch.zeros(n, 6)
        self.refresh_rigid_body_state_tensor(None)

    # ---- asset / actor creation (V5:378-556) ----
    def __getattr__(self, name):
        return lambda *a, **k: MagicMock(name=name)

    def load_asset(self, sim, root, file, options):
        return file

    def get_asset_dof_count(self, asset):
        return 6

    def get_asset_rigid_body_count(self, asset):
        return len(VINE_BODIES)

    def get_asset_dof_type(self, asset, i):
        return self.gymapi.DofType.DOF_TRANSLATION if i == 0 else self.gymapi.DofType.DOF_ROTATION

    def get_asset_dof_name(self, asset, i):
        return DOF_NAMES[i]

    def get_asset_dof_names(self, asset):
        return list(DOF_NAMES)

    def get_asset_dof_dict(self, asset):
        return {n: i for i, n in enumerate(DOF_NAMES)}

    def _dof_props(self):
        dt = np.dtype([("hasLimits", "?"), ("lower", "f4"), ("upper", "f4"), ("driveMode", "i4"),
                       ("velocity", "f4"), ("effort", "f4"), ("stiffness", "f4"), ("damping", "f4"),
                       ("friction", "f4"), ("armature", "f4")], align=True)
        p = np.zeros(6, dt)
        p["lower"][0], p["upper"][0] = self.vc.prismatic_lower, self.vc.prismatic_upper
        p["lower"][1:], p["upper"][1:] = self.vc.revolute_lower, self.vc.revolute_upper
        return p

    def get_asset_dof_properties(self, asset):
        return self._dof_props()

    def get_actor_dof_properties(self, env, handle):
        return self._dof_props()

    def set_actor_dof_properties(self, env, handle, props):
        self.applied_dof_props = props.copy()
        return True

    def get_actor_rigid_shape_properties(self, env, handle):
        return []

    def create_env(self, sim, lower, upper, per_row):
        self._env_counter += 1
        return self._env_counter - 1

    def create_actor(self, env, asset, pose, name, group=0, filter=0, segmentationId=0):
        return name

    def get_actor_index(self, env, handle, domain):
        i = self._actor_counter
        self._actor_counter += 1
        return i

    def find_actor_rigid_body_index(self, env, handle, name, domain):
        return self.bodies.index(name)

    # ---- tensor API (V5:299-318) ----
    def acquire_dof_state_tensor(self, sim):
        return self.dof_state

    def acquire_actor_root_state_tensor(self, sim):
        return self.root_state

    def acquire_rigid_body_state_tensor(self, sim):
        return self.rb_state

    def acquire_net_contact_force_tensor(self, sim):
        return self.contact

    def refresh_dof_state_tensor(self, sim):
        return True

    def refresh_actor_root_state_tensor(self, sim):
        return True

    def refresh_rigid_body_state_tensor(self, sim):
        rb = self.rb_state.view(self.n, self.nb, 13)
        tip, cart = self.bodies.index("tip"), self.bodies.index("cart")
        rb[:, tip, 0:3] = torch.from_numpy(self.c_tip)
        rb[:, tip, 7:10] = torch.from_numpy(self.c_tipvel)
        rb[:, cart, 1] = torch.from_numpy(self.c_cart_y)
        rb[:, cart, 2] = 0.975
        rb[:, cart, 8] = torch.from_numpy(self.c_cart_vy)
        return True

    def refresh_net_contact_force_tensor(self, sim):
        if self.vc.create_shelf:
            cf = self.contact.view(self.n, self.nb, 3)
            cf[:, self.bodies.index("shelf_link"), 1] = torch.from_numpy(self.c_lip)
        return True

    def set_dof_actuation_force_tensor(self, sim, t):
        self.efforts = t.detach().clone().float()
        return True

    def set_dof_state_tensor_indexed(self, sim, state, idx, n):
        return True  # dof_state is shared memory: nothing to copy; rigid bodies stay stale

    def set_actor_root_state_tensor_indexed(self, sim, state, idx, n):
        return True

    def simulate(self, sim):
        t = self.task
        n = self.n
        ds = self.dof_state.view(n, 6, 2)
        q = np.ascontiguousarray(ds[..., 0].numpy(), np.float32).copy()
        qd = np.ascontiguousarray(ds[..., 1].numpy(), np.float32).copy()
        use_smoothed = bool(t.cfg["env"]["USE_SMOOTHED_FPAM"])
        u_use = (t.smoothed_u_fpam if use_smoothed else t.u_fpam).detach().reshape(n).numpy().astype(np.float32).copy()
        arrays = {
            "dof_pos": q, "dof_vel": qd,
            "dof_efforts": np.ascontiguousarray(self.efforts.numpy(), np.float32),
            "dynamics_scaling": np.ascontiguousarray(self.feed.last_scale) if (t.vine_randomize and self.feed.last_scale is not None) else None,
            "u_fpam_to_use": u_use,
            "target_positions": np.ascontiguousarray(t.target_positions.numpy(), np.float32),
            "object_info": np.ascontiguousarray(t.object_info.numpy(), np.float32),
            "tip_positions": self.c_tip, "tip_velocities": self.c_tipvel, "shelf_contact_force": self.c_lip,
        }
        O.call_io("oracle_simulate", self.vc, n, abi.VineSimulateIO, arrays, int(self.use_f64))
        ds[..., 0] = torch.from_numpy(q)
        ds[..., 1] = torch.from_numpy(qd)
        self.c_cart_y[:] = q[:, 0]
        self.c_cart_vy[:] = qd[:, 0]
        return True


class ReferenceTask:
    """The reference's Vine5LinkMovingBase instance driven through the fake gym."""

    def __init__(self, task_cfg, seed=42, use_f64=True, global_env_offset=0):
        v5, vt, gymapi = load_reference()
        self.v5, self.vt = v5, vt
        self.task_cfg = task_cfg
        self.vc = vcfg.task_cfg_to_vine_config(task_cfg)
        n = int(task_cfg["env"]["numEnvs"])
        self.n = n
        self.feed = PhiloxFeed(seed, n, global_env_offset)
        self.gym = FakeGym(gymapi, self.vc, n, self.feed, use_f64)
        gymapi.acquire_gym = lambda: self.gym
        vt.EXISTING_SIM = None
        self._orig_ft, self._orig_randn = torch.FloatTensor, torch.randn_like
        cfg = {k: v for k, v in task_cfg.items()}
        cfg["env"] = dict(task_cfg["env"])
        cfg["env"]["CAPTURE_VIDEO"] = False
        cfg["sim"] = dict(task_cfg["sim"])
        cfg["sim"]["use_gpu_pipeline"] = False   # CPU tensors here; the step logic is identical
        self._patch()
        try:
            # the only draws in __init__ are sample_target_positions (V5:179): give them a reset context
            self.feed.reset_ids, self.feed.reset_k, self.feed.reset_step_flag = list(range(n)), 6, 0x40000000
            self.task = v5.Vine5LinkMovingBase(cfg=cfg, rl_device="cpu", sim_device="cpu", graphics_device_id=-1,
                                               headless=True, virtual_screen_capture=False, force_render=False)
        finally:
            self._unpatch()
        self.feed.reset_ids, self.feed.reset_step_flag = None, 0
        self.gym.task = self.task
        self.task.use_wandb = False
        # route reset_idx through a wrapper that tells the feed which envs are being reset
        ref_reset = self.task.reset_idx

        def reset_idx(env_ids):
            self.feed.reset_ids = [int(i) for i in env_ids]
            self.feed.reset_k = 0
            ref_reset(env_ids)
            self.feed.reset_ids = None
        self.task.reset_idx = reset_idx

    def _patch(self):
        torch.FloatTensor = self.feed.float_tensor
        torch.randn_like = self.feed.randn_like

    def _unpatch(self):
        torch.FloatTensor, torch.randn_like = self._orig_ft, self._orig_randn

    def step(self, actions):
        self.feed.sim_i = 0
        self._patch()
        try:
            out = self.task.step(torch.as_tensor(actions, dtype=torch.float32))
        finally:
            self._unpatch()
        self.feed.step += 1
        return out

    def reset_done_ids(self, env_ids):
        """reset_idx outside step (VT:412-427) for the given ids."""
        self.feed.reset_step_flag = 0x80000000
        self._patch()
        try:
            self.task.reset_idx(torch.as_tensor(env_ids, dtype=torch.long))
        finally:
            self._unpatch()
            self.feed.reset_step_flag = 0

    def snapshot(self):
        t = self.task
        f = lambda x: x.detach().cpu().numpy().copy()  # noqa: E731
        return {
            "obs_buf": f(t.obs_buf), "rew_buf": f(t.rew_buf), "reset_buf": f(t.reset_buf),
            "progress_buf": f(t.progress_buf), "timeout_buf": f(t.timeout_buf).astype(np.uint8),
            "obs_clamped": f(t.obs_dict["obs"]) if "obs" in t.obs_dict else f(t.obs_buf),
            "dof_pos": f(t.dof_pos), "dof_vel": f(t.dof_vel),
            "tip_positions": f(t.tip_positions), "tip_velocities": f(t.tip_velocities),
            "target_positions": f(t.target_positions), "object_info": f(t.object_info),
            "smoothed_u_fpam": f(t.smoothed_u_fpam).reshape(-1),
            "u_fpam": f(t.u_fpam).reshape(-1) if hasattr(t, "u_fpam") else np.zeros(self.n, np.float32),
            "u_rail_velocity": f(t.u_rail_velocity).reshape(-1) if hasattr(t, "u_rail_velocity") else np.zeros(self.n, np.float32),
            "prev_u_rail_velocity": f(t.prev_u_rail_velocity).reshape(-1),
            "rail_force": f(t.rail_force).reshape(-1) if hasattr(t, "rail_force") else np.zeros(self.n, np.float32),
            "prev_cart_vel": f(t.prev_cart_vel).reshape(-1), "prev_cart_vel_error": f(t.prev_cart_vel_error).reshape(-1),
            "aggregated_rew_buf": f(t.aggregated_rew_buf),
        }


def radians10():
    return math.radians(10)
