"""``Vine5LinkMovingBase`` — drop-in for the reference task class
(isaacgymenvs/tasks/Vine5LinkMovingBase.py, "V5") whose whole per-control-step pipeline runs in
one fused sm_100a CUDA kernel behind the C ABI of ``include/vine_b200.h``.

Same constructor signature as the reference (``utils/rlgames_utils.py:78-86`` builds it as
``Task(cfg=..., rl_device=..., sim_device=..., graphics_device_id=..., headless=...,
virtual_screen_capture=..., force_render=...)``), same ``cfg`` key set (V5:146-291), same buffers.
Viewer, video capture, wandb logging, keyboard events and .mat replay (V5:593-771, 947-982,
1122-1216) are outside the hot path and not provided.
"""
import ctypes as C
from enum import Enum

import torch

from .. import abi, config as vcfg
from .base.vec_task import VecTask

NUM_XYZ = 3
NUM_OBJECT_INFO = 2       # V5:51
N_REVOLUTE_DOFS = 5       # V5:54
N_PRESSURE_ACTIONS = 1    # V5:55
N_PRISMATIC_DOFS = 1      # V5:83
REWARD_NAMES = abi.REWARD_NAMES


class ObservationType(Enum):  # V5:67-73
    POS_ONLY = "POS_ONLY"
    POS_AND_VEL = "POS_AND_VEL"
    POS_AND_FD_VEL = "POS_AND_FD_VEL"
    POS_AND_PREV_POS = "POS_AND_PREV_POS"
    POS_AND_FD_VEL_AND_OBJ_INFO = "POS_AND_FD_VEL_AND_OBJ_INFO"
    TIP_AND_CART_AND_OBJ_INFO = "TIP_AND_CART_AND_OBJ_INFO"


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


class Vine5LinkMovingBase(VecTask):
    def __init__(self, cfg, rl_device, sim_device, graphics_device_id, headless,
                 virtual_screen_capture=False, force_render=False, *, global_env_offset=0):
        self.cfg = cfg
        self.max_episode_length = self.cfg["env"]["maxEpisodeLength"]
        self.vine_randomize = self.cfg["task"]["vine_randomize"]
        # numObservations / numActions follow OBSERVATION_TYPE (V5:152-171); KeyError like the reference
        observation_type = ObservationType[self.cfg["env"]["OBSERVATION_TYPE"]]
        self.cfg["env"]["numObservations"] = vcfg.num_observations(observation_type.value)
        self.cfg["env"]["numActions"] = N_PRESSURE_ACTIONS + N_PRISMATIC_DOFS
        if self.cfg["env"]["SCALE_OBSERVATIONS"] and observation_type not in (
                ObservationType.POS_AND_FD_VEL_AND_OBJ_INFO, ObservationType.TIP_AND_CART_AND_OBJ_INFO):
            raise NotImplementedError(f"Observation scaling not implemented for {observation_type}")  # V5:267-268
        self._seed = int(self.cfg.get("seed", 42))
        self._global_env_offset = int(global_env_offset)
        self._lib = abi.load_library()   # raises if the CUDA library is missing: there is no fallback
        self._h = None
        self._graph = None

        super().__init__(config=self.cfg, rl_device=rl_device, sim_device=sim_device,
                         graphics_device_id=graphics_device_id, headless=headless,
                         virtual_screen_capture=virtual_screen_capture, force_render=force_render)

        self.num_dof = N_REVOLUTE_DOFS + N_PRISMATIC_DOFS
        self.dt = self.cfg["sim"]["dt"]                       # V5:227
        self.control_dt = self.dt * self.control_freq_inv     # V5:228
        self.reward_weights = torch.tensor([[self.cfg["env"][k] for k in abi.REWARD_WEIGHT_KEYS]],
                                           device=self.device, dtype=torch.float)  # V5:186-203
        self.obs_scaling = torch.ones(self.num_obs, device=self.device)             # V5:241-266
        if self.cfg["env"]["SCALE_OBSERVATIONS"]:
            self.obs_scaling[:] = torch.tensor(_OBS_SCALING[observation_type], device=self.device)
        self.target_velocities = torch.zeros(self.num_envs, NUM_XYZ, device=self.device)  # V5:916-918
        self.index_to_view = int(0.1 * self.num_envs)
        self.num_steps = 0

    # ------------------------------------------------------------------ construction
    def create_sim(self):
        """V5:364-519 collapses to: bake constants, allocate the SoA state, bind the VecTask buffers."""
        self.up_axis = self.cfg["sim"]["up_axis"]
        assert self.up_axis == "z"                            # V5:441
        torch.cuda.set_device(self.device)
        self.vine_config = vcfg.task_cfg_to_vine_config(self.cfg)
        h = C.c_void_p()
        rc = self._lib.vine_create(C.byref(self.vine_config), self.num_envs, self._global_env_offset,
                                   self.device_id, self._seed, C.byref(h))
        if rc != abi.OK:
            msg = self._lib.vine_last_error(None).decode()
            if rc == abi.ERR_UNSUPPORTED:
                raise NotImplementedError(msg)
            raise RuntimeError(f"vine_create failed ({rc}): {msg}")
        self._h = h
        self.sim = h
        self.actions = torch.zeros(self.num_envs, self.num_actions, device=self.device, dtype=torch.float)
        self._obs_clamped = torch.zeros_like(self.obs_buf)
        self._bind()

    def _bind(self, actions=None):
        a = self.actions if actions is None else actions
        self._check(self._lib.vine_bind_io(self._h, _ptr(a), _ptr(self.obs_buf), _ptr(self.rew_buf),
                                           _ptr(self.reset_buf), _ptr(self.progress_buf),
                                           _ptr(self.timeout_buf), _ptr(self._obs_clamped)))

    def _check(self, rc):
        if rc != abi.OK:
            raise RuntimeError(f"libvine_b200 error {rc}: {self._lib.vine_last_error(self._h).decode()}")

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def __del__(self):
        try:
            if self._h is not None:
                self._lib.vine_destroy(self._h)
                self._h = None
        except Exception:
            pass

    # ------------------------------------------------------------------ VecTask contract
    def pre_physics_step(self, actions):
        """V5:922-945.  The action path itself runs inside the fused kernel; this stages the actions."""
        self.actions.copy_(actions, non_blocking=True)

    def post_physics_step(self):
        """V5:1110-1120 runs inside the fused kernel (progress, deferred reset, obs, reward, resets)."""
        self.num_steps += 1

    # ------------------------------------------------------------------ .mat trajectory replay (V5:281-297, 947-982)
    def read_mat_file(self, filename):
        """MAT_FILE of a recorded hardware run: cart_pos (1,T), Q (5,T), moving_target_pos (3,T), tip_pos (3,T), ..."""
        import scipy.io
        self.mat = scipy.io.loadmat(filename)
        return self.mat

    def overwrite_with_mat(self):
        """Overwrite every env's DOF positions (velocities 0), target and rigid-body tip with step
        ``num_steps % T`` of the recorded trajectory, like the reference's viewer replay (V5:947-982)."""
        if getattr(self, "mat", None) is None:
            self.read_mat_file(self.cfg["env"]["MAT_FILE"])
        m, n = self.mat, self.num_envs
        total = m["cart_pos"].shape[1]
        assert m["cart_pos"].shape == (1, total) and m["Q"].shape == (5, total)
        i = self.num_steps % total
        row = lambda a: torch.as_tensor(a, dtype=torch.float32, device=self.device).reshape(1, -1).repeat(n, 1)  # noqa: E731
        state = {"dof_pos": torch.cat([row(m["cart_pos"][:, i]), row(m["Q"][:, i])], dim=1),
                 "dof_vel": torch.zeros(n, 6, device=self.device),
                 "target_positions": row(m["moving_target_pos"][:, i])}
        if "tip_pos" in m:
            state["tip_positions"] = row(m["tip_pos"][:, i])
        self.set_state_dict(state)
        return i

    def step(self, actions):
        """VecTask.step (VT:319-380) as one kernel launch (or one CUDA-graph replay)."""
        self.pre_physics_step(actions)
        if self._graph is not None:
            self._graph.replay()
        else:
            self._check(self._lib.vine_step(self._h, self._stream()))
        self.post_physics_step()
        self.extras["time_outs"] = self.timeout_buf.to(self.rl_device)      # VT:372
        self.obs_dict["obs"] = self._obs_clamped.to(self.rl_device)         # VT:374
        if self.num_states > 0:
            self.obs_dict["states"] = self.get_state()
        return self.obs_dict, self.rew_buf.to(self.rl_device), self.reset_buf.to(self.rl_device), self.extras

    def step_device(self):
        """Hot loop for callers that write ``self.actions`` in place: launch only, no Python tensors."""
        if self._graph is not None:
            self._graph.replay()
        else:
            self._check(self._lib.vine_step(self._h, self._stream()))

    def step_host(self, actions_host, obs_host, rew_host, reset_host, timeout_host, chunks=8):
        """``step`` for callers whose policy lives on the HOST: pinned ``actions_host`` in, pinned
        ``obs_host`` (clamped, VT:374) / ``rew_host`` / ``reset_host`` / ``timeout_host`` out.

        The env range is cut into ``chunks`` pieces; each piece runs H2D(actions) -> fused step ->
        D2H(results) on its own stream, so the three phases of different pieces overlap (both copy
        engines and the SMs busy at once).  Envs are independent, so the result is bit-identical to
        ``step``.  Returns after everything has landed in the host buffers.
        """
        n = self.num_envs
        per = -(-n // max(int(chunks), 1))
        per = max(128, -(-per // 128) * 128)          # vine_step_range wants starts on CTA boundaries
        if not hasattr(self, "_chunk_streams"):
            self._chunk_streams = [torch.cuda.Stream(self.device) for _ in range(4)]
        cur = torch.cuda.current_stream(self.device)
        ready = torch.cuda.Event()
        ready.record(cur)
        first, k = 0, 0
        while first < n:
            cnt = min(per, n - first)
            s = self._chunk_streams[k % len(self._chunk_streams)]
            sl = slice(first, first + cnt)
            with torch.cuda.stream(s):
                if k < len(self._chunk_streams):
                    s.wait_event(ready)
                self.actions[sl].copy_(actions_host[sl], non_blocking=True)
                self._check(self._lib.vine_step_range(self._h, first, cnt, C.c_void_p(s.cuda_stream)))
                obs_host[sl].copy_(self._obs_clamped[sl], non_blocking=True)
                rew_host[sl].copy_(self.rew_buf[sl], non_blocking=True)
                reset_host[sl].copy_(self.reset_buf[sl], non_blocking=True)
                timeout_host[sl].copy_(self.timeout_buf[sl], non_blocking=True)
            first += cnt
            k += 1
        for s in self._chunk_streams:
            cur.wait_stream(s)
            s.synchronize()
        self.num_steps += 1

    def capture_graph(self):
        """Capture the step into a CUDA graph (the kernel neither allocates nor synchronises)."""
        torch.cuda.synchronize(self.device)
        s = torch.cuda.Stream(self.device)
        g = torch.cuda.CUDAGraph()
        state = self.get_state_dict()
        bufs = [b.clone() for b in (self.obs_buf, self.rew_buf, self.reset_buf, self.progress_buf, self.timeout_buf)]
        with torch.cuda.stream(s):
            with torch.cuda.graph(g, stream=s):
                self._check(self._lib.vine_step(self._h, C.c_void_p(s.cuda_stream)))
        torch.cuda.synchronize(self.device)
        # capture does not execute, but restore anyway so capture_graph() is side-effect free
        self.set_state_dict(state)
        for dst, src in zip((self.obs_buf, self.rew_buf, self.reset_buf, self.progress_buf, self.timeout_buf), bufs):
            dst.copy_(src)
        self._graph = g
        return g

    def reset_idx(self, env_ids):
        """V5:774-885 outside step (reset_done VT:412-427)."""
        ids = torch.as_tensor(env_ids, device=self.device, dtype=torch.long).contiguous()
        self._check(self._lib.vine_reset_idx(self._h, _ptr(ids), ids.numel(), self._stream()))

    def compute_observations(self, env_ids=None):
        """V5:1339: observations are produced by step(); returns the current buffer."""
        return self.obs_buf

    def compute_reward(self):
        """V5:1218: rewards are produced by step(); returns the current buffer."""
        return self.rew_buf

    def refresh_state_tensors(self):
        """V5:1333-1337: state lives in the library; the properties below read it on demand."""
        return None

    # ------------------------------------------------------------------ state access
    _STATE_FIELDS = {
        "dof_pos": (6,), "dof_vel": (6,), "tip_positions": (3,), "cart_body_vel_y": (), "target_positions": (3,),
        "object_info": (2,), "smoothed_u_fpam": (), "prev_cart_vel": (), "prev_cart_vel_error": (),
        "shelf_contact_force": (), "aggregated_rew_buf": (),
    }
    _DEBUG_FIELDS = {"u_rail_velocity": (), "u_fpam": (), "prev_u_rail_velocity": (), "rail_force": (),
                     "tip_velocities": (3,), "reward_matrix": (13,), "finite_difference_dof_vel": (6,),
                     "finite_difference_tip_velocities": (3,), "cart_body_pos_y": ()}

    def get_state_dict(self, debug=False):
        """Snapshot of the private SoA state as torch tensors named like the reference attributes."""
        n, dev = self.num_envs, self.device
        out = {k: torch.zeros((n,) + s, device=dev) for k, s in self._STATE_FIELDS.items()}
        D = int(self.cfg["env"]["ACTION_DELAY"])
        out["actions_history"] = torch.zeros(n, max(D, 1), 2, device=dev)
        out["step_count"] = torch.zeros(n, device=dev, dtype=torch.long)
        if debug:
            out.update({k: torch.zeros((n,) + s, device=dev) for k, s in self._DEBUG_FIELDS.items()})
        view = abi.VineStateView()
        for name, ftype in view._fields_:
            if name in out:
                setattr(view, name, C.cast(out[name].data_ptr(), ftype))
        self._check(self._lib.vine_get_state(self._h, C.byref(view), self._stream()))
        return out

    def set_state_dict(self, state):
        view = abi.VineStateView()
        keep = []
        for name, ftype in view._fields_:
            if name in state and name not in self._DEBUG_FIELDS:
                want = torch.long if name == "step_count" else torch.float
                t = state[name].to(device=self.device, dtype=want).contiguous()
                keep.append(t)
                setattr(view, name, C.cast(t.data_ptr(), ftype))
        self._check(self._lib.vine_set_state(self._h, C.byref(view), self._stream()))
        torch.cuda.current_stream(self.device).synchronize()
        del keep

    def enable_debug_outputs(self, enabled=True):
        self._check(self._lib.vine_set_debug_outputs(self._h, int(enabled)))
        self._debug_enabled = bool(enabled)

    def metrics_async(self):
        """Launch the metrics reduction (vine_metrics) for the state after the last step; returns pinned host tensors
        (sums f64[45], maxes f32[30]) and a CUDA event that fires when they are filled -- no host synchronisation."""
        if not getattr(self, "_debug_enabled", False):
            self.enable_debug_outputs(True)
            raise RuntimeError("metrics need the debug plane: enabled now, call again after the next step()")
        if not hasattr(self, "_metric_bufs"):
            self._metric_bufs = (torch.zeros(abi.METRIC_SUMS, dtype=torch.float64, device=self.device),
                                 torch.zeros(abi.METRIC_MAXES, dtype=torch.float32, device=self.device),
                                 torch.zeros(abi.METRIC_SUMS, dtype=torch.float64).pin_memory(),
                                 torch.zeros(abi.METRIC_MAXES, dtype=torch.float32).pin_memory())
        sums, maxes, sums_h, maxes_h = self._metric_bufs
        self._check(self._lib.vine_metrics(self._h, C.c_void_p(sums.data_ptr()), C.c_void_p(maxes.data_ptr()), self._stream()))
        sums_h.copy_(sums, non_blocking=True)
        maxes_h.copy_(maxes, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        return sums_h, maxes_h, ev

    def metrics(self):
        """The aggregate keys of the reference's ``wandb_dict`` (compute_reward, V5:1250-1322) with the same names."""
        sums_h, maxes_h, ev = self.metrics_async()
        ev.synchronize()
        n = float(self.num_envs)
        s, m = sums_h.tolist(), maxes_h.tolist()
        out = {k: s[i] / n for i, k in enumerate(abi.METRIC_SCALARS)}
        out.update({"max_abs_tip_y": m[0], "max_tip_z": m[1], "tip_velocities_max": m[2]})
        R = len(REWARD_NAMES)
        for i, name in enumerate(REWARD_NAMES):
            out[f"Mean {name} Reward"] = s[16 + i] / n
            out[f"Max {name} Reward"] = m[3 + i]
            out[f"Weighted Mean {name} Reward"] = s[16 + R + i] / n
            out[f"Weighted Max {name} Reward"] = m[3 + R + i]
        out["Mean Total Reward"], out["Max Total Reward"] = s[16 + 2 * R] / n, m[3 + 2 * R]
        mean = s[17 + 2 * R] / n
        std = max(s[18 + 2 * R] - n * mean * mean, 0.0) / max(n - 1.0, 1.0)
        out["Aggregated Reward"] = mean
        out["Aggregated Reward 1 Std Up"], out["Aggregated Reward 1 Std Down"] = mean + std ** 0.5, mean - std ** 0.5
        return out

    def view_traces(self, index=None):
        """The per-env traces the reference logs for ONE env, ``self.index_to_view`` (V5:1283-1312), with its key names:
        joint positions / velocities / finite-difference velocities, tip / cart / target positions and velocities, the
        commands, the rail force and the shelf contact force of the last step.  One state read, one host copy."""
        if not getattr(self, "_debug_enabled", False):
            self.enable_debug_outputs(True)
            raise RuntimeError("view traces need the debug plane: enabled now, call again after the next step()")
        i = self.index_to_view if index is None else int(index)
        st = {k: v[i].tolist() if v.dim() > 1 else float(v[i]) for k, v in self.get_state_dict(debug=True).items()
              if k not in ("actions_history", "step_count")}
        at = " at self.index_to_view"
        out = {f"prismatic_q0{at}": st["dof_pos"][0], f"prismatic_qd0{at}": st["dof_vel"][0],
               f"prismatic_finite_diff_qd0{at}": st["finite_difference_dof_vel"][0]}
        for j in range(N_REVOLUTE_DOFS):
            out[f"q{j}{at}"], out[f"qd{j}{at}"] = st["dof_pos"][j + 1], st["dof_vel"][j + 1]
            out[f"finite_diff_qd{j}{at}"] = st["finite_difference_dof_vel"][j + 1]
        cart_pos, cart_vel = [0.0, st["cart_body_pos_y"], 0.975], [0.0, st["cart_body_vel_y"], 0.0]
        for k, d in enumerate("xyz"):
            out[f"tip_vel_{d}{at}"], out[f"cart_vel_{d}{at}"], out[f"target_vel_{d}{at}"] = st["tip_velocities"][k], cart_vel[k], 0.0
            out[f"finite_diff_tip_vel_{d}{at}"] = st["finite_difference_tip_velocities"][k]
            out[f"tip_pos_{d}{at}"], out[f"cart_pos_{d}{at}"] = st["tip_positions"][k], cart_pos[k]
            out[f"target_pos_{d}{at}"] = st["target_positions"][k]
        contact = -st["reward_matrix"][12]                      # Contact Force term = -contact [contact > 0] (V5:1531)
        out.update({f"u_fpam{at}": st["u_fpam"], f"smoothed u_fpam{at}": st["smoothed_u_fpam"],
                    f"u_rail_velocity{at}": st["u_rail_velocity"], f"rail_force{at}": st["rail_force"],
                    f"contact_force{at}": contact, f"nonzero_contact_force{at}": float(contact > 0)})
        return out

    def wandb_dict(self):
        """Everything the reference's compute_reward puts into ``self.wandb_dict`` each step (V5:1250-1322), same keys:
        the aggregate entries (one reduction launch, ``metrics()``) plus the per-view-env traces (``view_traces()``)."""
        out = self.metrics()
        out.update(self.view_traces())
        return out

    def _get(self, name):
        return self.get_state_dict(debug=name in self._DEBUG_FIELDS)[name]

    def _view(self, name, shape=None, field=None, readonly=False):
        """The reference's state attributes are torch views of the simulator's tensors; here the state lives in the library's
        SoA planes, so an attribute read is a snapshot (one small kernel) wrapped in ``StateView``: reading works like a
        tensor, and the reference's in-place idioms -- ``env.dof_pos[ids] = x`` (V5:791-793), ``.copy_()``, ``.zero_()``,
        ``+=`` -- write THROUGH to the library state (``vine_set_state``).  A snapshot does not follow later steps: read the
        attribute again after ``step()``."""
        t = self._get(field or name)
        if shape is not None:
            t = t.reshape(shape)
        return StateView.wrap(t, self, field or name, readonly)

    dof_pos = property(lambda s: s._view("dof_pos"), lambda s, v: s.set_state_dict({"dof_pos": v}))
    dof_vel = property(lambda s: s._view("dof_vel"), lambda s, v: s.set_state_dict({"dof_vel": v}))
    tip_positions = property(lambda s: s._view("tip_positions"), lambda s, v: s.set_state_dict({"tip_positions": v}))
    target_positions = property(lambda s: s._view("target_positions"), lambda s, v: s.set_state_dict({"target_positions": v}))
    object_info = property(lambda s: s._view("object_info"), lambda s, v: s.set_state_dict({"object_info": v}))
    smoothed_u_fpam = property(lambda s: s._view("smoothed_u_fpam", (-1, 1)),
                               lambda s, v: s.set_state_dict({"smoothed_u_fpam": v.reshape(-1)}))
    prev_cart_vel = property(lambda s: s._view("prev_cart_vel", (-1, 1)),
                             lambda s, v: s.set_state_dict({"prev_cart_vel": v.reshape(-1)}))
    prev_cart_vel_error = property(lambda s: s._view("prev_cart_vel_error", (-1, 1)),
                                   lambda s, v: s.set_state_dict({"prev_cart_vel_error": v.reshape(-1)}))
    aggregated_rew_buf = property(lambda s: s._view("aggregated_rew_buf"), lambda s, v: s.set_state_dict({"aggregated_rew_buf": v}))
    # outputs of the last step (debug plane): read-only, an in-place write raises instead of silently doing nothing
    u_rail_velocity = property(lambda s: s._view("u_rail_velocity", (-1, 1), readonly=True))
    u_fpam = property(lambda s: s._view("u_fpam", (-1, 1), readonly=True))
    prev_u_rail_velocity = property(lambda s: s._view("prev_u_rail_velocity", (-1, 1), readonly=True))
    rail_force = property(lambda s: s._view("rail_force", (-1, 1), readonly=True))
    tip_velocities = property(lambda s: s._view("tip_velocities", readonly=True))

    @property
    def cart_positions(self):
        """Rigid-body view V5:359: (0, cart y, 0.975) (URDF:275); read-only (write dof_pos).  With the debug
        outputs enabled it is the body position the last step's reward used -- on a reset step still the old episode's (stale body
        views, V5:796) -- otherwise dof_pos[:, 0], which differs from it only on reset steps."""
        dbg = getattr(self, "_debug_enabled", False)
        y = self._get("cart_body_pos_y") if dbg else self._get("dof_pos")[:, 0]
        out = torch.zeros(self.num_envs, 3, device=self.device)
        out[:, 1] = y
        out[:, 2] = 0.975
        return StateView.wrap(out, self, "cart_positions", True)

    @property
    def cart_velocities(self):
        """Rigid-body view V5:362, read by the rail controller at V5:1069: (0, cart body velocity as of the last simulate, 0).
        Writable: the y column is the library's ``cart_body_vel_y`` plane (stale after a reset like the reference's)."""
        v = self._get("cart_body_vel_y")
        out = torch.zeros(self.num_envs, 3, device=self.device)
        out[:, 1] = v
        return StateView.wrap(out, self, "cart_body_vel_y", False, post=lambda t: t[:, 1].contiguous())


class StateView(torch.Tensor):
    """Snapshot of one state attribute that writes in-place modifications back to the library (see ``_view``)."""
    __torch_function__ = torch._C._disabled_torch_function_impl   # results of ordinary ops are plain tensors

    @staticmethod
    def wrap(data, env, name, readonly=False, post=None):
        t = torch.Tensor._make_subclass(StateView, data)
        t._env, t._name, t._readonly, t._post = env, name, readonly, post
        return t

    def _push(self):
        if self._readonly:
            raise RuntimeError(f"'{self._name}' is an output of the last step (read-only): writing it has no counterpart in the "
                               "library state; set dof_pos / dof_vel / tip_positions / ... instead")
        plain = self.as_subclass(torch.Tensor)
        plain = self._post(plain) if self._post is not None else plain
        n = self._env.num_envs
        self._env.set_state_dict({self._name: plain.reshape(n) if plain.dim() == 2 and plain.shape[1] == 1
                                  and self._name in self._env._STATE_FIELDS and self._env._STATE_FIELDS[self._name] == () else plain})

    def __setitem__(self, key, value):
        self.as_subclass(torch.Tensor).__setitem__(key, value)
        self._push()


def _write_through(name):
    plain = getattr(torch.Tensor, name)

    def method(self, *args, **kwargs):
        plain(self.as_subclass(torch.Tensor), *args, **kwargs)
        self._push()
        return self
    method.__name__ = name
    return method


for _m in ("copy_", "zero_", "fill_", "add_", "sub_", "mul_", "div_", "clamp_", "__iadd__", "__isub__", "__imul__", "__itruediv__"):
    setattr(StateView, _m, _write_through(_m))


_OBS_SCALING = {  # V5:246-266
    ObservationType.POS_AND_FD_VEL_AND_OBJ_INFO: [0.12, 0.269, 0.148, 0.249, 0.148, 0.344,
                                                  0.67, 2.22, 1.47, 1.14, 0.903, 0.716,
                                                  0.0656, 0.238, 0.0656, 0.732, 2.0, 0.732,
                                                  0.02, 0.0235, 0.02, 0.732, 2.0, 0.732,
                                                  0.845, 0.86, 0.0385, 0.5],
    ObservationType.TIP_AND_CART_AND_OBJ_INFO: [0.12, 0.67, 0.0656, 0.238, 0.0656, 0.732, 2.0, 0.732,
                                                0.02, 0.0235, 0.02, 0.732, 2.0, 0.732,
                                                0.845, 0.86, 0.0385, 0.5],
}
