"""ctypes wrapper of the CPU ORACLE (oracle/libvine_oracle.so).  TEST INFRASTRUCTURE ONLY.

Only tests/, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this module; the product package never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from vine_robot_isaacgymenvs_b200 import abi

_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_DIR, "libvine_oracle.so")
_lib = None

_fp = C.POINTER(C.c_float)
_i64p = C.POINTER(C.c_int64)
_u8p = C.POINTER(C.c_uint8)


class OracleArrays(C.Structure):
    _fields_ = [("n", C.c_int64), ("global_env_offset", C.c_int64), ("seed", C.c_uint64)] + [
        (k, _fp) for k in ("dof_pos", "dof_vel", "tip_body", "tipvel_body", "cart_body_y",
                           "cart_body_vy", "target", "object_info", "smoothed", "prev_cart_vel",
                           "prev_cart_vel_error", "lip_force", "history", "agg_rew")] + [
        ("step_count", _i64p), ("actions", _fp), ("obs", _fp), ("obs_clamped", _fp), ("rew", _fp),
        ("reset", _i64p), ("progress", _i64p), ("timeout", _u8p)] + [
        (k, _fp) for k in ("u_rail", "u_fpam", "prev_u_rail", "rail_force", "reward_matrix")]


def build(force=False):
    src = [os.path.join(_DIR, f) for f in ("vine_oracle.c", "vine_oracle_dyn.inc", "vine_oracle.h")]
    src.append(os.path.join(_DIR, "..", "include", "vine_b200.h"))
    if not force and os.path.exists(LIB_PATH):
        if all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(s) for s in src if os.path.exists(s)):
            return LIB_PATH
    subprocess.run(["make", "-C", _DIR, "-B", "libvine_oracle.so"], check=True, capture_output=True)
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        L.oracle_num_observations.argtypes = [C.c_int]
        L.oracle_config_defaults.argtypes = [C.POINTER(abi.VineConfig)]
        L.oracle_config_defaults.restype = None
        L.oracle_philox.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                    C.POINTER(C.c_uint32)]
        L.oracle_philox.restype = None
        for f in (L.oracle_uniform4, L.oracle_normal4):
            f.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, _fp]
            f.restype = None
        L.oracle_dynamics_uniforms.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, _fp]
        L.oracle_dynamics_uniforms.restype = None
        L.oracle_step.argtypes = [C.POINTER(abi.VineConfig), C.POINTER(OracleArrays), C.c_int, C.c_int]
        L.oracle_init.argtypes = [C.POINTER(abi.VineConfig), C.POINTER(OracleArrays)]
        L.oracle_reset_idx.argtypes = [C.POINTER(abi.VineConfig), C.POINTER(OracleArrays), _i64p, C.c_int64]
        L.oracle_pre_physics.argtypes = [C.POINTER(abi.VineConfig), C.c_int64, C.POINTER(abi.VinePrePhysicsIO)]
        L.oracle_actuation.argtypes = [C.POINTER(abi.VineConfig), C.c_int64, C.POINTER(abi.VineActuationIO)]
        L.oracle_simulate.argtypes = [C.POINTER(abi.VineConfig), C.c_int64, C.POINTER(abi.VineSimulateIO), C.c_int]
        L.oracle_post_physics.argtypes = [C.POINTER(abi.VineConfig), C.c_int64, C.POINTER(abi.VinePostPhysicsIO)]
        L.oracle_gae.argtypes = [_fp, _fp, _fp, _fp, _fp, C.c_int64, C.c_int64, C.c_double, C.c_double, _fp, _fp]
        L.oracle_mass_matrix.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.oracle_mass_matrix.restype = None
        L.oracle_energy.argtypes = [C.POINTER(abi.VineConfig), C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.oracle_energy.restype = C.c_double
        _lib = L
    return _lib


def default_config():
    cfg = abi.VineConfig()
    lib().oracle_config_defaults(C.byref(cfg))
    return cfg


def _ptr(a, ctype):
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"], "oracle arrays must be C-contiguous"
    return a.ctypes.data_as(C.POINTER(ctype))


def fill_struct(struct, arrays):
    """Point the pointer fields of a ctypes struct at numpy arrays (dict name -> array|None)."""
    for name, ftype in struct._fields_:
        if name in arrays and arrays[name] is not None:
            a = arrays[name]
            setattr(struct, name, a.ctypes.data_as(ftype))
    return struct


def philox(seed, gid, site, step, block):
    out = (C.c_uint32 * 4)()
    lib().oracle_philox(seed, gid, site, step, block, out)
    return np.array(list(out), dtype=np.uint32)


def uniform4(seed, gid, site, step, block):
    out = (C.c_float * 4)()
    lib().oracle_uniform4(seed, gid, site, step, block, out)
    return np.array(list(out), dtype=np.float32)


def dynamics_uniforms(seed, gid, step, sim_i):
    """The 24 16-bit uniforms of one sim step's dynamics-scaling draw (3 Philox blocks)."""
    out = (C.c_float * 24)()
    lib().oracle_dynamics_uniforms(seed, gid, step, sim_i, out)
    return np.array(list(out), dtype=np.float32)


def normal4(seed, gid, site, step, block):
    out = (C.c_float * 4)()
    lib().oracle_normal4(seed, gid, site, step, block, out)
    return np.array(list(out), dtype=np.float32)


SITE_ACTION_NOISE, SITE_DYNAMICS, SITE_OBS_NOISE, SITE_RESET = 1, 2, 3, 4


class OracleEnv:
    """Whole-env state holder mirroring the tensors of the reference task (numpy, host)."""

    STATE_F32 = {"dof_pos": 6, "dof_vel": 6, "tip_body": 3, "tipvel_body": 3, "cart_body_y": 0,
                 "cart_body_vy": 0, "target": 3, "object_info": 2, "smoothed": 0, "prev_cart_vel": 0,
                 "prev_cart_vel_error": 0, "lip_force": 0, "agg_rew": 0,
                 "u_rail": 0, "u_fpam": 0, "prev_u_rail": 0, "rail_force": 0, "reward_matrix": 13,
                 "rew": 0}

    def __init__(self, cfg, num_envs, seed=42, global_env_offset=0, use_f64=True, nthreads=0):
        self.cfg = cfg
        self.n = int(num_envs)
        self.seed = int(seed)
        self.use_f64 = int(use_f64)
        self.nthreads = int(nthreads)
        self.O = lib().oracle_num_observations(cfg.observation_type)
        self.D = cfg.action_delay
        n = self.n
        self.a = {}
        for k, w in self.STATE_F32.items():
            self.a[k] = np.zeros((n, w) if w else (n,), np.float32)
        self.a["history"] = np.zeros((n, max(self.D, 1), 2), np.float32)
        self.a["step_count"] = np.zeros(n, np.int64)
        self.a["actions"] = np.zeros((n, 2), np.float32)
        self.a["obs"] = np.zeros((n, self.O), np.float32)
        self.a["obs_clamped"] = np.zeros((n, self.O), np.float32)
        self.a["reset"] = np.ones(n, np.int64)      # VT:275
        self.a["progress"] = np.zeros(n, np.int64)
        self.a["timeout"] = np.zeros(n, np.uint8)
        self.arr = OracleArrays()
        self.arr.n = n
        self.arr.global_env_offset = global_env_offset
        self.arr.seed = self.seed
        fill_struct(self.arr, self.a)
        rc = lib().oracle_init(C.byref(self.cfg), C.byref(self.arr))
        if rc != 0:
            raise RuntimeError(f"oracle_init failed: {rc} (unsupported config?)")

    def __getattr__(self, k):
        a = self.__dict__.get("a")
        if a is not None and k in a:
            return a[k]
        raise AttributeError(k)

    def step(self, actions):
        self.a["actions"][...] = np.asarray(actions, np.float32)
        rc = lib().oracle_step(C.byref(self.cfg), C.byref(self.arr), self.use_f64, self.nthreads)
        if rc != 0:
            raise RuntimeError(f"oracle_step failed: {rc}")
        return self.a["obs"], self.a["rew"], self.a["reset"], self.a["timeout"]

    def reset_idx(self, env_ids):
        ids = np.ascontiguousarray(env_ids, np.int64)
        rc = lib().oracle_reset_idx(C.byref(self.cfg), C.byref(self.arr), _ptr(ids, C.c_int64), len(ids))
        if rc != 0:
            raise RuntimeError(f"oracle_reset_idx failed: {rc}")


def call_io(fn_name, cfg, n, io_struct, arrays, *extra):
    io = fill_struct(io_struct(), arrays)
    rc = getattr(lib(), fn_name)(C.byref(cfg), n, C.byref(io), *extra)
    if rc != 0:
        raise RuntimeError(f"{fn_name} failed: {rc}")


def gae(rewards, values, dones, last_values, last_dones, gamma, tau):
    T, N = rewards.shape
    adv = np.zeros((T, N), np.float32)
    ret = np.zeros((T, N), np.float32)
    f = lambda a: _ptr(np.ascontiguousarray(a, np.float32), C.c_float)  # noqa: E731
    r, v, d, lv, ld = (np.ascontiguousarray(x, np.float32) for x in (rewards, values, dones, last_values, last_dones))
    lib().oracle_gae(_ptr(r, C.c_float), _ptr(v, C.c_float), _ptr(d, C.c_float), _ptr(lv, C.c_float),
                     _ptr(ld, C.c_float), T, N, gamma, tau, _ptr(adv, C.c_float), _ptr(ret, C.c_float))
    del f
    return adv, ret


def mass_matrix(q):
    q = np.ascontiguousarray(q, np.float64)
    M = np.zeros(36, np.float64)
    lib().oracle_mass_matrix(_ptr(q, C.c_double), _ptr(M, C.c_double))
    return M.reshape(6, 6)


def energy(cfg, q, qd):
    q = np.ascontiguousarray(q, np.float64)
    qd = np.ascontiguousarray(qd, np.float64)
    return lib().oracle_energy(C.byref(cfg), _ptr(q, C.c_double), _ptr(qd, C.c_double))
