"""Pin the CPU oracle against the fixtures produced by executing the reference
(tests/golden/generate_golden.py).  CPU only.

The oracle restates the task logic in f32 with torch's CPU operation order, so everything
except the batched-GEMV torque sum (V5:1062) is compared BIT-EXACTLY.
"""
import numpy as np
import pytest

from conftest import golden_files, golden_id, load_golden
from vine_robot_isaacgymenvs_b200 import abi

STEP_KEYS_EXACT = ["obs_buf", "obs_clamped", "rew_buf", "reset_buf", "progress_buf", "timeout_buf", "dof_pos",
                   "dof_vel", "target_positions", "object_info", "smoothed_u_fpam", "u_fpam", "u_rail_velocity",
                   "prev_u_rail_velocity", "rail_force", "prev_cart_vel", "prev_cart_vel_error",
                   "aggregated_rew_buf", "tip_positions", "contact"]


def oracle_view(oe):
    return {"obs_buf": oe.obs, "obs_clamped": oe.obs_clamped, "rew_buf": oe.rew, "reset_buf": oe.reset,
            "progress_buf": oe.progress, "timeout_buf": oe.timeout, "dof_pos": oe.dof_pos, "dof_vel": oe.dof_vel,
            "target_positions": oe.target, "object_info": oe.object_info, "smoothed_u_fpam": oe.smoothed,
            "u_fpam": oe.u_fpam, "u_rail_velocity": oe.u_rail, "prev_u_rail_velocity": oe.prev_u_rail,
            "rail_force": oe.rail_force, "prev_cart_vel": oe.prev_cart_vel,
            "prev_cart_vel_error": oe.prev_cart_vel_error, "aggregated_rew_buf": oe.agg_rew,
            "tip_positions": oe.tip_body, "contact": oe.lip_force}


@pytest.mark.parametrize("path", golden_files("step_"), ids=golden_id)
def test_oracle_step_matches_reference_rollout(path, oracle_lib):
    """Whole VecTask.step rollouts (VT:319-380) of the reference's own task object."""
    g, vc, _ = load_golden(path)
    T, n = g["actions"].shape[:2]
    oe = oracle_lib.OracleEnv(vc, n, seed=int(g["seed"]), use_f64=True)
    for t in range(T):
        oe.step(g["actions"][t])
        view = oracle_view(oe)
        for k in STEP_KEYS_EXACT:
            if vc.torque_law_integration == abi.TORQUE_LAW_INTEGRATION["zoh"]:
                # ZOH feeds the reference's matmul torques (BLAS summation order, 1 ulp) into an
                # unstable integrator: compare with a tolerance instead of bit-exactly
                np.testing.assert_allclose(view[k], g[k][t], rtol=2e-2, atol=1e-3, err_msg=k)
                continue
            assert np.array_equal(view[k], g[k][t]), f"{k} differs at step {t}: max |d| = " \
                f"{np.abs(view[k].astype(np.float64) - g[k][t].astype(np.float64)).max()}"
        if int(g["reset_done_at"]) == t:
            oe.reset_idx(np.arange(0, n, 3))
    assert T <= 2 or g["reset_buf"].sum() > 0, "fixture exercises no reset"


@pytest.mark.parametrize("path", golden_files("fn_post_"), ids=golden_id)
def test_oracle_post_physics_matches_reference_functions(path, oracle_lib):
    """compute_observations V5:1339 + compute_reward V5:1218 / compute_reward_jit V5:1470 +
    compute_reset_jit V5:1540 + timeout VT:366 on synthetic inputs incl. threshold cases."""
    g, vc, _ = load_golden(path)
    n = g["dof_pos"].shape[0]
    O_ = abi.NUM_OBSERVATIONS[vc.observation_type]
    out = {"obs_buf": np.zeros((n, O_), np.float32), "rew_buf": np.zeros(n, np.float32),
           "reward_matrix": np.zeros((n, 13), np.float32), "reset_buf_out": np.zeros(n, np.int64),
           "timeout_buf": np.zeros(n, np.uint8)}
    arrays = {k: np.ascontiguousarray(v) for k, v in g.items()}
    arrays.update(out)
    if not vc.create_shelf:
        arrays["contact_force_norms"] = None
    oracle_lib.call_io("oracle_post_physics", vc, n, abi.VinePostPhysicsIO, arrays)
    assert np.array_equal(out["reset_buf_out"], g["reset_buf_out"])
    assert np.array_equal(out["timeout_buf"], g["timeout_buf"])
    assert np.array_equal(out["obs_buf"], g["obs_buf"])
    assert np.array_equal(out["reward_matrix"], g["reward_matrix"])
    assert np.array_equal(out["rew_buf"], g["rew_buf"])
    # the fixtures must actually straddle the thresholds they are there for
    assert 0 < g["target_reached"].sum() < n and 0 < g["limit_hit"].sum() < n


@pytest.mark.parametrize("path", golden_files("fn_pre_act_"), ids=golden_id)
def test_oracle_action_path_and_actuation_match_reference(path, oracle_lib):
    """pre_physics_step V5:927-940 (noise, rescale, delay ring, FORCE_*, smoothing) and
    compute_and_set_dof_actuation_force_tensor V5:1028-1106."""
    g, vc, _ = load_golden(path)
    T, n = g["actions"].shape[:2]
    D = max(vc.action_delay, 1)
    for t in range(T):
        out = {"history_out": np.zeros((n, D, 2), np.float32), "u_rail_velocity": np.zeros(n, np.float32),
               "u_fpam": np.zeros(n, np.float32), "smoothed_out": np.zeros(n, np.float32)}
        arrays = {"actions": g["actions"][t].copy(), "action_noise": g["action_noise"][t].copy() if vc.vine_randomize else None,
                  "history_in": g["history_in"][t].copy(), "smoothed_in": g["smoothed_in"][t].copy()}
        arrays.update(out)
        oracle_lib.call_io("oracle_pre_physics", vc, n, abi.VinePrePhysicsIO, arrays)
        for k in ("u_rail_velocity", "u_fpam", "smoothed_out"):
            assert np.array_equal(out[k], g[k][t]), k
        if vc.action_delay > 0:
            assert np.array_equal(out["history_out"], g["history_out"][t])
        out = {"dof_efforts": np.zeros((n, 6), np.float32), "prev_cart_vel_out": np.zeros(n, np.float32),
               "prev_cart_vel_error_out": np.zeros(n, np.float32)}
        arrays = {k: g[k][t].copy() for k in ("dof_pos", "dof_vel", "cart_vel_y", "u_rail_velocity", "u_fpam_to_use",
                                              "prev_cart_vel", "prev_cart_vel_error")}
        arrays["dynamics_scaling"] = g["dynamics_scaling"][t].copy() if vc.vine_randomize else None
        arrays.update(out)
        oracle_lib.call_io("oracle_actuation", vc, n, abi.VineActuationIO, arrays)
        # rail force: bit exact (pure elementwise f32); joint torques: 4-term dot product whose
        # summation order inside torch.matmul is a BLAS detail -> 1e-6 relative
        assert np.array_equal(out["dof_efforts"][:, 0], g["dof_efforts"][t][:, 0])
        np.testing.assert_allclose(out["dof_efforts"][:, 1:], g["dof_efforts"][t][:, 1:], rtol=2e-6, atol=1e-7)
        assert np.array_equal(out["prev_cart_vel_out"], g["prev_cart_vel_out"][t])
        assert np.array_equal(out["prev_cart_vel_error_out"], g["prev_cart_vel_error_out"][t])


def test_notebook_known_answer():
    """The only PhysX-generated numbers in the reference: visualize_observation_distribution.ipynb
    cell 3 shows consecutive POS_AND_FD_VEL rows; they pin joint_vel = (q_t - q_{t-1}) / (dt*4)
    (V5:1347, V5:228).  Values transcribed in SURVEY.md §4."""
    control_dt = np.float32(0.00833 * 4)
    fd = (np.float32(-0.115578) - np.float32(-0.101274)) / control_dt
    assert abs(float(fd) - (-0.429308)) < 2e-5


def test_dynamics_model_matches_survey_float64_probe(oracle_lib):
    """SURVEY Appendix E.3 (independent float64 Lagrangian model of the URDF): M(0) diagonal and
    the static equilibria for u = 0 and u = 3."""
    M = oracle_lib.mass_matrix(np.zeros(6))
    np.testing.assert_allclose(np.diag(M), [0.52, 1.68e-2, 1.01e-2, 5.1e-3, 1.9e-3, 3.0e-4], rtol=2e-2)
    assert np.allclose(M, M.T)
    cfg = oracle_lib.default_config()
    cfg.create_pipe = 0
    cfg.vine_randomize = 0
    cfg.randomize_dof_init = 0
    cfg.max_episode_length = 100000
    cfg.use_target_reached_reset = 0
    expect = {0.0: ((-0.0076, 0.5227), None),
              3.0: ((-0.093, 0.539), (-0.001, -0.076, -0.137, -0.092, -0.170))}
    for u, (tip, theta) in expect.items():
        env = oracle_lib.OracleEnv(cfg, 1)
        a1 = (u + 0.1) / 3.1 * 2 - 1
        for _ in range(400):
            env.step(np.array([[0.0, a1]], np.float32))
        np.testing.assert_allclose(env.tip_body[0, 1:], tip, atol=6e-4)
        if theta is not None:
            np.testing.assert_allclose(env.dof_pos[0, 1:], theta, atol=1.5e-3)
        assert np.abs(env.dof_vel).max() < 1e-4


def test_oracle_f32_tracks_f64(oracle_lib):
    """The f32 build of the dynamics (the CPU-baseline flavour) stays within the stated
    single-step tolerance of the f64 one (pos 1e-3, vel 1e-2 relative)."""
    g, vc, _ = load_golden(golden_files("step_c3_shelf")[0])
    n = g["actions"].shape[1]
    a = oracle_lib.OracleEnv(vc, n, use_f64=True)
    b = oracle_lib.OracleEnv(vc, n, use_f64=False)
    for t in range(3):
        a.step(g["actions"][t])
        b.step(g["actions"][t])
    np.testing.assert_allclose(b.dof_pos, a.dof_pos, rtol=1e-3, atol=1e-5)
    np.testing.assert_allclose(b.dof_vel, a.dof_vel, rtol=1e-2, atol=1e-4)
