"""GPU: the VecTask contract of the env class, CUDA-graph replay, and size-independent
properties at BASELINE's full sizes (sharding invariance, determinism, delay-ring semantics)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def fstr_env(n, extra=(), **kw):
    import vine_robot_isaacgymenvs_b200 as vine
    from vine_robot_isaacgymenvs_b200 import config as vcfg
    cfg = vcfg.compose(vcfg.FSTR_OVERRIDES + [f"num_envs={n}", "headless=True"] + list(extra))
    return vine.make(cfg=cfg, **kw)


def test_vectask_contract():
    n = 512
    env = fstr_env(n)
    assert (env.num_envs, env.num_obs, env.num_acts, env.num_states, env.num_agents) == (n, 18, 2, 0, 1)
    assert env.observation_space.shape == (18,) and env.action_space.shape == (2,)
    assert env.control_freq_inv == 4 and env.max_episode_length == 100 and env.clip_obs == 5.0 and env.clip_actions == 1.0
    assert abs(env.control_dt - 0.00833 * 4) < 1e-12 and env.reward_weights.shape == (1, 13) and env.obs_scaling.shape == (18,)
    for name, dt, shape in (("obs_buf", torch.float32, (n, 18)), ("states_buf", torch.float32, (n, 0)),
                            ("rew_buf", torch.float32, (n,)), ("reset_buf", torch.int64, (n,)),
                            ("progress_buf", torch.int64, (n,)), ("randomize_buf", torch.int64, (n,))):
        t = getattr(env, name)
        assert t.dtype == dt and tuple(t.shape) == shape and t.is_cuda, name
    assert bool((env.reset_buf == 1).all())                               # VT:275
    obs = env.reset()["obs"]                                              # VT:398-410: zeros before the first step
    assert tuple(obs.shape) == (n, 18) and float(obs.abs().max()) == 0.0
    assert tuple(env.zero_actions().shape) == (n, 2) and tuple(env.get_state().shape) == (n, 0)
    g = torch.Generator(device="cuda").manual_seed(0)
    for t in range(5):
        od, rew, reset, extras = env.step(torch.rand(n, 2, device="cuda", generator=g) * 4 - 2)
        assert od["obs"].dtype == torch.float32 and tuple(od["obs"].shape) == (n, 18)
        assert torch.equal(od["obs"], torch.clamp(env.obs_buf, -5.0, 5.0))   # VT:374
        assert float(od["obs"].abs().max()) <= 5.0 and float(env.obs_buf.abs().max()) > 5.0   # quirk D.11
        assert rew.dtype == torch.float32 and reset.dtype == torch.int64 and extras["time_outs"].dtype == torch.bool
        assert set(reset.unique().tolist()) <= {0, 1}
    # every env was reset during the first step (quirk D.4), so progress counts from there
    assert int(env.progress_buf.max()) <= 4
    for name, shape in (("dof_pos", (n, 6)), ("dof_vel", (n, 6)), ("tip_positions", (n, 3)), ("cart_positions", (n, 3)),
                        ("target_positions", (n, 3)), ("target_velocities", (n, 3)), ("object_info", (n, 2)),
                        ("smoothed_u_fpam", (n, 1)), ("aggregated_rew_buf", (n,))):
        assert tuple(getattr(env, name).shape) == shape, name
    # reset_done (VT:412-427)
    env.reset_buf[::7] = 1
    od, ids = env.reset_done()
    assert ids.numel() >= n // 7 and bool((env.reset_buf[ids] == 0).all()) and bool((env.progress_buf[ids] == 0).all())
    q = env.dof_pos
    assert float(q[ids, 1:].abs().max()) <= np.radians(10) + 1e-6 and float(env.dof_vel[ids].abs().max()) == 0.0
    t = env.target_positions[ids]
    assert float(t[:, 1].min()) >= -0.4 and float(t[:, 1].max()) <= 0.4 and float(t[:, 2].min()) >= 0.55


def test_config_errors_match_the_reference():
    import vine_robot_isaacgymenvs_b200 as vine
    with pytest.raises(KeyError):
        vine.make(num_envs=8, overrides=["OBSERVATION_TYPE=NOPE"])
    with pytest.raises(NotImplementedError):                       # V5:267-268
        vine.make(num_envs=8, overrides=["OBSERVATION_TYPE=POS_AND_FD_VEL"])
    with pytest.raises(RuntimeError, match="CUDA-only"):
        vine.make(num_envs=8, sim_device="cpu", overrides=["sim_device=cpu"])


def test_cuda_graph_replay_equals_eager_launches():
    n = 2048
    a, b = fstr_env(n), fstr_env(n)
    b.capture_graph()
    g = torch.Generator(device="cuda").manual_seed(1)
    for t in range(25):
        act = torch.rand(n, 2, device="cuda", generator=g) * 2 - 1
        a.step(act); b.step(act)
        for k in ("obs_buf", "rew_buf", "reset_buf", "progress_buf", "timeout_buf"):
            assert torch.equal(getattr(a, k), getattr(b, k)), (k, t)
    sa, sb = a.get_state_dict(), b.get_state_dict()
    for k in sa:
        assert torch.equal(sa[k], sb[k]), k


def test_pipelined_host_step_equals_plain_step():
    """env.step_host (chunked vine_step_range launches on several streams, pinned host I/O) must be
    bit-identical to env.step: envs are independent and Philox is keyed by the global env id."""
    n = 50_000                                     # not a multiple of the chunk / CTA size on purpose
    a, b = fstr_env(n), fstr_env(n)
    obs_h = torch.empty(n, 18).pin_memory(); rew_h = torch.empty(n).pin_memory()
    rst_h = torch.empty(n, dtype=torch.long).pin_memory(); to_h = torch.empty(n, dtype=torch.bool).pin_memory()
    g = torch.Generator().manual_seed(5)
    for t in range(12):
        act_h = (torch.rand(n, 2, generator=g) * 2 - 1).pin_memory()
        od, rew, rst, extras = a.step(act_h.cuda())
        b.step_host(act_h, obs_h, rew_h, rst_h, to_h, chunks=5)
        assert torch.equal(od["obs"].cpu(), obs_h) and torch.equal(rew.cpu(), rew_h), t
        assert torch.equal(rst.cpu(), rst_h) and torch.equal(extras["time_outs"].cpu(), to_h), t
        assert torch.equal(a.obs_buf, b.obs_buf) and torch.equal(a.progress_buf, b.progress_buf)


def test_sharding_invariance_and_determinism_at_full_size():
    """Philox keyed by the GLOBAL env id: 1 x 1,048,576 envs == 2 x 524,288 envs, bit for bit."""
    n = 1 << 20
    dr = ["task.task.randomization_parameters.OBSERVATION_NOISE_STD=0.01",
          "task.task.randomization_parameters.DYNAMICS_SCALING_MIN=0.9",
          "task.task.randomization_parameters.DYNAMICS_SCALING_MAX=1.1", "task.env.maxEpisodeLength=3"]
    whole = fstr_env(n, dr)
    lo = fstr_env(n // 2, dr, global_env_offset=0)
    hi = fstr_env(n // 2, dr, global_env_offset=n // 2)
    again = fstr_env(n, dr)
    g = torch.Generator(device="cuda").manual_seed(2)
    for t in range(5):
        act = torch.rand(n, 2, device="cuda", generator=g) * 2.4 - 1.2
        whole.step(act); again.step(act)
        lo.step(act[: n // 2]); hi.step(act[n // 2:])
        for k in ("obs_buf", "rew_buf", "reset_buf", "progress_buf", "timeout_buf"):
            w = getattr(whole, k)
            assert torch.equal(w, getattr(again, k)), ("determinism", k, t)
            assert torch.equal(w[: n // 2], getattr(lo, k)) and torch.equal(w[n // 2:], getattr(hi, k)), ("sharding", k, t)
        assert bool(torch.isfinite(whole.obs_buf).all())
    assert int(whole.reset_buf.sum()) > 0


@pytest.mark.parametrize("delay", [0, 1, 3])
def test_action_delay_is_exactly_k_control_steps(delay):
    n = 256
    env = fstr_env(n, [f"task.env.ACTION_DELAY={delay}", "vine_randomize=False", "task.env.maxEpisodeLength=1000",
                       "task.env.USE_TARGET_REACHED_RESET=False"])
    env.enable_debug_outputs(True)
    g = torch.Generator(device="cuda").manual_seed(3)
    sent = []
    for t in range(8):
        act = torch.rand(n, 2, device="cuda", generator=g) * 2 - 1
        sent.append(act.clone())
        env.step(act)
        u_rail = env.u_rail_velocity.reshape(-1)
        expect = sent[t - delay][:, 0] * 1.0 if t - delay >= 0 else torch.zeros(n, device="cuda")   # V5:288-291
        assert torch.equal(u_rail, expect), t


def test_long_random_rollout_stays_physical():
    n = 4096
    env = fstr_env(n)
    g = torch.Generator(device="cuda").manual_seed(4)
    resets = timeouts = 0
    for t in range(300):
        od, rew, reset, extras = env.step(torch.rand(n, 2, device="cuda", generator=g) * 2 - 1)
        resets += int(reset.sum()); timeouts += int(extras["time_outs"].sum())
    assert bool(torch.isfinite(env.obs_buf).all()) and bool(torch.isfinite(env.rew_buf).all())
    assert int(env.progress_buf.max()) <= 99                                   # maxEpisodeLength 100
    assert resets > n and 0 < timeouts < resets                               # limit hits + successes + timeouts
    assert float(env.dof_vel.abs().max()) < 200.0 and float(env.dof_pos[:, 1:].abs().max()) < 3.2
    tip = env.tip_positions
    assert float(tip[:, 2].max()) < 1.42 and float(tip[:, 2].min()) > 0.5      # within the chain's reach of the pivot


@pytest.mark.parametrize("preset,n", [("SHELF_OVERRIDES", 16384), ("PIPE_DR_OVERRIDES", 65536), ("SHELF_OVERRIDES", 131072),
                                      ("PIPE_DR_OVERRIDES", 131072)],
                         ids=["configs2_shelf", "configs3_pipe_dr", "shelf_binned_launch", "pipe_dr_binned_launch"])
def test_contact_configs_at_baseline_sizes_are_deterministic_shard_invariant_and_physical(preset, n):
    """BASELINE configs[2] / configs[3] at their full env counts, through size-independent properties: bit-identical
    reruns, bit-identical under 4-way sharding by global env id (the 8-GPU layout of configs[3]), finite and bounded
    state after a random rollout that pushes half the envs into the obstacles. At 131072 envs the whole batches are ROUTED
    (vine_bin_kernel -> near pass || far pass -> redo pass, forced with sim.vine_contact.binning=2) while the four 32768-env
    shards step as single launches in identity order, so the same comparison pins the routed step bit for bit against the
    single launch of the contact variant."""
    import vine_robot_isaacgymenvs_b200 as vine
    from vine_robot_isaacgymenvs_b200 import config as vcfg
    ov = getattr(vcfg, preset) + ["headless=True", "task.env.maxEpisodeLength=40"]
    make = lambda m, off=0, route=0: vine.make(  # noqa: E731
        cfg=vcfg.compose(ov + [f"num_envs={m}", f"+task.sim.vine_contact.binning={route}"]), global_env_offset=off)
    route = 2 if n >= 131072 else 1
    whole, again = make(n, route=route), make(n, route=route)
    shards = [make(n // 4, k * (n // 4)) for k in range(4)]
    g = torch.Generator(device="cuda").manual_seed(11)
    early_resets = 0
    for t in range(60):
        act = torch.rand(n, 2, device="cuda", generator=g) * 2.4 - 1.2
        act[: n // 2, 1] = act[: n // 2, 1].abs()                   # half the envs inflate: pushes the tip into the obstacles
        whole.step(act); again.step(act)
        for k, sh in enumerate(shards):
            sh.step(act[k * (n // 4):(k + 1) * (n // 4)])
        if t % 10 == 9 or t < 3:
            for name in ("obs_buf", "rew_buf", "reset_buf", "progress_buf", "timeout_buf"):
                w = getattr(whole, name)
                assert torch.equal(w, getattr(again, name)), ("determinism", name, t)
                assert torch.equal(w, torch.cat([getattr(sh, name) for sh in shards])), ("sharding", name, t)
        early_resets += int(((whole.reset_buf != 0) & (whole.progress_buf < 39)).sum())    # success / rail limit / contact
    assert bool(torch.isfinite(whole.obs_buf).all()) and bool(torch.isfinite(whole.rew_buf).all())
    assert float(whole.dof_vel.abs().max()) < 400.0 and float(whole.obs_dict["obs"].abs().max()) <= 5.0      # VT:374 clamp
    tip = whole.tip_positions
    assert float(tip[:, 2].max()) < 1.42 and float(tip[:, 2].min()) > 0.5
    assert early_resets > 0


@pytest.mark.parametrize("preset,n", [("SHELF_OVERRIDES", 3000), ("PIPE_DR_OVERRIDES", 20000 + 13)], ids=["shelf", "pipe_dr"])
def test_routed_obstacle_step_equals_the_single_launch_bit_for_bit(preset, n):
    """Obstacle variants: the routed step (bin -> near pass || far pass with the free-space integrator -> redo pass of the envs
    the far pass gave up) against ONE launch of the contact variant over all envs in identity order: every buffer and every
    state plane identical at every step, while envs keep crossing between far and near (half of them are driven into the
    obstacle, episodes end and restart)."""
    import vine_robot_isaacgymenvs_b200 as vine
    from vine_robot_isaacgymenvs_b200 import config as vcfg
    ov = getattr(vcfg, preset) + ["headless=True", "task.env.maxEpisodeLength=30", f"num_envs={n}"]
    routed = vine.make(cfg=vcfg.compose(ov + ["+task.sim.vine_contact.binning=2"]))
    single = vine.make(cfg=vcfg.compose(ov + ["+task.sim.vine_contact.binning=0"]))
    g = torch.Generator(device="cuda").manual_seed(n)
    touched = 0
    for t in range(70):
        act = torch.rand(n, 2, device="cuda", generator=g) * 2.4 - 1.2
        act[: n // 2, 1] = act[: n // 2, 1].abs()
        act[: n // 2, 0] = -act[: n // 2, 0].abs()                  # toward -y, where the obstacles are
        routed.step(act); single.step(act)
        for name in ("obs_buf", "rew_buf", "reset_buf", "progress_buf", "timeout_buf"):
            assert torch.equal(getattr(routed, name), getattr(single, name)), (name, t)
        assert torch.equal(routed.obs_dict["obs"], single.obs_dict["obs"])
        if preset == "SHELF_OVERRIDES":
            touched += int((routed.get_state_dict()["shelf_contact_force"] > 0).sum())
    sa, sb = routed.get_state_dict(), single.get_state_dict()
    for k in sa:
        assert torch.equal(sa[k], sb[k]), k
    assert preset != "SHELF_OVERRIDES" or touched > 0


def test_binned_contact_launch_with_a_ragged_tail_equals_unbinned_shards():
    """98,341 envs (not a multiple of the warp or of the 1024-env binning block) step through the routed step; two shards
    stepping as single launches cover the same global env ids in identity order. Bit-identical outputs and state."""
    import vine_robot_isaacgymenvs_b200 as vine
    from vine_robot_isaacgymenvs_b200 import config as vcfg
    n, n0 = 98304 + 37, 49152
    ov = vcfg.SHELF_OVERRIDES + ["headless=True", "task.env.maxEpisodeLength=30"]
    make = lambda m, off=0, route=0: vine.make(  # noqa: E731
        cfg=vcfg.compose(ov + [f"num_envs={m}", f"+task.sim.vine_contact.binning={route}"]), global_env_offset=off)
    whole, lo, hi = make(n, route=2), make(n0), make(n - n0, n0)
    g = torch.Generator(device="cuda").manual_seed(5)
    for t in range(45):
        act = torch.rand(n, 2, device="cuda", generator=g) * 2.4 - 1.2
        act[: n // 2, 1] = act[: n // 2, 1].abs()
        whole.step(act); lo.step(act[:n0]); hi.step(act[n0:])
    for name in ("obs_buf", "rew_buf", "reset_buf", "progress_buf", "timeout_buf", "dof_pos", "dof_vel", "tip_positions"):
        assert torch.equal(getattr(whole, name), torch.cat([getattr(lo, name), getattr(hi, name)])), name
    assert int((whole.reset_buf != 0).sum()) > 0


@pytest.mark.parametrize("preset", ["FSTR_OVERRIDES", "SHELF_OVERRIDES", "PIPE_DR_OVERRIDES"])
def test_ragged_and_single_env_batches_equal_the_head_of_a_larger_batch(preset):
    """Edge sizes: 1 env and 33 envs (one lane past a warp) produce, bit for bit, what envs 0 and 0..32 of a 200-env batch
    produce (same global env ids, same actions); reset_idx accepts an empty id list and a full one."""
    import vine_robot_isaacgymenvs_b200 as vine
    from vine_robot_isaacgymenvs_b200 import config as vcfg
    ov = getattr(vcfg, preset) + ["headless=True", "task.env.maxEpisodeLength=7"]
    make = lambda m: vine.make(cfg=vcfg.compose(ov + [f"num_envs={m}"]))  # noqa: E731
    big, one, ragged = make(200), make(1), make(33)
    g = torch.Generator(device="cuda").manual_seed(3)
    for t in range(20):
        act = torch.rand(200, 2, device="cuda", generator=g) * 2.4 - 1.2
        act[:, 1] = act[:, 1].abs()
        big.step(act); one.step(act[:1]); ragged.step(act[:33])
        if t == 9:   # an explicit reset of everything, and an empty one, in all three
            for env, m in ((big, 200), (one, 1), (ragged, 33)):
                env.reset_idx(torch.zeros(0, dtype=torch.long, device="cuda"))
                env.reset_idx(torch.arange(m, device="cuda"))
    for name in ("obs_buf", "rew_buf", "reset_buf", "progress_buf", "timeout_buf", "dof_pos", "dof_vel", "target_positions"):
        b = getattr(big, name)
        assert torch.equal(getattr(one, name), b[:1]), (name, "1 env")
        assert torch.equal(getattr(ragged, name), b[:33]), (name, "33 envs")
    assert bool(torch.isfinite(big.obs_buf).all())


def test_metrics_kernel_matches_the_reference_wandb_formulas():
    """vine_metrics vs the formulas of compute_reward's wandb_dict (V5:1250-1322) evaluated with torch on the exposed state."""
    import vine_robot_isaacgymenvs_b200 as vine
    from vine_robot_isaacgymenvs_b200 import config as vcfg
    from vine_robot_isaacgymenvs_b200.tasks.vine5link_moving_base import REWARD_NAMES
    n = 3000
    env = vine.make(cfg=vcfg.compose(vcfg.SHELF_OVERRIDES + [f"num_envs={n}", "headless=True", "task.env.maxEpisodeLength=50"]))
    env.enable_debug_outputs(True)
    g = torch.Generator(device="cuda").manual_seed(6)
    for t in range(25):
        act = torch.rand(n, 2, device="cuda", generator=g) * 2.4 - 1.2
        act[: n // 2, 1] = act[: n // 2, 1].abs()
        env.step(act)
    got = env.metrics()
    st = env.get_state_dict(debug=True)
    R, w = st["reward_matrix"], env.reward_weights.reshape(-1)
    tip, tipvel = st["tip_positions"], st["tip_velocities"]
    want = {
        "dist_tip_to_target": (-R[:, 0]).mean(), "target_reached": (R[:, 2] != 0).float().mean(),
        "limit_hit": (R[:, 9] != 0).float().mean(), "abs_tip_y": tip[:, 1].abs().mean(), "tip_z": tip[:, 2].mean(),
        "max_abs_tip_y": tip[:, 1].abs().max(), "max_tip_z": tip[:, 2].max(), "tip_velocities": tipvel.norm(dim=-1).mean(),
        "tip_velocities_max": tipvel.norm(dim=-1).max(), "u_rail_velocity": st["u_rail_velocity"].abs().mean(),
        "rail_force": st["rail_force"].abs().mean(), "u_fpam": st["u_fpam"].abs().mean(),
        "smoothed_u_fpam": st["smoothed_u_fpam"].abs().mean(), "progress_buf": env.progress_buf.float().mean(),
        "contact_forces": (-R[:, 12]).mean(), "nonzero_contact_force": (R[:, 12] != 0).float().mean(),
        "Mean Total Reward": env.rew_buf.mean(), "Max Total Reward": env.rew_buf.max(),
        "Aggregated Reward": st["aggregated_rew_buf"].mean(),
        "Aggregated Reward 1 Std Up": st["aggregated_rew_buf"].mean() + st["aggregated_rew_buf"].std(),
    }
    for i, name in enumerate(REWARD_NAMES):
        want[f"Mean {name} Reward"], want[f"Max {name} Reward"] = R[:, i].mean(), R[:, i].max()
        want[f"Weighted Mean {name} Reward"], want[f"Weighted Max {name} Reward"] = (R[:, i] * w[i]).mean(), (R[:, i] * w[i]).max()
    assert set(want) <= set(got)
    for k, v in want.items():
        assert abs(got[k] - float(v)) <= 1e-4 * abs(float(v)) + 1e-5, (k, got[k], float(v))
    assert got["contact_forces"] > 0 and got["target_reached"] >= 0


def test_mat_trajectory_replay_overwrites_state_like_the_reference(tmp_path):
    """overwrite_with_mat (V5:947-982): every env gets the recorded DOF positions / target / tip of step num_steps % T."""
    import numpy as np
    import scipy.io
    T = 7
    rng = np.random.default_rng(0)
    mat = {"cart_pos": rng.uniform(-0.2, 0.2, (1, T)), "Q": rng.uniform(-0.3, 0.3, (5, T)),
           "moving_target_pos": np.stack([np.zeros(T), rng.uniform(-0.4, 0.4, T), rng.uniform(0.55, 0.7, T)]),
           "target_vel": np.zeros((3, 1)), "tip_pos": rng.uniform(0.4, 0.6, (3, T)), "tip_vel": np.zeros((3, T))}
    path = str(tmp_path / "traj.mat")
    scipy.io.savemat(path, mat)
    env = fstr_env(64, [f"task.env.MAT_FILE={path}"])
    env.step(torch.zeros(64, 2, device="cuda"))
    env.step(torch.zeros(64, 2, device="cuda"))
    i = env.overwrite_with_mat()
    assert i == 2
    q = env.dof_pos
    assert torch.allclose(q[:, 0], torch.full((64,), float(mat["cart_pos"][0, i]), device="cuda"))
    assert torch.allclose(q[:, 1:], torch.tensor(mat["Q"][:, i], dtype=torch.float32, device="cuda").expand(64, 5))
    assert float(env.dof_vel.abs().max()) == 0.0
    assert torch.allclose(env.target_positions[:, 1:], torch.tensor(mat["moving_target_pos"][1:, i], dtype=torch.float32, device="cuda").expand(64, 2))


def test_state_attributes_write_through_like_the_reference_views():
    """The reference pokes simulator state through tensor views: ``self.dof_pos[env_ids, idx] = ...`` / ``self.dof_vel[env_ids] = 0``
    (V5:780-793) before ``set_dof_state_tensor_indexed``; ``cart_velocities`` (V5:362) feeds the rail controller (V5:1069).
    Here the attributes are snapshots that write in-place modifications back to the library state."""
    import vine_robot_isaacgymenvs_b200 as vine
    from vine_robot_isaacgymenvs_b200 import config as vcfg
    n = 300
    make = lambda: vine.make(cfg=vcfg.compose(vcfg.FSTR_OVERRIDES + [f"num_envs={n}", "headless=True"]))  # noqa: E731
    env, ref = make(), make()
    g = torch.Generator(device="cuda").manual_seed(1)
    a = torch.rand(n, 2, device="cuda", generator=g) * 2 - 1
    for e in (env, ref):
        e.step(a); e.step(a)
    ids = torch.tensor([3, 17, 256], device="cuda")
    new_q = torch.rand(3, 6, device="cuda", generator=g) * 0.2 - 0.1
    # the reference idiom, in place through the attribute
    env.dof_pos[ids] = new_q
    env.dof_vel[ids] = 0.0
    env.dof_pos[ids, 0] += 0.01
    env.smoothed_u_fpam[ids] = 1.5
    cv = env.cart_velocities
    assert cv.shape == (n, 3) and float(cv[:, 0].abs().max()) == 0.0 and float(cv[:, 2].abs().max()) == 0.0
    env.cart_velocities[ids, 1] = 0.25
    # the same through the explicit state API
    st = ref.get_state_dict()
    st["dof_pos"][ids] = new_q
    st["dof_pos"][ids, 0] += 0.01
    st["dof_vel"][ids] = 0.0
    st["smoothed_u_fpam"][ids] = 1.5
    st["cart_body_vel_y"][ids] = 0.25
    ref.set_state_dict(st)
    assert torch.equal(env.dof_pos, ref.dof_pos) and torch.equal(env.dof_pos[ids, 1:], new_q[:, 1:])
    assert torch.equal(env.cart_velocities[:, 1], ref.get_state_dict()["cart_body_vel_y"])
    untouched = torch.ones(n, dtype=torch.bool, device="cuda"); untouched[ids] = False
    for e in (env, ref):
        e.step(a)
    for name in ("obs_buf", "rew_buf", "reset_buf", "progress_buf"):
        assert torch.equal(getattr(env, name), getattr(ref, name)), name
    assert not torch.equal(env.obs_buf[ids], make().obs_buf[ids])
    # outputs of the last step are read-only: an in-place write raises instead of silently doing nothing
    env.enable_debug_outputs(True); env.step(a)
    with pytest.raises(RuntimeError):
        env.rail_force[ids] = 0.0
    with pytest.raises(RuntimeError):
        env.cart_positions[ids, 1] = 0.0
    kept = env.progress_buf > 0                       # an env reset in this step still shows the old episode's body position
    assert env.cart_positions.shape == (n, 3) and torch.equal(env.cart_positions[kept, 1], env.dof_pos[kept, 0])


def test_route_counts_account_for_every_env():
    """vine_route_counts: near + far == num_envs for a routed step, the redone envs are a subset of the far pass, and a step
    that is not routed (free space, or binning off) reports zeros."""
    import ctypes as C
    import vine_robot_isaacgymenvs_b200 as vine
    from vine_robot_isaacgymenvs_b200 import config as vcfg
    n = 20000
    counts = lambda env: (lambda c: (env._lib.vine_route_counts(env._h, c), list(c))[1])((C.c_int64 * 4)())  # noqa: E731
    routed = vine.make(cfg=vcfg.compose(vcfg.SHELF_OVERRIDES + [f"num_envs={n}", "headless=True", "+task.sim.vine_contact.binning=2"]))
    g = torch.Generator(device="cuda").manual_seed(0)
    seen_near = 0
    for t in range(40):
        act = torch.rand(n, 2, device="cuda", generator=g) * 2 - 1
        act[: n // 2, 0] = -act[: n // 2, 0].abs(); act[: n // 2, 1] = act[: n // 2, 1].abs()
        routed.step(act)
        near, far, redone, _ = counts(routed)
        assert near + far == n and 0 <= redone <= far
        seen_near = max(seen_near, near)
    assert seen_near > 0
    free = vine.make(cfg=vcfg.compose(vcfg.FSTR_OVERRIDES + ["num_envs=256", "headless=True"]))
    free.step(torch.zeros(256, 2, device="cuda"))
    assert counts(free) == [0, 0, 0, 0]


@pytest.mark.parametrize("preset", ["FSTR_OVERRIDES", "SHELF_OVERRIDES"])
def test_programmatic_dependent_launch_changes_nothing_but_the_launch(preset):
    """vine_set_programmatic_launch(1): the step kernels are launched with the programmatic-stream-serialization attribute and
    begin with griddepcontrol.wait; outputs and state over 60 back-to-back steps (resets included) must equal the plain
    stream-ordered launches bit for bit, eagerly and from a captured graph."""
    import vine_robot_isaacgymenvs_b200 as vine
    from vine_robot_isaacgymenvs_b200 import abi, config as vcfg
    lib = abi.load_library()
    ov = getattr(vcfg, preset) + ["headless=True", "task.env.maxEpisodeLength=25", "num_envs=4096"]
    plain, pdl = vine.make(cfg=vcfg.compose(ov)), vine.make(cfg=vcfg.compose(ov))
    g = torch.Generator(device="cuda").manual_seed(5)
    acts = [torch.rand(4096, 2, device="cuda", generator=g) * 2 - 1 for _ in range(60)]
    assert lib.vine_set_programmatic_launch(0) == 0          # the library default
    for a in acts[:30]:
        plain.step(a)
    assert lib.vine_set_programmatic_launch(1) == 0
    try:
        for a in acts[:30]:
            pdl.step(a)
        graph = pdl.capture_graph() if hasattr(pdl, "capture_graph") else None
    finally:
        assert lib.vine_set_programmatic_launch(0) == 1
    for a in acts[30:]:
        plain.step(a)
        if graph is not None:
            pdl.actions.copy_(a)
            graph.replay()
        else:
            pdl.step(a)
    torch.cuda.synchronize()
    for name in ("obs_buf", "rew_buf", "reset_buf", "progress_buf"):
        assert torch.equal(getattr(plain, name), getattr(pdl, name)), name
    assert torch.equal(plain.dof_pos, pdl.dof_pos) and torch.equal(plain.dof_vel, pdl.dof_vel)


@pytest.mark.parametrize("preset", ["FSTR_OVERRIDES", "SHELF_OVERRIDES", "PIPE_DR_OVERRIDES"], ids=["fstr", "shelf_routed", "pipe_dr_routed"])
def test_million_env_batch_equals_its_eight_shards_by_checksum(preset):
    """The bench's full size (1,048,576 envs on one GPU: one launch in free space, the routed near/far/redo step with obstacles)
    against the 8-GPU layout of the same global env ids (eight 131,072-env handles with global_env_offset, single launches):
    a checksum of checksums -- the integer sums of the raw bits of obs / rew / reset / progress of the shards add up to the
    whole batch's, at several steps of a rollout with resets and contacts.  (Sums of bit patterns are order independent, so
    the comparison needs no concatenation of 1 M-row buffers; any differing element changes them.)"""
    import vine_robot_isaacgymenvs_b200 as vine
    from vine_robot_isaacgymenvs_b200 import config as vcfg
    n, parts = 1 << 20, 8
    ov = getattr(vcfg, preset) + ["headless=True", "task.env.maxEpisodeLength=30"]
    whole = vine.make(cfg=vcfg.compose(ov + [f"num_envs={n}"]))
    shards = [vine.make(cfg=vcfg.compose(ov + [f"num_envs={n // parts}"]), global_env_offset=k * (n // parts)) for k in range(parts)]

    def checksum(env):
        f = lambda t: int(t.contiguous().view(torch.int32).to(torch.int64).sum())  # noqa: E731
        return (f(env.obs_buf), f(env.rew_buf), int(env.reset_buf.sum()), int(env.progress_buf.sum()), int(env.timeout_buf.sum()))
    g = torch.Generator(device="cuda").manual_seed(3)
    for t in range(45):
        act = torch.rand(n, 2, device="cuda", generator=g) * 2.4 - 1.2
        act[: n // 2, 1] = act[: n // 2, 1].abs()
        whole.step(act)
        for k, sh in enumerate(shards):
            sh.step(act[k * (n // parts):(k + 1) * (n // parts)])
        if t in (0, 1, 14, 31, 44):
            total = [sum(c) for c in zip(*[checksum(sh) for sh in shards])]
            assert list(checksum(whole)) == total, (preset, t)
    assert int(whole.reset_buf.sum()) > 0 and bool(torch.isfinite(whole.obs_buf).all())
    if preset != "FSTR_OVERRIDES":
        from vine_robot_isaacgymenvs_b200 import abi
        import ctypes as C
        counts = (C.c_int64 * 4)()
        lib = abi.load_library()
        lib.vine_route_counts.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
        assert lib.vine_route_counts(whole._h, counts) == 0
        assert counts[0] > 0 and counts[0] + counts[1] == n, list(counts)      # the whole batch really took the routed step
