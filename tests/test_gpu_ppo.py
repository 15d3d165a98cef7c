"""GPU: the PPO iteration (rollout + GAE + update) runs through the fused env step for both network variants,
eager and CUDA-graph-captured, stays finite, and checkpoints round-trip in rl_games' layout."""
import pytest
import torch

import vine_robot_isaacgymenvs_b200 as vine
from vine_robot_isaacgymenvs_b200 import config as vcfg
from vine_robot_isaacgymenvs_b200.ppo.ppo import PPOAgent

pytestmark = pytest.mark.gpu


def make_agent(extra, graphs, n=512, **kw):
    cfg = vcfg.compose(vcfg.FSTR_OVERRIDES + [f"num_envs={n}", "headless=True", "train.params.config.minibatch_size=4096"]
                       + extra)
    env = vine.make(cfg=cfg)
    return PPOAgent(env, cfg["train"], seed=1, use_graphs=graphs, **kw)


MLP = ["train.params.network.rnn=null"]


@pytest.mark.parametrize("variant", ["lstm_native", "lstm_torch", "mlp_fused", "mlp_torch"])
@pytest.mark.parametrize("graphs", [False, True], ids=["eager", "graphs"])
def test_ppo_iterations_are_finite_and_learn_something(variant, graphs):
    rnn = variant.startswith("lstm")
    native = variant in ("lstm_native", "mlp_fused")
    agent = make_agent([] if rnn else MLP, graphs, use_fused_update=native)
    assert agent.has_rnn == rnn and agent.seq_len == (4 if rnn else 1)
    assert agent.fused_update == native and agent.fused == (not rnn) and agent.native_lstm == (variant == "lstm_native")
    before = [p.detach().clone() for p in agent.model.parameters()]
    for _ in range(6):
        agent.train_epoch()
    torch.cuda.synchronize()
    st = agent.pop_stats()
    assert st["episodes"] > 0 and all(x == x and abs(x) < 1e6 for x in (st["a_loss"], st["c_loss"], st["kl"]))
    assert all(torch.isfinite(p).all() for p in agent.model.parameters())
    assert any(not torch.equal(a, b) for a, b in zip(before, agent.model.parameters()))
    assert float(agent.obs_rms.count) > 1.0 and agent.frames == (6 + (3 if graphs else 0)) * 16 * 512  # +3: graph warm-up
    assert agent.epoch == 6 + (3 if graphs else 0)          # warm-up iterations are real training iterations and are counted


def test_fused_update_moves_the_parameters_like_the_torch_update():
    """Same rollout, one minibatch, one Adam step: hand-written kernels vs torch autograd + torch.optim.Adam."""
    one = ["train.params.config.mini_epochs=1", "train.params.config.minibatch_size=8192"]
    a = make_agent(MLP + one, False, use_fused_update=True)
    b = make_agent(MLP + one, False, use_fused_update=False)
    b.load_state_dict(a.state_dict())
    start = a.flat.clone()
    a._rollout()
    for name in ("b_obs", "b_act", "b_mu", "b_nlp", "b_val", "b_ret", "b_adv", "b_done"):
        getattr(b, name).copy_(getattr(a, name))
    a._update_any()
    b._update_any()
    torch.cuda.synchronize()
    after_b = torch.cat([p.detach().reshape(-1) for p in b._param_order()])
    da, db = a.flat - start, after_b - start
    cos = float((da * db).sum() / (da.norm() * db.norm()))
    assert cos > 0.97, cos                                  # Adam's first step is sign(g) * lr: agreement of signs
    assert abs(float(da.abs().max()) - 3e-4) < 1e-5 and abs(float(db.abs().max()) - 3e-4) < 1e-5
    sa, sb = a.pop_stats(), b.pop_stats()
    for k in ("a_loss", "c_loss", "kl"):
        assert abs(sa[k] - sb[k]) <= 3e-2 * abs(sb[k]) + 1e-4, (k, sa[k], sb[k])
    assert torch.allclose(a.obs_rms.running_mean, b.obs_rms.running_mean, atol=1e-6)


@pytest.mark.parametrize("net", ["mlp", "lstm"])
def test_full_update_kl_and_learning_rate_follow_the_torch_update(net):
    """A whole update (4 mini-epochs x 2 minibatches) from the same rollout: the kernel paths and the torch path apply rl_games'
    update_mu_sigma after every minibatch (mini-epoch k measures its KL against mini-epoch k-1, not against the rollout), use
    the same policy_kl argument order, and therefore report the same mean KL and take the same adaptive-lr decisions."""
    two = ["train.params.config.minibatch_size=4096"]
    extra = (MLP if net == "mlp" else []) + two
    a = make_agent(extra, False, use_fused_update=True)
    b = make_agent(extra, False, use_fused_update=False)
    b.load_state_dict(a.state_dict())
    for _ in range(2):                                   # a second iteration: non-trivial normalisers and sigma
        a._rollout()
        for name in ("b_obs", "b_act", "b_mu", "b_nlp", "b_val", "b_ret", "b_adv", "b_done"):
            getattr(b, name).copy_(getattr(a, name))
        if net == "lstm":
            from vine_robot_isaacgymenvs_b200.ppo.tiles import from_tiles
            for ck in range(a.T // a.seq_len):
                b.b_h[ck].copy_(from_tiles(a._HH_saved[ck], a.n))
                b.b_c[ck].copy_(a._C_saved[ck])
        mu_rollout = a.b_mu.clone()
        a.pop_stats(); b.pop_stats()
        a._update_any()
        b._update_any()
        torch.cuda.synchronize()
        sa, sb = a.pop_stats(), b.pop_stats()
        assert sb["kl"] > 0
        assert abs(sa["kl"] - sb["kl"]) <= 0.15 * sb["kl"] + 2e-5, (sa["kl"], sb["kl"])
        # same number of x1.5 / /1.5 decisions (bf16 GEMMs may move one borderline minibatch across a threshold)
        import math
        assert abs(math.log(a.lr / b.lr) / math.log(1.5)) <= 1.01, (a.lr, b.lr)
        # update_mu_sigma happened: the "old" mean of the rows now is the last mini-epoch's, in both paths
        mu_a = a._scal[..., 2:4] if net == "lstm" else a.b_mu
        assert float((mu_a - mu_rollout).abs().max()) > 1e-4
        assert float((mu_a - b.b_mu).abs().mean()) < 3e-2
        # the sigma snapshot of every minibatch slot is the log-std BEFORE the last Adam step of that slot
        assert a._logstd_old.shape == (2, 2) and float((a._logstd_old - a.model.sigma).abs().max()) < 2e-2
        b.load_state_dict(a.state_dict())


def test_native_lstm_update_moves_the_parameters_like_the_torch_update():
    """Reference network, same rollout, one minibatch, one Adam step: kernels-only path vs torch autograd + torch Adam."""
    one = ["train.params.config.mini_epochs=1", "train.params.config.minibatch_size=8192"]
    a = make_agent(one, False, use_fused_update=True)
    b = make_agent(one, False, use_fused_update=False)
    assert a.native_lstm and not b.native_lstm
    b.load_state_dict(a.state_dict())
    start = torch.cat([a.flat[:a.flat.numel() - 197].clone(), a.flat_l.clone()])     # MLP half without the dummy heads
    a._rollout()
    for name in ("b_obs", "b_act", "b_mu", "b_nlp", "b_val", "b_ret", "b_adv", "b_done"):
        getattr(b, name).copy_(getattr(a, name))
    from vine_robot_isaacgymenvs_b200.ppo.tiles import from_tiles
    for ck in range(a.T // a.seq_len):                                               # LSTM state at the chunk starts
        b.b_h[ck].copy_(from_tiles(a._HH_saved[ck], a.n))
        b.b_c[ck].copy_(a._C_saved[ck])
    a._update_any()
    b._update_any()
    torch.cuda.synchronize()
    mb = [b.model.actor_mlp[0].weight, b.model.actor_mlp[0].bias, b.model.actor_mlp[2].weight, b.model.actor_mlp[2].bias,
          b.model.actor_mlp[4].weight, b.model.actor_mlp[4].bias]
    rb = b.model.rnn.rnn
    lb = [rb.weight_ih_l0, rb.weight_hh_l0, rb.bias_ih_l0, rb.bias_hh_l0, b.model.layer_norm.weight, b.model.layer_norm.bias,
          b.model.mu.weight, b.model.mu.bias, b.model.value.weight, b.model.value.bias, b.model.sigma]
    after_b = torch.cat([p.detach().reshape(-1) for p in mb + lb])
    after_a = torch.cat([a.flat[:a.flat.numel() - 197], a.flat_l])
    da, db = after_a - start, after_b - start
    cos = float((da * db).sum() / (da.norm() * db.norm()))
    assert cos > 0.9, cos                                      # first Adam step = lr * sign(g): agreement of gradient signs
    sa, sb = a.pop_stats(), b.pop_stats()
    for k in ("a_loss", "c_loss", "kl"):
        assert abs(sa[k] - sb[k]) <= 5e-2 * abs(sb[k]) + 2e-4, (k, sa[k], sb[k])


def test_checkpoint_roundtrip_uses_rl_games_layout(tmp_path):
    a = make_agent([], False)
    a.train_epoch()
    sd = a.state_dict()
    assert "a2c_network.rnn.rnn.weight_hh_l0" in sd["model"] and "running_mean_std.running_mean" in sd["model"]
    assert "value_mean_std.running_var" in sd["model"] and {"epoch", "frame", "optimizer", "last_lr"} <= set(sd)
    # rl_games 1.5.2 RunningMeanStd((value_size,)): the value normaliser is [1], its count a scalar
    assert sd["model"]["value_mean_std.running_mean"].shape == (1,) and sd["model"]["value_mean_std.count"].shape == ()
    assert sd["model"]["running_mean_std.running_mean"].shape == (18,) and "vine_rng_counter" in sd
    torch.save(sd, tmp_path / "ck.pth")
    b = make_agent([], False)
    b.load_state_dict(torch.load(tmp_path / "ck.pth", map_location="cuda:0"))
    for pa, pb in zip(a.model.parameters(), b.model.parameters()):
        assert torch.equal(pa, pb)
    assert torch.equal(a.obs_rms.running_mean, b.obs_rms.running_mean) and b.epoch == 1
    # fused-update agents: parameters are views of one flat vector, optimiser moments travel too
    c = make_agent(MLP, False)
    c.train_epoch()
    torch.save(c.state_dict(), tmp_path / "ck2.pth")
    d = make_agent(MLP, False)
    d.load_state_dict(torch.load(tmp_path / "ck2.pth", map_location="cuda:0"))
    assert torch.equal(c.flat, d.flat) and torch.equal(c.adam_m, d.adam_m) and torch.equal(c._packed, d._packed)
    assert float(d.ppo_state[1]) == float(c.ppo_state[1]) > 0 and abs(d.lr - c.lr) < 1e-12


def test_fused_lstm_cell_kernels_match_the_torch_restatement():
    """csrc/vine_lstm.cu + bf16 GEMMs (ppo/lstm_ops.py) vs the pure-torch fp32 loop of the same module: outputs, final
    state and every gradient (tolerance: bf16 GEMM operands, relative Frobenius error <= 2e-2)."""
    from vine_robot_isaacgymenvs_b200.ppo.ppo import ActorCritic
    rnn = {"name": "lstm", "units": 256, "layers": 1, "before_mlp": False, "concat_input": True, "layer_norm": True}
    torch.manual_seed(3)
    m = ActorCritic(18, 2, (256, 128, 64), rnn=rnn).cuda()
    L, S = 4, 2048
    obs = torch.randn(L, S, 18, device="cuda")
    h0 = (torch.randn(S, 256, device="cuda") * 0.5).requires_grad_(True)
    c0 = (torch.randn(S, 256, device="cuda") * 0.5).requires_grad_(True)
    nd = (torch.rand(L, S, device="cuda") > 0.2).float()
    wts = torch.randn(L * S, 3, device="cuda")

    def run(fused):
        m.fused_cell = fused
        m.zero_grad()
        for t in (h0, c0):
            t.grad = None
        mu, _, v, (h, c) = m(obs, (h0, c0), nd)
        loss = (mu * wts[:, :2]).sum() + (v.squeeze(-1) * wts[:, 2]).sum() + 0.3 * h.sum() - 0.2 * c.sum()
        loss.backward()
        grads = {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}
        grads["h0"], grads["c0"] = h0.grad.clone(), c0.grad.clone()
        return mu.detach(), v.detach(), h.detach(), c.detach(), grads

    mu_r, v_r, h_r, c_r, g_r = run(False)
    mu_f, v_f, h_f, c_f, g_f = run(True)
    rel = lambda a, b: float((a - b).norm() / (b.norm() + 1e-12))  # noqa: E731
    assert rel(mu_f, mu_r) < 2e-2 and rel(v_f, v_r) < 2e-2 and rel(h_f, h_r) < 1e-2 and rel(c_f, c_r) < 1e-2
    assert set(g_f) == set(g_r)
    errs = {k: rel(g_f[k], g_r[k]) for k in g_r}
    assert max(errs.values()) < 2e-2, errs


def test_train_entry_point_trains_saves_and_plays_back(tmp_path, monkeypatch):
    """``python -m vine_robot_isaacgymenvs_b200.train`` surface (reference train.py:35-171): a short training run writes
    runs/<name>/nn/<name>.pth in rl_games' layout, and ``test=True checkpoint=...`` replays the policy without learning."""
    from vine_robot_isaacgymenvs_b200 import train
    monkeypatch.chdir(tmp_path)
    base = vcfg.FSTR_OVERRIDES + ["num_envs=512", "headless=True", "train.params.config.minibatch_size=4096"]
    hist = train.launch(base + ["max_iterations=4"])
    assert len(hist) >= 1 and hist[-1]["epoch"] >= 4 and hist[-1]["frames"] > 0
    ck = tmp_path / "runs" / "Vine5LinkMovingBase" / "nn" / "Vine5LinkMovingBase.pth"
    assert ck.exists()
    sd = torch.load(ck, map_location="cpu")
    assert "a2c_network.rnn.rnn.weight_ih_l0" in sd["model"] and "running_mean_std.running_var" in sd["model"]
    stats = train.launch(base + ["test=True", f"checkpoint={ck}", "play_steps=64"])
    assert stats["episodes"] > 0 and 0.0 <= stats["success_rate"] <= 1.0
    # the deployment-side player restores the same file
    from vine_robot_isaacgymenvs_b200.player import PolicyPlayer
    pl = PolicyPlayer(18, device="cuda").restore(str(ck))
    a = pl.get_action(torch.zeros(18, device="cuda"), is_deterministic=True)
    assert a.shape == (2,) and bool(torch.isfinite(a).all())


def test_native_paths_run_on_the_default_task_config():
    """The reference's default task config (pipe obstacle, POS_AND_FD_VEL_AND_OBJ_INFO = 28 observations) through both
    kernel-only paths: a few PPO iterations stay finite and move the parameters."""
    for extra in ([], MLP):
        cfg = vcfg.compose(["num_envs=512", "headless=True", "train.params.config.minibatch_size=4096",
                            "task.env.maxEpisodeLength=50"] + extra)
        agent = PPOAgent(vine.make(cfg=cfg), cfg["train"], seed=3, use_graphs=False)
        assert agent.O == 28 and agent.fused_update and agent.native_lstm == (not extra)
        w0 = agent.model.actor_mlp[0].weight.detach().clone()
        for _ in range(3):
            agent.train_epoch()
        torch.cuda.synchronize()
        st = agent.pop_stats()
        assert all(x == x and abs(x) < 1e6 for x in (st["a_loss"], st["c_loss"], st["kl"]))
        assert all(torch.isfinite(p).all() for p in agent.model.parameters())
        assert not torch.equal(w0, agent.model.actor_mlp[0].weight)


@pytest.mark.parametrize("extra", [[], MLP], ids=["reference_network", "mlp"])
def test_fstr_training_reaches_the_success_band(extra):
    """Learning check (the only parity available for the PPO side: rl_games is not vendored): FSTR, 4096 envs, 250 iterations
    on the kernel-only paths.  Band: the torch-update agent reaches 0.80-0.85 success at this point
    (profiles/ppo_fstr_r01_training*.log); the kernel paths must be within 0.10 of that, i.e. >= 0.72, with episodes getting
    shorter than the 100-step time-out."""
    cfg = vcfg.compose(vcfg.FSTR_OVERRIDES + ["num_envs=4096", "headless=True"] + extra)
    agent = PPOAgent(vine.make(cfg=cfg), cfg["train"], seed=42, use_graphs=True)
    hist = agent.train(250, log_every=50, log=None)
    last = hist[-1]
    assert last["success_rate"] >= 0.72 and last["mean_length"] < 60 and last["mean_return"] > 600, last
    assert hist[0]["success_rate"] < last["success_rate"]
