"""GPU: the fused tcgen05 PPO minibatch kernel (forward + losses + backward + weight gradients), the gradient
reduction and the Adam kernel against a plain PyTorch fp32 reference (autograd + torch.optim.Adam) of the same op.

Tolerances: the kernel multiplies bf16 x bf16 -> f32 (weights, activations and back-propagated dz rounded to bf16);
against an fp32 autograd reference that uses the same bf16-rounded weights, every gradient tensor must agree to
3e-2 in relative Frobenius norm and the loss statistics to 2e-2 relative; Adam itself is f32: 1e-6 absolute.
"""
import ctypes as C
import math

import pytest
import torch

from vine_robot_isaacgymenvs_b200 import abi

pytestmark = pytest.mark.gpu
HYP = dict(e_clip=0.2, critic_coef=2.0, entropy_coef=0.0, bounds_loss_coef=1e-4)


def p(t):
    return C.c_void_p(t.data_ptr())


def make_problem(T, N, O, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    rnd = lambda *s, k=1.0: torch.randn(*s, device="cuda", generator=g) * k  # noqa: E731
    dims = [O, 256, 128, 64]
    W = []
    for i in range(3):
        W += [rnd(dims[i + 1], dims[i], k=(1.5 / dims[i]) ** 0.5), rnd(dims[i + 1], k=0.1)]
    W += [rnd(2, 64, k=0.15), rnd(2, k=0.1), rnd(1, 64, k=0.15), rnd(1, k=0.1), rnd(2, k=0.2)]   # ... logstd
    buf = dict(obs=rnd(T, N, O, k=2.0), act=rnd(T, N, 2), mu_old=rnd(T, N, 2, k=0.5), nlp_old=rnd(T, N, k=0.3) + 2.5,
               val_old=rnd(T, N), ret=rnd(T, N), adv=rnd(T, N))
    mean, inv_std = rnd(O, k=0.3), 1.0 / (0.5 + torch.rand(O, device="cuda", generator=g))
    logstd_old = rnd(2, k=0.2)
    return W, buf, mean, inv_std, logstd_old


def reference(W, buf, mean, inv_std, logstd_old, e0, E, hyp, bf16_weights=True):
    """fp32 autograd restatement of ppo.PPOAgent._update's loss for one env-slice minibatch."""
    r = (lambda t: t.bfloat16().float()) if bf16_weights else (lambda t: t)
    leaves = [w.clone().requires_grad_(True) for w in W]
    w1, b1, w2, b2, w3, b3, wmu, bmu, wv, bv, logstd = leaves
    sl = lambda t: t[:, e0:e0 + E].reshape(-1, *t.shape[2:])  # noqa: E731
    x = torch.clamp((sl(buf["obs"]) - mean) * inv_std, -5, 5)
    # weights enter the GEMMs rounded to bf16 (straight-through for the gradient)
    q = lambda w: w + (r(w) - w).detach()  # noqa: E731
    h = torch.nn.functional.elu(x @ q(w1).t() + b1)
    h = torch.nn.functional.elu(h @ q(w2).t() + b2)
    h = torch.nn.functional.elu(h @ q(w3).t() + b3)
    mu = h @ q(wmu).t() + bmu
    v = (h @ q(wv).t() + bv).squeeze(-1)
    sigma = torch.exp(logstd)
    act, adv, vo, ret = sl(buf["act"]), sl(buf["adv"]), sl(buf["val_old"]), sl(buf["ret"])
    nlp = 0.5 * (((act - mu) / sigma) ** 2).sum(-1) + 0.5 * math.log(2 * math.pi) * 2 + logstd.sum()
    ratio = torch.exp(sl(buf["nlp_old"]) - nlp)
    a_loss = torch.max(-adv * ratio, -adv * torch.clamp(ratio, 1 - hyp["e_clip"], 1 + hyp["e_clip"])).mean()
    v_clip = vo + (v - vo).clamp(-hyp["e_clip"], hyp["e_clip"])
    c_loss = torch.max((v - ret) ** 2, (v_clip - ret) ** 2).mean()
    b_loss = (torch.clamp_min(mu - 1.1, 0) ** 2 + torch.clamp_max(mu + 1.1, 0) ** 2).sum(-1).mean()
    entropy = (0.5 + 0.5 * math.log(2 * math.pi) + logstd).sum()
    loss = a_loss + 0.5 * c_loss * hyp["critic_coef"] - hyp["entropy_coef"] * entropy + b_loss * hyp["bounds_loss_coef"]
    loss.backward()
    # rl_games torch_ext.policy_kl(p0 = current, p1 = old), the restatement the torch update path calls
    from vine_robot_isaacgymenvs_b200.ppo.ppo import policy_kl
    so = torch.exp(logstd_old)
    kl = policy_kl(mu.detach(), sigma.detach().expand_as(mu), sl(buf["mu_old"]), so.expand_as(mu))
    return [l.grad for l in leaves], dict(a_loss=a_loss.item(), c_loss=c_loss.item(), kl=kl.item(), b_loss=b_loss.item()), \
        mu.detach(), v.detach()


def run_kernel(lib, W, buf, mean, inv_std, logstd_old, e0, E, hyp, O, T, N, debug=True):
    flat = torch.cat([w.reshape(-1) for w in W]).contiguous()
    P = lib.vine_ppo_num_params(O)
    assert P == flat.numel()
    packed = torch.zeros(abi.MLP_PACKED_BYTES, dtype=torch.uint8, device="cuda")
    assert lib.vine_mlp_pack(*[p(w.contiguous()) for w in W[:10]], O, p(packed), None) == 0
    ctas = lib.vine_ppo_max_ctas()
    ws = torch.zeros(ctas, abi.PPO_WS_FLOATS, device="cuda")
    state = torch.zeros(abi.PPO_STATE_FLOATS, device="cuda")
    state[0] = 3e-4
    dbg = torch.zeros(T * E, 4, device="cuda") if debug else None
    logstd = flat[P - 2:]
    mb = abi.VinePpoMinibatch(
        packed=packed.data_ptr(), obs=buf["obs"].data_ptr(), actions=buf["act"].data_ptr(), mu_old=buf["mu_old"].data_ptr(),
        neglogp_old=buf["nlp_old"].data_ptr(), values_old=buf["val_old"].data_ptr(), returns=buf["ret"].data_ptr(),
        advantages=buf["adv"].data_ptr(), obs_mean=mean.data_ptr(), obs_inv_std=inv_std.data_ptr(),
        logstd=logstd.data_ptr(), logstd_old=logstd_old.data_ptr(), workspace=ws.data_ptr(), state=state.data_ptr(),
        debug_out=dbg.data_ptr() if debug else None, horizon=T, num_envs=N, env_begin=e0, env_count=E, num_obs=O,
        workspace_ctas=ctas, adaptive_lr=1, kl_threshold=0.008, lr_min=1e-6, lr_max=1e-2, **hyp)
    mu_before = buf["mu_old"].clone()
    n_part = lib.vine_ppo_minibatch(C.byref(mb), None)
    assert n_part > 0, n_part
    out = torch.zeros(P + 4, device="cuda")
    ls_snapshot = torch.full((2,), 7.0, device="cuda")
    assert lib.vine_ppo_reduce(p(ws), n_part, O, p(out), p(logstd), p(ls_snapshot), None, None) == 0
    torch.cuda.synchronize()
    # rl_games dataset.update_mu_sigma: the kernel leaves this pass's mu in the minibatch's rows of mu_old (only there) and
    # the reduce launch snapshots the log-std the minibatch was evaluated with
    buf["mu_after"] = buf["mu_old"].clone()
    buf["mu_old"].copy_(mu_before)
    assert torch.equal(ls_snapshot, logstd)
    outside = torch.ones(N, dtype=torch.bool, device="cuda"); outside[e0:e0 + E] = False
    assert torch.equal(buf["mu_after"][:, outside], mu_before[:, outside])
    if debug:
        assert torch.equal(buf["mu_after"][:, e0:e0 + E].reshape(-1, 2), dbg[:, :2])
    return flat, packed, state, out, dbg, n_part


def split(flat, W):
    out, o = [], 0
    for w in W:
        out.append(flat[o:o + w.numel()].view_as(w))
        o += w.numel()
    return out


NAMES = ["W1", "b1", "W2", "b2", "W3", "b3", "Wmu", "bmu", "Wv", "bv", "logstd"]


@pytest.mark.parametrize("T,N,e0,E,O", [(16, 256, 0, 128, 18), (16, 4096, 2048, 2048, 18), (8, 600, 100, 333, 28)])
def test_fused_minibatch_gradients_match_autograd(T, N, e0, E, O):
    lib = abi.load_library()
    W, buf, mean, inv_std, logstd_old = make_problem(T, N, O, seed=T * N + O)
    flat, packed, state, out, dbg, n_part = run_kernel(lib, W, buf, mean, inv_std, logstd_old, e0, E, HYP, O, T, N)
    grads, stats, mu, v = reference(W, buf, mean, inv_std, logstd_old, e0, E, HYP)
    P = flat.numel()
    assert torch.isfinite(out).all()
    # forward (per sample): mu, v of the kernel vs the fp32 forward with bf16 weights (activations differ by bf16 rounding)
    assert float((dbg[:, :2] - mu).abs().mean()) < 1.5e-2 and float((dbg[:, 2] - v).abs().mean()) < 1.5e-2
    got = split(out[:P], W)
    errs = {}
    for name, g, r in zip(NAMES, got, grads):
        errs[name] = float((g - r).norm() / (r.norm() + 1e-12))
    print("relative gradient errors:", {k: f"{e:.2e}" for k, e in errs.items()}, "partials:", n_part)
    for k, e in errs.items():
        assert e < 3e-2, (k, e, errs)
    for j, k in enumerate(["a_loss", "c_loss", "kl", "b_loss"]):
        assert abs(float(out[P + j]) - stats[k]) <= 2e-2 * abs(stats[k]) + 1e-5, (k, float(out[P + j]), stats[k])
    assert float(state[1]) == 1.0   # the kernel counted one optimiser step


def test_minibatch_gradient_at_the_baseline_size_is_the_mean_of_its_halves():
    """BASELINE configs[1] at its full size (16 x 4096 rows, minibatch 32768 = horizon x 2048 envs) through a size-independent
    property instead of a 32768-row autograd reference: the loss is a mean over rows, so the gradient and the four loss statistics
    of a minibatch are the mean of those of its two env-halves (different tiles, different CTAs, different TMEM accumulation
    chains); tolerance 2e-5 of the gradient's norm (f32 accumulation order), against 3e-2 for the bf16 forward vs autograd."""
    lib = abi.load_library()
    T, N, O = 16, 4096, 18
    W, buf, mean, inv_std, logstd_old = make_problem(T, N, O, seed=123)
    outs = {}
    for name, (e0, E) in {"whole": (0, 2048), "a": (0, 1024), "b": (1024, 1024)}.items():
        flat, _, _, out, _, n_part = run_kernel(lib, W, buf, mean, inv_std, logstd_old, e0, E, HYP, O, T, N, debug=False)
        outs[name] = out.double()
        assert torch.isfinite(out).all() and n_part > 0
    P = lib.vine_ppo_num_params(O)
    mean_of_halves = 0.5 * (outs["a"] + outs["b"])
    g, h = outs["whole"][:P], mean_of_halves[:P]
    assert float(g.norm()) > 1e-3
    assert float((g - h).norm() / g.norm()) < 2e-5, float((g - h).norm() / g.norm())
    for j, k in enumerate(["a_loss", "c_loss", "kl", "b_loss"]):
        w, m = float(outs["whole"][P + j]), float(mean_of_halves[P + j])
        assert abs(w - m) <= 1e-5 * max(abs(w), 1e-3), (k, w, m)


def test_adam_kernel_matches_torch_adam_and_repacks_the_weights():
    lib = abi.load_library()
    T, N, O = 16, 512, 18
    W, buf, mean, inv_std, logstd_old = make_problem(T, N, O, seed=5)
    flat, packed, state, out, _, _ = run_kernel(lib, W, buf, mean, inv_std, logstd_old, 0, 256, HYP, O, T, N, debug=False)
    P = flat.numel()
    ref = flat.clone().requires_grad_(True)
    opt = torch.optim.Adam([ref], lr=3e-4, eps=1e-8)
    m, v = torch.zeros(P, device="cuda"), torch.zeros(P, device="cuda")
    for step in range(3):
        if step:   # what the next fused minibatch launch does: count the optimiser step
            state[1] += 1.0
        ref.grad = out[:P].clone()
        opt.step()
        assert lib.vine_ppo_adam(p(out), 1.0, p(flat), p(m), p(v), p(packed), p(state), O, 0.9, 0.999, 1e-8, 1, None, None) == 0
        torch.cuda.synchronize()
        assert float((flat - ref.detach()).abs().max()) < 1e-6
    assert float(state[3]) == 1.0 and abs(float(state[2]) - float(out[P + 2])) < 1e-7 and float(state[8]) == 3.0
    # the packed block now holds the updated weights: re-pack with vine_mlp_pack and compare byte for byte
    packed_ref = torch.zeros_like(packed)
    Wn = split(flat, W)
    assert lib.vine_mlp_pack(*[p(w.contiguous()) for w in Wn[:10]], O, p(packed_ref), None) == 0
    torch.cuda.synchronize()
    assert torch.equal(packed, packed_ref)


def _act_problem(n, O, seed=3):
    lib = abi.load_library()
    W, buf, mean, inv_std, _ = make_problem(1, n, O, seed=seed)
    packed = torch.zeros(abi.MLP_PACKED_BYTES, dtype=torch.uint8, device="cuda")
    assert lib.vine_mlp_pack(*[p(w.contiguous()) for w in W[:10]], O, p(packed), None) == 0
    return lib, W, buf["obs"][0].contiguous(), mean, inv_std, packed


def _act(lib, packed, obs, mean, inv_std, logstd, counter, n, O, offset=0, seed=1234):
    vstats = torch.tensor([0.3, 1.7], device="cuda")
    out = dict(mu=torch.zeros(n, 2, device="cuda"), value=torch.zeros(n, device="cuda"), actions=torch.zeros(n, 2, device="cuda"),
               neglogp=torch.zeros(n, device="cuda"), obs_copy=torch.zeros(n, O, device="cuda"),
               env_actions=torch.zeros(n, 2, device="cuda"))
    a = abi.VinePolicyAct(packed=packed.data_ptr(), obs=obs.data_ptr(), obs_mean=mean.data_ptr(), obs_inv_std=inv_std.data_ptr(),
                          value_stats=vstats.data_ptr(), logstd=logstd.data_ptr(), rng_counter=counter.data_ptr(), n=n, num_obs=O,
                          seed=seed, global_env_offset=offset, **{k: v.data_ptr() for k, v in out.items()})
    assert lib.vine_policy_act(C.byref(a), None) == 0
    torch.cuda.synchronize()
    return out


def test_policy_act_samples_gaussian_actions_keyed_by_global_env_id():
    n, O = 20000, 18
    lib, W, obs, mean, inv_std, packed = _act_problem(n, O)
    logstd = torch.tensor([-0.3, 0.2], device="cuda")
    counter = torch.zeros(1, dtype=torch.int32, device="cuda")
    o = _act(lib, packed, obs, mean, inv_std, logstd, counter, n, O)
    # the network part equals the inference entry point
    mu2, val2 = torch.zeros(n, 2, device="cuda"), torch.zeros(n, device="cuda")
    vstats = torch.tensor([0.3, 1.7], device="cuda")
    assert lib.vine_mlp_forward(p(packed), p(obs), p(mean), p(inv_std), n, O, p(vstats), p(mu2), p(val2), None) == 0
    torch.cuda.synchronize()
    assert torch.equal(o["mu"], mu2) and torch.equal(o["value"], val2) and torch.equal(o["obs_copy"], obs)
    z = (o["actions"] - o["mu"]) / torch.exp(logstd)
    assert abs(float(z.mean())) < 0.02 and abs(float(z.std()) - 1.0) < 0.02 and abs(float((z[:, 0] * z[:, 1]).mean())) < 0.02
    nlp = 0.5 * (z ** 2).sum(-1) + math.log(2 * math.pi) + logstd.sum()
    assert torch.allclose(o["neglogp"], nlp, atol=2e-4, rtol=1e-4)
    assert torch.equal(o["env_actions"], o["actions"].clamp(-1, 1))
    # same counter -> same noise; next counter -> fresh noise; a shard with the matching global offset -> the same slice
    o2 = _act(lib, packed, obs, mean, inv_std, logstd, counter, n, O)
    assert torch.equal(o2["actions"], o["actions"])
    shard = _act(lib, packed, obs[5000:9000].contiguous(), mean, inv_std, logstd, counter, 4000, O, offset=5000)
    assert torch.equal(shard["actions"], o["actions"][5000:9000])
    counter += 1
    o3 = _act(lib, packed, obs, mean, inv_std, logstd, counter, n, O)
    assert not torch.equal(o3["actions"], o["actions"]) and torch.equal(o3["mu"], o["mu"])


def test_rollout_post_matches_torch():
    lib = abi.load_library()
    n = 5000
    g = torch.Generator(device="cuda").manual_seed(1)
    rew = torch.randn(n, device="cuda", generator=g) * 400 + 300
    resets = (torch.rand(n, device="cuda", generator=g) < 0.3).long()
    timeouts = (torch.rand(n, device="cuda", generator=g) < 0.5) & (resets != 0)
    values = torch.randn(n, device="cuda", generator=g)
    ep_ret, ep_len = torch.randn(n, device="cuda", generator=g) * 10, torch.randint(0, 50, (n,), device="cuda").float()
    ep_ret0, ep_len0 = ep_ret.clone(), ep_len.clone()
    shaped, dones = torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    stats = torch.zeros(4, dtype=torch.float64, device="cuda")
    counter = torch.full((1,), 7, dtype=torch.int32, device="cuda")
    a = abi.VineRolloutPost(rewards=rew.data_ptr(), resets=resets.data_ptr(), timeouts=timeouts.data_ptr(), values=values.data_ptr(),
                            shaped_rewards=shaped.data_ptr(), dones_next=dones.data_ptr(), ep_return=ep_ret.data_ptr(),
                            ep_length=ep_len.data_ptr(), ep_stats=stats.data_ptr(), rng_counter=counter.data_ptr(), n=n,
                            reward_scale=0.01, gamma=0.99, value_bootstrap=1, success_reward_threshold=500.0)
    assert lib.vine_rollout_post(C.byref(a), None) == 0
    torch.cuda.synchronize()
    d = resets.float()
    assert torch.allclose(shaped, rew * 0.01 + 0.99 * values * timeouts.float(), atol=1e-6) and torch.equal(dones, d)
    er, el = ep_ret0 + rew, ep_len0 + 1
    assert torch.equal(ep_ret, er * (1 - d)) and torch.equal(ep_len, el * (1 - d)) and int(counter) == 8
    want = torch.stack([d.sum(), (d * (rew > 500)).sum(), (d * er).double().sum(), (d * el).sum()]).double()
    assert torch.allclose(stats, want, rtol=1e-6)


def test_update_prologue_matches_running_mean_std():
    from vine_robot_isaacgymenvs_b200.ppo.ppo import RunningMeanStd
    lib = abi.load_library()
    count, O = 16 * 1000, 18
    g = torch.Generator(device="cuda").manual_seed(2)
    f = lambda *s: torch.zeros(*s, device="cuda")  # noqa: E731
    obs_rms, val_rms = RunningMeanStd((O,)).cuda(), RunningMeanStd(()).cuda()
    ref_obs, ref_val = RunningMeanStd((O,)).cuda(), RunningMeanStd(()).cuda()
    moments = torch.zeros(2 * O + 4, dtype=torch.float64, device="cuda")
    mean_f, inv_f, vstats, astats, vn, rn, an = f(O), f(O), f(2), f(2), f(count), f(count), f(count)
    for it in range(3):
        obs = torch.randn(count, O, device="cuda", generator=g) * (1 + it) + it
        val, ret = torch.randn(count, device="cuda", generator=g) * 3 + 1, torch.randn(count, device="cuda", generator=g) * 2
        a = abi.VinePpoPrologue(
            obs=obs.data_ptr(), values=val.data_ptr(), returns=ret.data_ptr(), moments=moments.data_ptr(),
            obs_mean=obs_rms.running_mean.data_ptr(), obs_var=obs_rms.running_var.data_ptr(), obs_count=obs_rms.count.data_ptr(),
            val_mean=val_rms.running_mean.data_ptr(), val_var=val_rms.running_var.data_ptr(), val_count=val_rms.count.data_ptr(),
            obs_mean_f=mean_f.data_ptr(), obs_inv_std_f=inv_f.data_ptr(), value_stats=vstats.data_ptr(), adv_stats=astats.data_ptr(),
            values_n=vn.data_ptr(), returns_n=rn.data_ptr(), advantages_n=an.data_ptr(), count=count, num_obs=O, world=1,
            normalize_advantage=1)
        assert lib.vine_ppo_moments(C.byref(a), None) == 0 and lib.vine_ppo_finalize(C.byref(a), None) == 0
        torch.cuda.synchronize()
        ref_obs.update(obs)
        ref_val.update(torch.cat([val, ret]))
        assert torch.allclose(obs_rms.running_mean, ref_obs.running_mean, atol=1e-5)
        assert torch.allclose(obs_rms.running_var, ref_obs.running_var, rtol=1e-5)
        assert torch.allclose(val_rms.running_var, ref_val.running_var, rtol=1e-5) and float(obs_rms.count) == float(ref_obs.count)
        assert torch.allclose(vn, ref_val(val), atol=1e-5) and torch.allclose(rn, ref_val(ret), atol=1e-5)
        adv = ret - val
        assert torch.allclose(an, (adv - adv.mean()) / (adv.std() + 1e-8), atol=1e-4)
        assert torch.allclose(mean_f, ref_obs.running_mean.float(), atol=1e-5)
        assert torch.allclose(inv_f, torch.rsqrt(ref_obs.running_var.float() + 1e-5), rtol=1e-5)
        assert float(moments.abs().max()) == 0.0


def test_p2p_channel_on_one_rank_is_transparent():
    """world = 1: the peer-memory all-reduce (csrc/vine_p2p.cuh) degenerates to "the reduce kernel writes into the region's
    buffer seq & 1, the Adam kernel reads it back": parameters after several optimiser steps are bit-identical to the plain
    path, the sequence number advances once per exchange, nothing times out."""
    from vine_robot_isaacgymenvs_b200 import distributed as vd
    lib = abi.load_library()
    T, N, O = 16, 512, 18
    W, buf, mean, inv_std, logstd_old = make_problem(T, N, O, seed=9)
    results = []
    for use_channel in (False, True):
        flat = torch.cat([w.reshape(-1) for w in W]).contiguous()
        P = flat.numel()
        packed = torch.zeros(abi.MLP_PACKED_BYTES, dtype=torch.uint8, device="cuda")
        assert lib.vine_mlp_pack(*[p(w.contiguous()) for w in split(flat, W)[:10]], O, p(packed), None) == 0
        ctas = lib.vine_ppo_max_ctas()
        ws = torch.zeros(ctas, abi.PPO_WS_FLOATS, device="cuda")
        state = torch.zeros(abi.PPO_STATE_FLOATS, device="cuda"); state[0] = 3e-4
        m, v, out = torch.zeros(P, device="cuda"), torch.zeros(P, device="cuda"), torch.zeros(P + 4, device="cuda")
        ch = vd.P2PChannel(lib, P + 4, "cuda") if use_channel else None
        mu_old = buf["mu_old"].clone()
        for k in range(5):                                   # 5 exchanges: both buffers of the region get reused
            e0 = (k % 2) * 256
            mb = abi.VinePpoMinibatch(
                packed=packed.data_ptr(), obs=buf["obs"].data_ptr(), actions=buf["act"].data_ptr(), mu_old=mu_old.data_ptr(),
                neglogp_old=buf["nlp_old"].data_ptr(), values_old=buf["val_old"].data_ptr(), returns=buf["ret"].data_ptr(),
                advantages=buf["adv"].data_ptr(), obs_mean=mean.data_ptr(), obs_inv_std=inv_std.data_ptr(),
                logstd=flat[P - 2:].data_ptr(), logstd_old=logstd_old.data_ptr(), workspace=ws.data_ptr(), state=state.data_ptr(),
                debug_out=None, horizon=T, num_envs=N, env_begin=e0, env_count=256, num_obs=O, workspace_ctas=ctas, adaptive_lr=1,
                kl_threshold=0.008, lr_min=1e-6, lr_max=1e-2, **HYP)
            n_part = lib.vine_ppo_minibatch(C.byref(mb), None)
            assert n_part > 0
            assert lib.vine_ppo_reduce(p(ws), n_part, O, p(out), None, None, ch.ptr if ch else None, None) == 0
            assert lib.vine_ppo_adam(p(out), 1.0, p(flat), p(m), p(v), p(packed), p(state), O, 0.9, 0.999, 1e-8, 1,
                                     ch.ptr if ch else None, None) == 0
        torch.cuda.synchronize()
        if ch is not None:
            assert ch.status() == (5, False)
            ch.close()
        results.append((flat.clone(), m.clone(), v.clone(), packed.clone(), state.clone()))
    for a, b in zip(*results):
        assert torch.equal(a, b)
