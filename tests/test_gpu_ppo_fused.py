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
    so = torch.exp(logstd_old)
    kl = (torch.log(sigma / so + 1e-5) + (so ** 2 + (mu - sl(buf["mu_old"])) ** 2) / (2 * (sigma ** 2 + 1e-5)) - 0.5).sum(-1).mean()
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
    n_part = lib.vine_ppo_minibatch(C.byref(mb), None)
    assert n_part > 0, n_part
    out = torch.zeros(P + 4, device="cuda")
    assert lib.vine_ppo_reduce(p(ws), n_part, O, p(out), None) == 0
    torch.cuda.synchronize()
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
        assert lib.vine_ppo_adam(p(out), 1.0, p(flat), p(m), p(v), p(packed), p(state), O, 0.9, 0.999, 1e-8, None) == 0
        torch.cuda.synchronize()
        assert float((flat - ref.detach()).abs().max()) < 1e-6
    assert float(state[3]) == 1.0 and abs(float(state[2]) - float(out[P + 2])) < 1e-7 and float(state[8]) == 3.0
    # the packed block now holds the updated weights: re-pack with vine_mlp_pack and compare byte for byte
    packed_ref = torch.zeros_like(packed)
    Wn = split(flat, W)
    assert lib.vine_mlp_pack(*[p(w.contiguous()) for w in Wn[:10]], O, p(packed_ref), None) == 0
    torch.cuda.synchronize()
    assert torch.equal(packed, packed_ref)
