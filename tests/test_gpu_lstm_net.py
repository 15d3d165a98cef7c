"""GPU: the recurrent half of the reference network on hand-written kernels (vine_policy_act -> U tiles, vine_lstm_step,
vine_lstm_mask, vine_lstm_head) vs the plain PyTorch fp32 module (ppo.ActorCritic with the reference's rnn block).
Tolerance: bf16 GEMM operands and bf16-stored hidden state: |d| <= 3e-2 + 3e-2 |ref| on mu / value, 2e-2 abs on h, c."""
import ctypes as C
import math

import pytest
import torch

from vine_robot_isaacgymenvs_b200 import abi
from vine_robot_isaacgymenvs_b200.ppo.ppo import ActorCritic
from vine_robot_isaacgymenvs_b200.ppo.tiles import from_tiles, to_tiles

pytestmark = pytest.mark.gpu
RNN = {"name": "lstm", "units": 256, "layers": 1, "before_mlp": False, "concat_input": True, "layer_norm": True}


def p(t):
    return C.c_void_p(t.data_ptr())


def test_tile_conversion_roundtrip_and_layout():
    x = torch.randn(300, 256, device="cuda")
    t = to_tiles(x)
    assert t.shape == (3, 2, 128 * 128) and torch.equal(from_tiles(t, 300), x.bfloat16().float())
    # element (row 13, col 200) of tile 0 lives in part 1 at byte (13/8)*2048 + ((200-128)/8)*128 + (13%8)*16 + (200%8)*2
    off = (13 // 8) * 2048 + ((200 - 128) // 8) * 128 + (13 % 8) * 16 + (200 % 8) * 2
    assert float(t[0, 1, off // 2]) == float(x[13, 200].bfloat16())


@pytest.mark.parametrize("n,O", [(256, 18), (1000, 28)])
def test_native_lstm_forward_matches_torch_over_several_steps(n, O):
    lib = abi.load_library()
    torch.manual_seed(n)
    m = ActorCritic(O, 2, (256, 128, 64), rnn=RNN).cuda()
    with torch.no_grad():      # make LayerNorm and biases non-trivial
        m.layer_norm.weight.uniform_(0.5, 1.5); m.layer_norm.bias.uniform_(-0.2, 0.2)
        m.mu.weight.mul_(3.0); m.value.weight.mul_(3.0)
    mean, inv_std = torch.randn(O, device="cuda") * 0.3, 1.0 / (0.5 + torch.rand(O, device="cuda"))
    vstats = torch.tensor([0.3, 1.7], device="cuda")
    packed = torch.zeros(abi.MLP_PACKED_BYTES, dtype=torch.uint8, device="cuda")
    z = lambda *s: torch.zeros(*s, device="cuda")  # noqa: E731
    mlp = [m.actor_mlp[0].weight, m.actor_mlp[0].bias, m.actor_mlp[2].weight, m.actor_mlp[2].bias, m.actor_mlp[4].weight,
           m.actor_mlp[4].bias, z(2, 64), z(2), z(1, 64), z(1)]
    assert lib.vine_mlp_pack(*[p(t.detach().contiguous()) for t in mlp], O, p(packed), None) == 0
    lpacked = torch.zeros(abi.LSTM_PACKED_BYTES, dtype=torch.uint8, device="cuda")
    r = m.rnn.rnn
    lp = [r.weight_ih_l0, r.weight_hh_l0, r.bias_ih_l0, r.bias_hh_l0, m.layer_norm.weight, m.layer_norm.bias, m.mu.weight,
          m.mu.bias, m.value.weight, m.value.bias]
    assert lib.vine_lstm_pack(*[p(t.detach().contiguous()) for t in lp], O, p(lpacked), None) == 0
    tiles = (n + 127) // 128
    U = torch.zeros(tiles, abi.LSTM_TILE_BYTES // 2, dtype=torch.bfloat16, device="cuda")
    HM = torch.zeros(tiles, 2, abi.LSTM_TILE_BYTES // 2, dtype=torch.bfloat16, device="cuda")
    ACT = torch.zeros(tiles, 16, 128 * 64, dtype=torch.bfloat16, device="cuda")
    h_ref, c_ref = torch.randn(n, 256, device="cuda") * 0.5, torch.randn(n, 256, device="cuda") * 0.5
    HH = to_tiles(h_ref)
    c_k = c_ref.clone()
    mu_k, val_k = z(n, 2), z(n)
    for step in range(4):
        obs = torch.randn(n, O, device="cuda") * 2
        nd = (torch.rand(n, device="cuda") > 0.25).float()
        with torch.no_grad():
            x = torch.clamp((obs - mean) * inv_std, -5, 5)
            mu_r, _, v_r, (h_ref, c_ref) = m(x.unsqueeze(0), (h_ref, c_ref), nd.unsqueeze(0))
            val_r = torch.clamp(v_r.squeeze(-1), -5, 5) * 1.7 + 0.3
        act = abi.VinePolicyAct(packed=packed.data_ptr(), obs=obs.data_ptr(), obs_mean=mean.data_ptr(), obs_inv_std=inv_std.data_ptr(),
                                value_stats=vstats.data_ptr(), n=n, num_obs=O, u_out=U.data_ptr())
        assert lib.vine_policy_act(C.byref(act), None) == 0
        assert lib.vine_lstm_mask(p(HH), p(nd), n, p(HM), None) == 0
        c_new = torch.empty_like(c_k)
        HH_new = torch.zeros_like(HH)
        st = abi.VineLstmStep(params=lpacked.data_ptr(), u=U.data_ptr(), hm=HM.data_ptr(), c_prev=c_k.data_ptr(), not_done=nd.data_ptr(),
                              c=c_new.data_ptr(), hh=HH_new.data_ptr(), act=ACT.data_ptr(), n=n)
        assert lib.vine_lstm_step(C.byref(st), None) == 0
        hd = abi.VineLstmHead(params=lpacked.data_ptr(), hh=HH_new.data_ptr(), value_stats=vstats.data_ptr(), mu=mu_k.data_ptr(),
                              value=val_k.data_ptr(), n=n)
        assert lib.vine_lstm_head(C.byref(hd), None) == 0
        torch.cuda.synchronize()
        HH, c_k = HH_new, c_new
        h_k = from_tiles(HH, n)
        # U tile content: [h3 | x | 0], ones column at 95
        u = from_tiles(U.view(tiles, 1, -1), n)
        assert float((u[:, 64:64 + O] - x).abs().max()) < 2e-2 and float((u[:, 95] - 1).abs().max()) == 0 and float(u[:, 96:].abs().max()) == 0
        assert float((h_k - h_ref).abs().max()) < 3e-2 and float((c_k - c_ref).abs().max()) < 3e-2, step
        assert bool(((mu_k - mu_r).abs() <= 3e-2 + 3e-2 * mu_r.abs()).all()), (step, float((mu_k - mu_r).abs().max()))
        assert bool(((val_k - val_r).abs() <= 5e-2 + 3e-2 * val_r.abs()).all()), (step, float((val_k - val_r).abs().max()))
        h_ref, c_ref = h_k.clone(), c_k.clone()      # keep the two recurrences from drifting apart (bf16 state)
    # activated gates of the last step are consistent with c: c = f * c_prev * nd + i * g is checked implicitly; sanity: in (0,1) / (-1,1)
    a = ACT.view(tiles, 16, 16, 8, 8, 8).float()
    assert float(a.min()) >= -1.0 and float(a.max()) <= 1.0


def test_lstm_head_sampling_matches_policy_act_conventions():
    lib = abi.load_library()
    n = 4096
    torch.manual_seed(1)
    m = ActorCritic(18, 2, (256, 128, 64), rnn=RNN).cuda()
    lpacked = torch.zeros(abi.LSTM_PACKED_BYTES, dtype=torch.uint8, device="cuda")
    r = m.rnn.rnn
    lp = [r.weight_ih_l0, r.weight_hh_l0, r.bias_ih_l0, r.bias_hh_l0, m.layer_norm.weight, m.layer_norm.bias, m.mu.weight,
          m.mu.bias, m.value.weight, m.value.bias]
    assert lib.vine_lstm_pack(*[p(t.detach().contiguous()) for t in lp], 18, p(lpacked), None) == 0
    HH = to_tiles(torch.randn(n, 256, device="cuda"))
    vstats = torch.tensor([0.0, 1.0], device="cuda")
    logstd = torch.tensor([-0.2, 0.1], device="cuda")
    ctr = torch.zeros(1, dtype=torch.int32, device="cuda")
    out = {k: torch.zeros(n, 2, device="cuda") for k in ("mu", "actions", "env_actions")}
    out["neglogp"], out["value"] = torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    hd = abi.VineLstmHead(params=lpacked.data_ptr(), hh=HH.data_ptr(), value_stats=vstats.data_ptr(), logstd=logstd.data_ptr(),
                          rng_counter=ctr.data_ptr(), n=n, seed=99, global_env_offset=0, **{k: v.data_ptr() for k, v in out.items()})
    assert lib.vine_lstm_head(C.byref(hd), None) == 0
    torch.cuda.synchronize()
    zz = (out["actions"] - out["mu"]) / torch.exp(logstd)
    assert abs(float(zz.mean())) < 0.05 and abs(float(zz.std()) - 1) < 0.05
    assert torch.allclose(out["neglogp"], 0.5 * (zz ** 2).sum(-1) + math.log(2 * math.pi) + logstd.sum(), atol=2e-4)
    assert torch.equal(out["env_actions"], out["actions"].clamp(-1, 1))


def _packed_gate_perm():
    """index tensor: packed gate row R -> torch gate row (piece p row g*16+k = gate g of unit 16p+k)."""
    R = torch.arange(1024)
    p_, r = R // 64, R % 64
    return (r // 16) * 256 + 16 * p_ + (r % 16)


def test_native_lstm_step_backward_chain_matches_autograd():
    """head_train -> cell_bwd_tiles -> bwd_gemm for one time step vs torch autograd of the same step (fp32 math on the
    bf16-rounded weights/inputs).  Tolerance: 3e-2 relative Frobenius error per gradient tensor."""
    lib = abi.load_library()
    n, O = 640, 18
    torch.manual_seed(7)
    dev = "cuda"
    q = lambda t: t.bfloat16().float()  # noqa: E731
    m = ActorCritic(O, 2, (256, 128, 64), rnn=RNN).to(dev)
    with torch.no_grad():
        m.layer_norm.weight.uniform_(0.5, 1.5); m.layer_norm.bias.uniform_(-0.2, 0.2)
        m.mu.weight.mul_(3.0); m.value.weight.mul_(3.0); m.sigma.uniform_(-0.3, 0.1)
    r = m.rnn.rnn
    lpacked = torch.zeros(abi.LSTM_PACKED_BYTES, dtype=torch.uint8, device=dev)
    lp = [r.weight_ih_l0, r.weight_hh_l0, r.bias_ih_l0, r.bias_hh_l0, m.layer_norm.weight, m.layer_norm.bias, m.mu.weight,
          m.mu.bias, m.value.weight, m.value.bias]
    assert lib.vine_lstm_pack(*[p(t.detach().contiguous()) for t in lp], O, p(lpacked), None) == 0
    tiles = (n + 127) // 128
    # inputs of the step
    u_real = q(torch.randn(n, 64 + O, device=dev))
    u_full = torch.zeros(n, 128, device=dev); u_full[:, :64 + O] = u_real; u_full[:, 95] = 1.0
    h_prev = q(torch.randn(n, 256, device=dev) * 0.5)
    c_prev = torch.randn(n, 256, device=dev) * 0.5
    nd = (torch.rand(n, device=dev) > 0.3).float()
    scal = torch.randn(n, 8, device=dev); scal[:, 4] = scal[:, 4] * 0.3 + 2.5
    logstd_old = torch.tensor([-0.1, 0.05], device=dev)
    hyp = dict(e_clip=0.2, critic_coef=2.0, entropy_coef=0.0, bounds_loss_coef=1e-4)
    # ---- kernels ----
    U, HM = to_tiles(u_full).view(tiles, -1).contiguous(), to_tiles(h_prev * nd[:, None]).contiguous()
    HH, ACT = torch.zeros_like(HM), torch.zeros(tiles, 16, 128 * 64, dtype=torch.bfloat16, device=dev)
    c_new = torch.zeros(n, 256, device=dev)
    st = abi.VineLstmStep(params=lpacked.data_ptr(), u=U.data_ptr(), hm=HM.data_ptr(), c_prev=c_prev.data_ptr(), not_done=nd.data_ptr(),
                          c=c_new.data_ptr(), hh=HH.data_ptr(), act=ACT.data_ptr(), n=n)
    assert lib.vine_lstm_step(C.byref(st), None) == 0
    DH, gparts, dbg = torch.zeros_like(HH), torch.zeros(abi.LSTM_HEAD_GRAD_PARTS, abi.LSTM_HEAD_GRAD_FLOATS, device=dev), torch.zeros(n, 4, device=dev)
    ht = abi.VineLstmHeadTrain(params=lpacked.data_ptr(), hh=HH.data_ptr(), scalars=scal.data_ptr(), logstd=m.sigma.data_ptr(),
                               logstd_old=logstd_old.data_ptr(), dh=DH.data_ptr(), grads=gparts.data_ptr(), debug_out=dbg.data_ptr(),
                               n=n, inv_B=1.0 / n, **hyp)
    nparts = lib.vine_lstm_head_train(C.byref(ht), None)
    assert 0 < nparts <= abi.LSTM_HEAD_GRAD_PARTS
    torch.cuda.synchronize()
    grads = gparts[:nparts].sum(0)
    DG, dc_prev_k = torch.zeros_like(ACT), torch.zeros(n, 256, device=dev)
    cb = abi.VineLstmCellBwd(act=ACT.data_ptr(), c_prev=c_prev.data_ptr(), c=c_new.data_ptr(), not_done=nd.data_ptr(), dh=DH.data_ptr(),
                             dg=DG.data_ptr(), dc_prev=dc_prev_k.data_ptr(), n=n)
    assert lib.vine_lstm_cell_bwd_tiles(C.byref(cb), None) == 0
    dh3_k, DHREC = torch.zeros(n, 64, device=dev), torch.zeros_like(HH)
    bg = abi.VineLstmBwdGemm(params=lpacked.data_ptr(), dg=DG.data_ptr(), not_done=nd.data_ptr(), dh3=dh3_k.data_ptr(),
                             dh_rec=DHREC.data_ptr(), n=n)
    assert lib.vine_lstm_bwd_gemm(C.byref(bg), None) == 0
    torch.cuda.synchronize()
    # ---- torch reference ----
    u_t = u_real.clone().requires_grad_(True)
    hp_t = h_prev.clone().requires_grad_(True)
    cp_t = c_prev.clone().requires_grad_(True)
    wih, whh = q(r.weight_ih_l0.detach()), q(r.weight_hh_l0.detach())
    gates = u_t @ wih.t() + (hp_t * nd[:, None]) @ whh.t() + (r.bias_ih_l0 + r.bias_hh_l0).detach()
    gates.retain_grad()
    i, f, g, o = gates.chunk(4, -1)
    c = torch.sigmoid(f) * (cp_t * nd[:, None]) + torch.sigmoid(i) * torch.tanh(g)
    h = torch.sigmoid(o) * torch.tanh(c)
    hq = h + (q(h) - h).detach()                     # the kernels store h in bf16 before LayerNorm
    lng, lnb = m.layer_norm.weight.detach().clone().requires_grad_(True), m.layer_norm.bias.detach().clone().requires_grad_(True)
    wmu, bmu = m.mu.weight.detach().clone().requires_grad_(True), m.mu.bias.detach().clone().requires_grad_(True)
    wv, bv = m.value.weight.detach().clone().requires_grad_(True), m.value.bias.detach().clone().requires_grad_(True)
    logstd = m.sigma.detach().clone().requires_grad_(True)
    y = torch.nn.functional.layer_norm(hq, (256,), lng, lnb, 1e-5)
    mu, v = y @ wmu.t() + bmu, (y @ wv.t() + bv).squeeze(-1)
    act, muo, nlpo, vo, ret, adv = scal[:, :2], scal[:, 2:4], scal[:, 4], scal[:, 5], scal[:, 6], scal[:, 7]
    sigma = torch.exp(logstd)
    nlp = 0.5 * (((act - mu) / sigma) ** 2).sum(-1) + math.log(2 * math.pi) + logstd.sum()
    ratio = torch.exp(nlpo - nlp)
    a_loss = torch.max(-adv * ratio, -adv * torch.clamp(ratio, 0.8, 1.2)).mean()
    v_clip = vo + (v - vo).clamp(-0.2, 0.2)
    c_loss = torch.max((v - ret) ** 2, (v_clip - ret) ** 2).mean()
    b_loss = (torch.clamp_min(mu - 1.1, 0) ** 2 + torch.clamp_max(mu + 1.1, 0) ** 2).sum(-1).mean()
    loss = a_loss + 0.5 * c_loss * 2.0 + b_loss * 1e-4
    loss.backward()
    rel = lambda a, b: float((a - b).norm() / (b.norm() + 1e-12))  # noqa: E731
    assert float((dbg[:, :2] - mu.detach()).abs().max()) < 3e-2 and float((dbg[:, 2] - v.detach()).abs().max()) < 3e-2
    errs = {
        "ln_g": rel(grads[0:256], lng.grad), "ln_b": rel(grads[256:512], lnb.grad),
        "w_mu": rel(grads[512:1024].view(2, 256), wmu.grad), "w_v": rel(grads[1024:1280], wv.grad[0]),
        "b_mu": rel(grads[1280:1282], bmu.grad), "b_v": rel(grads[1282:1283], bv.grad),
        "logstd": rel(grads[1284:1286], logstd.grad),
        "dc_prev": rel(dc_prev_k, cp_t.grad),
        "dG": rel(from_tiles(DG, n, width=64), gates.grad[:, _packed_gate_perm().to(dev)]),
        "dh3": rel(dh3_k, u_t.grad[:, :64]),
        "dh_prev": rel(from_tiles(DHREC, n), hp_t.grad),
    }
    print({k: f"{e:.2e}" for k, e in errs.items()})
    assert abs(float(grads[1286]) - float(a_loss)) < 2e-2 * abs(float(a_loss)) + 1e-4 and abs(float(grads[1287]) - float(c_loss)) < 2e-2 * float(c_loss)
    assert max(errs.values()) < 3e-2, errs


def test_native_lstm_minibatch_gradients_match_autograd_end_to_end():
    """The whole kernel-only minibatch (MLP forward -> 4 LSTM steps with done masks -> LayerNorm/heads/PPO loss -> BPTT -> MLP
    backward -> weight gradients) vs torch autograd of ppo.ActorCritic on the same data.  Tolerance 4e-2 relative Frobenius
    error per parameter tensor (bf16 operands through a 4-step recurrence)."""
    from vine_robot_isaacgymenvs_b200.ppo.lstm_native import NativeLstmPath
    lib = abi.load_library()
    L, S, O, dev = 4, 512, 18, "cuda"
    torch.manual_seed(11)
    m = ActorCritic(O, 2, (256, 128, 64), rnn=RNN).to(dev)
    m.fused_cell = False
    with torch.no_grad():
        m.layer_norm.weight.uniform_(0.5, 1.5); m.layer_norm.bias.uniform_(-0.2, 0.2)
        m.mu.weight.mul_(3.0); m.value.weight.mul_(3.0); m.sigma.uniform_(-0.3, 0.1)
        for prm in m.parameters():     # the kernels see bf16 weights: compare against the same rounded weights
            if prm.dim() == 2 and prm.shape[0] > 3:
                prm.copy_(prm.bfloat16().float())
    z = lambda *s: torch.zeros(*s, device=dev)  # noqa: E731
    packed = torch.zeros(abi.MLP_PACKED_BYTES, dtype=torch.uint8, device=dev)
    mlp = [m.actor_mlp[0].weight, m.actor_mlp[0].bias, m.actor_mlp[2].weight, m.actor_mlp[2].bias, m.actor_mlp[4].weight,
           m.actor_mlp[4].bias, z(2, 64), z(2), z(1, 64), z(1)]
    assert lib.vine_mlp_pack(*[p(t.detach().contiguous()) for t in mlp], O, p(packed), None) == 0
    lpacked = torch.zeros(abi.LSTM_PACKED_BYTES, dtype=torch.uint8, device=dev)
    r = m.rnn.rnn
    lp = [r.weight_ih_l0, r.weight_hh_l0, r.bias_ih_l0, r.bias_hh_l0, m.layer_norm.weight, m.layer_norm.bias, m.mu.weight,
          m.mu.bias, m.value.weight, m.value.bias]
    assert lib.vine_lstm_pack(*[p(t.detach().contiguous()) for t in lp], O, p(lpacked), None) == 0
    hyp = dict(e_clip=0.2, critic_coef=2.0, entropy_coef=0.0, bounds_loss_coef=1e-4)
    path = NativeLstmPath(O, L, S, dev, hyp)
    obs = torch.randn(L, S, O, device=dev) * 2
    mean, inv_std = torch.randn(O, device=dev) * 0.3, 1.0 / (0.5 + torch.rand(O, device=dev))
    scal = torch.randn(L, S, 8, device=dev); scal[..., 4] = scal[..., 4] * 0.3 + 2.5
    nd = (torch.rand(L, S, device=dev) > 0.15).float()
    h0 = (torch.randn(S, 256, device=dev) * 0.5).bfloat16().float()
    c0 = torch.randn(S, 256, device=dev) * 0.5
    logstd_old = torch.tensor([-0.1, 0.05], device=dev)
    vstats, state = torch.tensor([0.0, 1.0], device=dev), torch.zeros(16, device=dev)
    path.HM[0].copy_(to_tiles(h0 * nd[0][:, None]))
    path.C0.copy_(c0)
    dbg = torch.zeros(L * S, 4, device=dev)
    path.gradients(packed, lpacked, obs, scal, nd, mean, inv_std, vstats, m.sigma.detach(), logstd_old, state, debug_out=dbg)
    torch.cuda.synchronize()
    # ---- torch autograd on the same minibatch ----
    x = torch.clamp((obs - mean) * inv_std, -5, 5)
    mu, logstd, v, _ = m(x, (h0, c0), nd)
    mu, v = mu.float(), v.float().squeeze(-1)
    sc = scal.reshape(L * S, 8)
    act, muo, nlpo, vo, ret, adv = sc[:, :2], sc[:, 2:4], sc[:, 4], sc[:, 5], sc[:, 6], sc[:, 7]
    sigma = torch.exp(m.sigma)
    nlp = 0.5 * (((act - mu) / sigma) ** 2).sum(-1) + math.log(2 * math.pi) + m.sigma.sum()
    ratio = torch.exp(nlpo - nlp)
    a_loss = torch.max(-adv * ratio, -adv * torch.clamp(ratio, 0.8, 1.2)).mean()
    v_clip = vo + (v - vo).clamp(-0.2, 0.2)
    c_loss = torch.max((v - ret) ** 2, (v_clip - ret) ** 2).mean()
    b_loss = (torch.clamp_min(mu - 1.1, 0) ** 2 + torch.clamp_max(mu + 1.1, 0) ** 2).sum(-1).mean()
    (a_loss + 0.5 * c_loss * 2.0 + b_loss * 1e-4).backward()
    assert float((dbg[:, :2] - mu.detach()).abs().mean()) < 2e-2 and float((dbg[:, 2] - v.detach()).abs().mean()) < 2e-2
    rel = lambda a, b: float((a - b).norm() / (b.norm() + 1e-12))  # noqa: E731
    gl, o = path.flat_g_lstm, 0
    errs = {}
    for name, prm in (("W_ih", r.weight_ih_l0), ("W_hh", r.weight_hh_l0), ("b_ih", r.bias_ih_l0), ("b_hh", r.bias_hh_l0),
                      ("ln_g", m.layer_norm.weight), ("ln_b", m.layer_norm.bias), ("W_mu", m.mu.weight), ("b_mu", m.mu.bias),
                      ("W_v", m.value.weight), ("b_v", m.value.bias), ("logstd", m.sigma)):
        errs[name] = rel(gl[o:o + prm.numel()].view_as(prm), prm.grad)
        o += prm.numel()
    assert o == path.P_lstm
    gm, o = path.flat_g_mlp, 0
    for name, prm in (("W1", m.actor_mlp[0].weight), ("b1", m.actor_mlp[0].bias), ("W2", m.actor_mlp[2].weight),
                      ("b2", m.actor_mlp[2].bias), ("W3", m.actor_mlp[4].weight), ("b3", m.actor_mlp[4].bias)):
        errs[name] = rel(gm[o:o + prm.numel()].view_as(prm), prm.grad)
        o += prm.numel()
    print({k: f"{e:.2e}" for k, e in errs.items()})
    assert abs(float(gl[path.P_lstm]) - float(a_loss)) < 3e-2 * abs(float(a_loss)) + 1e-4
    assert abs(float(gl[path.P_lstm + 1]) - float(c_loss)) < 3e-2 * float(c_loss)
    assert max(errs.values()) < 4e-2, errs
