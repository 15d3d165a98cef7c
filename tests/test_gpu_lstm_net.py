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
