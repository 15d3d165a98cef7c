"""GPU: fused tcgen05/TMEM actor-critic MLP forward vs a plain PyTorch fp32 reference of the same op.

The kernel computes in bf16 x bf16 -> f32 (weights and inter-layer activations rounded to bf16), the
reference in fp32.  Tolerances (outputs are O(1..5), three bf16-rounded hidden layers):
  * against an fp32 reference that applies the SAME bf16 roundings: mean |d| < 5e-4, max |d| < 6e-2 (the max is a rare
    rounding-boundary flip of one hidden activation, ELU goes through __expf);
  * against the plain fp32 reference: mean |d| < 1.5e-2 and every element |d| <= 0.1 + 0.05 |ref|.
"""
import ctypes as C

import pytest
import torch

from vine_robot_isaacgymenvs_b200 import abi

pytestmark = pytest.mark.gpu


def p(t):
    return C.c_void_p(t.data_ptr())


def reference(obs, mean, inv_std, W, val_mean, val_std, emulate_bf16):
    r = (lambda t: t.bfloat16().float()) if emulate_bf16 else (lambda t: t)
    x = r(torch.clamp((obs - mean) * inv_std, -5, 5))
    for w, b in W[:3]:
        x = r(torch.nn.functional.elu(x @ r(w).t() + b))
    mu = x @ r(W[3][0]).t() + W[3][1]
    v = x @ r(W[4][0]).t() + W[4][1]
    return mu, torch.clamp(v.squeeze(-1), -5, 5) * val_std + val_mean


@pytest.mark.parametrize("n,num_obs", [(128, 18), (4096, 18), (5000, 28), (300_000, 18)])
def test_fused_mlp_forward_matches_torch(n, num_obs):
    lib = abi.load_library()
    g = torch.Generator(device="cuda").manual_seed(n)
    rnd = lambda *s, k=1.0: (torch.randn(*s, device="cuda", generator=g) * k)  # noqa: E731
    dims = [num_obs, 256, 128, 64]
    W = [(rnd(dims[i + 1], dims[i], k=(1.5 / dims[i]) ** 0.5), rnd(dims[i + 1], k=0.1)) for i in range(3)]
    W += [(rnd(2, 64, k=0.15), rnd(2, k=0.1)), (rnd(1, 64, k=0.15), rnd(1, k=0.1))]
    obs = rnd(n, num_obs, k=3.0)
    mean, inv_std = rnd(num_obs, k=0.5), 1.0 / (0.5 + torch.rand(num_obs, device="cuda", generator=g))
    packed = torch.zeros(abi.MLP_PACKED_BYTES, dtype=torch.uint8, device="cuda")
    flat = [t.contiguous() for wb in W for t in wb]
    assert lib.vine_mlp_pack(*[p(t) for t in flat], num_obs, p(packed), None) == 0
    mu, val = torch.zeros(n, 2, device="cuda"), torch.zeros(n, device="cuda")
    vstats = torch.tensor([0.3, 1.7], device="cuda")
    assert lib.vine_mlp_forward(p(packed), p(obs), p(mean), p(inv_std), n, num_obs, p(vstats), p(mu), p(val), None) == 0
    torch.cuda.synchronize()
    mu_e, val_e = reference(obs, mean, inv_std, W, 0.3, 1.7, emulate_bf16=True)
    mu_f, val_f = reference(obs, mean, inv_std, W, 0.3, 1.7, emulate_bf16=False)
    assert torch.isfinite(mu).all() and torch.isfinite(val).all()
    # vs the same bf16 roundings: identical up to rare rounding-boundary flips of an activation (ELU via __expf)
    assert float((mu - mu_e).abs().max()) < 6e-2 and float((val - val_e).abs().max()) < 1e-1
    assert float((mu - mu_e).abs().mean()) < 5e-4 and float((val - val_e).abs().mean()) < 1e-3
    assert float((mu - mu_f).abs().mean()) < 1.5e-2 and float((val - val_f).abs().mean()) < 3e-2
    assert bool(((mu - mu_f).abs() <= 0.1 + 0.05 * mu_f.abs()).all())
    assert bool(((val - val_f).abs() <= 0.2 + 0.05 * val_f.abs()).all())
    assert float(mu_f.abs().mean()) > 0.05       # the comparison is not vacuous
