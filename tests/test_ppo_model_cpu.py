"""CPU: the actor-critic network of Vine5LinkMovingBasePPO.yaml:10-40 (host-side torch restatement used by the
PPO update) -- LSTM-with-dones semantics, rl_games checkpoint key names, minibatch sequence layout."""
import torch

from vine_robot_isaacgymenvs_b200 import config as vcfg
from vine_robot_isaacgymenvs_b200.ppo.ppo import ActorCritic

RNN = {"name": "lstm", "units": 256, "layers": 1, "before_mlp": False, "concat_input": True, "layer_norm": True}


def test_default_train_config_carries_the_reference_rnn_block_and_null_removes_it():
    c = vcfg.compose([])
    assert c["train"]["params"]["network"]["rnn"] == RNN
    assert c["train"]["params"]["config"]["seq_len"] == 4
    c = vcfg.compose(["train.params.network.rnn=null"])
    assert c["train"]["params"]["network"]["rnn"] is None


def test_lstm_forward_equals_torch_lstm_and_dones_cut_the_sequence():
    torch.manual_seed(0)
    m = ActorCritic(18, 2, (256, 128, 64), rnn=RNN)
    L, S = 4, 6
    obs, h, c = torch.randn(L, S, 18), torch.randn(S, 256), torch.randn(S, 256)
    mu, logstd, v, (h2, c2) = m(obs, (h, c), None)
    x = obs.reshape(L * S, 18)
    inp = torch.cat([m.actor_mlp(x), x], -1).view(L, S, -1)                # concat_input: True
    out, (hr, cr) = m.rnn.rnn(inp, (h[None], c[None]))
    ref = m.layer_norm(out.reshape(L * S, -1))
    assert torch.allclose(mu, m.mu(ref), atol=1e-5) and torch.allclose(v, m.value(ref), atol=1e-5)
    assert torch.allclose(h2, hr[0], atol=1e-6) and torch.allclose(c2, cr[0], atol=1e-6)
    assert logstd.shape == (L * S, 2) and float(logstd.abs().max()) == 0.0  # fixed_sigma, const_initializer 0
    # a done flag before step 2 of sequence 3 == restarting that sequence from a zero state at step 2
    nd = torch.ones(L, S)
    nd[2, 3] = 0.0
    mu_d, _, _, (h_d, _) = m(obs, (h, c), nd)
    mu_r, _, _, (h_r, _) = m(obs[2:, 3:4], (torch.zeros(1, 256), torch.zeros(1, 256)), None)
    got = mu_d.view(L, S, 2)[2:, 3]
    assert torch.allclose(got, mu_r.view(2, 2), atol=1e-5) and torch.allclose(h_d[3], h_r[0], atol=1e-6)
    assert torch.allclose(mu_d.view(L, S, 2)[:2], mu.view(L, S, 2)[:2], atol=1e-6)   # earlier steps untouched


def test_state_dict_keys_follow_rl_games_a2c_network():
    keys = set(ActorCritic(18, 2, (256, 128, 64), rnn=RNN).state_dict())
    assert {"sigma", "actor_mlp.0.weight", "actor_mlp.2.weight", "actor_mlp.4.bias", "rnn.rnn.weight_ih_l0",
            "rnn.rnn.weight_hh_l0", "rnn.rnn.bias_ih_l0", "rnn.rnn.bias_hh_l0", "layer_norm.weight", "layer_norm.bias",
            "mu.weight", "value.weight"} <= keys
    m = ActorCritic(18, 2, (256, 128, 64), rnn=None)
    assert not m.has_rnn and m.mu.in_features == 64
    assert m.rnn.rnn.input_size == 64 + 18 if m.has_rnn else True
