"""GPU: the PPO iteration (rollout + GAE + update) runs through the fused env step for both network variants,
eager and CUDA-graph-captured, stays finite, and checkpoints round-trip in rl_games' layout."""
import pytest
import torch

import vine_robot_isaacgymenvs_b200 as vine
from vine_robot_isaacgymenvs_b200 import config as vcfg
from vine_robot_isaacgymenvs_b200.ppo.ppo import PPOAgent

pytestmark = pytest.mark.gpu


def make_agent(extra, graphs, n=512):
    cfg = vcfg.compose(vcfg.FSTR_OVERRIDES + [f"num_envs={n}", "headless=True", "train.params.config.minibatch_size=4096"]
                       + extra)
    env = vine.make(cfg=cfg)
    return PPOAgent(env, cfg["train"], seed=1, use_graphs=graphs)


@pytest.mark.parametrize("rnn", [True, False], ids=["lstm", "mlp"])
@pytest.mark.parametrize("graphs", [False, True], ids=["eager", "graphs"])
def test_ppo_iterations_are_finite_and_learn_something(rnn, graphs):
    agent = make_agent([] if rnn else ["train.params.network.rnn=null"], graphs)
    assert agent.has_rnn == rnn and agent.seq_len == (4 if rnn else 1)
    before = [p.detach().clone() for p in agent.model.parameters()]
    for _ in range(6):
        agent.train_epoch()
    torch.cuda.synchronize()
    st = agent.pop_stats()
    assert st["episodes"] > 0 and all(x == x and abs(x) < 1e6 for x in (st["a_loss"], st["c_loss"], st["kl"]))
    assert all(torch.isfinite(p).all() for p in agent.model.parameters())
    assert any(not torch.equal(a, b) for a, b in zip(before, agent.model.parameters()))
    assert float(agent.obs_rms.count) > 1.0 and agent.frames == (6 + (3 if graphs else 0)) * 16 * 512  # +3: graph warm-up


def test_checkpoint_roundtrip_uses_rl_games_layout(tmp_path):
    a = make_agent([], False)
    a.train_epoch()
    sd = a.state_dict()
    assert "a2c_network.rnn.rnn.weight_hh_l0" in sd["model"] and "running_mean_std.running_mean" in sd["model"]
    assert "value_mean_std.running_var" in sd["model"] and {"epoch", "frame", "optimizer", "last_lr"} <= set(sd)
    torch.save(sd, tmp_path / "ck.pth")
    b = make_agent([], False)
    b.load_state_dict(torch.load(tmp_path / "ck.pth", map_location="cuda:0"))
    for pa, pb in zip(a.model.parameters(), b.model.parameters()):
        assert torch.equal(pa, pb)
    assert torch.equal(a.obs_rms.running_mean, b.obs_rms.running_mean) and b.epoch == 1
