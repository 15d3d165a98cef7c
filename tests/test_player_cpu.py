"""CPU: deployment-side player (vine_robot_test_model.py analogue) restores an rl_games-layout checkpoint, keeps LSTM
state across calls, and rescales actions to hardware ranges."""
import torch

from vine_robot_isaacgymenvs_b200.player import REFERENCE_RNN, PolicyPlayer, VineRobotControlModel
from vine_robot_isaacgymenvs_b200.ppo.ppo import ActorCritic, RunningMeanStd


def _checkpoint(rnn):
    torch.manual_seed(0)
    m = ActorCritic(18, 2, (256, 128, 64), rnn=REFERENCE_RNN if rnn else None)
    rms = RunningMeanStd((18,))
    rms.update(torch.randn(1000, 18) * 2 + 1)
    sd = {"a2c_network." + k: v for k, v in m.state_dict().items()}
    sd.update({"running_mean_std." + k: v for k, v in rms.state_dict().items()})
    return {"model": sd, "epoch": 3}, m, rms


def test_player_reproduces_the_network_and_carries_lstm_state(tmp_path):
    ck, m, rms = _checkpoint(True)
    torch.save(ck, tmp_path / "p.pth")
    pl = PolicyPlayer(18).restore(str(tmp_path / "p.pth"))
    assert pl.model.has_rnn
    obs = torch.randn(5, 18)
    acts = [pl.get_action(o, is_deterministic=True) for o in obs]
    with torch.no_grad():
        mu, _, _, _ = m(rms(obs).reshape(5, 1, 18), (torch.zeros(1, 256), torch.zeros(1, 256)), None)
    assert torch.allclose(torch.stack(acts), mu.clamp(-1, 1), atol=1e-5)     # 5 calls == one 5-step sequence
    pl.reset()
    assert torch.allclose(pl.get_action(obs[0], True), acts[0], atol=1e-6)   # reset() restarts the episode
    a = pl.get_action(obs[1], False)
    assert a.shape == (2,) and float(a.abs().max()) <= 1.0


def test_control_model_rescales_to_hardware_ranges(tmp_path):
    ck, m, rms = _checkpoint(False)
    cm = VineRobotControlModel(ck, x_range=(-2.0, 2.0), u_range=(0.0, 3.0), num_obs=18, deterministic=True)
    assert not cm.player.model.has_rnn
    parts = [torch.randn(6), torch.randn(6), torch.randn(3), torch.randn(3)]
    out = cm.get_action(*parts)
    with torch.no_grad():
        mu = m(rms(torch.cat(parts)[None]))[0][0].clamp(-1, 1)
    assert torch.allclose(out, torch.stack([(mu[0] + 1) * 2.0 - 2.0, (mu[1] + 1) * 1.5]), atol=1e-5)
