'''Host-side logic of the off-policy siblings (SAC / DDPG) on the CPU: replay-buffer ring semantics against a numpy
restatement of SACBuffer (sac_utils.py:283-366), the squashed-Gaussian log-probability against torch.distributions, the Polyak update, one
agent update step against hand-written losses, the exploration noise recursion.'''
import numpy as np
import pytest

torch = pytest.importorskip('torch')

from safe_control_gym_b200.controllers.ddpg.ddpg_utils import DDPGAgent, OUNoise, make_action_noise_process  # noqa: E402
from safe_control_gym_b200.controllers.sac.sac_utils import DeviceReplayBuffer, SACAgent, soft_update  # noqa: E402
from safe_control_gym_b200.spaces import Box  # noqa: E402

OBS, ACT = Box(-np.ones(6), np.ones(6)), Box(np.array([0.05, 0.05]), np.array([0.3, 0.3]))


def test_replay_buffer_ring_matches_numpy_restatement():
    buf = DeviceReplayBuffer(OBS, ACT, max_size=10, batch_size=4, device='cpu')
    ref = {k: np.zeros((10, w), np.float32) for k, w in (('obs', 6), ('next_obs', 6), ('act', 2), ('rew', 1))}
    ref['mask'] = np.ones((10, 1), np.float32)
    pos = size = 0
    rng = np.random.default_rng(0)
    for n in (4, 4, 4, 3, 7):                                        # wraps around twice
        batch = {'obs': rng.normal(size=(n, 6)), 'next_obs': rng.normal(size=(n, 6)), 'act': rng.normal(size=(n, 2)), 'rew': rng.normal(size=(n, 1)),
                 'mask': rng.integers(0, 2, size=(n, 1))}
        buf.push({k: torch.as_tensor(v) for k, v in batch.items()})
        for k, v in batch.items():
            v = v.astype(np.float32)
            if pos + n <= 10:
                ref[k][pos:pos + n] = v
            else:
                rem = pos + n - 10
                ref[k][pos:] = v[:-rem]; ref[k][:rem] = v[-rem:]
        if size < 10:
            size = min(10, pos + n)
        pos = (pos + n) % 10
        assert len(buf) == size and buf.pos == pos
        for k in ref:
            np.testing.assert_array_equal(buf.data[k].numpy(), ref[k])
    s = buf.sample(64, generator=torch.Generator().manual_seed(1))
    assert s['obs'].shape == (64, 6) and s['mask'].shape == (64, 1)
    st = buf.state_dict()
    buf2 = DeviceReplayBuffer(OBS, ACT, max_size=10, device='cpu'); buf2.load_state_dict(st)
    assert buf2.pos == buf.pos and len(buf2) == len(buf) and torch.equal(buf2.data['act'], buf.data['act'])


def test_squashed_gaussian_logprob_and_action_map():
    torch.manual_seed(0)
    ag = SACAgent(OBS, ACT, hidden_dim=32)
    obs = torch.randn(50, 6)
    torch.manual_seed(5)
    a, logp = ag.ac.actor(obs)
    assert float(a.min()) >= 0.05 and float(a.max()) <= 0.3
    # independent: tanh-transformed Normal from torch.distributions on the pre-squash sample recovered from the action
    feat = ag.ac.actor.net(obs)
    mu, ls = ag.ac.actor.mu_layer(feat), ag.ac.actor.log_std_layer(feat).clamp(-20, 2)
    y = (a - 0.05) / (0.3 - 0.05) * 2 - 1
    base = torch.distributions.Normal(mu, ls.exp())
    dist = torch.distributions.TransformedDistribution(base, [torch.distributions.transforms.TanhTransform(cache_size=1)])
    np.testing.assert_allclose(logp.detach().numpy()[:, 0], dist.log_prob(y.clamp(-0.999999, 0.999999)).sum(-1).detach().numpy(), rtol=2e-3, atol=2e-3)
    det, _ = ag.ac.actor(obs, deterministic=True, with_logprob=False)
    np.testing.assert_allclose(det.detach().numpy(), (0.05 + 0.5 * (torch.tanh(mu) + 1) * 0.25).detach().numpy(), rtol=1e-6)
    # reference parameter names (checkpoint format)
    assert {'actor.net.fcs.0.weight', 'actor.mu_layer.weight', 'actor.log_std_layer.bias', 'q1.q_net.fcs.2.weight', 'q2.q_net.fcs.0.bias'} <= set(ag.ac.state_dict())


def test_polyak_update_and_one_sac_step():
    torch.manual_seed(1)
    ag = SACAgent(OBS, ACT, hidden_dim=16, tau=0.25, use_entropy_tuning=True)
    ag.to('cpu')
    before_t = [p.clone() for p in ag.ac_targ.parameters()]
    before_s = [p.detach().clone() for p in ag.ac.parameters()]
    batch = {'obs': torch.randn(32, 6), 'act': torch.rand(32, 2) * 0.25 + 0.05, 'rew': torch.randn(32, 1), 'next_obs': torch.randn(32, 6), 'mask': torch.ones(32, 1)}
    # critic target by hand with the same noise
    torch.manual_seed(9)
    res = ag.update(batch)
    assert all(np.isfinite(float(v)) for v in res.values())
    for bt, s_new, t_new in zip(before_t, ag.ac.parameters(), ag.ac_targ.parameters()):
        np.testing.assert_allclose(t_new.detach().numpy(), (0.75 * bt + 0.25 * s_new.detach()).numpy(), rtol=1e-6, atol=1e-7)
    assert any(not torch.equal(a, b.detach()) for a, b in zip(before_s, ag.ac.parameters()))
    assert float(ag.log_alpha) != pytest.approx(np.log(0.2))         # temperature moved
    src, tgt = torch.nn.Linear(3, 3), torch.nn.Linear(3, 3)
    w0 = tgt.weight.detach().clone()
    soft_update(src, tgt, 0.1)
    np.testing.assert_allclose(tgt.weight.detach().numpy(), (0.9 * w0 + 0.1 * src.weight.detach()).numpy(), rtol=1e-6)


def test_ddpg_losses_by_hand():
    torch.manual_seed(2)
    ag = DDPGAgent(OBS, ACT, hidden_dim=16, gamma=0.9, tau=0.0)      # tau = 0: targets frozen, so the hand computation below stays valid
    batch = {'obs': torch.randn(20, 6), 'act': torch.rand(20, 2) * 0.25 + 0.05, 'rew': torch.randn(20, 1), 'next_obs': torch.randn(20, 6), 'mask': (torch.rand(20, 1) > 0.3).float()}
    with torch.no_grad():
        pl = -ag.ac.q(batch['obs'], ag.ac.actor(batch['obs'])).mean()
    res = ag.update(batch)
    assert float(res['policy_loss']) == pytest.approx(float(pl), rel=1e-6)
    assert np.isfinite(float(res['critic_loss']))
    a = ag.ac.act(batch['obs'])
    assert float(a.min()) >= 0.05 - 1e-6 and float(a.max()) <= 0.3 + 1e-6
    assert {'actor.net.fcs.2.bias', 'q.q_net.fcs.0.weight'} <= set(ag.ac.state_dict())


def test_ou_noise_recursion_and_schedule():
    g = torch.Generator().manual_seed(3)
    p = make_action_noise_process(dict(func='OrnsteinUhlenbeckProcess', std=dict(func='LinearSchedule', args=0.2, end=0.1, steps=10)), 5, ACT, 'cpu', g)
    g2 = torch.Generator().manual_seed(3)
    x = np.zeros((5, 2))
    for t in range(12):
        std = 0.2 + (0.1 - 0.2) * min(1.0, t / 10)
        n = torch.randn(5, 2, generator=g2).numpy()
        x = x + 0.15 * (0 - x) * 1e-2 + std * np.sqrt(1e-2) * n       # random_processes.py:41-42
        np.testing.assert_allclose(p.sample().numpy(), x, rtol=1e-5, atol=1e-7)
    assert isinstance(p, OUNoise) and make_action_noise_process(None, 5, ACT, 'cpu', g) is None
