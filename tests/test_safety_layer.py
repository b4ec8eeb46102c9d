'''SafeExplorer safety layer on the CPU: the torch `SafetyLayer` against golden vectors of the reference class
(tests/golden/safety_layer.npz, tools/make_golden_safety.py), and the ORACLE's in-rollout projection (oracle/scg_oracle.c) against the
torch layer -- which pins the checker the GPU parity test (tests/test_gpu_safe_explorer.py) relies on.'''
import os

import numpy as np
import pytest

import helpers as H

torch = pytest.importorskip('torch')
GOLD = os.path.join(os.path.dirname(__file__), 'golden', 'safety_layer.npz')


class _Space:
    def __init__(self, n):
        self.shape = (n,)


def _golden_layer(which='w0'):
    from safe_control_gym_b200.controllers.safe_explorer.safe_explorer_utils import SafetyLayer
    g = np.load(GOLD)
    OBS, ACT, C, Hd = (int(v) for v in g['dims'])
    layer = SafetyLayer(_Space(OBS), _Space(ACT), hidden_dim=Hd, num_constraints=C, lr=1e-3, slack=g['slack'].tolist())
    sd = {k[len(which) + 2:].replace('__', '.'): torch.from_numpy(g[k]) for k in g.files if k.startswith(which + '__')}
    layer.constraint_models.load_state_dict(sd)
    return layer, g


def test_projection_and_regression_match_the_reference_class():
    layer, g = _golden_layer()
    obs, act, c, c_next = (torch.from_numpy(g[k]) for k in ('obs', 'act', 'c', 'c_next'))
    with torch.no_grad():
        safe = layer.get_safe_action(obs, act, c)                       # batched; the reference evaluates one row at a time
    np.testing.assert_allclose(safe.numpy(), g['safe'], rtol=1e-5, atol=5e-6)       # batched vs row-by-row GEMM rounding
    assert float((np.abs(g['safe'] - g['act']).max(-1) > 1e-7).mean()) > 0.5       # the fixture exercises the projection
    batch = {'obs': obs, 'act': act, 'c': c, 'c_next': c_next}
    losses = [float(l.detach()) for l in layer.compute_loss(batch)]
    np.testing.assert_allclose(losses, g['losses'], rtol=1e-5)
    res = layer.update(batch)                                            # one Adam step per constraint model
    np.testing.assert_allclose([res[f'constraint_{i}_loss'] for i in range(len(losses))], g['update_losses'], rtol=1e-5)
    after, _ = _golden_layer('w1')
    for (k, a), b in zip(layer.constraint_models.state_dict().items(), after.constraint_models.state_dict().values()):
        np.testing.assert_allclose(a.numpy(), b.numpy(), rtol=1e-5, atol=1e-7, err_msg=k)


def test_pack_layout():
    layer, g = _golden_layer()
    p = layer.pack('cpu')
    OBS, ACT, C, Hd = (int(v) for v in g['dims'])
    per = Hd * OBS + Hd + ACT * Hd + ACT
    assert p['C'] == C and p['H'] == Hd and p['blob'].numel() == C * per and p['blob'].dtype == torch.float32
    m2 = layer.constraint_models[2]
    blk = p['blob'][2 * per:3 * per]
    assert torch.equal(blk[:Hd * OBS].reshape(Hd, OBS), m2.fcs[0].weight.detach())
    assert torch.equal(blk[Hd * OBS + Hd:Hd * OBS + Hd + ACT * Hd].reshape(ACT, Hd), m2.fcs[1].weight.detach())


@pytest.mark.parametrize('name', ['quadrotor_2D_track', 'cartpole_stab'])
def test_oracle_rollout_projection_matches_the_torch_layer(name):
    from safe_control_gym_b200.controllers.safe_explorer.safe_explorer_utils import SafetyLayer
    env = H.make_env(name)
    N, T = 48, 40
    C = env.constraints.num_state_constraints
    torch.manual_seed(9)
    layer = SafetyLayer(env.observation_space, env.action_space, hidden_dim=10, num_constraints=C, lr=1e-3, slack=0.6)
    with torch.no_grad():
        for m in layer.constraint_models:
            m.fcs[1].weight.mul_(3.0)
    p = layer.pack('cpu')
    pol = H.random_policy(env.obs_dim, env.action_dim, 64, seed=2)
    o = H.make_oracle(env, N, auto_reset=True, rng_mode='philox', seed=4)
    o.reset()
    rec = o.rollout(pol, T, deterministic=True, safety=dict(p, blob=p['blob'].numpy()), trace=True)
    obs, act, c = (torch.from_numpy(rec[k].reshape(T * N, -1)).float() for k in ('obs', 'act', 'c'))
    # c is the state-constraint part of info['constraint_values'] of the CURRENT state (safe_ppo.py:123,344): A x - b, the state being the
    # leading part of the (noise-free) observation
    A, b = env.constraints.state_constraints[0].A, env.constraints.state_constraints[0].b
    st = rec['obs'].reshape(T * N, -1)[:, :A.shape[1]].astype(np.float64)
    np.testing.assert_allclose(c.numpy(), (st @ A.T - b)[:, :C], rtol=0, atol=1e-5)
    h = obs
    for i, (w_, b_) in enumerate(zip(pol.aw, pol.ab)):
        h = h @ torch.from_numpy(w_).T + torch.from_numpy(b_)
        if i < 2:
            h = torch.tanh(h)
    with torch.no_grad():
        safe = layer.get_safe_action(obs, h, c)
        g = torch.stack([m(obs) for m in layer.constraint_models], 1)
    # mu_safe = mu - num * g / (g.g + 1e-8) is ill-conditioned where the winning sensitivity is ~0 (1-D actions: multipliers of 1e3 and
    # more, float32 rounding of g amplified by 1/|g|): those samples get a relative bound scaled by 1/|g|, the rest the tight one
    lam = torch.relu(((g * h.unsqueeze(1)).sum(-1) + c + 0.6) / ((g * g).sum(-1) + 1e-8))
    gwin = torch.gather((g * g).sum(-1), 1, lam.argmax(-1, keepdim=True))[:, 0].sqrt()
    d = ((safe - act).abs() / (1 + act.abs())).max(-1).values
    tol = 2e-5 + 2e-6 / gwin.clamp_min(1e-6)
    assert float((d > tol).float().mean()) < 2e-3, float(d.max())          # (arg-max flips between nearly equal multipliers excepted)
    assert float((gwin > 0.05).float().mean()) > 0.8
    assert float(((safe - h).abs().max(-1).values > 1e-6).float().mean()) > 0.05
