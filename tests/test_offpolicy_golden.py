'''SAC / DDPG agents against golden vectors produced by the UNMODIFIED reference classes (tools/make_golden_offpolicy.py imports
safe_control_gym/controllers/sac/sac_utils.py and .../ddpg/ddpg_utils.py from /root/reference; tests/golden/offpolicy.npz is its committed output).

The agents here keep the reference's parameter names, so the golden initial weights load by name; then, with torch seeded as the generator script
seeded it, ONE and TWO `update(batch)` steps must reproduce the reference's losses and every weight afterwards (actor, critics, Polyak-averaged
targets, temperature).  CPU, float32: the tolerance covers only the different association of a few float32 sums.'''
import os

import numpy as np
import pytest

torch = pytest.importorskip('torch')

from safe_control_gym_b200.controllers.ddpg.ddpg_utils import DDPGAgent  # noqa: E402
from safe_control_gym_b200.controllers.sac.sac_utils import DeviceReplayBuffer, SACAgent  # noqa: E402
from safe_control_gym_b200.spaces import Box  # noqa: E402

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'offpolicy.npz'))
OBS = Box(-np.ones(6), np.ones(6))
ACT = Box(np.array([0.05, 0.05], np.float32), np.array([0.3, 0.3], np.float32))
TOL = dict(rtol=2e-5, atol=2e-6)


def _sd(prefix):
    p = prefix + '__'
    return {k[len(p):].replace('__', '.'): torch.as_tensor(G[k]) for k in G.files if k.startswith(p)}


def _batch(tag):
    return {k: torch.as_tensor(G[f'{tag}_batch_{k}']) for k in ('obs', 'act', 'rew', 'next_obs', 'mask')}


def _check_weights(module, prefix):
    want = _sd(prefix)
    got = module.state_dict()
    assert set(want) == set(got), (sorted(set(want) ^ set(got)))
    for k, v in want.items():
        np.testing.assert_allclose(got[k].detach().numpy(), v.numpy(), err_msg=f'{prefix}.{k}', **TOL)


@pytest.mark.parametrize('tag,tuning', [('sac', False), ('sact', True)])
def test_sac_agent_reproduces_the_reference_update(tag, tuning):
    ag = SACAgent(OBS, ACT, hidden_dim=32, use_entropy_tuning=tuning, init_temperature=0.2, actor_lr=3e-3, critic_lr=2e-3, entropy_lr=5e-3)
    ag.to('cpu')
    ag.ac.load_state_dict(_sd(f'{tag}_ac0'))                          # same parameter names as the reference (strict load)
    ag.ac_targ.load_state_dict(_sd(f'{tag}_targ0'))
    batch = _batch(tag)
    ag.train()
    with torch.no_grad():
        det = ag.ac.actor(batch['obs'], deterministic=True, with_logprob=False)[0]
        np.testing.assert_allclose(det.numpy(), G[f'{tag}_act_det'], **TOL)
        torch.manual_seed(102)                                        # the sampled action consumes the generator exactly like Normal.rsample
        a, lp = ag.ac.actor(batch['obs'])
        np.testing.assert_allclose(a.numpy(), G[f'{tag}_act_sample'], **TOL)
        np.testing.assert_allclose(lp.numpy(), G[f'{tag}_logp_sample'], rtol=2e-5, atol=2e-5)
    torch.manual_seed(103)
    res = ag.update(batch)
    np.testing.assert_allclose([float(res['policy_loss']), float(res['critic_loss']), float(res['entropy_loss'])], G[f'{tag}_update_res'], rtol=2e-5, atol=2e-6)
    _check_weights(ag.ac, f'{tag}_ac1')
    _check_weights(ag.ac_targ, f'{tag}_targ1')
    np.testing.assert_allclose(float(ag.log_alpha.detach()), float(G[f'{tag}_log_alpha1']), rtol=1e-6, atol=1e-7)
    if not tuning:
        assert float(ag.log_alpha.detach()) == pytest.approx(np.log(0.2), rel=1e-6)     # the temperature only moves with entropy tuning
    torch.manual_seed(104)
    res = ag.update(batch)                                            # second step: Adam moments carried over
    np.testing.assert_allclose([float(res['policy_loss']), float(res['critic_loss']), float(res['entropy_loss'])], G[f'{tag}_update_res2'], rtol=5e-5, atol=5e-6)
    _check_weights(ag.ac, f'{tag}_ac2')
    np.testing.assert_allclose(float(ag.log_alpha.detach()), float(G[f'{tag}_log_alpha2']), rtol=1e-6, atol=1e-7)


def test_ddpg_agent_reproduces_the_reference_update():
    ag = DDPGAgent(OBS, ACT, hidden_dim=32, actor_lr=3e-3, critic_lr=2e-3)
    ag.to('cpu')
    ag.ac.load_state_dict(_sd('ddpg_ac0'))
    ag.ac_targ.load_state_dict(_sd('ddpg_targ0'))
    batch = _batch('ddpg')
    ag.train()
    with torch.no_grad():
        np.testing.assert_allclose(ag.ac.actor(batch['obs']).numpy(), G['ddpg_act'], **TOL)
    for i in (1, 2):
        res = ag.update(batch)
        np.testing.assert_allclose([float(res['policy_loss']), float(res['critic_loss'])], G[f'ddpg_update_res{i}'], rtol=5e-5, atol=5e-6)
        _check_weights(ag.ac, f'ddpg_ac{i}')
        _check_weights(ag.ac_targ, f'ddpg_targ{i}')
    # the first update's policy loss is compute_policy_loss on the initial weights (nothing is sampled in DDPG)
    assert float(G['ddpg_update_res1'][0]) == pytest.approx(float(G['ddpg_losses0'][0]), rel=1e-6)


def test_replay_buffer_matches_the_reference_ring():
    '''SACBuffer.push (sac_utils.py:343-362) on a push sequence that wraps around: same arrays, position and size.'''
    buf = DeviceReplayBuffer(OBS, ACT, max_size=10, batch_size=4, device='cpu')
    o = 0
    for n in G['buf_sizes']:
        buf.push({k: torch.as_tensor(G[f'buf_push_{k}'][o:o + n]) for k in ('obs', 'next_obs', 'act', 'rew', 'mask')})
        o += int(n)
    assert [buf.pos, len(buf)] == list(G['buf_pos_size'])
    for k in ('obs', 'next_obs', 'act', 'rew', 'mask'):
        np.testing.assert_array_equal(buf.data[k].numpy(), G[f'buf_final_{k}'].astype(np.float32).reshape(buf.data[k].shape), err_msg=k)
