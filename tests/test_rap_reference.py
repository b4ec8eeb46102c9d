'''`RAP.collect_rollouts` of THIS repo (controllers/rarl/rap.py: the population collector composed from replayed single-adversary launches) against the
buffers the reference's own `RAP.collect_rollouts()` returned (controllers/rarl/rap.py:348-470 of /root/reference on a DummyVecEnv of unmodified reference
envs, recorded by tools/make_golden_rap_loop.py; tests/golden/rap_loop_*.npz).

The composition runs here on the CPU: the engine behind it is an adapter over the ORACLE (same `get_state` / `set_state` / `rollout(adversary=, adv_record=)`
contract as EngineHandle; the CUDA collector is compared with the oracle's two-player rollout on the GPU, tests/test_gpu_extended.py), fed with the
reference run's weights, per-env adversary draw and normals.  Required: the protagonist's buffer over all envs and, for every adversary that was drawn, its
columns of the adversaries' buffer -- its own actions / values / log-probabilities, `-rew`, terminal and bootstrap values from ITS critic, returns, and the
advantages normalised over the WHOLE batch before the split -- as the reference returned them.'''
import json
import os

import numpy as np
import pytest

import helpers as H
from safe_control_gym_b200.controllers.rarl import rap as R
from safe_control_gym_b200.utils.registration import make

torch = pytest.importorskip('torch')
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


class OracleEngine:
    '''EngineHandle's contract as far as RAP uses it, on the CPU oracle; the normals are the recorded ones instead of Philox draws'''

    def __init__(self, oracle, eps, eps_adv):
        self.o, self.eps, self.eps_adv = oracle, eps, eps_adv

    def get_state(self):
        return self.o.get_state()

    def set_state(self, s):
        self.o.set_state(s)

    def rollout(self, pol, dims, T, activation='tanh', gamma=0.99, gae_lambda=0.95, use_gae=False, out=None, want=(), adversary=None, adv_record=False):
        adv_pol, adv_dims = adversary
        assert adv_dims == (dims[0], adv_pol.act_dim, dims[2])
        r = self.o.rollout(pol, T, eps=self.eps, eps_adv=self.eps_adv, gamma=gamma, gae_lambda=gae_lambda, use_gae=use_gae, adversary=adv_pol, adv_record=adv_record)
        return {k: torch.as_tensor(v) for k, v in r.items()}


def _policy(G, who):
    p = who + '__'
    return H.orc.OraclePolicy.from_state_dict({k[len(p):].replace('__', '.'): torch.as_tensor(G[k]) for k in G.files if k.startswith(p)}, activation='tanh')


@pytest.mark.parametrize('case,tol', [('quad2d_dynamics', 5e-5)])
def test_rap_composition_returns_the_buffers_the_reference_rap_returned(case, tol):
    G = np.load(os.path.join(GOLD, f'rap_loop_{case}.npz'))
    cfg = json.loads(str(G['task_config']))
    env = make(str(G['env_id']), **cfg)
    T, N = G['pro_rew'].shape[:2]
    K = int(G['num_adversaries'])
    o = H.make_oracle(env, N, auto_reset=True, rng_mode='table')
    o.set_episode_table(H.episode_table(env, N, K=int((1 - G['pro_mask']).sum(0).max()) + 4, seed=int(G['seed'])))
    o.set_stale_force_point(True)                              # dynamics adversary on the quadrotor (tests/test_refloop_golden.py)
    o.reset()                                                  # the controller's reset(); collect_rollouts resets once more (rap.py:360)
    pol = _policy(G, 'agent')
    dims = (pol.obs_dim, pol.act_dim, pol.hidden)
    advs = [_policy(G, f'adversary{i}') for i in range(K)]
    ctl = R.RAP.__new__(R.RAP)                                  # (the constructor opens a CUDA env; the collector logic is what is under test)
    eng = OracleEngine(o, G['eps'], G['eps_adv'])
    ctl.env = type('Env', (), {'unwrapped': type('U', (), {'engine': eng})(), 'reset': lambda self: (o.reset(), None)})()
    ctl.rollout_steps, ctl.rollout_batch_size, ctl.num_adversaries = T, N, K
    ctl.activation, ctl.gamma, ctl.gae_lambda, ctl.use_gae = 'tanh', float(G['gamma']), float(G['gae_lambda']), True
    ctl.total_steps, ctl._rap_out = 0, {}
    ctl._adv_rng = type('Rng', (), {'randint': lambda self, k, size: np.array(G['adv_indices'])})()     # the reference's draw (numpy global generator)
    ctl._policy_blob = lambda: (pol, dims)
    mk = lambda a: type('Agent', (), {'fused': type('F', (), {'rollout_blob': lambda self: (a, (a.obs_dim, a.act_dim, a.hidden))})()})()
    ctl.adversaries = [mk(a) for a in advs]
    rollouts, splits, stats = ctl.collect_rollouts()
    np.testing.assert_array_equal(rollouts['mask'].numpy(), G['pro_mask'])
    for k in ('obs', 'act', 'rew', 'v', 'logp', 'terminal_v'):
        np.testing.assert_allclose(rollouts[k].numpy(), G[f'pro_{k}'], rtol=tol, atol=tol, err_msg='protagonist ' + k)
    np.testing.assert_allclose(rollouts['ret'].numpy(), G['pro_ret'], rtol=10 * tol, atol=10 * tol)
    np.testing.assert_allclose(rollouts['adv'].numpy(), G['pro_adv'], rtol=0, atol=5e-3, err_msg='protagonist normalised adv')
    drawn = sorted(set(int(g) for g in G['adv_indices']))
    assert [g for g, _ in splits] == drawn
    for g, d in splits:
        cols = np.nonzero(G['adv_indices'] == g)[0]
        assert d['act'].shape == G[f'adv{g}_act'].shape == (T, len(cols), advs[g].act_dim)
        np.testing.assert_array_equal(d['mask'].numpy(), G[f'adv{g}_mask'])
        for k in ('obs', 'act', 'rew', 'v', 'logp', 'terminal_v'):
            np.testing.assert_allclose(d[k].numpy(), G[f'adv{g}_{k}'], rtol=tol, atol=tol, err_msg=f'adversary {g} {k}')
        np.testing.assert_allclose(d['ret'].numpy(), G[f'adv{g}_ret'], rtol=10 * tol, atol=10 * tol)
        np.testing.assert_allclose(d['adv'].numpy(), G[f'adv{g}_adv'], rtol=0, atol=5e-3, err_msg=f'adversary {g} normalised adv')
    assert int(stats['ep_count'].sum()) == int((1 - G['pro_mask']).sum())
