'''The oracle's two-player rollout against the PPOBuffers the UNMODIFIED reference RARL controller filled: `RARL.collect_rollouts(adversary=False)`
then `(adversary=True)` of /root/reference (controllers/rarl/rarl.py:349-418) on a DummyVecEnv of unmodified reference envs, recorded by
tools/make_golden_rarl_loop.py (tests/golden/rarl_loop_*.npz).

Same weights for both actor-critics, the normals of both `ac.step` calls of every step, env i seeded `seed + i`.  Required back: the protagonist's
record of the first rollout and -- continuing from where it stopped -- the adversary's record of the second: ITS actions (own width), values and
log-probabilities, `-rew`, terminal values of truncated episodes and the bootstrap value from ITS critic, returns / advantages on those; the episode
boundaries identical.  This pins the extended collector's record semantics (scg_rollout_args.d_adv_policy / adv_record) on the reference's own code;
the CUDA kernel is compared with the oracle on the GPU (tests/test_gpu_extended.py).'''
import json
import os

import numpy as np
import pytest

import helpers as H
from safe_control_gym_b200.utils.registration import make

torch = pytest.importorskip('torch')
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def _policy(G, who):
    p = who + '__'
    sd = {k[len(p):].replace('__', '.'): torch.as_tensor(G[k]) for k in G.files if k.startswith(p)}
    return H.orc.OraclePolicy.from_state_dict(sd, activation='tanh')


@pytest.mark.parametrize('case,tol', [('quad2d_dynamics', 2e-5), ('cartpole_action', 2e-4)])
def test_oracle_two_player_rollout_fills_the_buffers_the_reference_rarl_filled(case, tol):
    G = np.load(os.path.join(GOLD, f'rarl_loop_{case}.npz'))
    cfg = json.loads(str(G['task_config']))
    env = make(str(G['env_id']), **cfg)
    T, N = G['pro_rew'].shape[:2]
    K = int((1 - G['pro_mask']).sum(0).max() + (1 - G['adv_mask']).sum(0).max()) + 2
    o = H.make_oracle(env, N, auto_reset=True, rng_mode='table')
    o.set_episode_table(H.episode_table(env, N, K=K, seed=int(G['seed'])))
    if str(G['env_id']) == 'quadrotor' and cfg.get('adversary_disturbance') == 'dynamics':
        o.set_stale_force_point(True)                      # base_aviary.py:268-279 (see tests/test_refloop_golden.py)
    r = o.reset()
    np.testing.assert_allclose(r['obs'], G['obs0'], rtol=1e-9, atol=1e-9)
    pol, adv = _policy(G, 'agent'), _policy(G, 'adversary')
    kw = dict(gamma=float(G['gamma']), gae_lambda=float(G['gae_lambda']), use_gae=True, adversary=adv)
    outs = [o.rollout(pol, T, eps=G['eps'][0], eps_adv=G['eps_adv'][0], adv_record=False, **kw),
            o.rollout(pol, T, eps=G['eps'][1], eps_adv=G['eps_adv'][1], adv_record=True, **kw)]
    for tag, out in zip(('pro', 'adv'), outs):
        np.testing.assert_array_equal(out['mask'], G[f'{tag}_mask'], err_msg=tag + ' mask')
        assert np.array_equal(out['terminal_v'] != 0, G[f'{tag}_terminal_v'] != 0), tag + ' truncation pattern'
        assert out['act'].shape == G[f'{tag}_act'].shape
        for k in ('obs', 'act', 'rew', 'v', 'logp', 'terminal_v'):
            np.testing.assert_allclose(out[k], G[f'{tag}_{k}'], rtol=tol, atol=tol, err_msg=f'{tag} {k}')
        np.testing.assert_allclose(out['ret'], G[f'{tag}_ret'], rtol=10 * tol, atol=10 * tol, err_msg=tag + ' ret')
        a = out['adv'].astype(np.float64)
        np.testing.assert_allclose((a - a.mean()) / (a.std() + 1e-6), G[f'{tag}_adv'], rtol=0, atol=3e-3, err_msg=tag + ' normalised adv')      # rarl.py:417
    # the adversary is paid -rew (rarl.py:394); compute_returns_and_advantages then adds gamma * terminal_v IN PLACE to the buffer's rew (ppo_utils.py:386)
    assert float(G['adv_rew'][G['adv_terminal_v'] == 0].max()) <= 0.0 or not env.rew_exponential
    np.testing.assert_allclose(outs[1]['last_obs'], G['last_obs'], rtol=tol, atol=tol)
