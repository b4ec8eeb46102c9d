'''The oracle's rollout with the safety layer against the SafePPOBuffer the UNMODIFIED reference SafeExplorerPPO controller filled:
`SafeExplorerPPO.train_step()` of /root/reference (controllers/safe_explorer/safe_ppo.py:299-355; the projection of safe_explorer_utils.py between the
actor's sample and the env, fed with the constraint values of the previous step -- of the reset state after an auto-reset) on a DummyVecEnv of
unmodified reference envs, recorded by tools/make_golden_safe_loop.py (tests/golden/safe_loop_*.npz).

Same actor-critic and constraint-model weights, normals and per-env seeds; required back: obs, the PROJECTED action the env received, rew, mask
(identical), v, logp, terminal_v, the constraint values `c` the layer saw at every step, ret and the normalised adv.  This pins the SafeExplorer
collector (scg_rollout_args.d_safety) on the reference's own controller code; the CUDA kernel is compared with the oracle on the GPU
(tests/test_gpu_safe_explorer.py).'''
import json
import os

import numpy as np
import pytest

import helpers as H
from safe_control_gym_b200.controllers.safe_explorer.safe_explorer_utils import SafetyLayer
from safe_control_gym_b200.utils.registration import make

torch = pytest.importorskip('torch')
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


@pytest.mark.parametrize('case,tol', [('quad2d_track', 5e-5), ('cartpole_stab', 5e-4)])
def test_oracle_safe_rollout_fills_the_buffer_the_reference_controller_filled(case, tol):
    G = np.load(os.path.join(GOLD, f'safe_loop_{case}.npz'))
    cfg = json.loads(str(G['task_config']))
    env = make(str(G['env_id']), **cfg)
    T, N = G['buf_rew'].shape[:2]
    C = int(G['num_constraints'])
    K = int((1 - G['buf_mask']).sum(0).max()) + 2
    o = H.make_oracle(env, N, auto_reset=True, rng_mode='table')
    o.set_episode_table(H.episode_table(env, N, K=K, seed=int(G['seed'])))
    r = o.reset()
    np.testing.assert_allclose(r['obs'], G['obs0'], rtol=1e-9, atol=1e-9)
    sd = {k[len('ac__'):].replace('__', '.'): torch.as_tensor(G[k]) for k in G.files if k.startswith('ac__')}
    pol = H.orc.OraclePolicy.from_state_dict(sd, activation='tanh')
    layer = SafetyLayer(env.observation_space, env.action_space, hidden_dim=int(G['constraint_hidden_dim']), num_constraints=C, lr=1e-3, slack=float(G['slack']))
    layer.constraint_models.load_state_dict({k[len('safety__'):].replace('__', '.'): torch.as_tensor(G[k]) for k in G.files if k.startswith('safety__')})
    p = layer.pack('cpu')
    out = o.rollout(pol, T, eps=G['eps'].astype(np.float32), gamma=float(G['gamma']), gae_lambda=float(G['gae_lambda']), use_gae=True,
                    safety=dict(p, blob=p['blob'].numpy()))
    np.testing.assert_array_equal(out['mask'], G['buf_mask'])
    assert np.array_equal(out['terminal_v'] != 0, G['buf_terminal_v'] != 0)
    np.testing.assert_allclose(out['c'][0], G['c0'], rtol=tol, atol=tol, err_msg='constraint values of the reset state')
    for k in ('obs', 'act', 'c', 'rew', 'v', 'logp', 'terminal_v'):
        np.testing.assert_allclose(out[k], G[f'buf_{k}'], rtol=tol, atol=tol, err_msg=k)
    np.testing.assert_allclose(out['ret'], G['buf_ret'], rtol=10 * tol, atol=10 * tol, err_msg='ret')
    a = out['adv'].astype(np.float64)
    np.testing.assert_allclose((a - a.mean()) / (a.std() + 1e-6), G['buf_adv'], rtol=0, atol=5e-3, err_msg='normalised adv')
    # the projection was active: the recorded action is not the plain sample mean + sd * eps everywhere
    plain = o2 = None
    fresh = H.make_oracle(env, N, auto_reset=True, rng_mode='table')
    fresh.set_episode_table(H.episode_table(env, N, K=K, seed=int(G['seed'])))
    fresh.reset()
    plain = fresh.rollout(pol, T, eps=G['eps'].astype(np.float32), gamma=float(G['gamma']), gae_lambda=float(G['gae_lambda']), use_gae=True)
    assert np.abs(plain['act'][0] - G['buf_act'][0]).max() > 1e-3
