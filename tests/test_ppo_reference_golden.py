'''PPO pieces against golden vectors produced by the UNMODIFIED reference module (tools/make_golden_ppo.py imports
safe_control_gym/controllers/ppo/ppo_utils.py from /root/reference; tests/golden/ppo_reference.npz is its committed output):

* `compute_returns_and_advantages` (ppo_utils.py:374-400) -> the torch restatement here, and the ORACLE's in-rollout GAE through it;
* `MLPActorCritic.step` (ppo_utils.py:224-231): sampled action, value, log-probability with torch seeded alike;
* `PPOAgent.update` (ppo_utils.py:82-146) with its sampler's minibatch order: `PPOAgent.update_torch` here -- the parity reference of the fused sm_100a
  update kernel (tests/test_gpu_ppo_update.py) -- must reproduce the returned loss averages and every weight, plain and clipped value loss, with the KL
  gate closing part of the way through.'''
import os

import numpy as np
import pytest

torch = pytest.importorskip('torch')

import helpers as H  # noqa: E402
from safe_control_gym_b200.controllers.ppo.ppo_utils import MLPActorCritic, PPOAgent, compute_returns_and_advantages  # noqa: E402
from safe_control_gym_b200.spaces import Box  # noqa: E402

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'ppo_reference.npz'))
OBS, ACT = Box(-np.ones(12), np.ones(12)), Box(-np.ones(2), np.ones(2))


def _sd(prefix):
    p = prefix + '__'
    return {k[len(p):].replace('__', '.'): torch.as_tensor(G[k]) for k in G.files if k.startswith(p)}


@pytest.mark.parametrize('tag,use_gae', [('mc', False), ('gae', True)])
def test_returns_and_advantages_match_the_reference_function(tag, use_gae):
    t = lambda k: torch.as_tensor(G[k].copy())
    ret, adv = compute_returns_and_advantages(t('gae_rews'), t('gae_vals'), t('gae_masks'), t('gae_term'), t('gae_last'), gamma=0.99, use_gae=use_gae, gae_lambda=0.95)
    np.testing.assert_allclose(ret.numpy(), G[f'gae_ret_{tag}'], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(adv.numpy(), G[f'gae_adv_{tag}'], rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize('use_gae', [False, True])
def test_oracle_rollout_gae_equals_the_reference_recursion(use_gae):
    '''The oracle's in-rollout returns / advantages (the kernels are compared with them on the GPU) recomputed from its own record by the restatement
    the golden above pins on the reference function.'''
    env = H.make_env('quadrotor_2D_track')
    N, T = 12, 90
    o = H.make_oracle(env, N, auto_reset=True)
    o.set_episode_table(H.episode_table(env, N, K=6, seed=3))
    o.reset()
    pol = H.random_policy(env.obs_dim, env.action_dim, 64, seed=1)
    eps = np.random.default_rng(2).normal(size=(T, N, env.action_dim)).astype(np.float32)
    r = o.rollout(pol, T, eps=eps, gamma=0.99, gae_lambda=0.95, use_gae=use_gae)
    assert float((1 - r['mask']).sum()) > 0                                  # episodes end inside the horizon
    t = lambda a: torch.as_tensor(np.array(a, dtype=np.float32))
    ret, adv = compute_returns_and_advantages(t(r['rew']), t(r['v']), t(r['mask']), t(r['terminal_v']), t(r['last_val']), gamma=0.99, use_gae=use_gae, gae_lambda=0.95)
    np.testing.assert_allclose(r['ret'], ret.numpy(), rtol=2e-6, atol=2e-6)
    np.testing.assert_allclose(r['adv'], adv.numpy(), rtol=2e-5, atol=2e-5)


def test_actor_critic_step_matches_the_reference():
    ac = MLPActorCritic(OBS, ACT, hidden_dims=[64, 64], activation='tanh')
    ac.load_state_dict(_sd('step_ac'))                                       # same parameter names as the reference (strict load)
    obs = torch.as_tensor(G['step_obs'])
    torch.manual_seed(6)
    a, v, logp = ac.step(obs)
    np.testing.assert_allclose(a.numpy(), G['step_act'], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(v.numpy(), G['step_v'], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(logp.numpy(), G['step_logp'], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(ac.act(obs).numpy(), G['step_act_det'], rtol=1e-6, atol=1e-6)
    # and the oracle's float32 forward on the same weights (mean action and value): the policy the fused collector evaluates
    pol = H.orc.OraclePolicy.from_state_dict(_sd('step_ac'), activation='tanh')
    mu, val = pol.forward(G['step_obs'])
    np.testing.assert_allclose(mu, G['step_act_det'], rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(np.asarray(val).reshape(-1), G['step_v'].reshape(-1), rtol=2e-5, atol=2e-6)


@pytest.mark.parametrize('tag,clipped', [('upd', False), ('updc', True)])
def test_update_matches_the_reference_agent(tag, clipped, monkeypatch):
    ag = PPOAgent(OBS, ACT, hidden_dim=64, use_clipped_value=clipped, clip_param=0.2, target_kl=0.004, entropy_coef=0.01, actor_lr=3e-3, critic_lr=3e-3,
                  opt_epochs=4, mini_batch_size=32, activation='tanh', fused_update=False)
    ag.to('cpu')
    ag.ac.load_state_dict(_sd(f'{tag}_ac0'))
    rollouts = {k: torch.as_tensor(G[f'{tag}_buf_{k}']) for k in ('obs', 'act', 'logp', 'adv', 'ret', 'v')}
    perms = [torch.as_tensor(p) for p in G[f'{tag}_perms']]                  # the reference sampler's np.random.permutation results, in order
    it = iter(perms)
    monkeypatch.setattr(torch, 'randperm', lambda n, **kw: next(it))
    ag.train()
    res = ag.update_torch(rollouts, 'cpu')
    np.testing.assert_allclose([res['policy_loss'], res['value_loss'], res['entropy_loss'], res['approx_kl']], G[f'{tag}_res'], rtol=2e-4, atol=2e-6)
    want, got = _sd(f'{tag}_ac1'), ag.ac.state_dict()
    assert set(want) == set(got)
    for k, v in want.items():
        np.testing.assert_allclose(got[k].detach().numpy(), v.numpy(), rtol=2e-4, atol=2e-6, err_msg=k)
    # the KL gate closed part of the way through (approx_kl average far above 1.5 * target_kl): the actor moved less than the critic
    assert float(G[f'{tag}_res'][3]) > 1.5 * 0.004
