'''The oracle's fused rollout against the PPOBuffer the UNMODIFIED reference controller filled: `PPO.train_step()` of /root/reference
(controllers/ppo/ppo.py:259-303) on a DummyVecEnv of unmodified reference envs, recorded by tools/make_golden_ppo_loop.py (PyBullet replaced by the
stand-in of tools/ref_harness.py; tests/golden/ppo_loop_*.npz is the committed output).

The oracle gets what the reference run had: the actor-critic weights, the standard normals `Normal.sample()` drew, env i seeded `seed + i`
(vectorized_env/__init__.py:31-35) -- through the host env's own draw functions -- and must return the same buffer: obs, act, rew, mask, v, logp,
terminal_v (critic of the terminal observation where the time limit truncated, ppo.py:274-283), ret and the batch-normalised adv (ppo.py:288-300), and
the observation the next rollout starts from.  This is the collector loop the fused CUDA kernel replaces, pinned on the
reference's own controller code; the kernels are compared with the oracle on the GPU.

Tolerances: the policy is float32 on both sides but torch sums its GEMMs in another order than the oracle's loops (1e-7 relative on the mean action);
the closed loop carries that through the horizon, so trajectories are compared at 2e-5 (50-step episodes) or at the north-star tolerance 1e-4 (250-step episodes: 5 of 7200 observation entries exceed
2e-5, the largest by 3e-5), and the open-loop-unstable cart-pole freely over its first steps and teacher-forced over the whole buffer.'''
import json
import os

import numpy as np
import pytest

import helpers as H
from safe_control_gym_b200.utils.registration import make

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


@pytest.mark.parametrize('case,tol,horizon', [('quad2d_track', 1e-4, None), ('quad2d_track_short', 2e-5, None), ('cartpole_stab', 2e-4, 12)])
def test_oracle_rollout_fills_the_buffer_the_reference_controller_filled(case, tol, horizon):
    G = np.load(os.path.join(GOLD, f'ppo_loop_{case}.npz'))
    cfg = json.loads(str(G['task_config']))
    env = make(str(G['env_id']), **cfg)
    T, N = G['buf_rew'].shape[:2]
    seed = int(G['seed'])
    K = int((1 - G['buf_mask']).sum(0).max()) + 2                    # episodes any env needs
    o = H.make_oracle(env, N, auto_reset=True, rng_mode='table')
    o.set_episode_table(H.episode_table(env, N, K=K, seed=seed))     # env i draws with its own generator seeded seed + i, in the reference's order
    r = o.reset()
    np.testing.assert_allclose(r['obs'], G['obs0'], rtol=1e-9, atol=1e-9, err_msg='observations after PPO.reset()')
    sd = {k[len('ac__'):].replace('__', '.'): G[k] for k in G.files if k.startswith('ac__')}
    import torch
    pol = H.orc.OraclePolicy.from_state_dict({k: torch.as_tensor(v) for k, v in sd.items()}, activation='tanh')
    use_gae = bool(int(G['use_gae']))
    out = o.rollout(pol, T, eps=G['eps'].astype(np.float32), gamma=float(G['gamma']), gae_lambda=float(G['gae_lambda']), use_gae=use_gae)
    # episode boundaries and truncation bootstraps: identical
    if horizon is None:
        np.testing.assert_array_equal(out['mask'], G['buf_mask'])
        assert np.array_equal(out['terminal_v'] != 0, G['buf_terminal_v'] != 0)
    h = T if horizon is None else horizon
    for k in ('obs', 'act', 'rew', 'v', 'logp', 'terminal_v'):
        np.testing.assert_allclose(out[k][:h], G[f'buf_{k}'][:h], rtol=tol, atol=tol, err_msg=k)
    if horizon is None:
        np.testing.assert_allclose(out['ret'], G['buf_ret'], rtol=10 * tol, atol=10 * tol, err_msg='ret')
        adv = out['adv'].astype(np.float64)
        np.testing.assert_allclose((adv - adv.mean()) / (adv.std() + 1e-6), G['buf_adv'], rtol=0, atol=2e-3, err_msg='normalised adv')     # ppo.py:300
        np.testing.assert_allclose(out['last_obs'], G['last_obs'], rtol=tol, atol=tol, err_msg='observation the next rollout starts from')
    else:
        # beyond the comparable horizon of the unstable plant: what the reference recorded is still a valid record of ITS trajectory -- replay it
        # (teacher-forced one-step predictions) and require every transition of the whole buffer
        fresh = H.make_oracle(env, N, auto_reset=True, rng_mode='table')
        fresh.set_episode_table(H.episode_table(env, N, K=K, seed=seed))
        fresh.reset()
        rp = fresh.rollout(pol, T, eps=G['eps'].astype(np.float32), gamma=float(G['gamma']), gae_lambda=float(G['gae_lambda']), use_gae=use_gae,
                           replay={'obs': G['buf_obs'], 'act': G['buf_act'], 'mask': G['buf_mask']})
        np.testing.assert_array_equal(rp['mask'], G['buf_mask'])                       # the oracle ends the episodes where the reference did
        for k in ('rew', 'v', 'logp', 'act'):
            np.testing.assert_allclose(rp[k], G[f'buf_{k}'], rtol=2e-5, atol=2e-5, err_msg='replay ' + k)
        cont = G['buf_mask'][:-1, :, 0] == 1                                            # next observation where the episode went on
        np.testing.assert_allclose(rp['next_obs'][:-1][cont], G['buf_obs'][1:][cont], rtol=2e-5, atol=2e-5, err_msg='replay next_obs')


@pytest.mark.parametrize('case,tol', [('quad2d_track', 2e-4), ('quad2d_track_short', 2e-4), ('cartpole_stab', 2e-3)])
def test_oracle_deterministic_rollout_reproduces_the_reference_evaluation(case, tol):
    '''`PPO.run(env=eval_env, n_episodes=4)` of the reference (ppo.py:210-257: mean action, one env seeded seed * 111, RecordEpisodeStatistics' per-episode
    return / length / constraint-violation / mse sums) against an oracle deterministic rollout of the same policy on one env seeded the same way.
    (`PPO.run` of this repo is compared with that oracle rollout on the GPU: tests/test_gpu_api.py.)'''
    import torch
    G = np.load(os.path.join(GOLD, f'ppo_loop_{case}.npz'))
    cfg = json.loads(str(G['task_config']))
    env = make(str(G['env_id']), **cfg)
    n_ep = len(G['eval_ep_returns'])
    T = int(G['eval_ep_lengths'].sum())
    o = H.make_oracle(env, 1, auto_reset=True, rng_mode='table')
    o.set_episode_table(H.episode_table(env, 1, K=n_ep + 1, seed=int(G['seed']) * 111))
    o.reset()
    sd = {k[len('ac1__'):].replace('__', '.'): torch.as_tensor(G[k]) for k in G.files if k.startswith('ac1__')}
    pol = H.orc.OraclePolicy.from_state_dict(sd, activation='tanh')
    out = o.rollout(pol, T, deterministic=True, gamma=0.0)            # gamma 0: `rew` stays the env's reward (no gamma * terminal_v added, ppo_utils.py:386)
    ends = np.nonzero(out['mask'][:, 0, 0] == 0)[0]
    assert len(ends) == n_ep and ends[-1] == T - 1, (ends, G['eval_ep_lengths'])
    starts = np.concatenate([[0], ends[:-1] + 1])
    np.testing.assert_array_equal(ends - starts + 1, G['eval_ep_lengths'].astype(int))
    for key, rec in (('ep_returns', 'rew'), ('constraint_violation', 'cviol'), ('mse', 'mse')):
        got = np.array([out[rec][s:e + 1, 0, 0].astype(np.float64).sum() for s, e in zip(starts, ends)])
        np.testing.assert_allclose(got, G[f'eval_{key}'], rtol=tol, atol=tol, err_msg=key)


@pytest.mark.parametrize('case', ['quad2d_track', 'quad2d_track_short', 'cartpole_stab'])
def test_episode_statistics_of_the_rollout_match_the_reference_wrapper(case):
    '''`VecRecordEpisodeStatistics` of the reference (record_episode_statistics.py:139-166) logged one return / length / constraint-violation count per episode
    that ended inside the recorded rollout; the fused rollout hands back per-env sums and counts (scg_rollout_out.d_ep_*).  Recomputed from the reference's own
    buffer -- which the oracle reproduces, test above -- both views must agree: the multiset of per-episode values, and the oracle's per-env sums.'''
    import torch
    G = np.load(os.path.join(GOLD, f'ppo_loop_{case}.npz'))
    cfg = json.loads(str(G['task_config']))
    env = make(str(G['env_id']), **cfg)
    T, N = G['buf_rew'].shape[:2]
    gamma = float(G['gamma'])
    rew = G['buf_rew'][..., 0].astype(np.float64) - gamma * G['buf_terminal_v'][..., 0].astype(np.float64)        # undo `rews += gamma * terminal_vals`
    done = G['buf_mask'][..., 0] == 0
    rets, lens, per_env_ret, per_env_cnt = [], [], np.zeros(N), np.zeros(N, int)
    for n in range(N):
        acc, ln = 0.0, 0
        for t in range(T):
            acc += rew[t, n]; ln += 1
            if done[t, n]:
                rets.append(acc); lens.append(ln); per_env_ret[n] += acc; per_env_cnt[n] += 1
                acc, ln = 0.0, 0
    np.testing.assert_allclose(sorted(rets), sorted(G['stat_returns']), rtol=1e-5, atol=1e-5)
    assert sorted(lens) == sorted(int(x) for x in G['stat_lengths'])
    # the oracle's per-env sums for the same rollout
    K = int(done.sum(0).max()) + 2
    o = H.make_oracle(env, N, auto_reset=True, rng_mode='table')
    o.set_episode_table(H.episode_table(env, N, K=K, seed=int(G['seed'])))
    o.reset()
    sd = {k[len('ac__'):].replace('__', '.'): torch.as_tensor(G[k]) for k in G.files if k.startswith('ac__')}
    pol = H.orc.OraclePolicy.from_state_dict(sd, activation='tanh')
    out = o.rollout(pol, T, eps=G['eps'].astype(np.float32), gamma=gamma, gae_lambda=float(G['gae_lambda']), use_gae=bool(int(G['use_gae'])))
    if case != 'cartpole_stab':                                  # (the cart-pole's free-running trajectories only agree over the first steps)
        np.testing.assert_array_equal(out['ep_count'], per_env_cnt)
        np.testing.assert_allclose(out['ep_return_sum'], per_env_ret, rtol=2e-4, atol=2e-4)
        assert float(out['ep_length_sum'].sum()) == float(sum(lens))
        assert float(out['ep_cviol_sum'].sum()) == float(G['stat_cviol'].sum())
