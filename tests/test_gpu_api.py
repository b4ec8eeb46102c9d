'''GPU tests of the reference-facing Python surface: make()/reset()/step(), CudaVecEnv, the PPO controller.'''
import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu
torch = pytest.importorskip('torch')


def test_single_env_gym_api_matches_oracle_stream():
    '''BenchmarkEnv.reset()/step() (batch of one on the GPU) against the oracle fed with the same numpy Generator stream.'''
    env = H.make_env('quadrotor_2D_track', seed=11)
    obs, info = env.reset()
    assert obs.shape == (12,) and obs.dtype == np.float64 and info['current_step'] == 0 and info['x_reference'].shape == (251, 6)
    assert info['constraint_values'].shape == (12,)          # state constraints only at reset (benchmark_env.py:355-357)
    ref_env = H.make_env('quadrotor_2D_track', seed=11)
    o = H.make_oracle(ref_env, 1, auto_reset=False)
    o.set_episode_table(np.asarray(ref_env.draw_episode_row()).reshape(1, 1, -1))
    r = o.reset()
    np.testing.assert_allclose(obs, r['obs'][0], atol=1e-12)
    rng = np.random.default_rng(0)
    for t in range(30):
        a = rng.uniform(-1, 1, size=2)
        obs, rew, done, info = env.step(a)
        so = o.step(a.reshape(1, 2))
        np.testing.assert_allclose(obs, so['obs'][0], rtol=1e-9, atol=1e-10)
        assert abs(rew - so['rew'][0]) < 1e-10 and done == bool(so['done'][0]) and info['constraint_violation'] == int(so['cviol'][0])
        np.testing.assert_allclose(info['constraint_values'], so['cvals'][0], atol=2e-8)
        assert set(info) >= {'mse', 'out_of_bounds', 'current_step', 'constraint_values', 'constraint_violation'}
        if done:
            break
    env.close()
    with pytest.raises(RuntimeError, match='reset'):
        H.make_env('cartpole').step(np.zeros(1))


def test_cartpole_c1_anchor_linear_feedback_200_steps():
    '''BASELINE config 1: CartPole stabilisation, 1 env, 200 steps, zero disturbance, fixed linear state feedback
    (the reference PID cannot drive CartPole, pid.py:59-60).  GPU (f64) vs oracle, rtol 1e-4 per the north star (observed ~1e-12).'''
    K = np.array([-3.1623, -4.2691, -37.3716, -8.5106])      # LQR-like gain, applied identically on both sides
    kw = dict(randomized_init=False, init_state=np.array([0.1, 0.0, 0.05, 0.0]), task_info={'stabilization_goal': [0], 'stabilization_goal_tolerance': 0.0},
              ctrl_freq=50, pyb_freq=50, episode_len_sec=5)
    env = H.make_env('cartpole', **kw)
    o = H.make_oracle(H.make_env('cartpole', **kw), 1, auto_reset=False)
    o.set_episode_table(np.zeros((1, 1, 8)))
    obs, _ = env.reset()
    o.reset()
    for t in range(200):
        u = float(-K @ obs)
        obs, rew, done, info = env.step(np.array([u]))
        so = o.step(np.array([[u]]))
        np.testing.assert_allclose(obs, so['obs'][0], rtol=1e-4, atol=1e-9)
        np.testing.assert_allclose(rew, so['rew'][0], rtol=1e-4)
        assert done == bool(so['done'][0])
    assert abs(obs[2]) < 0.05 and not done                     # the pole is balanced
    env.close()


def test_cuda_vec_env_api_and_terminal_infos():
    from functools import partial

    from safe_control_gym_b200 import configs
    from safe_control_gym_b200.envs.env_wrappers.vectorized_env import make_vec_envs
    from safe_control_gym_b200.utils.registration import make
    env_func = partial(make, 'quadrotor', **{k: v for k, v in configs.get('quadrotor_2D_track').items() if k != 'seed'})
    venv = make_vec_envs(env_func, None, batch_size=64, n_processes=1, seed=3)
    obs, info = venv.reset()
    assert obs.is_cuda and tuple(obs.shape) == (64, 12) and len(info['n']) == 64
    saw_done = False
    for t in range(40):
        act = torch.rand(64, 2, device=obs.device) * 2 - 1
        obs, rew, done, info = venv.step(act)
        assert rew.is_cuda and done.dtype == torch.bool and tuple(rew.shape) == (64,)
        if bool(done.any()):
            i = int(torch.nonzero(done)[0])
            inf = info['n'][i]
            assert 'terminal_observation' in inf and 'terminal_info' in inf and inf['terminal_observation'].shape == (12,)
            assert inf['terminal_info']['current_step'] >= 1
            saw_done = True
    assert saw_done
    assert venv.get_attr('CTRL_FREQ') == [50] * 64
    st = venv.get_env_random_state()
    a = venv.step(torch.zeros(64, 2, device=obs.device))[0].clone()
    venv.set_env_random_state(st)
    b = venv.step(torch.zeros(64, 2, device=obs.device))[0]
    assert torch.equal(a, b)
    venv.close()


def test_kernel_policy_forward_matches_torch_on_shipped_checkpoint():
    '''128-128 tanh actor-critic shipped with the reference (E=16 tile path): deterministic rollout => act = mean, v = value.'''
    from oracle import oracle as orc
    from safe_control_gym_b200.engine import EngineHandle
    g = np.load(f'{H.GOLDEN}/ppo_quadrotor_2D_track.npz')
    w = lambda k: g[k.replace('.', '__')]
    pol = orc.OraclePolicy([w(f'actor.pi_net.fcs.{i}.weight') for i in range(3)], [w(f'actor.pi_net.fcs.{i}.bias') for i in range(3)],
                           [w(f'critic.v_net.fcs.{i}.weight') for i in range(3)], [w(f'critic.v_net.fcs.{i}.bias') for i in range(3)],
                           w('actor.logstd'), 'tanh')
    env = H.make_env('quadrotor_2D_track')
    N, T = 48, 25
    eng = EngineHandle(env, num_envs=N, auto_reset=True, rng_mode='philox', seed=4)
    eng.reset()
    blob, dims = H.policy_blob(pol, eng.device)
    out = eng.rollout(blob, dims, T, deterministic=True)
    obs = out['obs'].cpu()
    sd = {k: torch.from_numpy(w(k)) for k in ['actor.pi_net.fcs.0.weight', 'actor.pi_net.fcs.0.bias', 'actor.pi_net.fcs.1.weight',
                                              'actor.pi_net.fcs.1.bias', 'actor.pi_net.fcs.2.weight', 'actor.pi_net.fcs.2.bias',
                                              'critic.v_net.fcs.0.weight', 'critic.v_net.fcs.0.bias', 'critic.v_net.fcs.1.weight',
                                              'critic.v_net.fcs.1.bias', 'critic.v_net.fcs.2.weight', 'critic.v_net.fcs.2.bias']}

    def mlp(prefix, x):
        for i in range(3):
            x = torch.nn.functional.linear(x, sd[f'{prefix}.fcs.{i}.weight'], sd[f'{prefix}.fcs.{i}.bias'])
            if i < 2:
                x = torch.tanh(x)
        return x
    mean, val = mlp('actor.pi_net', obs), mlp('critic.v_net', obs)
    np.testing.assert_allclose(out['act'].cpu().numpy(), mean.numpy(), rtol=1e-5, atol=3e-6)      # rtol 1e-5
    np.testing.assert_allclose(out['v'].cpu().numpy(), val.numpy(), rtol=1e-5, atol=2e-5)
    eng.close()


def test_ppo_train_step_run_and_checkpoint(tmp_path):
    from functools import partial

    from safe_control_gym_b200 import configs
    from safe_control_gym_b200.utils.registration import make
    task = {k: v for k, v in configs.get('quadrotor_2D_track').items() if k != 'seed'}
    env_func = partial(make, 'quadrotor', **task)
    ctrl = make('ppo', env_func, training=True, checkpoint_path=str(tmp_path / 'model_latest.pt'), output_dir=str(tmp_path), seed=7,
                rollout_batch_size=256, rollout_steps=50, mini_batch_size=1024, opt_epochs=2, use_gae=True, hidden_dim=64)
    ctrl.reset()
    r1 = ctrl.train_step()
    assert ctrl.total_steps == 256 * 50 and np.isfinite([r1['policy_loss'], r1['value_loss'], r1['entropy_loss'], r1['approx_kl']]).all()
    ev = ctrl.run(n_episodes=8)
    assert ev['ep_returns'].shape == (8,) and (ev['ep_lengths'] >= 1).all()
    a = ctrl.select_action(np.zeros(12))
    assert a.shape == (2,)
    ctrl.save(str(tmp_path / 'ck.pt'))
    sd = torch.load(str(tmp_path / 'ck.pt'), weights_only=False)
    assert set(sd['agent']['ac']) >= {'actor.logstd', 'actor.pi_net.fcs.0.weight', 'critic.v_net.fcs.2.bias'}   # reference checkpoint keys
    ref_obs = ctrl.train_step()
    o1 = ctrl.obs.clone()
    ctrl.load(str(tmp_path / 'ck.pt'))
    ctrl.train_step()
    assert torch.allclose(ctrl.obs, o1)       # resume reproduces the same rollout (env state + RNG counters restored)
    ctrl.close()


def test_value_improves_on_cartpole_training():
    '''A few PPO iterations on the fused collector must not diverge and should raise the mean episode return.'''
    from functools import partial

    from safe_control_gym_b200.utils.registration import make
    env_func = partial(make, 'cartpole', normalized_rl_action_space=True, rew_state_weight=[1, 1, 1, 1], rew_act_weight=0.1,
                       task_info={'stabilization_goal': [0], 'stabilization_goal_tolerance': 0.0})
    ctrl = make('ppo', env_func, seed=1, rollout_batch_size=512, rollout_steps=100, mini_batch_size=2048, opt_epochs=5, use_gae=True,
                hidden_dim=64, target_kl=0.02)
    ctrl.reset()
    lens = []
    for it in range(12):
        ctrl.train_step()
        lens.append(np.mean(ctrl.env.length_queue) if len(ctrl.env.length_queue) else 0)
    assert lens[-1] > lens[1] * 1.2, lens
    ctrl.close()


def test_host_collector_pipelined_records_equal_direct_rollouts():
    '''engine.HostCollector (host parameters -> scg_policy_pack -> H2D -> fused rollout -> pinned host record, two records in flight)
    returns bit-identical records to the same sequence of direct device rollouts.'''
    from safe_control_gym_b200.controllers.ppo.ppo_utils import MLPActorCritic
    from safe_control_gym_b200.engine import EngineHandle, HostCollector, pack_policy
    env = H.make_env('quadrotor_2D_track')
    N, T, K = 96, 40, 5
    torch.manual_seed(3)
    ac = MLPActorCritic(env.observation_space, env.action_space, hidden_dims=[64, 64], activation='tanh')
    recs = []
    for mode in ('direct', 'host', 'host_update_keys'):
        eng = EngineHandle(env, num_envs=N, auto_reset=True, rng_mode='philox', seed=21)
        eng.reset()
        got = []
        if mode == 'direct':
            blob, dims = pack_policy(ac.to(eng.device), eng.device)
            ac.cpu()
            for _ in range(K):
                out = eng.rollout(blob, dims, T, use_gae=True, want=('ret', 'adv'))
                got.append({k: out[k].cpu().numpy().copy() for k, _ in out['_layout']})
        else:
            keys = HostCollector.UPDATE_KEYS if mode == 'host_update_keys' else None     # only what PPOAgent.update reads travels to the host
            hc = HostCollector(eng, (12, 2, 64), T, depth=2, use_gae=True, want=('ret', 'adv'), keys=keys)
            for i in range(K):
                hc.submit(ac)
                if i >= 1:
                    got.append({k: v.copy() for k, v in hc.fetch().items()})
            got.append({k: v.copy() for k, v in hc.fetch().items()})
            with pytest.raises(RuntimeError):
                hc.fetch()
            assert hc.h2d_bytes > 0 and hc.d2h_bytes == 4 * T * N * (12 + 2 + (7 if keys is None else 4))
            assert set(got[0]) == (set(keys) if keys else {'obs', 'act', 'rew', 'mask', 'v', 'logp', 'terminal_v', 'ret', 'adv'})
        eng.close()
        recs.append(got)
    assert len(recs[0]) == len(recs[1]) == len(recs[2]) == K
    for a, b, c in zip(*recs):
        for k in a:
            np.testing.assert_array_equal(a[k], b[k], err_msg=k)
        for k in c:
            np.testing.assert_array_equal(a[k], c[k], err_msg=k)


def test_ppo_run_matches_an_oracle_deterministic_rollout(tmp_path):
    '''PPO.run (ppo.py:210-257: deterministic actions, first episode of each eval env) as one fused rollout against the same evaluation
    done by the oracle: returns, lengths, constraint violations and mean squared error per episode.'''
    from functools import partial

    from oracle import oracle as orc
    from safe_control_gym_b200 import configs
    from safe_control_gym_b200.utils.registration import make
    task = {k: v for k, v in configs.get('quadrotor_2D_track').items() if k != 'seed'}
    env_func = partial(make, 'quadrotor', **task)
    ctrl = make('ppo', env_func, training=True, output_dir=str(tmp_path), seed=7, rollout_batch_size=64, rollout_steps=20, mini_batch_size=256,
                opt_epochs=1, hidden_dim=64)
    ctrl.reset()
    ctrl.train_step()
    n_ep = 24
    ev = ctrl.run(n_episodes=n_ep)
    env = H.make_env('quadrotor_2D_track')
    T = int(env.CTRL_STEPS)
    pol = orc.OraclePolicy.from_state_dict({k: v.detach().cpu() for k, v in ctrl.agent.ac.state_dict().items()}, 'tanh')
    o = H.make_oracle(env, n_ep, auto_reset=True, rng_mode='philox', seed=ctrl.seed * 111)
    o.reset()
    oo = o.rollout(pol, T, deterministic=True)
    mask = oo['mask'][..., 0]
    first_done = np.where((mask == 0).any(0), (mask == 0).argmax(0), T - 1)
    live = (np.arange(T)[:, None] <= first_done[None, :]).astype(np.float64)
    np.testing.assert_array_equal(ev['ep_lengths'], first_done + 1)
    np.testing.assert_allclose(ev['ep_returns'], (oo['rew'][..., 0] * live).sum(0), rtol=2e-4, atol=1e-4)
    np.testing.assert_allclose(ev['constraint_violation'], (oo['cviol'][..., 0] * live).sum(0), rtol=0, atol=0)
    np.testing.assert_allclose(ev['mse'], (oo['mse'][..., 0] * live).sum(0) / (first_done + 1), rtol=2e-4, atol=1e-5)
    ctrl.close()


def test_batched_experiment_metrics_match_the_oracle(tmp_path):
    '''BaseExperiment.run_evaluation + MetricExtractor (base_experiment.py:49-165,380-492) on the fused collector: every metric against the
    same evaluation done step by step by the oracle.'''
    from functools import partial

    from oracle import oracle as orc
    from safe_control_gym_b200 import configs
    from safe_control_gym_b200.experiments.base_experiment import BaseExperiment, MetricExtractor
    from safe_control_gym_b200.utils.registration import make
    task = {k: v for k, v in configs.get('quadrotor_2D_track').items() if k != 'seed'}
    ctrl = make('ppo', partial(make, 'quadrotor', **task), training=True, output_dir=str(tmp_path), seed=3, rollout_batch_size=64, rollout_steps=20,
                mini_batch_size=256, opt_epochs=1)
    ctrl.reset()
    ctrl.train_step()
    n_ep = 32
    exp = BaseExperiment(ctrl)
    data, metrics = exp.run_evaluation(n_episodes=n_ep)
    env = H.make_env('quadrotor_2D_track')
    T = int(env.CTRL_STEPS)
    pol = orc.OraclePolicy.from_state_dict({k: v.detach().cpu() for k, v in ctrl.agent.ac.state_dict().items()}, 'tanh')
    o = H.make_oracle(env, n_ep, auto_reset=True, rng_mode='philox', seed=ctrl.seed * 111)
    o.reset()
    oo = o.rollout(pol, T, deterministic=True)
    mask = oo['mask'][..., 0]
    first = np.where((mask == 0).any(0), (mask == 0).argmax(0), T - 1)
    ref = {'length': [np.ones(f + 1) for f in first], 'reward': [oo['rew'][:f + 1, i, 0].astype(np.float64) for i, f in enumerate(first)],
           'mse': [oo['mse'][:f + 1, i, 0].astype(np.float64) for i, f in enumerate(first)],
           'constraint_violation': [oo['cviol'][:f + 1, i, 0].astype(np.float64) for i, f in enumerate(first)]}
    mref = MetricExtractor().compute_metrics(ref)
    for k in ('average_length', 'failure_rate', 'average_constraint_violation'):
        assert metrics[k] == mref[k], k
    for k in ('average_return', 'average_rmse', 'rmse_std', 'worst_case_rmse_at_0.5'):
        np.testing.assert_allclose(metrics[k], mref[k], rtol=2e-4, atol=1e-6, err_msg=k)
    assert len(data['obs']) == n_ep and data['obs'][0].shape[1] == 12
    exp.close()
    ctrl.close()
