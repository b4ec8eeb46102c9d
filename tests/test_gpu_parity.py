'''Parity tests proper: the CUDA path (through the C-ABI) against the CPU oracle on identical inputs.

Tolerances: the engine in SCG_PREC_F64 uses the reference's arithmetic (float64 env, float32 policy); it differs from
the oracle only by libm ulps / FMA contraction / summation order => states agree to ~1e-9 and float32 outputs to
~1e-5 relative.  SCG_PREC_F32 is characterised against the north-star tolerance rtol = 1e-4 over 200 steps.
'''
import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu

torch = pytest.importorskip('torch')


def _engine(env, N, **kw):
    from safe_control_gym_b200.engine import EngineHandle
    return EngineHandle(env, num_envs=N, **kw)


def _np(t):
    return t.detach().cpu().numpy()


def test_philox_device_matches_known_answers_and_oracle():
    import ctypes as C

    from oracle import oracle as orc
    from safe_control_gym_b200 import _abi
    rng = np.random.default_rng(0)
    ctr = rng.integers(0, 2**32, size=(257, 4), dtype=np.uint64).astype(np.uint32)
    key = rng.integers(0, 2**32, size=(257, 2), dtype=np.uint64).astype(np.uint32)
    ctr[0] = 0; key[0] = 0
    out = np.empty_like(ctr)
    rc = _abi.lib().scg_philox_selftest(0, ctr.ctypes.data_as(C.c_void_p), key.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), len(ctr))
    assert rc == 0
    assert tuple(int(x) for x in out[0]) == (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)
    np.testing.assert_array_equal(out, orc.philox(ctr, key))


def test_cartpole_golden_pybullet_transitions_through_the_c_abi():
    '''The real-PyBullet transitions, this time stepped by the CUDA kernel (state injected via scg_set_state).'''
    g = np.load(f'{H.GOLDEN}/cartpole_pybullet_transitions.npz')
    S, A, blk = g['state'].astype(np.float64), g['act'].astype(np.float64), int(g['block'])
    idx = np.asarray([k for ep in range(len(S) // blk) for k in range(ep * blk, ep * blk + blk - 1)])
    N = len(idx)
    env = H.make_env('cartpole', ctrl_freq=int(g['ctrl_freq']), pyb_freq=int(g['pyb_freq']), cost='quadratic', task='stabilization',
                     task_info={'stabilization_goal': [0, 0], 'stabilization_goal_tolerance': 0.05}, randomized_init=True,
                     init_state_randomization_info=None)
    for prec, tol in (('f64', 1e-6), ('f32', 2e-5)):
        eng = _engine(env, N, auto_reset=False, precision=prec)
        # the initial states enter as a K=1 episode table of additive samples on a zero nominal state
        tab = np.zeros((1, N, eng.P))
        tab[0, :, eng.P - 4:] = S[idx]
        eng.set_episode_table(tab)
        r = eng.reset()
        np.testing.assert_allclose(_np(r['state_f64']), S[idx], atol=1e-6 if prec == 'f32' else 0)
        out = eng.step(A[idx + 1].reshape(-1, 1))
        err = np.abs(_np(out['state_f64']) - S[idx + 1])
        assert err.max() < tol, (prec, err.max(0))
        eng.close()


CASES = [
    # name, N, T, hidden, extra env kwargs
    ('cartpole_stab', 96, 64, 64, {}),
    ('cartpole', 64, 120, 32, dict(task='traj_tracking', obs_goal_horizon=2, cost='rl_reward', ctrl_freq=50, pyb_freq=50)),
    ('quadrotor_2D_track', 128, 200, 64, {}),
    ('quadrotor_3D_track', 96, 200, 64, {}),
    ('quadrotor_2D_track', 40, 60, 128, dict(done_on_violation=True, use_constraint_penalty=True, constraint_penalty=0.5)),
    # other CTA shapes of the fused collector (envs per CTA): 16 x 128 threads, 8 x 64 threads (owner lanes do everything)
    ('quadrotor_2D_track', 72, 100, 64, {}, 16),
    ('quadrotor_2D_track', 44, 100, 64, {}, 8),
    ('cartpole_stab', 50, 64, 32, {}, 16),
    # 1-D quadrotor (quadrotor.py quad_type 1) through the fused collector
    ('quadrotor', 48, 120, 64, dict(quad_type=1, randomized_init=True, task='traj_tracking', obs_goal_horizon=1, cost='rl_reward', ctrl_freq=50,
                                    pyb_freq=1000, normalized_rl_action_space=True, done_on_out_of_bound=True,
                                    task_info=dict(trajectory_type='figure8', num_cycles=1, trajectory_plane='xz',
                                                   trajectory_position_offset=[0, 1], trajectory_scale=0.5),
                                    init_state=dict(init_x=1.0, init_x_dot=0.0))),
]


def _compare_rollouts(o, g, N, T, obs_tol, state_tol, oracle=None):
    '''o: oracle dict (numpy, trace=True), g: engine dict (torch).  Per env up to the first `mask` mismatch; every mismatch needs a
    proof that it is a float tie at a termination threshold (helpers.compare_rollouts).'''
    return H.compare_rollouts(o, g, oracle.cfg, oracle.x_goal, N, T, obs_tol)[0]


@pytest.mark.parametrize('prec', ['f64', 'f32'])
@pytest.mark.parametrize('case', CASES, ids=[f'{c[0]}-h{c[3]}-{i}' for i, c in enumerate(CASES)])
def test_fused_rollout_matches_oracle(case, prec):
    name, N, T, hid, kw = case[:5]
    epb = case[5] if len(case) > 5 else 0
    env = H.make_env(name, **kw)
    seed = 1000
    tab = H.episode_table(env, N, K=8, seed=seed)
    pol = H.random_policy(env.obs_dim, env.action_dim, hid, seed=0, activation='tanh')
    eps = torch.randn(T, N, env.action_dim, generator=torch.Generator().manual_seed(1))
    # oracle
    o = H.make_oracle(env, N, auto_reset=True, precision=prec)
    o.set_episode_table(tab)
    o.reset()
    oo = o.rollout(pol, T, eps=eps.numpy(), gamma=0.99, gae_lambda=0.95, use_gae=True, trace=True)
    # engine
    eng = _engine(env, N, auto_reset=True, precision=prec)
    eng.set_episode_table(tab)
    eng.reset()
    blob, dims = H.policy_blob(pol, eng.device)
    gg = eng.rollout(blob, dims, T, eps=eps.to(eng.device), gamma=0.99, gae_lambda=0.95, use_gae=True, envs_per_block=epb,
                     want=('ret', 'adv', 'cviol', 'mse', 'last', 'episode'))
    torch.cuda.synchronize()
    if prec == 'f64':
        _compare_rollouts(oo, gg, N, T, obs_tol=2e-5, state_tol=1e-8, oracle=o)
    else:
        _compare_rollouts(oo, gg, N, T, obs_tol=1e-4, state_tol=1e-4, oracle=o)   # north-star tolerance
    assert eng.launch_count() == 2   # reset + ONE fused rollout launch for the whole horizon
    eng.close()


@pytest.mark.parametrize('name,N,steps', [('cartpole_stab', 33, 40), ('quadrotor_2D_track', 70, 60), ('quadrotor_3D_track', 50, 60)])
def test_step_api_matches_oracle_with_auto_reset(name, N, steps):
    '''DummyVecEnv semantics: caller-provided actions, auto-reset with terminal observations, float64 outputs.'''
    env = H.make_env(name)
    tab = H.episode_table(env, N, K=16, seed=7)
    o = H.make_oracle(env, N, auto_reset=True)
    o.set_episode_table(tab)
    eng = _engine(env, N, auto_reset=True)
    eng.set_episode_table(tab)
    ro, rg = o.reset(), eng.reset()
    np.testing.assert_allclose(_np(rg['obs_f64']), ro['obs'], rtol=0, atol=1e-12)
    rng = np.random.default_rng(3)
    ndone = 0
    for t in range(steps):
        act = rng.uniform(-1.5, 1.5, size=(N, env.action_dim))
        so, sg = o.step(act), eng.step(act)
        np.testing.assert_array_equal(_np(sg['done']), so['done'])
        np.testing.assert_array_equal(_np(sg['truncated']), so['truncated'])
        np.testing.assert_allclose(_np(sg['state_f64']), so['state'], rtol=1e-9, atol=1e-10)
        np.testing.assert_allclose(_np(sg['obs_f64']), so['obs'], rtol=1e-9, atol=1e-10)
        np.testing.assert_allclose(_np(sg['rew_f64']), so['rew'], rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(_np(sg['mse']), so['mse'], rtol=1e-8, atol=1e-12)
        np.testing.assert_array_equal(_np(sg['cviol']), so['cviol'])
        np.testing.assert_allclose(_np(sg['cvals']), so['cvals'], rtol=0, atol=2e-8)    # values are rounded to 1e-8
        np.testing.assert_allclose(_np(sg['terminal_obs_f64']), so['terminal_obs'], rtol=1e-9, atol=1e-10)
        np.testing.assert_array_equal(_np(sg['current_step']), so['current_step'])
        np.testing.assert_array_equal(_np(sg['out_of_bounds']), so['out_of_bounds'])
        ndone += int(so['done'].sum())
    assert ndone > 0            # the auto-reset path was exercised
    eng.close()


def test_disturbances_and_inertial_randomisation_match_oracle():
    '''BASELINE config 4 ingredients: parametric (inertial) randomisation + white-noise / impulse / step disturbances.'''
    dist = {
        'action': [{'disturbance_func': 'white_noise', 'std': 0.002}],
        'dynamics': [{'disturbance_func': 'white_noise', 'std': 0.01},
                     {'disturbance_func': 'impulse', 'magnitude': 0.02, 'duration': 6, 'decay_rate': 0.8}],
        'observation': [{'disturbance_func': 'uniform', 'low': [-0.01] * 12, 'high': [0.01] * 12, 'mask': [1, 0] * 6}],
    }
    info = {'M': {'distrib': 'uniform', 'low': -0.004, 'high': 0.004}, 'Ixx': {'distrib': 'uniform', 'low': -1e-6, 'high': 1e-6},
            'Iyy': {'distrib': 'uniform', 'low': -1e-6, 'high': 1e-6}, 'Izz': {'distrib': 'normal', 'loc': 0.0, 'scale': 5e-7}}
    env = H.make_env('quadrotor_3D_track', obs_goal_horizon=0, disturbances=dist, randomized_inertial_prop=True,
                     inertial_prop_randomization_info=info, honor_randomization_info=True)
    N, steps = 48, 50
    tab = H.episode_table(env, N, K=8, seed=11)
    ntab = H.noise_table(env, N, T=steps, seed=5)
    o = H.make_oracle(env, N, auto_reset=True)
    eng = _engine(env, N, auto_reset=True)
    for h in (o, eng):
        h.set_episode_table(tab)
        h.set_noise_table(ntab)
    ro, rg = o.reset(), eng.reset()
    np.testing.assert_allclose(_np(rg['state_f64']), ro['state'], atol=1e-12)
    rng = np.random.default_rng(0)
    for t in range(steps):
        act = rng.uniform(-1, 1, size=(N, 4))
        so, sg = o.step(act), eng.step(act)
        np.testing.assert_array_equal(_np(sg['done']), so['done'])
        np.testing.assert_allclose(_np(sg['state_f64']), so['state'], rtol=1e-9, atol=1e-10)
        np.testing.assert_allclose(_np(sg['obs_f64']), so['obs'], rtol=1e-9, atol=1e-10)
        np.testing.assert_allclose(_np(sg['rew_f64']), so['rew'], rtol=1e-9, atol=1e-12)
    eng.close()


def test_philox_mode_statistics_and_oracle_agreement():
    '''In-kernel randomness: same integer streams as the oracle, transforms to libm tolerance.'''
    env = H.make_env('quadrotor_2D_track', disturbances={'dynamics': [{'disturbance_func': 'white_noise', 'std': 0.005}]})
    N, T, hid = 64, 40, 64
    pol = H.random_policy(env.obs_dim, env.action_dim, hid, seed=2)
    o = H.make_oracle(env, N, auto_reset=True, rng_mode='philox', seed=77)
    o.reset()
    oo = o.rollout(pol, T, eps=None, trace=True)
    eng = _engine(env, N, auto_reset=True, rng_mode='philox', seed=77)
    eng.reset()
    blob, dims = H.policy_blob(pol, eng.device)
    gg = eng.rollout(blob, dims, T, eps=None)
    _compare_rollouts(oo, gg, N, T, obs_tol=5e-5, state_tol=1e-6, oracle=o)
    eng.close()


def test_step_before_reset_raises_like_the_reference():
    env = H.make_env('quadrotor_2D_track')
    eng = _engine(env, 4)
    with pytest.raises(RuntimeError, match='reset'):
        eng.step(np.zeros((4, 2)))
    eng.close()


def test_state_checkpoint_roundtrip():
    env = H.make_env('quadrotor_2D_track')
    N = 32
    eng = _engine(env, N, auto_reset=True)
    eng.set_episode_table(H.episode_table(env, N, K=4, seed=1))
    eng.reset()
    rng = np.random.default_rng(0)
    acts = rng.uniform(-1, 1, size=(20, N, 2))
    for t in range(10):
        eng.step(acts[t])
    snap = eng.get_state()
    a = [_np(eng.step(acts[t])['state_f64']).copy() for t in range(10, 20)]
    eng.set_state(snap)
    b = [_np(eng.step(acts[t])['state_f64']).copy() for t in range(10, 20)]
    np.testing.assert_array_equal(np.asarray(a), np.asarray(b))
    eng.close()


def test_fused_peer_gather_writes_identical_copies():
    '''Fused rollout + all-gather (scg_rollout_args.n_peers): every PPOBuffer store is repeated at pointer + peer offset.  On one
    GPU the "peers" are two more slots of the same allocation; all copies must be bit-identical and equal to a plain rollout.'''
    from safe_control_gym_b200.controllers.ppo.ppo import rollout_flat_numel
    N, T, hid = 80, 60, 64
    env = H.make_env('quadrotor_2D_track')
    pol = H.random_policy(env.obs_dim, env.action_dim, hid, seed=0, activation='tanh')
    numel = rollout_flat_numel(env.obs_dim, env.action_dim, T, N)
    slot = ((numel * 4 + 127) // 128) * 32
    outs = []
    for peers in (0, 2):
        eng = _engine(env, N, auto_reset=True, precision='f64', rng_mode='philox', seed=7)
        eng.reset()
        blob, dims = H.policy_blob(pol, eng.device)
        buf = torch.zeros(3 * slot, dtype=torch.float32, device=eng.device)
        offs = [slot * 4 * (g + 1) for g in range(peers)]
        out = eng.rollout(blob, dims, T, use_gae=True, flat=buf[:numel], peer_byte_offsets=offs, want=('ret', 'adv'))
        torch.cuda.synchronize()
        outs.append(buf.clone())
        if peers:
            for g in range(peers):
                assert torch.equal(buf[:numel], buf[slot * (g + 1): slot * (g + 1) + numel])
        else:
            assert float(buf[slot:].abs().max()) == 0.0
        eng.close()
    assert torch.equal(outs[0][:numel], outs[1][:numel])
