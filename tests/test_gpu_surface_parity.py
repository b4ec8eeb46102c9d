'''GPU parity cases for the rest of the env surface (VERDICT r01 "next" 8): quadratic constraint rows, `cost: quadratic`,
`obs_wrap_angle`, `step` / `periodic` disturbances, the CartPole tap force -- step API and fused collector against the oracle.'''
import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu
torch = pytest.importorskip('torch')


def _np(t):
    return t.detach().cpu().numpy()


def _step_pair(env, N, steps, seed=3, act_scale=1.0, noise=False):
    from safe_control_gym_b200.engine import EngineHandle
    tab = H.episode_table(env, N, K=16, seed=seed)
    o = H.make_oracle(env, N, auto_reset=True)
    eng = EngineHandle(env, num_envs=N, auto_reset=True)
    for h in (o, eng):
        h.set_episode_table(tab)
    if noise:
        ntab = H.noise_table(env, N, T=steps, seed=seed + 1)
        for h in (o, eng):
            h.set_noise_table(ntab)
    ro, rg = o.reset(), eng.reset()
    np.testing.assert_allclose(_np(rg['obs_f64']), ro['obs'], rtol=0, atol=1e-12)
    rng = np.random.default_rng(seed)
    ndone = nviol = 0
    for t in range(steps):
        act = rng.uniform(-act_scale, act_scale, size=(N, env.action_dim))
        so, sg = o.step(act), eng.step(act)
        np.testing.assert_array_equal(_np(sg['done']), so['done'])
        np.testing.assert_allclose(_np(sg['state_f64']), so['state'], rtol=1e-9, atol=1e-10)
        np.testing.assert_allclose(_np(sg['obs_f64']), so['obs'], rtol=1e-9, atol=1e-10)
        np.testing.assert_allclose(_np(sg['rew_f64']), so['rew'], rtol=1e-9, atol=1e-12)
        np.testing.assert_array_equal(_np(sg['cviol']), so['cviol'])
        if o.n_con:
            np.testing.assert_allclose(_np(sg['cvals']), so['cvals'], rtol=0, atol=2e-8)
        ndone += int(so['done'].sum()); nviol += int(so['cviol'].sum())
    eng.close()
    return ndone, nviol


def test_quadratic_constraint_rows():
    '''QuadraticConstraint x'F'PFx <= b (constraints.py:181-231) on the state and on state + input, next to linear rows.'''
    P6 = np.diag([1.0, 0.1, 1.0, 0.1, 4.0, 0.05]); P6[0, 2] = P6[2, 0] = 0.3
    cons = [dict(constraint_form='quadratic_constraint', constrained_variable='state', P=P6.tolist(), b=1.5),
            dict(constraint_form='quadratic_constraint', constrained_variable='input_and_state', active_dims=[0, 2, 6, 7],
                 P=(np.eye(4) + 0.2).tolist(), b=2.0, strict=True),
            dict(constraint_form='default_constraint', constrained_variable='input')]
    env = H.make_env('quadrotor_2D_track', constraints=cons)
    assert env.constraints.num_constraints == 1 + 1 + 4
    ndone, nviol = _step_pair(env, 64, 50)
    assert nviol > 0 and ndone > 0


def test_quadratic_cost_tracking_and_stabilisation():
    '''`cost: quadratic` reward through the diagonal Q / R (symbolic_systems.py:112-121; reference index quirks quadrotor.py:858 / cartpole.py:648).'''
    for name, kw in (('quadrotor_2D_track', dict(cost='quadratic', obs_goal_horizon=0)),
                     ('quadrotor_3D_track', dict(cost='quadratic', obs_goal_horizon=0)),
                     ('cartpole', dict(cost='quadratic', task='traj_tracking', ctrl_freq=50, pyb_freq=50)),
                     ('cartpole', dict(cost='quadratic', task='stabilization', task_info={'stabilization_goal': [0.2], 'stabilization_goal_tolerance': 0.02}))):
        env = H.make_env(name, **kw)
        ndone, _ = _step_pair(env, 40, 40)
        assert ndone >= 0


def test_cartpole_obs_wrap_angle_and_tap_force():
    '''obs_wrap_angle (cartpole.py:598-599) and the dynamics disturbance applied as a tap force at the pole (cartpole.py:552-566).'''
    dist = {'dynamics': [{'disturbance_func': 'white_noise', 'std': 0.5}, {'disturbance_func': 'step', 'magnitude': 0.3, 'step_offset': 5}]}
    env = H.make_env('cartpole', obs_wrap_angle=True, disturbances=dist, ctrl_freq=25, pyb_freq=1000, done_on_out_of_bound=False,
                     init_state=dict(init_x=0.0, init_x_dot=0.0, init_theta=3.0, init_theta_dot=1.0), episode_len_sec=2,
                     task_info={'stabilization_goal': [0], 'stabilization_goal_tolerance': 0.0})
    ndone, _ = _step_pair(env, 48, 60, act_scale=8.0, noise=True)
    assert ndone > 0      # time-limit resets


def test_step_periodic_impulse_disturbances_on_all_modes():
    dist = {'action': [{'disturbance_func': 'step', 'magnitude': 0.01, 'step_offset': 10},
                       {'disturbance_func': 'periodic', 'scale': 0.005, 'frequency': 2.0}],
            'dynamics': [{'disturbance_func': 'periodic', 'scale': 0.02, 'frequency': 1.0},
                         {'disturbance_func': 'impulse', 'magnitude': 0.05, 'duration': 4, 'decay_rate': 0.5}],
            'observation': [{'disturbance_func': 'step', 'magnitude': 0.01}, {'disturbance_func': 'white_noise', 'std': 0.005}]}
    env = H.make_env('quadrotor_2D_track', disturbances=dist, obs_goal_horizon=0)     # observation noise needs obs_dim == state_dim, as in the reference
    ndone, _ = _step_pair(env, 48, 50, noise=True)
    assert ndone > 0


@pytest.mark.parametrize('engine', ['ffma2', 'tcgen05'])
def test_fused_collector_with_quadratic_rows_and_quadratic_cost(engine):
    '''The same surface through the fused collector (constraint VALUES in the record, serial schedule because of done_on_violation).'''
    from safe_control_gym_b200.engine import EngineHandle
    P6 = np.diag([0.5, 0.05, 0.5, 0.05, 2.0, 0.02])
    cons = [dict(constraint_form='quadratic_constraint', constrained_variable='state', P=P6.tolist(), b=2.5),
            dict(constraint_form='default_constraint', constrained_variable='input')]
    for kw in (dict(constraints=cons), dict(constraints=cons, done_on_violation=True), dict(cost='quadratic', obs_goal_horizon=0)):
        env = H.make_env('quadrotor_2D_track', **kw)
        N, T = 96, 80
        pol = H.random_policy(env.obs_dim, env.action_dim, 64, seed=1)
        o = H.make_oracle(env, N, auto_reset=True, rng_mode='philox', seed=9)
        o.reset()
        oo = o.rollout(pol, T, eps=None, use_gae=True, trace=True)
        eng = EngineHandle(env, num_envs=N, auto_reset=True, rng_mode='philox', seed=9)
        eng.reset()
        blob, dims = H.policy_blob(pol, eng.device)
        gg = eng.rollout(blob, dims, T, eps=None, use_gae=True, want=('ret', 'adv', 'cviol', 'cvals', 'mse', 'last', 'episode'), mlp_engine=engine)
        H.compare_rollouts(oo, gg, o.cfg, o.x_goal, N, T, obs_tol=5e-5, keys_extra=('cvals',))
        eng.close()
