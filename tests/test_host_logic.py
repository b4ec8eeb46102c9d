'''Host-side logic of the drop-in boundary: registry, config resolution, constraints, disturbances, RNG draw order.'''
import numpy as np
import pytest

import helpers as H
from safe_control_gym_b200 import _abi, configs
from safe_control_gym_b200.utils.registration import get_config, make, register, spec


def test_registry_matches_the_reference_ids_and_defaults():
    assert spec('cartpole').idx == 'cartpole' and spec('quadrotor').idx == 'quadrotor' and spec('ppo').idx == 'ppo'
    q = get_config('quadrotor')
    assert (q['ctrl_freq'], q['pyb_freq'], q['quad_type'], q['constraint_penalty'], q['norm_act_scale']) == (60, 240, 2, -1, 0.1)
    c = get_config('cartpole')
    assert (c['ctrl_freq'], c['pyb_freq'], c['randomized_init'], c['rew_act_weight']) == (50, 50, True, 0.0001)
    p = get_config('ppo')
    assert (p['hidden_dim'], p['rollout_steps'], p['gamma'], p['use_gae']) == (64, 100, 0.99, False)
    with pytest.raises(Exception, match='Cannot re-register'):
        register(idx='cartpole', entry_point='x:y')
    env = make('quadrotor', **get_config('quadrotor'))
    assert env.spec.idx == 'quadrotor' and env.NAME == 'quadrotor'


def test_invalid_frequencies_raise_value_error():
    with pytest.raises(ValueError, match='pyb_freq is not divisible'):
        make('cartpole', ctrl_freq=30, pyb_freq=50)


def test_quadrotor_constants_and_float32_bounds():
    env = H.make_env('quadrotor_2D_track')
    # float32 thrust bounds 0.056323376 / 0.2966829 N (2D), hover 0.1323 N per input
    assert env.physical_action_bounds[0].dtype == np.float32
    np.testing.assert_allclose(env.physical_action_bounds[0], 0.056323376, rtol=1e-7)
    np.testing.assert_allclose(env.physical_action_bounds[1], 0.2966829, rtol=1e-7)
    assert abs(env.hover_thrust - 0.1323) < 1e-12 and env.obs_dim == 12 and env.state_dim == 6 and env.action_dim == 2
    cfg = env.build_config(num_envs=3)
    assert cfg.act_low[0] == float(np.float32(env.physical_action_bounds[0][0]))
    assert cfg.state_high[4] == float(np.float32(85 * np.pi / 180))
    assert (cfg.ctrl_steps, cfg.pyb_freq // cfg.ctrl_freq, cfg.arm) == (250, 20, 0.028)
    env3 = H.make_env('quadrotor_3D_track')
    np.testing.assert_allclose(env3.physical_action_bounds[0], 0.028161688, rtol=1e-7)
    assert env3.obs_dim == 24 and env3.X_GOAL.shape == (251, 12)


def test_randomisation_info_quirk_of_the_reference_is_kept():
    '''quadrotor.py:208/233 overwrite user ranges with the class defaults; honor_randomization_info opts out.'''
    env = H.make_env('quadrotor_2D_track')
    assert env.INIT_STATE_RAND_INFO['init_x'] == {'distrib': 'uniform', 'low': -0.5, 'high': 0.5}
    env2 = H.make_env('quadrotor_2D_track', honor_randomization_info=True)
    assert env2.INIT_STATE_RAND_INFO['init_x']['low'] == -2
    cart = H.make_env('cartpole_stab')      # CartPole has no such overwrite (cartpole.py)
    assert cart.INIT_STATE_RAND_INFO['init_x']['low'] == -2


def test_constraint_rows_follow_reference_numerics():
    env = H.make_env('quadrotor_2D_track')
    kind, strict, ptr, col, val, b = env.constraints.csr()
    assert len(kind) == 16 and env.constraints.num_constraints == 16 and list(np.diff(ptr)) == [1] * 16
    # rows: [-I; I] x <= [-lb; ub] for the state (6) then the input (2); b is float32-rounded (constraints.py:267-268)
    assert list(col[:12]) == list(range(6)) * 2 and list(col[12:]) == [6, 7, 6, 7]
    assert list(val[:6]) == [-1.0] * 6 and b[14] == float(np.float32(0.29)) and b[12] == float(np.float32(-0.06))
    # get_value mirrors np.round(A x - b, 8)
    env.state = np.array([0.1, 0, 1.0, 0, 0.25, 0])
    env.current_noisy_physical_action = np.array([0.1, 0.3])
    g = env.constraints.get_values(env)
    assert g.shape == (16,) and g[10] == round(0.25 - float(np.float32(0.2)), 8) and env.constraints.is_violated(env, c_value=g)
    cart = H.make_env('cartpole', constraints=[{'constraint_form': 'abs_bound', 'constrained_variable': 'state', 'bound': [0.5], 'active_dims': [0]}])
    k2 = cart.constraints.csr()
    assert list(k2[0]) == [_abi.SCG_CON_ABS] and list(k2[3]) == [0] and list(k2[5]) == [0.5]


def test_disturbance_records_and_draw_order():
    dist = {'observation': [{'disturbance_func': 'white_noise', 'std': 0.1}],
            'action': [{'disturbance_func': 'impulse', 'magnitude': 2.0, 'duration': 4, 'decay_rate': 0.5},
                       {'disturbance_func': 'uniform', 'low': -0.1, 'high': 0.1}]}
    env = H.make_env('cartpole', disturbances=dist, seed=5)
    cfg = env.build_config()
    assert list(cfg.n_dist) == [2, 0, 1]
    assert cfg.dist[0][0].kind == _abi.SCG_DIST_IMPULSE and cfg.dist[0][0].step_offset == -1 and cfg.dist[0][1].kind == _abi.SCG_DIST_UNIFORM
    assert cfg.dist[2][0].kind == _abi.SCG_DIST_WHITE and cfg.dist[2][0].dim == 4
    # reference stream of one env: offsets (YAML order of the modes) -> init state; per step: action -> dynamics -> observation
    rng = np.random.Generator(np.random.PCG64(np.random.SeedSequence(5)))
    row = env.draw_episode_row()
    exp_off = float(rng.integers(250))
    exp_init = [rng.uniform(-0.05, 0.05) for _ in range(4)]
    assert row[0] == exp_off and np.allclose(row[-4:], exp_init, rtol=0, atol=0)
    step = env.draw_step_row()
    exp_step = list(rng.uniform(0, 1, size=1)) + list(rng.normal(0, 1, size=4))
    assert np.array_equal(step, exp_step)
    with pytest.raises(ValueError, match='low must be specified'):
        H.make_env('cartpole', disturbances={'action': [{'disturbance_func': 'uniform', 'low': 1}]})


def test_episode_table_reproduces_the_reference_generator_stream():
    env = H.make_env('quadrotor_2D_track')
    tab = H.episode_table(env, num_envs=3, K=2, seed=1000)
    assert tab.shape == (2, 3, 4 + 6)
    rng = np.random.Generator(np.random.PCG64(np.random.SeedSequence(1001)))   # env 1 is seeded seed + 1
    info = env.INIT_STATE_RAND_INFO
    exp = [[rng.uniform(info[k]['low'], info[k]['high']) for k in env._init_keys()] for _ in range(2)]
    np.testing.assert_array_equal(tab[:, 1, 4:], exp)
    assert np.all(tab[:, :, :4] == 0)


def test_trajectory_generators():
    env = H.make_env('cartpole', task='traj_tracking', task_info={'trajectory_type': 'circle', 'num_cycles': 1, 'trajectory_plane': 'zx',
                                                                  'trajectory_position_offset': [0, 0], 'trajectory_scale': 0.2,
                                                                  'stabilization_goal_tolerance': 0.0})
    assert env.X_GOAL.shape == (251, 4)
    t = np.arange(0, 5 + 0.02, 0.02)
    np.testing.assert_allclose(env.X_GOAL[:, 0], 0.2 * np.sin(2 * np.pi / 5 * t), atol=1e-12)   # plane 'zx': x is the second coordinate
    sq = H.make_env('quadrotor', task='traj_tracking', quad_type=2, task_info={'trajectory_type': 'square', 'num_cycles': 2,
                    'trajectory_plane': 'xz', 'trajectory_position_offset': [0, 1], 'trajectory_scale': 1.0})
    assert sq.X_GOAL.shape == (251, 6) and abs(sq.X_GOAL[0, 2] - 1.0) < 1e-12
    with pytest.raises(ValueError, match='Trajectory type'):
        H.make_env('quadrotor', task='traj_tracking', task_info={'trajectory_type': 'spiral', 'num_cycles': 1, 'trajectory_plane': 'xz',
                                                                   'trajectory_position_offset': [0, 0], 'trajectory_scale': 1})


def test_metric_extractor_on_reference_style_and_batched_data():
    '''MetricExtractor (base_experiment.py:380-492): the same numbers from the reference's layout (lists of per-step info dicts) and from the
    batched harness' per-episode arrays; compute_cvar as performance_metrics.py:6-33.'''
    from safe_control_gym_b200.experiments.base_experiment import MetricExtractor, compute_cvar
    rng = np.random.default_rng(0)
    lens = [5, 9, 3, 7]
    eps = [{'rew': rng.random(L), 'mse': rng.random(L), 'cv': (rng.random(L) < 0.3).astype(float)} for L in lens]
    batched = {'length': [np.ones(L, dtype=np.int64) for L in lens], 'reward': [e['rew'] for e in eps], 'mse': [e['mse'] for e in eps],
               'constraint_violation': [e['cv'] for e in eps]}
    ref_style = {'length': [np.ones(L) for L in lens], 'reward': [e['rew'] for e in eps],
                 'info': [[{'mse': e['mse'][t], 'constraint_violation': e['cv'][t]} for t in range(L)] for e, L in zip(eps, lens)]}
    ma, mb = MetricExtractor().compute_metrics(batched), MetricExtractor().compute_metrics(ref_style)
    for k in ma:
        np.testing.assert_allclose(np.asarray(ma[k], dtype=float), np.asarray(mb[k], dtype=float), rtol=1e-12)
    rmse = np.array([np.sqrt(e['mse'].mean()) for e in eps])
    np.testing.assert_allclose(ma['average_rmse'], rmse.mean())
    np.testing.assert_allclose(ma['worst_case_rmse_at_0.5'], np.sort(rmse)[-2:].mean())
    np.testing.assert_allclose(ma['failure_rate'], np.mean([float(e['cv'].any()) for e in eps]))
    assert compute_cvar(np.arange(10.0), 0.3) == 1.0


def test_physics_variants_and_adversary_reach_the_engine_config():
    '''Host resolution of the ABI-v4 fields: `physics` (Physics enum, base_aviary.py:32-40) with the URDF coefficients and the ground-effect clip
    height (base_aviary.py:130-131), `adversary_disturbance` with the adversary spaces of benchmark_env.py:288-295.'''
    from safe_control_gym_b200 import _abi
    env = H.make_env('quadrotor_2D_track', physics='pyb_gnd_drag_dw', adversary_disturbance='dynamics', adversary_disturbance_scale=0.02, adversary_disturbance_offset=0.001)
    cfg = env.build_config(num_envs=4)
    assert cfg.physics == _abi.SCG_PHYS_PYB_GND_DRAG_DW and cfg.adversary_mode == _abi.SCG_ADV_DYNAMICS
    assert (cfg.adversary_scale, cfg.adversary_offset) == (0.02, 0.001)
    assert cfg.gnd_eff_coeff == 11.36859 and cfg.prop_radius == 2.31348e-2 and cfg.drag_coeff_xy == 9.1785e-7 and cfg.drag_coeff_z == 10.311e-7
    # GND_EFF_H_CLIP = 0.25 PROP_RADIUS sqrt(15 MAX_RPM^2 KF GND_EFF_COEFF / MAX_THRUST) with MAX_THRUST = 4 KF MAX_RPM^2
    assert cfg.gnd_eff_h_clip == pytest.approx(0.25 * 2.31348e-2 * np.sqrt(15 * 11.36859 / 4), rel=1e-12)
    assert env.adversary_action_space.shape == (2,) and env.adversary_observation_space is env.observation_space
    for name, mode, dim in (('quadrotor_3D_track', 'dynamics', 3), ('quadrotor_3D_track', 'action', 4), ('cartpole', 'dynamics', 2), ('cartpole', 'action', 1)):
        e = H.make_env(name, adversary_disturbance=mode)
        assert e.adversary_action_space.shape == (dim,), (name, mode)
        assert np.all(e.adversary_action_space.low == -1) and np.all(e.adversary_action_space.high == 1)
    plain = H.make_env('cartpole')
    assert plain.build_config().adversary_mode == _abi.SCG_ADV_NONE and plain.build_config().physics == _abi.SCG_PHYS_PYB
    with pytest.raises(RuntimeError, match='adversary_disturbance does not exist'):
        plain.set_adversary_control(np.zeros(1))
    with pytest.raises(NotImplementedError):
        H.make_env('quadrotor_2D_track', physics='dyn')                 # dead code in the reference too (base_aviary.py:495-498)
    with pytest.raises(NotImplementedError):
        H.make_env('cartpole', adversary_disturbance='observation')


def test_sibling_controllers_are_registered_with_the_reference_defaults():
    '''controllers/__init__.py of the reference registers ppo, sac, ddpg, rarl, rap, safe_explorer_ppo (...): the ones built here carry the yaml defaults.'''
    from safe_control_gym_b200.utils.registration import get_config
    assert get_config('rarl')['agent_iterations'] == 10 and get_config('rarl')['adversary_iterations'] == 10 and get_config('rarl')['rollout_steps'] == 100
    assert get_config('sac')['hidden_dim'] == 256 and get_config('sac')['warm_up_steps'] == 1000 and get_config('sac')['init_temperature'] == 0.2
    assert get_config('ddpg')['warm_up_steps'] == 10000 and get_config('ddpg')['random_process']['func'] == 'OrnsteinUhlenbeckProcess'
    assert get_config('safe_explorer_ppo')['constraint_hidden_dim'] == 10 and get_config('ppo')['mini_batch_size'] == 64
