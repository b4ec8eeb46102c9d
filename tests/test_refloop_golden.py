'''The oracle against trajectories recorded from the UNMODIFIED reference environments run in the loop (tools/make_golden_refloop.py +
tools/ref_harness.py: `Quadrotor.reset()/step()` of the reference executed here with PyBullet replaced by a stand-in rigid body; the files under
tests/golden/refloop_*.npz are its committed output).

Every case rebuilds the recorded task config through this repo's host env (`make('quadrotor', **cfg)`), draws the episode table and the per-step
noise table with the host env's own generator seeded like the reference env (same draws in the same order, interleaved the way the reference
interleaves them: reset draws, then one row per step), hands them to the CPU oracle and steps it with the recorded actions.  Every number the
reference returned must come back: observations (incl. the terminal one and the one after each reset), reward, done, TimeLimit.truncated, info['mse'],
out_of_bounds, goal_reached, constraint values / violation flag, the step counter, the state, X_GOAL.

Pinned by this: everything the reference computes around the integrator (action map, PWM / RPM conversion, force application, observation assembly,
goal look-ahead, reward, termination, constraints, disturbance flow, randomisation order and arithmetic, time limit, reset).  NOT pinned: PyBullet's
integrator itself (the stand-in integrates the way the oracle's restatement does).'''
import json
import os

import numpy as np
import pytest

import helpers as H
from safe_control_gym_b200.envs import disturbances as dist_mod
from safe_control_gym_b200.utils.registration import make

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
CASES = ['quad2d_track', 'quad2d_track_short', 'quad3d_track', 'quad3d_track_noise', 'quad2d_stab_noise', 'quad2d_gnd', 'quad2d_drag', 'quad3d_gnd_drag_dw',
         'quad2d_adv_dyn', 'quad3d_adv_act', 'quad1d_stab_dist', 'quad2d_done_on_violation', 'quad2d_penalty', 'quad2d_custom_constraints', 'quad2d_circle', 'quad3d_square', 'quad2d_quadratic_constraint', 'cartpole_abs_bound', 'cartpole_stab', 'cartpole_track', 'cartpole_noise', 'cartpole_c2', 'cartpole_adv_dyn', 'cartpole_adv_act']
ATOL = 2e-9        # float64 on both sides; the two integrators associate their sums differently, the horizon amplifies that a little


def _close(a, b, what, t, atol=ATOL):
    np.testing.assert_allclose(np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64), rtol=1e-9, atol=atol, err_msg=f'{what} at step {t}')


@pytest.mark.parametrize('case', CASES)
def test_oracle_reproduces_the_reference_in_the_loop(case):
    G = np.load(os.path.join(GOLD, f'refloop_{case}.npz'))
    cfg = json.loads(str(G['task_config']))
    env = make(str(G['env_id']), **cfg)
    T = len(G['rew'])
    # (a16) the reference trajectory generator, whole table
    np.testing.assert_allclose(np.asarray(env._x_goal_table(), dtype=np.float64).reshape(G['x_goal'].shape), G['x_goal'], rtol=0, atol=1e-12)
    assert int(env.CTRL_STEPS) == int(G['ctrl_steps'])
    # the public surface of the env object (benchmark_env.py / quadrotor.py / cartpole.py attributes a controller reads): spaces incl. dtype, bounds, goals, counts
    for mine, key in ((env.observation_space.low, 'obs_low'), (env.observation_space.high, 'obs_high'), (env.action_space.low, 'act_low'), (env.action_space.high, 'act_high'),
                      (env.state_space.low, 'state_low'), (env.state_space.high, 'state_high'), (env.physical_action_bounds[0], 'phys_act_low'),
                      (env.physical_action_bounds[1], 'phys_act_high')):
        assert np.asarray(mine).dtype == G['surf_' + key].dtype, key
        np.testing.assert_array_equal(np.asarray(mine), G['surf_' + key], err_msg=key)
    np.testing.assert_allclose(np.asarray(env.U_GOAL, dtype=np.float64), G['surf_u_goal'], rtol=0, atol=1e-15)
    assert float(env.CTRL_TIMESTEP) == float(G['surf_ctrl_timestep']) and int(env.PYB_STEPS_PER_CTRL) == int(G['surf_pyb_steps_per_ctrl'])
    if env.constraints is not None:
        assert [env.constraints.num_constraints, env.constraints.num_state_constraints, env.constraints.num_input_constraints] == list(G['surf_n_con'])
    # the reference's draws, re-made by the host env's own draw functions on a generator seeded the same way, in the reference's interleaving
    env.seed(cfg['seed'])
    obs_dist = {'observation': env.disturbances['observation']} if env.disturbances.get('observation') is not None else None
    ep_rows, step_rows, reset_draws = [], [], []

    def draw_reset():
        ep_rows.append(env.draw_episode_row())
        # the reference also disturbs the observation that reset() returns (quadrotor.py:387 -> _get_observation): one more draw per reset
        reset_draws.append(dist_mod.draw_step_noise(obs_dist, env.np_random) if obs_dist else [])
    draw_reset()
    for t in range(T):
        step_rows.append(env.draw_step_row())
        if G['done'][t]:
            draw_reset()
    o = H.make_oracle(env, 1, auto_reset=True, rng_mode='table')
    o.set_episode_table(np.asarray(ep_rows, dtype=np.float64)[:, None, :])
    if o.D > 0:
        o.set_noise_table(np.asarray(step_rows, dtype=np.float64)[:, None, :])
    if str(G['env_id']) == 'quadrotor' and (env.disturbances.get('dynamics') is not None or cfg.get('adversary_disturbance') == 'dynamics'):
        # the reference applies the disturbance force at the position the drone had when the control step began (base_aviary.py:268-279 with
        # physics 'pyb'): a small torque.  The oracle has that behind a switch; the CUDA path does not have it.
        o.set_stale_force_point(True)
    nx = env.state_dim

    def check_reset(obs, ep, t):
        # TABLE mode of the oracle / the kernels carries no noise for the reset observation (the episode row has no slot for it; Philox mode draws it
        # from its own stream): the oracle returns the undisturbed one.  So: the oracle's reset observation = the reference's reset STATE (+ goal), and
        # the reference's reset observation = that + the white noise made of the draws the host generator produced at this point of the stream.
        want = np.array(G['reset_obs'][ep], dtype=np.float64)
        if obs_dist:
            noise, at_zero, o_ = np.zeros(nx), np.zeros(nx), 0
            for d in env.disturbances['observation'].disturbances:        # white: std * z (disturbances.py:215-223); uniform: low + (high - low) * u (:187-192)
                z = np.asarray(reset_draws[ep][o_:o_ + d.dim]); o_ += d.dim
                noise += d.std * z if d.step_draws == 'normal' else d.low + (d.high - d.low) * z
                at_zero += 0.0 if d.step_draws == 'normal' else d.low
            _close(G['reset_state'][ep] + noise, want[:nx], 'reference reset observation = state + noise made of the host draw', t)
            # table mode evaluates the reset-observation noise with the draw replaced by 0: nothing for white noise, `low` for uniform noise
            want[:nx] = G['reset_state'][ep] + at_zero
        _close(obs, want, 'observation after a reset', t)
    r = o.reset()
    check_reset(r['obs'][0], 0, -1)
    ep, nc = 0, G['cvals'].shape[1]
    for t in range(T):
        if int(G['adv_dim']) > 0:
            o.set_adversary_control(G['adv'][t][None, :])          # benchmark_env.py:216-228, before every step (rarl.py:367)
        s = o.step(G['act'][t][None, :])
        assert bool(s['done'][0]) == bool(G['done'][t]), f'done at step {t}'
        done = bool(G['done'][t])
        _close(s['terminal_obs'][0] if done else s['obs'][0], G['obs'][t], 'observation', t)
        _close(s['terminal_state'][0] if done else s['state'][0], G['state'][t], 'state', t)
        _close(s['rew'][0], G['rew'][t], 'reward', t)
        _close(s['mse'][0], G['mse'][t], 'mse', t)
        assert bool(s['truncated'][0]) == bool(G['truncated'][t]), f'TimeLimit.truncated at step {t}'
        if s['goal_reached'][0]:
            # stabilisation goal reached: quadrotor.py:871-874 returns before `self.out_of_bounds` is refreshed, so info['out_of_bounds'] of such a
            # step is whatever an earlier step left there (and info has no 'goal_reached' key unless the cost is quadratic, :905-906): the episode must
            # end, the flag itself is not comparable
            assert done, f'goal reached without done at step {t}'
        else:
            assert bool(s['out_of_bounds'][0]) == bool(G['out_of_bounds'][t]), f'out_of_bounds at step {t}'
        assert int(s['current_step'][0]) == int(G['current_step'][t]), f'current_step at step {t}'
        if env.constraints is not None:
            _close(s['cvals'][0][:nc], G['cvals'][t], 'constraint values', t, atol=2e-8)        # (the reference rounds them to 8 decimals: constraints.py:104)
            assert float(s['cviol'][0]) == float(G['cviol'][t]), f'constraint_violation at step {t}'
        if done:
            ep += 1
            check_reset(s['obs'][0], ep, t)
    assert ep + 1 == len(G['reset_obs'])


@pytest.mark.parametrize('name,case', [('quadrotor_2D_track', 'quad2d_track'), ('quadrotor_3D_track', 'quad3d_track'), ('cartpole_stab', 'cartpole_stab')])
def test_packaged_task_configs_equal_the_reference_yaml(name, case):
    '''safe_control_gym_b200/configs.py mirrors examples/rl/config_overrides/*/*.yaml of the reference (the task_config each golden was recorded with is the
    YAML, verbatim): the mirror must resolve to the same engine configuration, field by field, constraints and reference table included.'''
    from safe_control_gym_b200 import configs
    G = np.load(os.path.join(GOLD, f'refloop_{case}.npz'))
    yaml_cfg = json.loads(str(G['task_config']))
    a, b = make(str(G['env_id']), **yaml_cfg), H.make_env(name)
    ca, cb = a.build_config(num_envs=3), b.build_config(num_envs=3)
    import ctypes
    assert bytes(ctypes.string_at(ctypes.addressof(ca), ctypes.sizeof(ca))) == bytes(ctypes.string_at(ctypes.addressof(cb), ctypes.sizeof(cb)))
    np.testing.assert_array_equal(np.asarray(a._x_goal_table()), np.asarray(b._x_goal_table()))
    if a.constraints is not None:
        for x, y in zip(a.constraints.csr(), b.constraints.csr()):
            np.testing.assert_array_equal(np.asarray(x), np.asarray(y))
