'''Pins the CPU oracle (and the host-side table generation) against artefacts shipped by the reference.'''
import numpy as np
import pytest

import helpers as H
from oracle import oracle as orc


def test_cartpole_matches_real_pybullet_transitions():
    '''5208 (state, action) -> next-state transitions produced by the reference's real PyBullet path.

    Fixture: tests/golden/cartpole_pybullet_transitions.npz (tools/make_golden.py).  float32 storage => 1e-6 abs.'''
    g = np.load(f'{H.GOLDEN}/cartpole_pybullet_transitions.npz')
    S, A, blk = g['state'].astype(np.float64), g['act'].astype(np.float64), int(g['block'])
    env = H.make_env('cartpole', ctrl_freq=int(g['ctrl_freq']), pyb_freq=int(g['pyb_freq']), cost='quadratic',
                     task='stabilization', task_info={'stabilization_goal': [0, 0], 'stabilization_goal_tolerance': 0.05},
                     randomized_init=False, done_on_out_of_bound=True)
    idx = [k for ep in range(len(S) // blk) for k in range(ep * blk, ep * blk + blk - 1)]
    o = H.make_oracle(env, num_envs=len(idx), auto_reset=False)
    for j, k in enumerate(idx):
        o.set_env_state(j, S[k])
    out = o.step(A[np.asarray(idx) + 1].reshape(-1, 1))
    err = np.abs(out['state'] - S[np.asarray(idx) + 1])
    assert err.max() < 1e-6, err.max(0)
    # the rod inertia m(2l)^2/12 written into the URDF by the reference (cartpole.py:296) does NOT reproduce the data:
    # guard that the box inertia is what is being tested (5e-5 gap in theta_dot)
    assert np.median(err[:, 3]) < 1e-8


@pytest.mark.parametrize('name,act', [('cartpole_stab', 'leaky_relu'), ('quadrotor_2D_track', 'tanh'), ('quadrotor_3D_track', 'tanh')])
def test_mlp_forward_matches_torch_on_shipped_checkpoints(name, act):
    g = np.load(f'{H.GOLDEN}/ppo_{name}.npz')
    w = lambda k: g[k.replace('.', '__')]
    pol = orc.OraclePolicy([w(f'actor.pi_net.fcs.{i}.weight') for i in range(3)], [w(f'actor.pi_net.fcs.{i}.bias') for i in range(3)],
                           [w(f'critic.v_net.fcs.{i}.weight') for i in range(3)], [w(f'critic.v_net.fcs.{i}.bias') for i in range(3)],
                           w('actor.logstd'), act)
    mean, val = pol.forward(g['x'])
    np.testing.assert_allclose(mean, g['mean'], rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(val, g['value'][:, 0], rtol=1e-5, atol=2e-5)


def test_x_goal_matches_observations_stored_in_checkpoints():
    '''The goal half of the last training observations (real PyBullet env) are rows of X_GOAL.'''
    for name, nx in (('quadrotor_2D_track', 6), ('quadrotor_3D_track', 12)):
        env = H.make_env(name)
        obs = np.load(f'{H.GOLDEN}/ppo_{name}.npz')['ckpt_obs']
        d = np.abs(env.X_GOAL[:, None, :] - obs[None, :, nx:]).max(-1)
        assert d.min(0).max() < 1e-12
    env = H.make_env('quadrotor_2D_track')
    np.testing.assert_allclose(env.X_GOAL[150], [-0.587785, -1.016641, 1.475528, 0.388322, 0, 0], atol=5e-7)
    env3 = H.make_env('quadrotor_3D_track')
    # velocity rows carry the plane offset (transformations.py:123-124): y_dot + z_dot = 0.5
    np.testing.assert_allclose(env3.X_GOAL[:, 3] + env3.X_GOAL[:, 5], 0.5, atol=1e-12)


def test_philox_known_answers():
    '''Random123 known-answer vectors for philox4x32-10.'''
    kat = [
        ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
        ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
        ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
    ]
    out = orc.philox([k[0] for k in kat], [k[1] for k in kat])
    for row, k in zip(out, kat):
        assert tuple(int(x) for x in row) == k[2]
