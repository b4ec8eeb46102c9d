'''Replay-buffer rows of the UNMODIFIED reference SAC controller (the DDPG collector is the same code) against the oracle's step API:
`SAC.train_step()` of /root/reference (controllers/sac/sac.py:269-333) on a DummyVecEnv of unmodified reference envs during its warm-up, recorded by
tools/make_golden_offpolicy_loop.py (tests/golden/offpolicy_loop_*.npz).

The oracle is stepped with the recorded actions (env i seeded `seed + i`); every row the reference pushed must follow from its step outputs by the rule
`controllers/sac/sac.py` of this repo implements on batched device tensors (`OffPolicyController.train_step`: `truncated & done` selects the terminal
observation and keeps the mask at 1; a true termination stores the reset observation the vec env returned and mask 0):

    next_obs = terminal_obs if truncated else obs_returned_by_step        mask = 1 if truncated else 1 - done

tests/test_gpu_extended.py checks the controller's device buffer against the same rule evaluated on the oracle.'''
import json
import os

import numpy as np
import pytest

import helpers as H
from safe_control_gym_b200.utils.registration import make

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


# Tolerance of the quadrotor case: the controller hands the env FLOAT32 actions (gymnasium samples / the policy network), and with a float32 input the
# reference's whole action map -- (1 + 0.1 a) * hover thrust, clip, cmd2pwm, pwm2rpm -- stays float32 (numpy keeps the array's dtype against Python
# floats): its motor speeds carry ~7 digits.  Oracle and kernels promote the float32 action to float64 first.  The difference (1e-7 relative on the
# thrusts, a few 1e-6 on the body rate after some steps) is far inside the north-star tolerance; CartPole's map is a single product and shows nothing.
@pytest.mark.parametrize('case,tol', [('quad2d_track', 2e-5), ('cartpole_stab', 2e-6)])
def test_replay_rows_of_the_reference_collector_follow_from_the_oracle_step(case, tol):
    G = np.load(os.path.join(GOLD, f'offpolicy_loop_{case}.npz'))
    cfg = json.loads(str(G['task_config']))
    env = make(str(G['env_id']), **cfg)
    M, N = G['rew'].shape[:2]
    K = int((G['mask'] == 0).sum(0).max()) + M // int(env.CTRL_STEPS) + 3
    o = H.make_oracle(env, N, auto_reset=True, rng_mode='table')
    o.set_episode_table(H.episode_table(env, N, K=K, seed=int(G['seed'])))
    obs = o.reset()['obs']
    np.testing.assert_allclose(obs, G['obs0'], rtol=1e-9, atol=1e-9)
    n_trunc = n_term = 0
    for t in range(M):
        np.testing.assert_allclose(G['obs'][t], obs, rtol=tol, atol=tol, err_msg=f'obs row {t}')           # (the buffer is float32)
        s = o.step(G['act'][t].astype(np.float64))
        done, trunc = s['done'].astype(bool), s['truncated'].astype(bool) & s['done'].astype(bool)
        want_next = np.where(trunc[:, None], s['terminal_obs'], s['obs'])
        want_mask = np.where(trunc, 1.0, 1.0 - done)
        np.testing.assert_allclose(G['next_obs'][t], want_next, rtol=tol, atol=tol, err_msg=f'next_obs row {t}')
        np.testing.assert_array_equal(G['mask'][t][:, 0], want_mask, err_msg=f'mask row {t}')
        np.testing.assert_allclose(G['rew'][t][:, 0], s['rew'], rtol=tol, atol=tol, err_msg=f'rew row {t}')
        n_trunc += int(trunc.sum()); n_term += int((done & ~trunc).sum())
        obs = s['obs']
    assert n_trunc > 0 and n_term > 0, (n_trunc, n_term)                   # both kinds of episode end occurred
