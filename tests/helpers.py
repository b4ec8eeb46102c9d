'''Shared helpers of the test-suite: build the same inputs for the CUDA engine and for the CPU oracle.'''
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import safe_control_gym_b200  # noqa: E402,F401  (registers envs)
from oracle import oracle as orc  # noqa: E402
from safe_control_gym_b200 import configs  # noqa: E402
from safe_control_gym_b200.utils.registration import make  # noqa: E402

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def make_env(name, **overrides):
    '''name: 'cartpole' | 'quadrotor' | a key of safe_control_gym_b200.configs.'''
    if name in ('cartpole', 'quadrotor'):
        return make(name, **overrides)
    cfg = configs.get(name)
    cfg.update(overrides)
    return make('cartpole' if name.startswith('cartpole') else 'quadrotor', **cfg)


def episode_table(env, num_envs, K, seed):
    '''[K][N][P] additive-sample table drawn with each env's own numpy Generator (seed + i), in the reference's order.'''
    rows = np.zeros((K, num_envs, 0))
    out = []
    for i in range(num_envs):
        env.seed(seed + i)
        out.append([env.draw_episode_row() for _ in range(K)])
    rows = np.asarray(out, dtype=np.float64).transpose(1, 0, 2)     # [K][N][P]
    return np.ascontiguousarray(rows)


def noise_table(env, num_envs, T, seed):
    rng_rows = []
    for i in range(num_envs):
        env.seed(10_000 + seed + i)
        rng_rows.append([env.draw_step_row() for _ in range(T)])
    tab = np.asarray(rng_rows, dtype=np.float64).transpose(1, 0, 2)
    return np.ascontiguousarray(tab)


def make_oracle(env, num_envs, auto_reset=True, rng_mode='table', seed=0, precision='f64', env_index_offset=0):
    cfg = env.build_config(num_envs=num_envs, auto_reset=auto_reset, rng_mode=rng_mode, seed=seed, precision=precision,
                           env_index_offset=env_index_offset)
    csr = env.constraints.csr() if env.constraints is not None else None
    return orc.Oracle(cfg, env._x_goal_table(), csr)


def random_policy(obs_dim, act_dim, hidden, seed=0, activation='tanh', scale=1.0):
    '''nn.Linear default init (neural_networks.py:41-45, init_weights=False), logstd = -0.5 (ppo_utils.py:166).'''
    import torch
    torch.manual_seed(seed)
    dims = [obs_dim, hidden, hidden]
    a = [torch.nn.Linear(dims[0], hidden), torch.nn.Linear(hidden, hidden), torch.nn.Linear(hidden, act_dim)]
    c = [torch.nn.Linear(dims[0], hidden), torch.nn.Linear(hidden, hidden), torch.nn.Linear(hidden, 1)]
    g = lambda t: (t.detach().numpy() * scale).astype(np.float32)
    return orc.OraclePolicy([g(l.weight) for l in a], [g(l.bias) for l in a], [g(l.weight) for l in c], [g(l.bias) for l in c],
                            -0.5 * np.ones(act_dim, np.float32), activation)


def policy_blob(pol, device):
    '''Packs an OraclePolicy into the engine's device blob through the C-ABI packer (scg_policy_pack).'''
    import ctypes as C

    import torch

    from safe_control_gym_b200 import _abi
    lib = _abi.lib()
    n = lib.scg_policy_blob_floats(pol.obs_dim, pol.act_dim, pol.hidden)
    out = np.zeros(n, np.float32)
    arr3 = lambda lst: (C.c_void_p * 3)(*[a.ctypes.data for a in lst])
    aw, ab, cw, cb = arr3(pol.aw), arr3(pol.ab), arr3(pol.cw), arr3(pol.cb)
    rc = lib.scg_policy_pack(pol.obs_dim, pol.act_dim, pol.hidden, aw, ab, pol.logstd.ctypes.data_as(C.c_void_p), cw, cb,
                             out.ctypes.data_as(C.c_void_p))
    assert rc == 0
    return torch.from_numpy(out).to(device), (pol.obs_dim, pol.act_dim, pol.hidden)


# ------------------------------------------------------------------------------------------------------------------
# strict rollout comparison with FORK PROOFS
def _termination_margin(cfg, x_goal, state, cvals_row):
    '''Distance of one env.state (oracle, float64) to the nearest termination threshold of `_get_done`
    (quadrotor.py:864-894, cartpole.py:654-672) / done_on_violation (benchmark_env.py:476-481).'''
    from safe_control_gym_b200 import _abi
    m = []
    if cfg.task == _abi.SCG_TASK_STABILIZATION:
        nx = len(state)
        m.append(abs(float(np.linalg.norm(state - np.asarray(x_goal).reshape(-1)[:nx])) - cfg.goal_tolerance))
    if cfg.done_on_out_of_bound:
        if cfg.system == _abi.SCG_SYS_CARTPOLE:
            m += [abs(abs(state[0]) - cfg.x_threshold), abs(abs(state[2]) - cfg.theta_threshold)]
        else:
            mask = {_abi.SCG_SYS_QUAD1D: [1, 0], _abi.SCG_SYS_QUAD2D: [1, 0, 1, 0, 1, 0],
                    _abi.SCG_SYS_QUAD3D: [1, 0, 1, 0, 1, 0, 1, 1, 1, 0, 0, 0]}[cfg.system]
            for k, on in enumerate(mask):
                if on:
                    m += [abs(state[k] - cfg.state_low[k]), abs(state[k] - cfg.state_high[k])]
    if cfg.done_on_violation and cvals_row is not None:
        m.append(float(np.abs(cvals_row).min()))
    return min(m) if m else np.inf


def compare_rollouts(o, g, cfg, x_goal, N, T, obs_tol, max_fork_frac=0.02, keys_extra=(), early=None, horizon_check=True):
    '''o: oracle record (numpy, produced with trace=True), g: engine record (torch or numpy).  Every key is compared per env up
    to the first `mask` mismatch.  A mismatch ("fork") is only accepted with a PROOF that it is a float tie at a termination
    threshold: at the fork step the oracle's env.state must lie within the already observed oracle-vs-kernel trajectory
    deviation of that env (x10, + float32 record quantisation) of the threshold that fired; anything else fails the test.
    ret / adv / episode statistics are compared on the un-forked envs.  `early` = (steps, tol): a tighter bound on the first steps of the
    horizon -- the float32 policy's rounding differences are amplified by the closed loop over the horizon, so the full-horizon bound
    (the north-star rtol 1e-4) alone would not tell amplified rounding from a small systematic error.  Returns the proven forks and
    the largest relative deviation per key.'''
    tonp = lambda t: t.detach().cpu().numpy() if hasattr(t, 'detach') else np.asarray(t)
    gm, om = tonp(g['mask'])[..., 0], o['mask'][..., 0]
    mism = gm != om
    first_bad = np.where(mism.any(0), mism.argmax(0), T)            # [N]
    forked = np.nonzero(first_bad < T)[0]
    nx = o['state_trace'].shape[-1]
    go, oo = tonp(g['obs']), o['obs']
    proofs = []
    for i in forked:
        ts = int(first_bad[i])
        dev = float(np.abs(go[:ts + 1, i, :nx] - oo[:ts + 1, i, :nx]).max())
        tol = min(10.0 * dev + 2e-6, 1e-3)
        cv = o['cvals'][ts, i] if 'cvals' in o else None
        margin = _termination_margin(cfg, x_goal, o['state_trace'][ts, i], cv)
        assert margin <= tol, (f'env {i} forked at step {ts} (oracle done={1 - om[ts, i]:.0f}, kernel done={1 - gm[ts, i]:.0f}) but the oracle state is '
                               f'{margin:.3e} away from the nearest termination threshold (trajectory deviation so far {dev:.3e}): not a float tie')
        proofs.append((int(i), ts, margin, dev))
    assert len(forked) <= max(1, int(max_fork_frac * N)), f'{len(forked)} of {N} envs forked (all proven ties, but too many): {proofs[:8]}'
    valid = np.arange(T)[:, None] < first_bad[None, :]                # [T][N] steps strictly before the fork
    tols = [('obs', obs_tol), ('act', obs_tol), ('rew', obs_tol), ('v', 5e-5), ('logp', 2e-4), ('terminal_v', 5e-5), ('mse', obs_tol)]
    tols += [(k, obs_tol) for k in keys_extra]
    worst = {}
    for k, tol in tols:
        if k not in g or k not in o:
            continue
        a, b = tonp(g[k]), o[k]
        d = np.abs(a - b) / (1.0 + np.abs(b))
        d = d.reshape(T, N, -1).max(-1)
        worst[k] = float(d[valid].max()) if valid.any() else 0.0
        if horizon_check:
            # over the whole horizon the closed loop amplifies the float32 policy's rounding differences: all but a handful of the T x N
            # samples must be inside `tol`, and even the worst chaotic env inside 10 x tol
            n_out = int((d[valid] > tol).sum())
            worst[k + '_over_tol'] = n_out
            assert n_out <= max(2, int(2e-5 * valid.sum())) and worst[k] <= 10 * tol, (k, worst[k], n_out, np.unravel_index(np.argmax(np.where(valid, d, 0)), d.shape))
        if early is not None and k in ('obs', 'act', 'rew', 'mse') + tuple(keys_extra):
            ve = valid[:early[0]]
            worst[k + '_early'] = float(d[:early[0]][ve].max())
            assert worst[k + '_early'] <= early[1], (k, 'first steps', worst[k + '_early'])
    # constraint_violation is a threshold function of the state too: a mismatch needs the same kind of proof -- the oracle's
    # constraint value nearest to zero must be within the trajectory deviation (values are rounded to 8 decimals, constraints.py:109)
    cv_tie = np.zeros(N, bool)
    if 'cviol' in g and 'cviol' in o:
        cm = (tonp(g['cviol'])[..., 0] != o['cviol'][..., 0]) & valid
        for t, i in zip(*np.nonzero(cm)):
            dev = float(np.abs(go[:t + 2, i, :nx] - oo[:t + 2, i, :nx]).max())
            margin = float(np.abs(o['cvals'][t, i]).min()) if 'cvals' in o else np.inf
            assert margin <= min(10.0 * dev + 2e-6, 1e-3), f'cviol of env {i} differs at step {t} but the nearest constraint value is {margin:.3e} (deviation {dev:.3e})'
            cv_tie[i] = True
        assert cv_tie.sum() <= max(1, int(max_fork_frac * N)), f'{cv_tie.sum()} envs with constraint-violation ties'
        worst['cviol_ties'] = int(cv_tie.sum())
    full = first_bad == T
    if not horizon_check:
        return proofs, worst
    for k in ('ret', 'adv'):
        if k in g and k in o:
            np.testing.assert_allclose(tonp(g[k])[:, full], o[k][:, full], rtol=2e-4, atol=2e-4, err_msg=k)
    if 'ep_count' in g:
        np.testing.assert_array_equal(tonp(g['ep_count'])[full], o['ep_count'][full])
        np.testing.assert_allclose(tonp(g['ep_return_sum'])[full], o['ep_return_sum'][full], rtol=1e-4, atol=1e-4)
        np.testing.assert_allclose(tonp(g['ep_length_sum'])[full], o['ep_length_sum'][full], rtol=0, atol=0)
        np.testing.assert_allclose(tonp(g['ep_cviol_sum'])[full & ~cv_tie], o['ep_cviol_sum'][full & ~cv_tie], rtol=0, atol=0)
    if 'last_obs' in g:
        np.testing.assert_allclose(tonp(g['last_obs'])[full], o['last_obs'][full], rtol=obs_tol * 10, atol=obs_tol * 10)
    return proofs, worst


def compare_replay(oracle, policy, g, N, T, tol=2e-5, gamma=0.99, gae_lambda=0.95, use_gae=True, max_tie_frac=0.002, act_tol=None, self_check=None, **oracle_kw):
    '''TEACHER-FORCED parity: `oracle` (a fresh, reset Oracle of the same configuration) replays the engine's record `g` -- before every
    step its env state is overwritten with the record's observation and it is stepped with the record's action -- and every output
    of every one of the T x N transitions is compared with the oracle's one-step answer: sampled action (mean + sd * Philox normal),
    value, log-probability, next observation, reward, mse, constraint values and flags, done flag, and the GAE columns.  Unlike a
    trajectory comparison this bound does not grow with the horizon (no closed-loop amplification of the float32 policy's rounding
    differences), so it is tight at the full benchmark size.  done / constraint-violation mismatches need a tie proof.
    act_tol: optional per-sample bound [T][N][1] on the sampled action (default 3e-5 everywhere); used where the action mean goes through an
    ill-conditioned map whose condition number the caller knows (SafeExplorer projection: 1 / |g|).'''
    tonp = lambda t: t.detach().cpu().numpy() if hasattr(t, 'detach') else np.asarray(t)
    rec = {k: tonp(g[k]) for k in ('obs', 'act', 'mask')}
    p = oracle.rollout(policy, T, eps=None, gamma=gamma, gae_lambda=gae_lambda, use_gae=use_gae, replay=rec, **oracle_kw)
    cfg, nx = oracle.cfg, oracle.nx
    worst = {}

    def rel(a, b):
        return np.abs(a - b) / (1.0 + np.abs(b))
    # The replay rebuilds the physical state from the OBSERVED state.  For the quadrotor that map is only injective inside the Euler chart:
    # a tumbling drone whose pitch passes pi/2 BETWEEN two control steps (15 degrees per step at 13 rad/s: it can jump over the 85 degree
    # termination band) comes back with the angle folded (asin) and the rest of the rotation in roll / yaw = pi, which the observation does
    # not carry.  Such transitions cannot be teacher-forced -- the oracle replaying ITS OWN record disagrees with itself there -- so, when the
    # caller passes `self_check = (make_oracle, oracle_record)`, they are found exactly that way and excluded (the trajectory comparison
    # covers them); they must stay a small fraction.
    chart = np.ones((T, N), bool)
    if self_check is not None:
        mk_, own = self_check
        so = mk_(); so.reset()
        ps = so.rollout(policy, T, eps=None, gamma=gamma, gae_lambda=gae_lambda, use_gae=use_gae, replay={k: own[k] for k in ('obs', 'act', 'mask')}, **oracle_kw)
        om = own['mask'][..., 0]
        bad = np.zeros((T, N), bool)
        bad[:-1] = (rel(own['obs'][1:], ps['next_obs'][:-1]).max(-1) > tol) & (om[:-1] == 1)
        # a folded attitude stays folded until the episode ends
        for i in np.nonzero(bad.any(0))[0]:
            t = 0
            while t < T:
                if bad[t, i]:
                    e = t
                    while e < T and om[e, i] == 1:
                        e += 1
                    bad[t:min(e + 1, T), i] = True
                    t = e + 1
                else:
                    t += 1
        chart = ~bad
        worst['outside_euler_chart'] = int(bad.sum())
        assert bad.mean() <= 5e-3, f'{bad.sum()} transitions the oracle cannot teacher-force on its own record'
    for k, t_ in (('act', 3e-5), ('v', 1e-5), ('logp', 1e-4), ('rew', tol), ('mse', tol), ('cvals', tol), ('c', tol)):
        if k in g and k in p:
            d = rel(tonp(g[k]), p[k])
            if k in ('rew', 'mse', 'cvals'):
                d = np.where(chart[..., None], d, 0.0)
            if k == 'act' and act_tol is not None:
                d = d / (np.asarray(act_tol) / t_)          # per-sample bound, reported in units of the default one
            worst[k] = float(d.max())
            assert worst[k] <= t_, (k, worst[k], np.unravel_index(np.argmax(d), d.shape))
    gm, pm = rec['mask'][..., 0], p['mask'][..., 0]
    # next observation: record obs[t + 1] against the prediction, wherever the record did not reset
    cont = (gm[:-1] == 1) & chart[:-1]
    d = rel(rec['obs'][1:], p['next_obs'][:-1]).max(-1)
    worst['next_obs'] = float(d[cont].max())
    assert worst['next_obs'] <= tol, ('next_obs', worst['next_obs'], np.unravel_index(np.argmax(np.where(cont, d, 0)), d.shape))
    # done flags: a mismatch is only accepted as a tie at a termination threshold (the predicted state sits on it)
    ties = 0
    for t, i in zip(*np.nonzero((gm != pm) & chart)):
        cv = p['cvals'][t, i] if 'cvals' in p else None
        margin = _termination_margin(cfg, oracle.x_goal, p['next_obs'][t, i, :nx].astype(np.float64), cv)
        assert margin <= 20 * tol, f'done flag of env {i} differs at step {t}; the predicted state is {margin:.3e} away from the nearest threshold'
        ties += 1
    if 'cviol' in g:
        for t, i in zip(*np.nonzero((tonp(g['cviol'])[..., 0] != p['cviol'][..., 0]) & chart)):
            margin = float(np.abs(p['cvals'][t, i]).min())
            assert margin <= 20 * tol, f'cviol of env {i} differs at step {t}; nearest constraint value {margin:.3e}'
            ties += 1
    assert ties <= max(2, int(max_tie_frac * N * T / 50)), f'{ties} threshold ties'
    worst['ties'] = ties
    # terminal values (critic on the terminal observation of truncated envs) where both sides agree on done
    same = (gm == pm) & chart
    d = rel(tonp(g['terminal_v'])[..., 0], p['terminal_v'][..., 0])
    worst['terminal_v'] = float(d[same].max())
    assert worst['terminal_v'] <= 5e-5, ('terminal_v', worst['terminal_v'])
    if 'ret' in g and ties == 0 and chart.all():
        for k in ('ret', 'adv'):
            d = rel(tonp(g[k]), p[k])
            worst[k] = float(d.max())
            assert worst[k] <= 2e-4, (k, worst[k])
    return worst
