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
