'''ctypes binding of the CPU oracle (oracle/scg_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import this
module.  It takes the same `scg_config` POD (include/scg.h) and the same tables as the CUDA engine, so a
parity test feeds identical inputs to both and compares outputs.
'''

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'libscg_oracle.so')


def build(force=False):
    '''Compiles the C restatement with the Makefile next to it (gcc, seconds).'''
    src = os.path.join(HERE, 'scg_oracle.c')
    hdr = os.path.join(HERE, '..', 'include', 'scg.h')
    if (not force and os.path.exists(LIB_PATH) and os.path.getmtime(LIB_PATH) >= max(os.path.getmtime(src), os.path.getmtime(hdr))):
        return LIB_PATH
    subprocess.run(['make', '-C', HERE, '-B', 'libscg_oracle.so'], check=True, capture_output=True)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        _lib = C.CDLL(LIB_PATH)
        _lib.scg_oracle_create.restype = C.c_void_p
        _lib.scg_oracle_create.argtypes = [C.c_void_p]
        _lib.scg_oracle_destroy.argtypes = [C.c_void_p]
        _lib.scg_oracle_rollout.restype = C.c_int
        _lib.scg_oracle_step.restype = C.c_int
        _lib.scg_oracle_num_threads.restype = C.c_int
        _lib.scg_oracle_adversary_dim.restype = C.c_int
        _lib.scg_oracle_adversary_dim.argtypes = [C.c_void_p]
        _lib.scg_oracle_set_adversary_control.restype = C.c_int
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class _StepOut(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ('obs', 'state', 'rew', 'done', 'truncated', 'cviol', 'cvals', 'mse', 'out_of_bounds',
                                         'goal_reached', 'current_step', 'terminal_obs', 'terminal_state')]


class _Policy(C.Structure):
    _fields_ = [('obs_dim', C.c_int), ('act_dim', C.c_int), ('hidden', C.c_int), ('activation', C.c_int),
                ('aw', C.c_void_p * 3), ('ab', C.c_void_p * 3), ('cw', C.c_void_p * 3), ('cb', C.c_void_p * 3), ('logstd', C.c_void_p)]


class _RolloutOut(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ('obs', 'act', 'rew', 'mask', 'v', 'logp', 'terminal_v', 'ret', 'adv', 'cviol', 'cvals', 'mse',
                                         'last_obs', 'last_val', 'ep_return_sum', 'ep_length_sum', 'ep_cviol_sum', 'ep_count',
                                         'state_trace', 'replay_obs', 'replay_act', 'replay_mask', 'next_obs', 'norm', 'obs0', 'obs_raw', 'rew_raw')] + \
        [('safety_C', C.c_int), ('safety_H', C.c_int), ('safety', C.c_void_p), ('safety_slack', C.c_void_p), ('c', C.c_void_p),
         ('adv_policy', C.c_void_p), ('eps_adv', C.c_void_p), ('adv_record', C.c_int)]


class OraclePolicy:
    '''Actor-critic weights as float32 numpy arrays in nn.Linear layout ([out][in]).'''

    ACT = {'tanh': 0, 'relu': 1, 'leaky_relu': 2}

    def __init__(self, aw, ab, cw, cb, logstd, activation='tanh'):
        self.aw = [np.ascontiguousarray(w, dtype=np.float32) for w in aw]
        self.ab = [np.ascontiguousarray(b, dtype=np.float32) for b in ab]
        self.cw = [np.ascontiguousarray(w, dtype=np.float32) for w in cw]
        self.cb = [np.ascontiguousarray(b, dtype=np.float32) for b in cb]
        self.logstd = np.ascontiguousarray(logstd, dtype=np.float32)
        self.activation = activation
        self.obs_dim, self.hidden, self.act_dim = self.aw[0].shape[1], self.aw[0].shape[0], self.aw[2].shape[0]

    @classmethod
    def from_state_dict(cls, sd, activation='tanh'):
        '''sd: the `agent.ac` state dict of a reference PPO checkpoint (ppo_utils.py:66-70).'''
        g = lambda k: sd[k].detach().cpu().numpy()
        return cls([g(f'actor.pi_net.fcs.{i}.weight') for i in range(3)], [g(f'actor.pi_net.fcs.{i}.bias') for i in range(3)],
                   [g(f'critic.v_net.fcs.{i}.weight') for i in range(3)], [g(f'critic.v_net.fcs.{i}.bias') for i in range(3)],
                   g('actor.logstd'), activation)

    def c_struct(self):
        s = _Policy()
        s.obs_dim, s.act_dim, s.hidden, s.activation = self.obs_dim, self.act_dim, self.hidden, self.ACT[self.activation]
        for i in range(3):
            s.aw[i], s.ab[i] = self.aw[i].ctypes.data, self.ab[i].ctypes.data
            s.cw[i], s.cb[i] = self.cw[i].ctypes.data, self.cb[i].ctypes.data
        s.logstd = self.logstd.ctypes.data
        return s

    def forward(self, obs):
        obs = np.ascontiguousarray(obs, dtype=np.float32).reshape(-1, self.obs_dim)
        n = obs.shape[0]
        mean = np.empty((n, self.act_dim), np.float32)
        val = np.empty((n,), np.float32)
        s = self.c_struct()
        lib().scg_oracle_policy_forward(C.byref(s), _p(obs), C.c_int(n), _p(mean), _p(val))
        return mean, val


class Oracle:
    def __init__(self, cfg, x_goal, constraints_csr=None):
        '''cfg: a ctypes scg_config (include/scg.h); x_goal: [L][nx] float64; constraints_csr: the 6 CSR arrays or None.'''
        self.lib = lib()
        self.cfg = cfg
        self.h = C.c_void_p(self.lib.scg_oracle_create(C.byref(cfg)))
        if not self.h:
            raise ValueError('scg_oracle_create failed (config size mismatch?)')
        self.N = cfg.num_envs
        xg = np.ascontiguousarray(np.atleast_2d(x_goal), dtype=np.float64)
        self.x_goal = xg
        self.lib.scg_oracle_set_reference(self.h, _p(xg), C.c_int(xg.shape[0]))
        if constraints_csr is not None:
            kind, strict, ptr, col, val, b = constraints_csr
            self.lib.scg_oracle_set_constraints(self.h, C.c_int(len(kind)), _p(kind), _p(strict), _p(ptr), _p(col), _p(val), _p(b))
        d = [C.c_int() for _ in range(6)]
        self.lib.scg_oracle_dims(self.h, *[C.byref(x) for x in d])
        self.nx, self.nu, self.obs_dim, self.n_con, self.P, self.D = [x.value for x in d]

    def close(self):
        if self.h:
            self.lib.scg_oracle_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_episode_table(self, tab):
        tab = np.ascontiguousarray(tab, dtype=np.float64)
        assert tab.shape[1:] == (self.N, self.P), (tab.shape, self.N, self.P)
        self.lib.scg_oracle_set_episode_table(self.h, _p(tab), C.c_int(tab.shape[0]))

    def set_noise_table(self, tab):
        tab = np.ascontiguousarray(tab, dtype=np.float64)
        assert tab.shape[1:] == (self.N, self.D), (tab.shape, self.N, self.D)
        self.lib.scg_oracle_set_noise_table(self.h, _p(tab), C.c_int(tab.shape[0]))

    def get_state(self):
        '''Snapshot of all envs (+ the noise cursor), like EngineHandle.get_state.'''
        self.lib.scg_oracle_state_bytes.restype = C.c_size_t
        buf = np.empty(self.lib.scg_oracle_state_bytes(self.h), np.uint8)
        self.lib.scg_oracle_get_state(self.h, _p(buf))
        return buf

    def set_state(self, buf):
        buf = np.ascontiguousarray(buf, dtype=np.uint8)
        self.lib.scg_oracle_set_state(self.h, _p(buf))

    def set_stale_force_point(self, on=True):
        '''Reference quirk (off by default, absent from the CUDA path): see scg_oracle_set_stale_force_point in scg_oracle.c.'''
        self.lib.scg_oracle_set_stale_force_point(self.h, C.c_int(int(bool(on))))

    def set_env_state(self, i, state, inertial=None):
        s = np.ascontiguousarray(state, dtype=np.float64)
        inr = None if inertial is None else np.ascontiguousarray(inertial, dtype=np.float64)
        self.lib.scg_oracle_set_env_state(self.h, C.c_int(i), _p(s), _p(inr))

    def reset(self, mask=None):
        obs = np.empty((self.N, self.obs_dim), np.float64)
        st = np.empty((self.N, self.nx), np.float64)
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        self.lib.scg_oracle_reset(self.h, _p(m), _p(obs), _p(st))
        return {'obs': obs, 'state': st}

    @property
    def adversary_dim(self):
        return int(self.lib.scg_oracle_adversary_dim(self.h))

    def set_adversary_control(self, adv):
        '''BenchmarkEnv.set_adversary_control for all envs (raw adversary actions [N][adversary_dim]).'''
        adv = np.ascontiguousarray(adv, dtype=np.float64).reshape(self.N, -1)
        rc = self.lib.scg_oracle_set_adversary_control(self.h, _p(adv))
        if rc == -3:
            raise RuntimeError('[ERROR] adversary_disturbance does not exist, env.set_adversary_control() cannot be called.')

    def step(self, act):
        act = np.ascontiguousarray(act, dtype=np.float64).reshape(self.N, self.nu)
        N, od, nx, nc = self.N, self.obs_dim, self.nx, max(self.n_con, 1)
        o = {
            'obs': np.empty((N, od)), 'state': np.empty((N, nx)), 'rew': np.empty(N), 'done': np.empty(N, np.uint8),
            'truncated': np.empty(N, np.uint8), 'cviol': np.empty(N), 'cvals': np.zeros((N, nc)), 'mse': np.empty(N),
            'out_of_bounds': np.empty(N, np.uint8), 'goal_reached': np.empty(N, np.uint8), 'current_step': np.empty(N, np.int32),
            'terminal_obs': np.empty((N, od)), 'terminal_state': np.empty((N, nx)),
        }
        so = _StepOut()
        for k, v in o.items():
            setattr(so, k, v.ctypes.data)
        rc = self.lib.scg_oracle_step(self.h, _p(act), C.byref(so))
        if rc == -3:
            raise RuntimeError('[ERROR] You must call env.reset() at least once before using env.step().')
        return o

    def rollout(self, policy, T, eps=None, deterministic=False, gamma=0.99, gae_lambda=0.95, use_gae=False, trace=False, replay=None, norm=None, obs0=None, safety=None,
                adversary=None, eps_adv=None, adv_record=False):
        '''replay = {'obs', 'act', 'mask'} (float32 [T][N][.] arrays of another collector's record): teacher-forced one-step predictions,
        see oracle_rollout_out in scg_oracle.c.'''
        N, od, A, nc = self.N, self.obs_dim, self.nu, self.n_con
        f = np.float32
        Arec = adversary.act_dim if (adversary is not None and adv_record) else A      # rarl.py:352: the adversary's PPOBuffer
        o = {
            'obs': np.empty((T, N, od), f), 'act': np.empty((T, N, Arec), f), 'rew': np.empty((T, N, 1), f), 'mask': np.empty((T, N, 1), f),
            'v': np.empty((T, N, 1), f), 'logp': np.empty((T, N, 1), f), 'terminal_v': np.empty((T, N, 1), f),
            'ret': np.empty((T, N, 1), f), 'adv': np.empty((T, N, 1), f), 'cviol': np.empty((T, N, 1), f), 'mse': np.empty((T, N, 1), f),
            'last_obs': np.empty((N, od), f), 'last_val': np.empty((N, 1), f),
            'ep_return_sum': np.empty(N, f), 'ep_length_sum': np.empty(N, f), 'ep_cviol_sum': np.empty(N, f), 'ep_count': np.empty(N, np.int32),
        }
        if nc > 0:
            o['cvals'] = np.empty((T, N, nc), f)
        if trace:
            o['state_trace'] = np.empty((T, N, self.nx), np.float64)
        keep = []
        if safety is not None:          # {'blob': float32 [C * per], 'C', 'H', 'slack': [C]}
            o['c'] = np.empty((T, N, safety['C']), f)
        if norm is not None:
            o['obs_raw'] = np.empty((T, N, od), f)
            o['rew_raw'] = np.empty((T, N, 1), f)
        if replay is not None:
            o['next_obs'] = np.empty((T, N, od), f)
            for k in ('obs', 'act', 'mask'):
                keep.append(np.ascontiguousarray(replay[k], dtype=np.float32).reshape(T, N, -1))
        ro = _RolloutOut()
        for k, v in o.items():
            setattr(ro, k, v.ctypes.data)
        if replay is not None:
            ro.replay_obs, ro.replay_act, ro.replay_mask = keep[0].ctypes.data, keep[1].ctypes.data, keep[2].ctypes.data
        if norm is not None:
            keep.append(np.ascontiguousarray(norm, dtype=np.float64)); ro.norm = keep[-1].ctypes.data
        if obs0 is not None:
            keep.append(np.ascontiguousarray(obs0, dtype=np.float32)); ro.obs0 = keep[-1].ctypes.data
        if safety is not None:
            keep.append(np.ascontiguousarray(safety['blob'], dtype=np.float32)); ro.safety = keep[-1].ctypes.data
            sl = np.zeros(16); sl[:safety['C']] = safety['slack']
            keep.append(sl); ro.safety_slack = keep[-1].ctypes.data
            ro.safety_C, ro.safety_H = int(safety['C']), int(safety['H'])
        e = None if eps is None else np.ascontiguousarray(eps, dtype=np.float32).reshape(T, N, A)
        ps = policy.c_struct()
        if adversary is not None:       # an OraclePolicy with act_dim = adversary_dim (RARL / RAP, rarl.py:349-418)
            assert adversary.act_dim == self.adversary_dim and adversary.obs_dim == od
            qs = adversary.c_struct(); keep.append(qs)
            ro.adv_policy = C.addressof(qs); ro.adv_record = int(bool(adv_record))
            if eps_adv is not None:
                keep.append(np.ascontiguousarray(eps_adv, dtype=np.float32).reshape(T, N, adversary.act_dim)); ro.eps_adv = keep[-1].ctypes.data
        rc = self.lib.scg_oracle_rollout(self.h, C.byref(ps), _p(e), C.c_int(T), C.c_int(int(deterministic)), C.c_double(gamma),
                                         C.c_double(gae_lambda), C.c_int(int(use_gae)), C.byref(ro))
        if rc != 0:
            raise RuntimeError(f'scg_oracle_rollout failed with {rc}')
        return o


def philox(ctr, key):
    ctr = np.ascontiguousarray(ctr, dtype=np.uint32).reshape(-1, 4)
    key = np.ascontiguousarray(key, dtype=np.uint32).reshape(-1, 2)
    out = np.empty_like(ctr)
    lib().scg_oracle_philox(_p(ctr), _p(key), _p(out), C.c_int(ctr.shape[0]))
    return out


def num_threads():
    return lib().scg_oracle_num_threads()


def set_threads(n):
    lib().scg_oracle_set_threads(C.c_int(n))
