'''EngineHandle: thin Python owner of one `scg_env` (a batch of N envs on one GPU).

torch is used for device memory and streams only (plumbing): every output buffer is a torch.cuda tensor
allocated here and handed to the C-ABI as a raw pointer; results are returned as those tensors.
'''

import ctypes as C
import os

import numpy as np
import torch

from safe_control_gym_b200 import _abi
from safe_control_gym_b200.envs import disturbances as dist_mod


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _np_ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def policy_blob_layout(obs_dim, act_dim, hid):
    '''Float offsets of the packed policy blob (csrc/scg_kernels.cuh: blob_layout).'''
    def pad4(x):
        return (x + 3) & ~3
    w1t = 0
    b1 = w1t + obs_dim * 2 * hid
    w2t = b1 + 2 * hid
    b2 = w2t + 2 * hid * hid
    w3 = b2 + 2 * hid
    b3 = w3 + (act_dim + 1) * hid
    logstd = b3 + pad4(act_dim + 1)
    total = logstd + pad4(act_dim)
    return dict(w1t=w1t, b1=b1, w2t=w2t, b2=b2, w3=w3, b3=b3, logstd=logstd, total=total)


def pack_policy(ac, device):
    '''MLPActorCritic (ppo_utils.py:197-238 layout: actor.pi_net.fcs.{0,1,2}, actor.logstd, critic.v_net.fcs.{0,1,2})
    -> packed device blob, built with torch ops on the GPU (no host round trip).'''
    a, c = ac.actor.pi_net.fcs, ac.critic.v_net.fcs
    if len(a) != 3 or len(c) != 3:
        raise NotImplementedError('the fused collector supports two hidden layers (hidden_dims=[h, h])')
    obs_dim, hid, act_dim = a[0].in_features, a[0].out_features, a[2].out_features
    if obs_dim % 2 or hid % 4:
        raise NotImplementedError('the fused collector needs an even observation width and hidden_dim % 4 == 0')
    L = policy_blob_layout(obs_dim, act_dim, hid)
    blob = torch.zeros(L['total'], dtype=torch.float32, device=device)
    with torch.no_grad():
        f = lambda t: t.detach().to(device=device, dtype=torch.float32)
        # pair-major layouts (csrc/scg_kernels.cuh): W1P [obs/2][2H][2], W2P [net][H/2][H][2]
        w1 = torch.cat([f(a[0].weight).t(), f(c[0].weight).t()], dim=1)                                          # [obs][2H]
        blob[L['w1t']:L['b1']] = w1.reshape(obs_dim // 2, 2, 2 * hid).permute(0, 2, 1).reshape(-1)
        blob[L['b1']:L['w2t']] = torch.cat([f(a[0].bias), f(c[0].bias)])
        w2 = torch.stack([f(a[1].weight).t(), f(c[1].weight).t()])                                               # [2][k][u]
        blob[L['w2t']:L['b2']] = w2.reshape(2, hid // 2, 2, hid).permute(0, 1, 3, 2).reshape(-1)
        blob[L['b2']:L['w3']] = torch.cat([f(a[1].bias), f(c[1].bias)])
        blob[L['w3']:L['b3']] = torch.cat([f(a[2].weight), f(c[2].weight)], dim=0).reshape(-1)                 # [A+1][H]
        blob[L['b3']:L['b3'] + act_dim + 1] = torch.cat([f(a[2].bias), f(c[2].bias)])
        blob[L['logstd']:L['logstd'] + act_dim] = f(ac.actor.logstd)
    return blob, (obs_dim, act_dim, hid)


class EngineHandle:
    def __init__(self, env, num_envs=1, auto_reset=False, rng_mode='table', seed=0, env_index_offset=0, device=None, precision=None):
        self.lib = _abi.lib()
        self.env = env
        self.N = int(num_envs)
        self.cfg = env.build_config(num_envs=num_envs, auto_reset=auto_reset, env_index_offset=env_index_offset,
                                    rng_mode=rng_mode, seed=seed, device=device, precision=precision)
        self.rng_mode = rng_mode
        self.device_index = int(self.cfg.device)
        self.device = torch.device('cuda', self.device_index)
        h = C.c_void_p()
        _abi.check(self.lib.scg_create(C.byref(self.cfg), C.byref(h)), None)
        self.h = h
        dims = [C.c_int32() for _ in range(6)]
        _abi.check(self.lib.scg_dims(self.h, *[C.byref(d) for d in dims]), self.h)
        self.nx, self.nu, self.obs_dim, _, self.P, self.D = [d.value for d in dims]
        self.adv_dim = int(self.lib.scg_adversary_dim(self.h))
        xg = env._x_goal_table()
        _abi.check(self.lib.scg_set_reference(self.h, _np_ptr(xg), xg.shape[0]), self.h)
        self.n_con = 0
        if env.constraints is not None:
            kind, strict, ptr, col, val, b = env.constraints.csr()
            self.n_con = len(kind)
            _abi.check(self.lib.scg_set_constraints(self.h, self.n_con, _np_ptr(kind), _np_ptr(strict), _np_ptr(ptr), _np_ptr(col),
                                                    _np_ptr(val), _np_ptr(b)), self.h)
        self._step_bufs = None

    # -- tables ---------------------------------------------------------------------------------
    def set_episode_table(self, tab, restart=False):
        tab = np.ascontiguousarray(tab, dtype=np.float64)
        assert tab.ndim == 3 and tab.shape[1] == self.N and tab.shape[2] == self.P, (tab.shape, self.N, self.P)
        _abi.check(self.lib.scg_set_episode_table(self.h, _np_ptr(tab), tab.shape[0]), self.h)

    def set_noise_table(self, tab):
        tab = np.ascontiguousarray(tab, dtype=np.float64)
        assert tab.ndim == 3 and tab.shape[1] == self.N and tab.shape[2] == self.D, (tab.shape, self.N, self.D)
        _abi.check(self.lib.scg_set_noise_table(self.h, _np_ptr(tab), tab.shape[0]), self.h)

    # -- env API --------------------------------------------------------------------------------
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def reset(self, mask=None):
        dev = self.device
        obs = torch.empty(self.N, self.obs_dim, dtype=torch.float32, device=dev)
        obs64 = torch.empty(self.N, self.obs_dim, dtype=torch.float64, device=dev)
        st64 = torch.empty(self.N, self.nx, dtype=torch.float64, device=dev)
        m = None
        if mask is not None:
            m = torch.as_tensor(mask, dtype=torch.uint8, device=dev).contiguous()
        _abi.check(self.lib.scg_reset(self.h, _ptr(m), _ptr(obs), _ptr(obs64), _ptr(st64), self._stream()), self.h)
        return {'obs': obs, 'obs_f64': obs64, 'state_f64': st64}

    def set_adversary_control(self, adv):
        '''BenchmarkEnv.set_adversary_control for all N envs (benchmark_env.py:216-228): raw adversary actions [N][adv_dim]; the engine
        clips / scales / offsets them and applies the result at the next step().'''
        if isinstance(adv, np.ndarray):
            adv = torch.from_numpy(np.ascontiguousarray(adv))
        adv = torch.as_tensor(adv).to(self.device, torch.float32).reshape(self.N, -1).contiguous()
        if self.adv_dim and adv.shape[1] != self.adv_dim:
            raise ValueError(f'adversary action must have {self.adv_dim} components')
        _abi.check(self.lib.scg_set_adversary_control(self.h, _ptr(adv), self._stream()), self.h)

    def _alloc_step(self):
        dev, N, od, nx, nc = self.device, self.N, self.obs_dim, self.nx, max(self.n_con, 1)
        f32, f64, u8, i32 = torch.float32, torch.float64, torch.uint8, torch.int32
        b = {
            'obs': torch.empty(N, od, dtype=f32, device=dev), 'obs_f64': torch.empty(N, od, dtype=f64, device=dev),
            'state_f64': torch.empty(N, nx, dtype=f64, device=dev),
            'rew': torch.empty(N, dtype=f32, device=dev), 'rew_f64': torch.empty(N, dtype=f64, device=dev),
            'done': torch.empty(N, dtype=u8, device=dev), 'truncated': torch.empty(N, dtype=u8, device=dev),
            'cviol': torch.empty(N, dtype=f32, device=dev), 'cvals': torch.zeros(N, nc, dtype=f64, device=dev),
            'mse': torch.empty(N, dtype=f64, device=dev),
            'out_of_bounds': torch.empty(N, dtype=u8, device=dev), 'goal_reached': torch.empty(N, dtype=u8, device=dev),
            'current_step': torch.empty(N, dtype=i32, device=dev),
            'terminal_obs': torch.empty(N, od, dtype=f32, device=dev), 'terminal_obs_f64': torch.empty(N, od, dtype=f64, device=dev),
            'terminal_state_f64': torch.empty(N, nx, dtype=f64, device=dev),
        }
        return b

    def step(self, actions, fresh_buffers=True):
        '''actions: [N][nu] numpy / torch (float32 or float64).  Returns a dict of torch.cuda tensors.'''
        dev = self.device
        if isinstance(actions, np.ndarray):
            actions = torch.from_numpy(np.ascontiguousarray(actions))
        actions = actions.to(dev).reshape(self.N, self.nu).contiguous()
        a32 = a64 = None
        if actions.dtype == torch.float64:
            a64 = actions
        else:
            a32 = actions.to(torch.float32)
        b = self._alloc_step() if (fresh_buffers or self._step_bufs is None) else self._step_bufs
        self._step_bufs = b
        out = _abi.scg_step_out()
        for name in ('obs', 'obs_f64', 'state_f64', 'rew', 'rew_f64', 'done', 'truncated', 'cviol', 'mse', 'out_of_bounds',
                     'goal_reached', 'current_step', 'terminal_obs', 'terminal_obs_f64', 'terminal_state_f64'):
            setattr(out, 'd_' + name, b[name].data_ptr())
        out.d_cvals = b['cvals'].data_ptr() if self.n_con > 0 else None
        _abi.check(self.lib.scg_step(self.h, _ptr(a32), _ptr(a64), C.byref(out), self._stream()), self.h)
        return b

    def rollout(self, policy_blob, dims, T, eps=None, activation='tanh', gamma=0.99, gae_lambda=0.95, use_gae=False,
                deterministic=False, envs_per_block=0, want=('ret', 'adv', 'cviol', 'mse', 'last', 'episode'), out=None,
                flat=None, peer_byte_offsets=None, multicast_byte_offset=None, mlp_engine=None, norm=None, obs0=None, safety=None,
                adversary=None, eps_adv=None, adv_record=False):
        '''One fused launch: T steps of policy + env for all N envs.  Returns PPOBuffer-shaped torch.cuda tensors.

        flat: optional caller-owned float32 buffer for the PPOBuffer keys (e.g. this rank's slot of a peer-mapped gathered buffer);
        peer_byte_offsets: byte distances from that slot to the same slot in each peer GPU's gathered buffer -- the kernel then
        repeats every record store there (fused rollout + all-gather over NVLink, see PeerGatheredRollout in controllers/ppo);
        multicast_byte_offset: byte distance from that slot to the NVSwitch multicast mapping of the gathered buffer -- one
        `multimem.st` per record entry then replaces the per-peer copies.
        adversary: (blob, (obs_dim, adv_dim, hidden)) of the RARL / RAP adversary's actor-critic; adv_record: the record (act, v, logp, -rew,
        terminal / bootstrap values) is the adversary's (rarl.py:349-418).'''
        dev, N = self.device, self.N
        obs_dim, act_dim, hid = dims
        rec_act = act_dim
        if adversary is not None:
            adv_blob, adv_dims = adversary
            assert adv_dims[0] == obs_dim and adv_dims[2] == hid, 'protagonist and adversary share the observation and the hidden width'
            if adv_record:
                rec_act = int(adv_dims[1])
        f32 = torch.float32
        if out is None:
            # the PPOBuffer keys that feed the update live back to back in ONE flat allocation ([key][T][N][width]): multi-GPU
            # training exchanges it with a single all-gather, no packing copy (controllers/ppo/ppo.py: all_gather_rollout)
            widths = [('obs', obs_dim), ('act', rec_act), ('rew', 1), ('mask', 1), ('v', 1), ('logp', 1), ('terminal_v', 1)]
            if 'ret' in want or 'adv' in want:
                widths += [('ret', 1), ('adv', 1)]
            numel = sum(w for _, w in widths) * T * N
            if flat is None:
                flat = torch.empty(numel, dtype=f32, device=dev)
            assert flat.is_cuda and flat.dtype == f32 and flat.is_contiguous() and flat.numel() == numel, (flat.shape, numel)
            out, o = {'_flat': flat, '_layout': [(k, w) for k, w in widths], '_TN': (T, N)}, 0
            for k, w in widths:
                out[k] = flat[o:o + T * N * w].view(T, N, w)
                o += T * N * w
            if 'cviol' in want:
                out['cviol'] = torch.empty(T, N, 1, dtype=f32, device=dev)
            if 'cvals' in want and self.n_con > 0:
                out['cvals'] = torch.empty(T, N, self.n_con, dtype=f32, device=dev)
            if 'mse' in want:
                out['mse'] = torch.empty(T, N, 1, dtype=f32, device=dev)
            if 'last' in want:
                out['last_obs'] = torch.empty(N, obs_dim, dtype=f32, device=dev)
                out['last_val'] = torch.empty(N, 1, dtype=f32, device=dev)
            if safety is not None:     # SafePPOBuffer 'c': the constraint values the safety layer saw at every step
                out['c'] = torch.empty(T, N, int(safety['C']), dtype=f32, device=dev)
            if 'raw' in want:          # un-normalised observation / reward streams for the controller's running normalisers
                out['obs_raw'] = torch.empty(T, N, obs_dim, dtype=f32, device=dev)
                out['rew_raw'] = torch.empty(T, N, 1, dtype=f32, device=dev)
            if 'episode' in want:
                out['ep_return_sum'] = torch.empty(N, dtype=f32, device=dev)
                out['ep_length_sum'] = torch.empty(N, dtype=f32, device=dev)
                out['ep_cviol_sum'] = torch.empty(N, dtype=f32, device=dev)
                out['ep_count'] = torch.empty(N, dtype=torch.int32, device=dev)
        ro = _abi.scg_rollout_out()
        for name in ('obs', 'act', 'rew', 'mask', 'v', 'logp', 'terminal_v', 'ret', 'adv', 'cviol', 'cvals', 'mse', 'last_obs', 'last_val',
                     'ep_return_sum', 'ep_length_sum', 'ep_cviol_sum', 'ep_count', 'obs_raw', 'rew_raw', 'c'):
            setattr(ro, 'd_' + name, out[name].data_ptr() if name in out else None)
        args = _abi.scg_rollout_args()
        args.T, args.obs_dim, args.act_dim, args.hidden_dim = int(T), int(obs_dim), int(act_dim), int(hid)
        args.activation = _abi.ACTIVATIONS[activation] if isinstance(activation, str) else int(activation)
        args.use_gae, args.deterministic, args.envs_per_block = int(bool(use_gae)), int(bool(deterministic)), int(envs_per_block)
        args.gamma, args.gae_lambda = float(gamma), float(gae_lambda)
        args.d_policy = policy_blob.data_ptr()
        eng_name = mlp_engine if mlp_engine is not None else os.environ.get('SCG_MLP_ENGINE', 'auto')
        args.mlp_engine = _abi.MLP_ENGINES[eng_name]
        if eps is not None:
            assert eps.dtype == torch.float32 and eps.is_cuda and tuple(eps.shape) == (T, N, act_dim), (eps.dtype, eps.shape)
            eps = eps.contiguous()
        args.d_eps = eps.data_ptr() if eps is not None else None
        if norm is not None:
            assert norm.dtype == torch.float64 and norm.is_cuda and norm.numel() == 2 * obs_dim + 5, (norm.dtype, norm.shape)
            norm = norm.contiguous()
        args.d_norm = norm.data_ptr() if norm is not None else None
        if obs0 is not None:
            assert obs0.dtype == torch.float32 and obs0.is_cuda and tuple(obs0.shape) == (N, obs_dim), (obs0.dtype, obs0.shape)
            obs0 = obs0.contiguous()
        args.d_obs0 = obs0.data_ptr() if obs0 is not None else None
        if safety is not None:
            blob_s = safety['blob']
            Cn, Hn = int(safety['C']), int(safety['H'])
            assert blob_s.dtype == torch.float32 and blob_s.is_cuda and blob_s.is_contiguous() and blob_s.numel() == Cn * (Hn * obs_dim + Hn + act_dim * Hn + act_dim)
            args.safety_constraints, args.safety_hidden, args.d_safety = Cn, Hn, blob_s.data_ptr()
            for i, v in enumerate(safety['slack']):
                args.safety_slack[i] = float(v)
        if adversary is not None:
            assert adv_blob.dtype == torch.float32 and adv_blob.is_cuda and adv_blob.is_contiguous()
            assert tuple(out['act'].shape) == (T, N, rec_act), (out['act'].shape, rec_act)
            args.d_adv_policy, args.adv_act_dim, args.adv_record = adv_blob.data_ptr(), int(adv_dims[1]), int(bool(adv_record))
            if eps_adv is not None:
                assert eps_adv.dtype == torch.float32 and eps_adv.is_cuda and tuple(eps_adv.shape) == (T, N, int(adv_dims[1]))
                eps_adv = eps_adv.contiguous()
                args.d_eps_adv = eps_adv.data_ptr()
        offs = list(peer_byte_offsets) if peer_byte_offsets is not None else out.get('_peer_byte_offsets', [])
        if len(offs) > _abi.SCG_MAX_PEERS:
            raise ValueError(f'at most {_abi.SCG_MAX_PEERS} peers')
        args.n_peers = len(offs)
        for g, v in enumerate(offs):
            args.peer_byte_offset[g] = int(v)
        if offs:
            out['_peer_byte_offsets'] = offs
        mc = multicast_byte_offset if multicast_byte_offset is not None else out.get('_multicast_byte_offset', 0)
        args.multicast_byte_offset = int(mc or 0)
        if mc:
            out['_multicast_byte_offset'] = int(mc)
        _abi.check(self.lib.scg_rollout(self.h, C.byref(args), C.byref(ro), self._stream()), self.h)
        return out

    # -- checkpoint -----------------------------------------------------------------------------
    def get_state(self):
        n = C.c_size_t()
        _abi.check(self.lib.scg_state_bytes(self.h, C.byref(n)), self.h)
        buf = np.empty(n.value, dtype=np.uint8)
        _abi.check(self.lib.scg_get_state(self.h, _np_ptr(buf), n.value, self._stream()), self.h)
        return buf

    def set_state(self, buf):
        buf = np.ascontiguousarray(buf, dtype=np.uint8)
        _abi.check(self.lib.scg_set_state(self.h, _np_ptr(buf), buf.size, self._stream()), self.h)

    def launch_count(self):
        n = C.c_uint64()
        _abi.check(self.lib.scg_launch_count(self.h, C.byref(n)), self.h)
        return n.value

    def close(self):
        if getattr(self, 'h', None) is not None and self.h:
            self.lib.scg_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def pack_policy_host(ac_or_params, dims, blob_out):
    '''Host-side packer (C-ABI `scg_policy_pack`): parameters in host memory -> packed blob in (pinned) host memory.
    `ac_or_params`: an MLPActorCritic on the CPU, or a dict {'actor_w': [3], 'actor_b': [3], 'logstd', 'critic_w': [3], 'critic_b': [3]}
    of contiguous float32 CPU tensors.'''
    if not isinstance(ac_or_params, dict):
        a, c = ac_or_params.actor.pi_net.fcs, ac_or_params.critic.v_net.fcs
        ac_or_params = {'actor_w': [l.weight for l in a], 'actor_b': [l.bias for l in a], 'logstd': ac_or_params.actor.logstd,
                        'critic_w': [l.weight for l in c], 'critic_b': [l.bias for l in c]}
    P = ac_or_params
    obs_dim, act_dim, hid = dims

    def arr(ts):
        ts = [t.detach() for t in ts]
        assert all(t.device.type == 'cpu' and t.dtype == torch.float32 and t.is_contiguous() for t in ts)
        return (C.c_void_p * 3)(*[t.data_ptr() for t in ts]), ts

    aw, k1 = arr(P['actor_w']); ab, k2 = arr(P['actor_b']); cw, k3 = arr(P['critic_w']); cb, k4 = arr(P['critic_b'])
    ls = P['logstd'].detach()
    assert blob_out.device.type == 'cpu' and blob_out.dtype == torch.float32 and blob_out.is_contiguous()
    assert blob_out.numel() >= _abi.lib().scg_policy_blob_floats(obs_dim, act_dim, hid)
    _abi.check(_abi.lib().scg_policy_pack(obs_dim, act_dim, hid, aw, ab, ls.data_ptr(), cw, cb, blob_out.data_ptr()))
    return blob_out


def bind_to_gpu_numa_node(device_index):
    '''Pins the calling process to the CPUs of the NUMA node its GPU hangs off (sysfs), so that pinned host buffers allocated afterwards
    are node-local (first touch) and the D2H DMA of the rollout record does not cross the socket interconnect.  One process per GPU
    (torchrun) otherwise floats over both sockets and eight concurrent record copies collapse (SCALE_r01: 8-GPU e2e below 4-GPU).
    Returns the node index, or None when the topology cannot be read (nothing is changed then).'''
    try:
        pci = torch.cuda.get_device_properties(device_index)
        bus_id = f'{pci.pci_domain_id:04x}:{pci.pci_bus_id:02x}:{pci.pci_device_id:02x}.0'
        with open(f'/sys/bus/pci/devices/{bus_id}/numa_node') as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open(f'/sys/devices/system/node/node{node}/cpulist') as f:
            cpus = set()
            for part in f.read().strip().split(','):
                lo, _, hi = part.partition('-')
                cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except Exception:   # noqa: BLE001 -- best effort: containers often hide sysfs
        return None


class HostCollector:
    '''The fused collector for HOST-side consumers (the reference keeps policy parameters and the PPOBuffer in host memory,
    ppo_utils.py:241-355): `submit(params)` packs the parameters on the host, copies the blob to the GPU, launches the fused rollout and
    queues the device->host copy of the whole PPOBuffer record into pinned memory; `fetch()` returns the oldest record as numpy-backed
    views.  `depth` records are in flight: the D2H copy of rollout k runs on a copy stream underneath rollout k+1.
    `keys`: which record keys travel to the host (default: all of them).  The host-side `PPOAgent.update` of the reference consumes
    obs, act, logp, adv, ret, v (ppo_utils.py:56-80,82-146); rew / mask / terminal_v are only inputs of compute_returns_and_advantages,
    which the collector kernel has already evaluated.'''

    UPDATE_KEYS = ('obs', 'act', 'logp', 'adv', 'ret', 'v')

    def __init__(self, eng, dims, T, depth=2, numa_bind=True, keys=None, **rollout_kw):
        self.eng, self.dims, self.T, self.depth, self.kw = eng, dims, int(T), int(depth), rollout_kw
        self.keys = None if keys is None else tuple(keys)
        self._plan = None
        dev = eng.device
        self.numa_node = bind_to_gpu_numa_node(eng.device_index) if numa_bind else None
        nblob = _abi.lib().scg_policy_blob_floats(*dims)
        self.blob_host = [torch.zeros(nblob, dtype=torch.float32).pin_memory() for _ in range(depth)]
        self.blob_dev = [torch.zeros(nblob, dtype=torch.float32, device=dev) for _ in range(depth)]
        self.out = [None] * depth
        self.host = [None] * depth
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.ev_done = [torch.cuda.Event() for _ in range(depth)]
        self.ev_copied = [torch.cuda.Event() for _ in range(depth)]
        self.submitted = self.fetched = 0

    @property
    def h2d_bytes(self):
        return self.blob_host[0].numel() * 4

    @property
    def d2h_bytes(self):
        return self.host[0].numel() * 4 if self.host[0] is not None else 0

    def submit(self, params):
        if self.submitted - self.fetched >= self.depth:
            raise RuntimeError('HostCollector: fetch() the oldest record before submitting another rollout')
        s = self.submitted % self.depth
        pack_policy_host(params, self.dims, self.blob_host[s])
        cur = torch.cuda.current_stream(self.eng.device)
        if self.submitted >= self.depth:
            cur.wait_event(self.ev_copied[s])            # the record buffer is free once its previous D2H has finished
        self.blob_dev[s].copy_(self.blob_host[s], non_blocking=True)
        self.out[s] = self.eng.rollout(self.blob_dev[s], self.dims, self.T, out=self.out[s], **self.kw)
        self.ev_done[s].record(cur)
        flat = self.out[s]['_flat']
        if self._plan is None:            # [(key, width, device offset, host offset)] + merged copy segments (adjacent keys travel in one DMA)
            T_, N_ = self.out[s]['_TN']
            plan, segs, o, ho = [], [], 0, 0
            for k, w in self.out[s]['_layout']:
                n = T_ * N_ * w
                if self.keys is None or k in self.keys:
                    plan.append((k, w, o, ho))
                    if segs and segs[-1][0] + segs[-1][2] == o:
                        segs[-1][2] += n
                    else:
                        segs.append([o, ho, n])
                    ho += n
                o += n
            if self.keys is not None and len(plan) != len(self.keys):
                raise KeyError(f'HostCollector: unknown record key in {self.keys}')
            self._plan, self._segs, self._host_numel = plan, segs, ho
        if self.host[s] is None:
            self.host[s] = torch.empty(self._host_numel, dtype=torch.float32).pin_memory()
        self.copy_stream.wait_event(self.ev_done[s])
        with torch.cuda.stream(self.copy_stream):
            for o, ho, n in self._segs:
                self.host[s][ho:ho + n].copy_(flat[o:o + n], non_blocking=True)
            self.ev_copied[s].record(self.copy_stream)
        self.submitted += 1

    def fetch(self):
        '''Blocks until the oldest in-flight record is in host memory; returns {key: numpy view [T][N][width]} (valid until the slot is reused).'''
        if self.fetched >= self.submitted:
            raise RuntimeError('HostCollector: nothing in flight')
        s = self.fetched % self.depth
        self.ev_copied[s].synchronize()
        self.fetched += 1
        T, N = self.out[s]['_TN']
        res, h = {}, self.host[s].numpy()
        for k, w, _, ho in self._plan:
            res[k] = h[ho:ho + T * N * w].reshape(T, N, w)
        return res


# ------------------------------------------------------------------------------------------------------------------
# fused PPO update (include/scg.h: scg_ppo_update; reference ppo_utils.py:82-146)
TRAIN_SECTIONS = ('actor.pi_net.fcs.0.weight', 'actor.pi_net.fcs.0.bias', 'actor.pi_net.fcs.1.weight', 'actor.pi_net.fcs.1.bias',
                  'actor.pi_net.fcs.2.weight', 'actor.pi_net.fcs.2.bias', 'actor.logstd',
                  'critic.v_net.fcs.0.weight', 'critic.v_net.fcs.0.bias', 'critic.v_net.fcs.1.weight', 'critic.v_net.fcs.1.bias',
                  'critic.v_net.fcs.2.weight', 'critic.v_net.fcs.2.bias')


def ppo_train_layout(obs_dim, act_dim, hid):
    '''{parameter name: float offset} of the flat training vector + 'critic0' (first critic float) + 'total' (C: scg_ppo_param_layout).'''
    o = (C.c_int32 * 15)()
    _abi.check(_abi.lib().scg_ppo_param_layout(obs_dim, act_dim, hid, C.byref(o)))
    L = {name: int(o[i]) for i, name in enumerate(TRAIN_SECTIONS)}
    L['critic0'], L['total'] = int(o[13]), int(o[14])
    return L


class FusedPPOState:
    '''Flat training state of an MLPActorCritic on the GPU: parameters (the module's nn.Parameters become views into `params`),
    Adam moments and step counts, in the train layout of include/scg.h.'''

    def __init__(self, ac, device):
        a, c = ac.actor.pi_net.fcs, ac.critic.v_net.fcs
        if len(a) != 3 or len(c) != 3:
            raise NotImplementedError('the fused PPO update supports two hidden layers (hidden_dims=[h, h])')
        self.dims = (a[0].in_features, a[2].out_features, a[0].out_features)
        self.device = torch.device(device)
        self.layout = ppo_train_layout(*self.dims)
        n = self.layout['total']
        self.params = torch.zeros(n, dtype=torch.float32, device=self.device)
        self.exp_avg = torch.zeros_like(self.params)
        self.exp_avg_sq = torch.zeros_like(self.params)
        self.steps = torch.zeros(2, dtype=torch.int64, device=self.device)      # actor, critic optimizer step counts
        self.stats = torch.zeros(8, dtype=torch.float32, device=self.device)
        self.workspace = None
        self.named = dict(ac.named_parameters())
        assert set(self.named) == set(TRAIN_SECTIONS), sorted(self.named)
        with torch.no_grad():
            for name in TRAIN_SECTIONS:
                p, o = self.named[name], self.layout[name]
                view = self.params[o:o + p.numel()].view(p.shape)
                view.copy_(p.detach().to(self.device, torch.float32))
                p.data = view                       # the module now reads / writes the flat vector
        self.blob = torch.zeros(_abi.lib().scg_policy_blob_floats(*self.dims), dtype=torch.float32, device=self.device)

    def views(self, flat, name):
        p, o = self.named[name], self.layout[name]
        return flat[o:o + p.numel()].view(p.shape)

    def rollout_blob(self):
        '''Packed policy blob for scg_rollout, refreshed from the training vector by one small kernel.'''
        _abi.check(_abi.lib().scg_policy_pack_device(*self.dims, self.params.data_ptr(), self.blob.data_ptr(),
                                                    torch.cuda.current_stream(self.device).cuda_stream))
        return self.blob, self.dims

    def update(self, batch, perm, mini_batch_size, activation='tanh', clip_param=0.2, target_kl=0.01, entropy_coef=0.01,
               use_clipped_value=False, actor_lr=3e-4, critic_lr=1e-3, betas=(0.9, 0.999), eps=1e-8, grad_out=None, num_mini_batch=None):
        '''batch: {'obs','act','logp','adv','ret','v'} float32 device tensors [B][R][w] (stride(1) == w, stride(2) == 1; any block stride);
        perm: int32 device tensor [opt_epochs][>= num_mini_batch * mini_batch_size] of row indices into the flattened [B*R] rows.
        Returns the number of minibatches per epoch.  One kernel launch; results accumulate into self.stats.'''
        a = _abi.scg_ppo_args()
        a.obs_dim, a.act_dim, a.hidden_dim = self.dims
        a.activation = _abi.ACTIVATIONS[activation]
        widths = {'obs': self.dims[0], 'act': self.dims[1], 'logp': 1, 'adv': 1, 'ret': 1, 'v': 1}
        B, R = batch['obs'].shape[0], batch['obs'].shape[1]
        keep = []
        for i, k in enumerate(_abi.PPO_KEYS):
            t = batch[k]
            if t.dtype != torch.float32 or t.device != self.device or tuple(t.shape) != (B, R, widths[k]):
                raise ValueError(f'fused PPO update: bad record tensor {k}: {t.dtype} {tuple(t.shape)} {t.device}')
            if (t.stride(2) != 1 and widths[k] > 1) or t.stride(1) != widths[k]:
                t = t.contiguous()
            keep.append(t)
            a.d_key[i] = t.data_ptr()
            a.block_stride[i] = t.stride(0)
        a.block_rows = R
        total = B * R
        a.mini_batch_size = int(mini_batch_size)
        a.num_mini_batch = total // int(mini_batch_size) if num_mini_batch is None else int(num_mini_batch)     # ppo_utils.py:93
        if perm.dtype != torch.int32 or perm.dim() != 2 or not perm.is_contiguous() or perm.device != self.device:
            raise ValueError('fused PPO update: perm must be a contiguous int32 device tensor [opt_epochs][rows]')
        a.opt_epochs, a.perm_stride, a.d_perm = perm.shape[0], perm.shape[1], perm.data_ptr()
        a.use_clipped_value = int(bool(use_clipped_value))
        a.clip_param, a.target_kl, a.entropy_coef = float(clip_param), float(target_kl), float(entropy_coef)
        a.actor_lr, a.critic_lr, a.beta1, a.beta2, a.adam_eps = float(actor_lr), float(critic_lr), float(betas[0]), float(betas[1]), float(eps)
        a.d_params, a.d_exp_avg, a.d_exp_avg_sq = self.params.data_ptr(), self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr()
        a.d_adam_steps, a.d_stats = self.steps.data_ptr(), self.stats.data_ptr()
        a.d_grad_out = grad_out.data_ptr() if grad_out is not None else None
        if a.num_mini_batch < 1:
            raise AssertionError('num_mini_batch is 0')            # ppo_utils.py:94
        need = _abi.lib().scg_ppo_workspace_bytes(C.byref(a))
        if need == 0:
            _abi.check(_abi.SCG_EINVAL)
        if self.workspace is None or self.workspace.numel() * 4 < need:
            self.workspace = torch.zeros((need + 3) // 4, dtype=torch.float32, device=self.device)     # pad floats are never written: keep them 0
        a.d_workspace, a.workspace_bytes = self.workspace.data_ptr(), self.workspace.numel() * 4
        _abi.check(_abi.lib().scg_ppo_update(C.byref(a), torch.cuda.current_stream(self.device).cuda_stream))
        return a.num_mini_batch
