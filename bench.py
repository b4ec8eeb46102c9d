#!/usr/bin/env python
'''Benchmark of the hot path: env-steps/s of the fused PPO rollout collector (policy forward + env.step + PPOBuffer write).

    python bench.py --gpus N --steps K --warmup W [--config c2|c3|c4] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

One "step" = one pass of the hot path over one batch = ONE fused rollout launch over a T = 200 control-step horizon:
  c3 (default, the configuration BASELINE.json's metric is quoted on): 4096 Quadrotor-2D trajectory-tracking envs per GPU
      (examples/rl/config_overrides/quadrotor_2D/quadrotor_2D_track.yaml), 64-64 tanh actor-critic, 20 physics sub-steps per step;
  c2: 4096 CartPole envs (cartpole.yaml defaults), 64-64 policy;
  c4: 8192 Quadrotor-3D figure-8 envs with Philox mass / inertia randomisation, white-noise dynamics disturbance and the 32 constraint
      rows, constraint values recorded.
float64 env arithmetic + float32 policy exactly like the reference (--precision f32 = throughput mode).

Printed JSON (rank 0, one line): `value` = device-resident throughput (inputs in HBM), `e2e` = the same metric through the
reference-facing collector call with HOST buffers (policy weights H2D + PPOBuffer record D2H inside the timed region),
`roofline` (HBM bound, algorithmic bytes of SURVEY.md 8d), `cpu_baseline` (the C oracle port on the host cores), `train_e2e`
(BASELINE config 5 semantics: rollout + gather + advantage normalisation + fused PPO update per train step).
`--impl reference` times the reference's CPU algorithm (oracle port, all host threads) on the same config.
'''
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

T_HORIZON = 200
METRIC = 'env-steps/sec (4096 envs, 200-step horizon)'

# algorithmic bytes per env-step = 4 * (obs_dim + act_dim + 6) -- fp32 obs, act, rew, mask, v, logp, terminal_v, cviol
# (+ 4 * n_con when the constraint VALUES are recorded: c4)
WORKLOADS = {
    'c2': dict(env='cartpole', task='cartpole_c2', envs=4096, obs=4, act=1, substeps=1, bytes=44, sys='CARTPOLE',
               text='CartPole stabilisation (cartpole.yaml defaults), PPO 64-64 tanh actor-critic, fused rollout'),
    'c3': dict(env='quadrotor', task='quadrotor_2D_track', envs=4096, obs=12, act=2, substeps=20, bytes=80, sys='QUAD2D',
               text='Quadrotor-2D trajectory tracking (quadrotor_2D_track.yaml), PPO 64-64 tanh actor-critic, fused rollout'),
    'c4': dict(env='quadrotor', task='quadrotor_3D_track_c4', envs=8192, obs=24, act=4, substeps=20, bytes=136 + 128, sys='QUAD3D',
               text='Quadrotor-3D figure-8 tracking (quadrotor_3D_track.yaml) + Philox M/Ixx/Iyy/Izz randomisation + white-noise dynamics '
                    'sigma 0.01 + 32 constraint rows (values recorded), PPO 64-64 tanh actor-critic, fused rollout'),
}


def mlp_flops(obs, act, hid):
    return 2 * ((obs * hid + hid * hid + hid * act) + (obs * hid + hid * hid + hid))


def config_dict(args, world):
    '''The SAME dict in both arms (the driver compares them): what is measured, not how an arm runs it.'''
    w = WORKLOADS[args.config]
    return {'workload': w['text'], 'config': args.config, 'envs_per_gpu': args.envs_per_gpu, 'horizon': T_HORIZON, 'hidden_dim': args.hidden,
            'physics_substeps_per_step': w['substeps'], 'one_bench_step': 'one fused rollout launch (envs_per_gpu x horizon env-steps per GPU)',
            'rng': 'philox seed 1337', 'precision': args.precision, 'parallelism': f'env-sharded x{world}',
            'l2': 'flushed with a 256 MiB write before every timed step; per-step CUDA-event intervals summed'}


def measured_peak_hbm():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        return float(json.load(open(p))['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


class ClockSampler:
    '''Samples nvidia-smi clocks / throttle reasons while the timed region runs.'''
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits', '-lms', '100', '-i', str(self.gpu)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[5:9]):
                    if v.lower().startswith('active'):
                        reasons.add(name)
            except Exception:
                pass
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None, 'samples': len(sm),
                'reasons': sorted(reasons)}


def build_policy(env, hidden, seed=0):
    import torch

    from safe_control_gym_b200.controllers.ppo.ppo_utils import MLPActorCritic
    torch.manual_seed(seed)
    return MLPActorCritic(env.observation_space, env.action_space, hidden_dims=[hidden, hidden], activation='tanh')


def run_train_e2e(args, world, rank, local_rank):
    '''BASELINE config 5 semantics through the controller API: K x PPO.train_step() = fused rollout (+ in-kernel all-gather) +
    advantage normalisation + fused PPO update (opt_epochs x minibatches in one launch) + re-pack of the policy blob.'''
    from functools import partial

    import torch
    import torch.distributed as dist

    from safe_control_gym_b200 import configs
    from safe_control_gym_b200.utils.registration import make
    w = WORKLOADS[args.config]
    N = args.envs_per_gpu
    task = {k: v for k, v in configs.get(w['task']).items() if k != 'seed'}
    env_func = partial(make, w['env'], **task)
    mb, epochs = args.train_mini_batch, args.train_epochs
    ctrl = make('ppo', env_func, training=True, output_dir='/tmp/scg_bench_train', seed=1337, rollout_batch_size=world * N, rollout_steps=T_HORIZON,
                mini_batch_size=mb, opt_epochs=epochs, use_gae=True, hidden_dim=args.hidden, device=local_rank, precision=args.precision)
    ctrl.reset()
    for _ in range(2):
        ctrl.train_step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    K = args.train_steps
    t0 = time.perf_counter()
    upd_ms = 0.0
    for _ in range(K):
        r = ctrl.train_step()
        upd_ms += r.get('update_ms', 0.0)
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=torch.device('cuda', local_rank))
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    dt = float(dt)
    rows = world * N * T_HORIZON
    res = {'value': rows * K / dt, 'unit': 'env-steps/s', 'ms_per_train_step': dt / K * 1e3, 'train_steps': K,
           'rows_per_update': rows, 'mini_batch_size': mb, 'opt_epochs': epochs, 'minibatches_per_train_step': epochs * (rows // mb),
           'update_ms_per_train_step': upd_ms / K if upd_ms else None,
           'update_samples_per_s': (epochs * (rows // mb) * mb) / (upd_ms / K * 1e-3) if upd_ms else None,
           'actor_steps_last': r.get('actor_steps'), 'approx_kl_last': r.get('approx_kl'),
           'what': 'PPO.train_step through make(\'ppo\', ...): fused rollout' + (' + in-kernel all-gather over NVLink' if world > 1 else '')
                   + ' + advantage normalisation + scg_ppo_update (all epochs x minibatches in one cooperative launch) + policy re-pack; '
                     'wall clock over K train steps incl. the host-side statistics readback'}
    # ppo.yaml default minibatch (64) on a bounded sample: one epoch of `n` minibatches of the same record
    try:
        agent = ctrl.agent
        if agent.fused is not None and rank == 0:
            out = ctrl._rollout_out[T_HORIZON]
            flat = {k: out[k].reshape(1, -1, out[k].shape[-1]) for k in ('obs', 'act', 'logp', 'ret', 'v')}
            adv = out['adv'].reshape(1, -1, 1)
            flat['adv'] = (adv - adv.mean()) / (adv.std(unbiased=False) + 1e-6)
            n_rows = flat['obs'].shape[1]
            n_mb = min(2000, n_rows // 64)
            perm = torch.randperm(n_rows, device=adv.device)[:n_mb * 64].to(torch.int32).reshape(1, -1).contiguous()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            agent.fused.update(flat, perm, 64, num_mini_batch=n_mb)
            torch.cuda.synchronize()
            e0.record()
            agent.fused.update(flat, perm, 64, num_mini_batch=n_mb)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            res['update_mb64'] = {'samples_per_s': n_mb * 64 / (ms * 1e-3), 'us_per_minibatch': ms * 1e3 / n_mb, 'minibatches': n_mb,
                                  'what': 'scg_ppo_update, mini_batch_size 64 (ppo.yaml default), one epoch over a bounded sample of the C3 record'}
    except Exception as e:   # noqa: BLE001 -- the headline numbers must survive a failure of this auxiliary leg
        res['update_mb64'] = {'error': f'{type(e).__name__}: {e}'}
    ctrl.close()
    return res


def run_ours(args):
    import torch
    import torch.distributed as dist

    import safe_control_gym_b200  # noqa: F401
    from safe_control_gym_b200 import configs
    from safe_control_gym_b200.controllers.ppo.ppo import PeerGatheredRollout, rollout_flat_numel
    from safe_control_gym_b200.engine import EngineHandle, pack_policy
    from safe_control_gym_b200.utils.registration import make

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    w = WORKLOADS[args.config]
    N = args.envs_per_gpu
    env = make(w['env'], **configs.get(w['task']))
    eng = EngineHandle(env, num_envs=N, auto_reset=True, rng_mode='philox', seed=1337, env_index_offset=rank * N, device=local_rank,
                       precision=args.precision)
    eng.reset()
    ac = build_policy(env, args.hidden).to(dev)
    blob, dims = pack_policy(ac, dev)
    want = ('ret', 'adv', 'cviol', 'last', 'episode') + (('cvals',) if args.config == 'c4' else ())
    # N > 1: the rollout record is all-gathered INSIDE the collector kernel (stores into every rank's gathered buffer over NVLink,
    # torch symmetric memory); the per-step collective is one barrier.  Falls back to the NCCL all-gather if P2P is unavailable.
    pg = PeerGatheredRollout(rollout_flat_numel(dims[0], dims[1], T_HORIZON, N), dev) if world > 1 else None
    out = eng.rollout(blob, dims, T_HORIZON, use_gae=True, want=want, envs_per_block=args.envs_per_block, **(pg.rollout_kwargs() if pg else {}))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)          # > 126 MB L2

    def one_step():
        if world > 1:
            pg.begin()                 # nobody may still be reading the gathered buffer of the previous step
        o = eng.rollout(blob, dims, T_HORIZON, use_gae=True, want=want, out=out, envs_per_block=args.envs_per_block)
        if world > 1:
            pg.finish(o)               # gathered record complete on every rank (barrier; NCCL all-gather in the fallback)
        return o

    sampler = ClockSampler(local_rank)      # samples from the warm-up to the end of the end-to-end loop (the timed region alone is ~ms)
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        flush.fill_(1)
        one_step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    l0 = eng.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
          for _ in range(args.steps)]
    torch.cuda.synchronize()
    t_wall = time.perf_counter()
    for a, k, b in ev:
        flush.fill_(1)                 # L2 flush between timed iterations (not inside the event-timed interval)
        if world > 1:
            pg.begin()
        a.record()
        o = eng.rollout(blob, dims, T_HORIZON, use_gae=True, want=want, out=out, envs_per_block=args.envs_per_block)
        k.record()                     # kernel-only interval for the roofline
        if world > 1:
            pg.finish(o)
        b.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t_wall = time.perf_counter() - t_wall
    launches = eng.launch_count() - l0
    step_ms = [a.elapsed_time(b) for a, _, b in ev]
    kern_ms = [a.elapsed_time(k) for a, k, _ in ev]
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms)
    env_steps = world * N * T_HORIZON * args.steps
    value = env_steps / (total_ms * 1e-3)

    # ---- end-to-end through the collector with HOST buffers (engine.HostCollector): every step the policy parameters come from host
    # memory (host pack -> one H2D) and the PPOBuffer record goes back to pinned host memory (one D2H).  Two records in flight:
    # the D2H of rollout k overlaps rollout k+1.  `serial_value` is the same with one record in flight (no overlap).
    from safe_control_gym_b200.engine import HostCollector
    ac_host = build_policy(env, args.hidden)

    def e2e_rate(depth, n_steps, keys=None):
        hc = HostCollector(eng, dims, T_HORIZON, depth=depth, use_gae=True, want=want, envs_per_block=args.envs_per_block, keys=keys)
        probe = 'rew' if keys is None else 'ret'
        acc = 0.0
        for _ in range(depth + 1):                     # warm-up (allocates the pinned buffers)
            hc.submit(ac_host)
            acc += float(hc.fetch()[probe][0, 0, 0])
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for i in range(n_steps):
            hc.submit(ac_host)
            if i >= depth - 1:
                acc += float(hc.fetch()[probe][-1, -1, 0])      # the host consumes every record
        while hc.fetched < hc.submitted:
            acc += float(hc.fetch()[probe][-1, -1, 0])
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        return world * N * T_HORIZON * n_steps / float(dt), hc.h2d_bytes, hc.d2h_bytes, acc

    e2e_steps = max(8, min(args.steps, 32))
    e2e_serial, h2d_bytes, d2h_bytes, _ = e2e_rate(1, e2e_steps)
    e2e_value, _, _, _ = e2e_rate(2, e2e_steps)
    # the same with only the keys a host-side PPOAgent.update reads (obs, act, logp, adv, ret, v): 6 of the 9 PPOBuffer keys
    e2e_upd, _, d2h_upd, _ = e2e_rate(2, e2e_steps, keys=HostCollector.UPDATE_KEYS)
    # the timed regions above last milliseconds: keep the same kernel running for ~2 s so that nvidia-smi (100 ms period)
    # records the clocks and throttle reasons UNDER THIS LOAD
    t_end = time.perf_counter() + 2.0
    while time.perf_counter() < t_end:
        for _ in range(20):
            one_step()
        torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None
    eng.close()

    train = None
    if not args.no_train:
        try:
            train = run_train_e2e(args, world, rank, local_rank)
        except Exception as e:   # noqa: BLE001
            train = {'error': f'{type(e).__name__}: {e}'}

    res = None
    if rank == 0:
        peak, peak_src = measured_peak_hbm()
        kern_s = float(np.mean(kern_ms)) * 1e-3
        achieved = w['bytes'] * N * T_HORIZON / kern_s / 1e9
        flops = mlp_flops(w['obs'], w['act'], args.hidden) * N * T_HORIZON / kern_s / 1e12
        prof = os.path.join(ROOT, 'profiles', 'ncu_summary.json')
        traffic, traffic_src = None, None
        if args.config == 'c3' and args.hidden == 64 and os.path.exists(prof):
            traffic = json.load(open(prof)).get('dram_bytes_per_launch')
            traffic_src = 'profiles/ncu_summary.json (ncu --set full capture of this kernel, dram__bytes_read.sum + dram__bytes_write.sum)'
        res = {
            'metric': METRIC, 'value': value, 'unit': 'env-steps/s', 'n_gpus': world,
            'steps': args.steps, 'warmup': max(args.warmup, 3), 'ms_per_step': total_ms / args.steps, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64 env state + f32 policy (as the reference)' if args.precision == 'f64' else 'f32',
            'data': f'synthetic (random-init {args.hidden}-{args.hidden} actor-critic, Philox-seeded initial states)',
            'config': config_dict(args, world),
            'gather': None if world == 1 else (
                'rollout record all-gathered inside the collector kernel (NVLink multimem / peer stores, torch symmetric memory) + two barriers per step'
                if pg.fused else f'one NCCL all-gather of the rollout record per step (peer memory unavailable: {pg.why})'),
            'gpu_launches': int(launches),
            'e2e': {'value': e2e_value, 'unit': 'env-steps/s', 'h2d_bytes_per_step': int(h2d_bytes), 'd2h_bytes_per_step': int(d2h_bytes),
                    'serial_value': e2e_serial, 'steps': e2e_steps,
                    'update_keys': {'value': e2e_upd, 'd2h_bytes_per_step': int(d2h_upd),
                                    'what': 'only obs, act, logp, adv, ret, v travel to the host: what the reference PPOAgent.update reads '
                                            '(ppo_utils.py:56-80); rew / mask / terminal_v feed compute_returns_and_advantages, done in the kernel'},
                    'what': 'engine.HostCollector: host policy parameters -> scg_policy_pack -> H2D, fused rollout, full PPOBuffer record '
                            '(obs, act, rew, mask, v, logp, terminal_v, ret, adv) -> pinned host, consumed by the host every step; 2 records in flight '
                            '(D2H of rollout k overlaps rollout k+1); serial_value = 1 record in flight'},
            'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak, 'traffic': traffic,
                         'traffic_source': traffic_src, 'peak_source': peak_src,
                         'kernel': f'rollout_kernel<{w["sys"]}, {"double" if args.precision == "f64" else "float"}, {args.hidden}, ...>', 'kernel_ms': kern_s * 1e3,
                         'algorithmic_bytes_per_env_step': w['bytes'],
                         'note': 'the path is latency / FP32-issue bound, not HBM bound (SURVEY.md 8d): see fp32 below',
                         'fp32': {'achieved_tflops': flops, 'peak_tflops': 74.4, 'frac': flops / 74.4,
                                  'peak_source': 'computed: 148 SMs x 128 lanes x 2 x 1.965 GHz'}},
            'train_e2e': train,
            'clocks': clocks,
            'wall_s_timed_region': t_wall,
        }
    if world > 1:
        dist.destroy_process_group()
    return res


def host_threads():
    '''Threads the CPU legs use: physical cores inside this process's affinity mask (torchrun exports OMP_NUM_THREADS=1, so ask
    explicitly; SMT siblings only slow the FP64 loops down).'''
    try:
        import psutil
        ncpu = psutil.cpu_count(logical=False) or os.cpu_count() or 1
    except Exception:
        ncpu = os.cpu_count() or 1
    aff = len(os.sched_getaffinity(0))
    return max(1, min(ncpu, aff)), aff, os.cpu_count() or 1


def oracle_rate(cfg_name, n_envs, hidden, steps, threads=None, precision='f64'):
    '''env-steps/s of the CPU oracle port on `n_envs` envs x T_HORIZON steps, `steps` times (bounded sample of the same workload).'''
    import helpers as H
    from oracle import oracle as orc
    orc.build()
    if threads:
        orc.set_threads(threads)
    env = H.make_env(WORKLOADS[cfg_name]['task'])
    o = H.make_oracle(env, n_envs, auto_reset=True, rng_mode='philox', seed=1337, precision=precision)
    o.reset()
    pol = H.random_policy(env.obs_dim, env.action_dim, hidden, seed=0)
    o.rollout(pol, 10, use_gae=True)
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        o.rollout(pol, T_HORIZON, use_gae=True)
        ts.append(time.perf_counter() - t0)
    return n_envs * T_HORIZON * len(ts) / sum(ts), sum(ts), orc.num_threads()


def cpu_baseline_dict(args, rate, secs, threads, reps, n_env):
    _, aff, logical = host_threads()
    return {'value': rate, 'unit': 'env-steps/s', 'cores': threads, 'kind': 'port',
            'threads_used': threads, 'affinity_cpus': aff, 'logical_cpus': logical, 'per_core_rate': rate / max(threads, 1),
            'sample': f'{n_env} envs x {T_HORIZON} control steps x {reps} repeats ({secs:.1f} s), C oracle (oracle/scg_oracle.c) port of the '
                      f'reference path, OpenMP over envs, config {args.config}'}


def run_reference(args):
    '''The reference's own CPU algorithm for the path (PyBullet is not installable here: the C oracle port is timed, with all the
    host threads OpenMP gives it), same config / metric / unit.'''
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return None
    n = args.envs_per_gpu
    ncpu, _, _ = host_threads()
    for _ in range(args.warmup):
        oracle_rate(args.config, n, args.hidden, 1, threads=ncpu, precision=args.precision)
    rate, secs, threads = oracle_rate(args.config, n, args.hidden, args.steps, threads=ncpu, precision=args.precision)
    world = int(os.environ.get('WORLD_SIZE', '1'))
    return {
        'impl': 'reference', 'metric': METRIC, 'value': rate, 'unit': 'env-steps/s',
        'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': secs / args.steps * 1e3,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f64 env state + f32 policy (as the reference)' if args.precision == 'f64' else 'f32', 'data': 'synthetic',
        'config': config_dict(args, world),
        'cpu_baseline': cpu_baseline_dict(args, rate, secs, threads, args.steps, n),
        'e2e': {'value': rate, 'unit': 'env-steps/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
        'note': 'rank 0 runs one batch of envs_per_gpu envs on the host cores; published reference rate for context: ~464 env-steps/s '
                '(quadrotor, 1 laptop core, README.md:188; different hardware)',
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--config', default='c3', choices=sorted(WORKLOADS))
    ap.add_argument('--envs-per-gpu', type=int, default=None)
    ap.add_argument('--hidden', type=int, default=64, choices=[32, 64, 128])
    ap.add_argument('--precision', default='f64', choices=['f64', 'f32'])
    ap.add_argument('--envs-per-block', type=int, default=0, help='developer knob: CTA shape of the collector (0 = library default)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-train', action='store_true', help='skip the train_e2e leg')
    ap.add_argument('--train-mini-batch', type=int, default=32768, help='mini_batch_size of the train_e2e leg (SURVEY.md 8d C5: scaled to the batch)')
    ap.add_argument('--train-epochs', type=int, default=10)
    ap.add_argument('--train-steps', type=int, default=3)
    args = ap.parse_args()
    if args.envs_per_gpu is None:
        args.envs_per_gpu = WORKLOADS[args.config]['envs']
    if args.impl == 'reference':
        res = run_reference(args)
        if res is not None:
            print(json.dumps(res), flush=True)
        return
    res = run_ours(args)
    if res is not None:
        if int(os.environ.get('WORLD_SIZE', '1')) == 1 and not args.no_cpu_baseline:
            ncpu, _, _ = host_threads()
            n_env = args.envs_per_gpu                  # the batch the metric is quoted on
            _, t1, _ = oracle_rate(args.config, n_env, args.hidden, 1, threads=ncpu, precision=args.precision)
            reps = max(2, min(20, int(15.0 / max(t1, 1e-3))))      # bounded sample: ~15 s of CPU work on any core count
            rate, secs, threads = oracle_rate(args.config, n_env, args.hidden, reps, threads=ncpu, precision=args.precision)
            res['cpu_baseline'] = cpu_baseline_dict(args, rate, secs, threads, reps, n_env)
        print(json.dumps(res), flush=True)


if __name__ == '__main__':
    main()
