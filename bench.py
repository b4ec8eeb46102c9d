#!/usr/bin/env python
'''Benchmark of the hot path: env-steps/s of the fused PPO rollout collector (policy forward + env.step + PPOBuffer write).

    python bench.py --gpus N --steps K --warmup W [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

One "step" = one pass of the hot path over one batch = ONE fused rollout launch: `envs_per_gpu` Quadrotor-2D trajectory-tracking
envs (examples/rl/config_overrides/quadrotor_2D/quadrotor_2D_track.yaml), 64-64 tanh actor-critic, horizon T = 200 control
steps (each 20 physics sub-steps), float64 env arithmetic + float32 policy exactly like the reference.  BASELINE.json metric:
env-steps/sec (4096 envs, 200-step horizon).

Printed JSON (rank 0, one line): `value` = device-resident throughput (inputs in HBM), `e2e` = the same metric through the
reference-facing collector call with HOST buffers (policy weights H2D + full PPOBuffer record D2H inside the timed region),
`roofline` (HBM bound, algorithmic bytes of SURVEY.md 8d), `cpu_baseline` (the C oracle port on the host cores).
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
HIDDEN = 64
WORKLOAD = 'quadrotor_2D_track'
BYTES_PER_ENV_STEP = 80            # 4 * (obs_dim 12 + act_dim 2 + 6) -- obs, act, rew, mask, v, logp, terminal_v, cviol


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


def build_policy(env, seed=0):
    import torch

    from safe_control_gym_b200.controllers.ppo.ppo_utils import MLPActorCritic
    torch.manual_seed(seed)
    return MLPActorCritic(env.observation_space, env.action_space, hidden_dims=[HIDDEN, HIDDEN], activation='tanh')


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
    N = args.envs_per_gpu
    env = make('quadrotor', **configs.get(WORKLOAD))
    eng = EngineHandle(env, num_envs=N, auto_reset=True, rng_mode='philox', seed=1337, env_index_offset=rank * N, device=local_rank,
                       precision=args.precision)
    eng.reset()
    ac = build_policy(env).to(dev)
    blob, dims = pack_policy(ac, dev)
    want = ('ret', 'adv', 'cviol', 'last', 'episode')
    # N > 1: the rollout record is all-gathered INSIDE the collector kernel (peer stores over NVLink into every rank's gathered
    # buffer, torch symmetric memory); the per-step collective is one barrier.  Falls back to the NCCL all-gather if P2P is unavailable.
    pg = PeerGatheredRollout(rollout_flat_numel(dims[0], dims[1], T_HORIZON, N), dev) if world > 1 else None
    out = eng.rollout(blob, dims, T_HORIZON, use_gae=True, want=want, **(pg.rollout_kwargs() if pg else {}))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)          # > 126 MB L2

    def one_step():
        o = eng.rollout(blob, dims, T_HORIZON, use_gae=True, want=want, out=out)
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
        a.record()
        o = eng.rollout(blob, dims, T_HORIZON, use_gae=True, want=want, out=out)
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
    # memory (host pack -> one H2D) and the full PPOBuffer record goes back to pinned host memory (one D2H).  Two records in flight:
    # the D2H of rollout k overlaps rollout k+1.  `e2e_serial` is the same with one record in flight (no overlap).
    from safe_control_gym_b200.engine import HostCollector
    ac_host = build_policy(env)

    def e2e_rate(depth, n_steps):
        hc = HostCollector(eng, dims, T_HORIZON, depth=depth, use_gae=True, want=want)
        acc = 0.0
        for _ in range(depth + 1):                     # warm-up (allocates the pinned buffers)
            hc.submit(ac_host)
            acc += float(hc.fetch()['rew'][0, 0, 0])
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for i in range(n_steps):
            hc.submit(ac_host)
            if i >= depth - 1:
                acc += float(hc.fetch()['rew'][-1, -1, 0])      # the host consumes every record
        while hc.fetched < hc.submitted:
            acc += float(hc.fetch()['rew'][-1, -1, 0])
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        return world * N * T_HORIZON * n_steps / float(dt), hc.h2d_bytes, hc.d2h_bytes, acc

    e2e_steps = max(8, min(args.steps, 32))
    e2e_serial, h2d_bytes, d2h_bytes, _ = e2e_rate(1, e2e_steps)
    e2e_value, _, _, _ = e2e_rate(2, e2e_steps)
    # the timed regions above last milliseconds: keep the same kernel running for ~2 s so that nvidia-smi (100 ms period)
    # records the clocks and throttle reasons UNDER THIS LOAD
    t_end = time.perf_counter() + 2.0
    while time.perf_counter() < t_end:
        for _ in range(20):
            one_step()
        torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None

    res = None
    if rank == 0:
        peak, peak_src = measured_peak_hbm()
        kern_s = float(np.mean(kern_ms)) * 1e-3
        achieved = BYTES_PER_ENV_STEP * N * T_HORIZON / kern_s / 1e9
        flops = 2 * ((12 * 64 + 64 * 64 + 64 * 2) + (12 * 64 + 64 * 64 + 64)) * N * T_HORIZON / kern_s / 1e12
        prof = os.path.join(ROOT, 'profiles', 'ncu_summary.json')
        traffic = json.load(open(prof)).get('dram_bytes_per_launch') if os.path.exists(prof) else None
        res = {
            'metric': 'env-steps/sec (4096 envs, 200-step horizon)', 'value': value, 'unit': 'env-steps/s', 'n_gpus': world,
            'steps': args.steps, 'warmup': max(args.warmup, 3), 'ms_per_step': total_ms / args.steps, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64 env state + f32 policy (as the reference)' if args.precision == 'f64' else 'f32',
            'data': 'synthetic (random-init 64-64 actor-critic, Philox-seeded initial states)',
            'config': {'workload': 'Quadrotor-2D trajectory tracking (quadrotor_2D_track.yaml), PPO 64-64 tanh actor-critic, fused rollout',
                       'envs_per_gpu': N, 'horizon': T_HORIZON, 'physics_substeps_per_step': 20, 'one_bench_step': 'one fused rollout launch',
                       'parallelism': f'env-sharded x{world}' + ('' if world == 1 else (
                           ', rollout record all-gathered inside the collector kernel (NVLink peer stores, torch symmetric memory) + one barrier per step'
                           if pg.fused else f', one NCCL all-gather of the rollout record per step (peer memory unavailable: {pg.why})')),
                       'l2': 'flushed with a 256 MiB write before every timed step; per-step CUDA-event intervals summed'},
            'gpu_launches': int(launches),
            'e2e': {'value': e2e_value, 'unit': 'env-steps/s', 'h2d_bytes_per_step': int(h2d_bytes), 'd2h_bytes_per_step': int(d2h_bytes),
                    'serial_value': e2e_serial, 'steps': e2e_steps,
                    'what': 'engine.HostCollector: host policy parameters -> scg_policy_pack -> H2D, fused rollout, full PPOBuffer record '
                            '(obs, act, rew, mask, v, logp, terminal_v, ret, adv) -> pinned host, consumed by the host every step; 2 records in flight '
                            '(D2H of rollout k overlaps rollout k+1); serial_value = 1 record in flight'},
            'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak, 'traffic': traffic,
                         'peak_source': peak_src, 'kernel': 'rollout_kernel<QUAD2D, double, 64, 32, 256>', 'kernel_ms': kern_s * 1e3,
                         'algorithmic_bytes_per_env_step': BYTES_PER_ENV_STEP,
                         'note': 'the path is latency / FP32-issue bound, not HBM bound (SURVEY.md 8d): see fp32 below',
                         'fp32': {'achieved_tflops': flops, 'peak_tflops': 74.4, 'frac': flops / 74.4,
                                  'peak_source': 'computed: 148 SMs x 128 lanes x 2 x 1.965 GHz'}},
            'clocks': clocks,
            'wall_s_timed_region': t_wall,
        }
    eng.close()
    if world > 1:
        dist.destroy_process_group()
    return res


def oracle_rate(n_envs, steps, threads=None):
    '''env-steps/s of the CPU oracle port on `n_envs` envs x T_HORIZON steps, `steps` times (bounded sample of the same workload).'''
    import helpers as H
    from oracle import oracle as orc
    orc.build()
    if threads:
        orc.set_threads(threads)
    env = H.make_env(WORKLOAD)
    o = H.make_oracle(env, n_envs, auto_reset=True, rng_mode='philox', seed=1337)
    o.reset()
    pol = H.random_policy(env.obs_dim, env.action_dim, HIDDEN, seed=0)
    o.rollout(pol, 10, use_gae=True)
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        o.rollout(pol, T_HORIZON, use_gae=True)
        ts.append(time.perf_counter() - t0)
    return n_envs * T_HORIZON * len(ts) / sum(ts), sum(ts), orc.num_threads()


def run_reference(args):
    '''The reference's own CPU algorithm for the path (PyBullet + torch are not installable here: the C oracle port is timed,
    with all the host threads OpenMP gives it), same config / metric / unit.'''
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return None
    n = args.ref_envs
    # torchrun exports OMP_NUM_THREADS=1: ask for every physical core explicitly (SMT siblings only slow the FP64 loops down)
    try:
        import psutil
        ncpu = psutil.cpu_count(logical=False) or os.cpu_count() or 1
    except Exception:
        ncpu = os.cpu_count() or 1
    ncpu = max(1, min(ncpu, len(os.sched_getaffinity(0))))
    for _ in range(args.warmup):
        oracle_rate(n, 1, threads=ncpu)
    rate, secs, threads = oracle_rate(n, args.steps, threads=ncpu)
    return {
        'impl': 'reference', 'metric': 'env-steps/sec (4096 envs, 200-step horizon)', 'value': rate, 'unit': 'env-steps/s',
        'n_gpus': int(os.environ.get('WORLD_SIZE', '1')), 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': secs / args.steps * 1e3,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64 env state + f32 policy', 'data': 'synthetic',
        'config': {'workload': 'Quadrotor-2D trajectory tracking (quadrotor_2D_track.yaml), PPO 64-64 tanh actor-critic, rollout collector',
                   'horizon': T_HORIZON, 'sample': f'{n} envs x {T_HORIZON} steps per bench step'},
        'cpu_baseline': {'value': rate, 'unit': 'env-steps/s', 'cores': threads, 'kind': 'port',
                         'sample': f'{n} envs x {T_HORIZON} control steps x {args.steps} repeats, C oracle (oracle/scg_oracle.c), OpenMP over envs'},
        'e2e': {'value': rate, 'unit': 'env-steps/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
        'note': 'published reference rate for context: ~464 env-steps/s (quadrotor, 1 laptop core, README.md:188; different hardware)',
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--envs-per-gpu', type=int, default=4096)
    ap.add_argument('--precision', default='f64', choices=['f64', 'f32'])
    ap.add_argument('--ref-envs', type=int, default=4096)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    args = ap.parse_args()
    if args.impl == 'reference':
        res = run_reference(args)
        if res is not None:
            print(json.dumps(res), flush=True)
        return
    res = run_ours(args)
    if res is not None:
        if int(os.environ.get('WORLD_SIZE', '1')) == 1 and not args.no_cpu_baseline:
            ncpu = os.cpu_count() or 8
            n_env, reps = 32 * ncpu, 20            # bounded sample, ~10-20 s of CPU work on any core count
            rate, secs, threads = oracle_rate(n_env, reps)
            res['cpu_baseline'] = {'value': rate, 'unit': 'env-steps/s', 'cores': threads, 'kind': 'port',
                                   'sample': f'{n_env} envs x {T_HORIZON} control steps x {reps} repeats ({secs:.1f} s), C oracle port of the reference path, OpenMP over envs'}
        print(json.dumps(res), flush=True)


if __name__ == '__main__':
    main()
