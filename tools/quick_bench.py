'''Developer timing script (GPU box): times the fused rollout for a few configurations with CUDA events.'''
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests'))
import numpy as np
import torch
import helpers as H
from safe_control_gym_b200.engine import EngineHandle


def run(name, N, T, hid, prec, epb, iters=5, **kw):
    env = H.make_env(name, **kw)
    eng = EngineHandle(env, num_envs=N, auto_reset=True, rng_mode='philox', seed=1, precision=prec)
    eng.reset()
    pol = H.random_policy(env.obs_dim, env.action_dim, hid, seed=0)
    blob, dims = H.policy_blob(pol, eng.device)
    out = None
    ts = []
    for i in range(iters + 2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = eng.rollout(blob, dims, T, eps=None, use_gae=True, envs_per_block=epb, out=out)
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = float(np.median(ts[2:]))
    done_frac = 1.0 - float(out['mask'].mean())
    print(f'{name:22s} N={N:6d} T={T} hid={hid:3d} {prec} E={epb:2d}: {ms:8.3f} ms  {N*T/ms/1e3:9.2f} M env-steps/s  done_frac={done_frac:.4f}', flush=True)
    eng.close()


if __name__ == '__main__':
    T = 200
    for prec in ('f64', 'f32'):
        for epb in (32, 16, 8):
            run('quadrotor_2D_track', 4096, T, 64, prec, epb)
    for epb in (32, 16, 8):
        run('quadrotor_2D_track', 32768, T, 64, 'f64', epb)
    run('cartpole', 4096, T, 64, 'f64', 32)
    run('cartpole', 4096, T, 64, 'f64', 16)
    run('cartpole', 4096, T, 64, 'f64', 8)
    run('cartpole_stab', 4096, T, 64, 'f64', 16)
    for epb in (32, 16, 8):
        run('quadrotor_3D_track', 8192, T, 64, 'f64', epb)
    run('quadrotor_3D_track', 8192, T, 64, 'f32', 16)
    run('quadrotor_2D_track', 4096, T, 128, 'f64', 16)
    run('quadrotor_2D_track', 4096, T, 128, 'f64', 8)
