'''The reference SAC controller's collector (the DDPG one is the same code), recorded.  Run in the build container:
python tools/make_golden_offpolicy_loop.py -> tests/golden/offpolicy_loop_*.npz

`SAC.train_step()` of /root/reference (controllers/sac/sac.py:269-333, unmodified) M times on a DummyVecEnv of unmodified reference envs
(tools/ref_harness.py), inside the warm-up (uniform random actions, no gradient step).  Stored: the replay buffer rows it pushed -- obs, act, rew, next_obs,
mask in push order -- i.e. the reference's own answer to "what is the next observation and the mask of a transition that ended an episode" (the reset
observation and 0 when the episode truly terminated, the TERMINAL observation and 1 when the time limit truncated it).
'''
import json
import os
import sys
import tempfile

import numpy as np
import yaml

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ref_harness as RH  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests', 'golden')


def run_case(name, env_id, task_cfg, seed, N, M):
    cfg = yaml.safe_load(open(os.path.join(RH.REF, 'controllers/sac/sac.yaml')))
    cfg.update(rollout_batch_size=N, seed=seed, warm_up_steps=10 ** 9, max_buffer_size=N * M)
    ctl = RH.make_sac(env_id, task_cfg, tempfile.mkdtemp(), **cfg)
    ctl.reset()
    obs0 = np.asarray(ctl.obs).copy()
    for _ in range(M):
        ctl.train_step()
    b = ctl.buffer
    arrs = {k: np.asarray(getattr(b, k))[:N * M].reshape(M, N, -1).copy() for k in ('obs', 'act', 'rew', 'next_obs', 'mask')}
    path = os.path.join(OUT, f'offpolicy_loop_{name}.npz')
    np.savez_compressed(path, task_config=np.array(json.dumps(task_cfg)), env_id=np.array(env_id), seed=np.int32(seed), obs0=obs0, **arrs)
    print(f'{name}: M={M} N={N} mask==0: {int((arrs["mask"] == 0).sum())} -> {os.path.getsize(path)} bytes')
    ctl.close()


def main():
    c3 = RH.load_task_config('examples/rl/config_overrides/quadrotor_2D/quadrotor_2D_track.yaml')
    run_case('quad2d_track', 'quadrotor', dict(c3, episode_len_sec=0.5), seed=12, N=6, M=80)
    cs = RH.load_task_config('examples/rl/config_overrides/cartpole/cartpole_stab.yaml')
    run_case('cartpole_stab', 'cartpole', dict(cs, episode_len_sec=1), seed=13, N=6, M=60)


if __name__ == '__main__':
    main()
