'''The reference RARL controller's collector, recorded.  Run in the build container:  python tools/make_golden_rarl_loop.py
-> tests/golden/rarl_loop_*.npz

`RARL.collect_rollouts(adversary=False)` followed by `RARL.collect_rollouts(adversary=True)` of /root/reference (controllers/rarl/rarl.py:349-418,
unmodified) on a DummyVecEnv of unmodified reference envs with `adversary_disturbance` set (tools/ref_harness.py).  Stored: both PPOBuffers the method
returns (the protagonist's record, then -- from where the first rollout stopped -- the adversary's: its actions, values, log-probabilities, -rew, terminal
and bootstrap values from ITS critic), both actor-critics' weights and the standard normals the two `ac.step` calls of every step drew (re-made from the
torch seed: protagonist first, adversary second).
'''
import json
import os
import sys
import tempfile

import numpy as np
import torch
import yaml

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ref_harness as RH  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests', 'golden')


def run_case(name, env_id, task_cfg, seed, N, T, torch_seed):
    cfg = yaml.safe_load(open(os.path.join(RH.REF, 'controllers/rarl/rarl.yaml')))
    cfg.update(rollout_batch_size=N, rollout_steps=T, mini_batch_size=32, opt_epochs=1, use_gae=True, seed=seed)
    ctl = RH.make_rarl(env_id, task_cfg, tempfile.mkdtemp(), **cfg)
    ctl.reset()
    with torch.no_grad():                       # make the adversary's exploration noise differ from the protagonist's
        ctl.adversary.ac.actor.logstd.fill_(-1.0)
    obs0 = np.asarray(ctl.obs).copy()
    A, A2 = ctl.env.action_space.shape[0], ctl.adv_act_space.shape[0]
    torch.manual_seed(torch_seed)
    bufs = [ctl.collect_rollouts(adversary=False), ctl.collect_rollouts(adversary=True)]
    torch.manual_seed(torch_seed)
    eps, eps_adv = np.zeros((2, T, N, A), np.float32), np.zeros((2, T, N, A2), np.float32)
    for r in range(2):
        for t in range(T):
            eps[r, t], eps_adv[r, t] = torch.randn(N, A).numpy(), torch.randn(N, A2).numpy()
    arrs = {}
    for r, (tag, b) in enumerate(zip(('pro', 'adv'), bufs)):
        for k in b.scheme:
            arrs[f'{tag}_{k}'] = np.asarray(getattr(b, k)).copy()
    for who, ag in (('agent', ctl.agent), ('adversary', ctl.adversary)):
        arrs.update({f'{who}__{k.replace(".", "__")}': v.detach().numpy().copy() for k, v in ag.ac.state_dict().items()})
    path = os.path.join(OUT, f'rarl_loop_{name}.npz')
    np.savez_compressed(path, task_config=np.array(json.dumps(task_cfg)), env_id=np.array(env_id), seed=np.int32(seed), gamma=np.float64(ctl.gamma),
                        gae_lambda=np.float64(ctl.gae_lambda), eps=eps, eps_adv=eps_adv, obs0=obs0, last_obs=np.asarray(ctl.obs), **arrs)
    print(f'{name}: T={T} N={N} A2={A2} dones={[int((1 - arrs[t + "_mask"]).sum()) for t in ("pro", "adv")]} '
          f'truncated={[int((arrs[t + "_terminal_v"] != 0).sum()) for t in ("pro", "adv")]} -> {os.path.getsize(path)} bytes')
    ctl.close()


def main():
    c3 = RH.load_task_config('examples/rl/config_overrides/quadrotor_2D/quadrotor_2D_track.yaml')
    run_case('quad2d_dynamics', 'quadrotor', dict(c3, episode_len_sec=0.6, adversary_disturbance='dynamics', adversary_disturbance_scale=0.03), seed=4, N=6, T=60,
             torch_seed=31)
    cs = RH.load_task_config('examples/rl/config_overrides/cartpole/cartpole_stab.yaml')
    run_case('cartpole_action', 'cartpole', dict(cs, episode_len_sec=1, adversary_disturbance='action', adversary_disturbance_scale=1.0), seed=6, N=6, T=40,
             torch_seed=32)


if __name__ == '__main__':
    main()
