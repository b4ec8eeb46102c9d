'''The reference PPO controller's own collector loop, recorded.  Run in the build container:  python tools/make_golden_ppo_loop.py
-> tests/golden/ppo_loop_*.npz

`PPO.train_step()` of /root/reference (controllers/ppo/ppo.py:259-303, unmodified) runs on a DummyVecEnv + VecRecordEpisodeStatistics of unmodified
reference envs (tools/ref_harness.py: PyBullet replaced by the stand-in).  `agent.update` of the INSTANCE is wrapped by a recorder that copies the
PPOBuffer it receives -- obs, act, rew, mask, v, logp, terminal_v, ret, the normalised adv -- and the actor-critic weights the rollout was collected
with; the standard normals `Normal.sample()` drew are re-made afterwards from the same torch seed (the rollout is the only consumer of torch's
generator).  tests/test_ppo_loop_golden.py runs the ORACLE's fused rollout on the same policy, normals and per-env seeds and requires the same buffer.
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


def run_case(name, env_id, task_cfg, seed, N, T, use_gae, torch_seed):
    ppo_cfg = yaml.safe_load(open(os.path.join(RH.REF, 'controllers/ppo/ppo.yaml')))
    ppo_cfg.update(rollout_batch_size=N, rollout_steps=T, mini_batch_size=32, opt_epochs=1, use_gae=use_gae, seed=seed, deque_size=1000)
    ctl = RH.make_ppo(env_id, task_cfg, tempfile.mkdtemp(), **ppo_cfg)
    ctl.reset()
    cap = {}
    orig = ctl.agent.update

    def recorder(rollouts, device='cpu'):
        cap['ac'] = {k: v.detach().numpy().copy() for k, v in ctl.agent.ac.state_dict().items()}
        cap['buf'] = {k: np.asarray(getattr(rollouts, k)).copy() for k in rollouts.scheme}
        return orig(rollouts, device)
    ctl.agent.update = recorder
    obs0 = np.asarray(ctl.obs).copy()
    torch.manual_seed(torch_seed)
    ctl.train_step()
    # what VecRecordEpisodeStatistics logged for the episodes that ended inside the rollout (record_episode_statistics.py:139-166), in completion order
    stats = dict(stat_returns=np.asarray(ctl.env.return_queue, dtype=np.float64), stat_lengths=np.asarray(ctl.env.length_queue, dtype=np.float64),
                 stat_cviol=np.asarray(ctl.env.queued_stats['constraint_violation'], dtype=np.float64))
    # deterministic evaluation with the UPDATED policy on the controller's own eval env (seeded seed * 111), as learn() calls it (ppo.py:175-178, 210-257)
    ev = ctl.run(env=ctl.eval_env, n_episodes=4)
    ev_arrs = {f'eval_{k}': np.asarray(v, dtype=np.float64) for k, v in ev.items() if k in ('ep_returns', 'ep_lengths', 'constraint_violation', 'mse')}
    ac1 = {f'ac1__{k.replace(".", "__")}': v.detach().numpy().copy() for k, v in ctl.agent.ac.state_dict().items()}
    A = cap['buf']['act'].shape[-1]
    torch.manual_seed(torch_seed)
    eps = np.stack([torch.randn(N, A).numpy() for _ in range(T)])         # Normal.sample() = normal_(0, 1) * std + mean, once per step
    path = os.path.join(OUT, f'ppo_loop_{name}.npz')
    np.savez_compressed(path, task_config=np.array(json.dumps(task_cfg)), env_id=np.array(env_id), seed=np.int32(seed), use_gae=np.int32(use_gae),
                        gamma=np.float64(ctl.gamma), gae_lambda=np.float64(ctl.gae_lambda), eps=eps, obs0=obs0, last_obs=np.asarray(ctl.obs),
                        **ev_arrs, **ac1, **stats, **{f'ac__{k.replace(".", "__")}': v for k, v in cap['ac'].items()}, **{f'buf_{k}': v for k, v in cap['buf'].items()})
    print('   eval:', {k: np.round(v, 4).tolist() for k, v in ev_arrs.items()})
    b = cap['buf']
    print(f'{name}: T={T} N={N} dones={int((1 - b["mask"]).sum())} truncated(terminal_v != 0)={int((b["terminal_v"] != 0).sum())} -> {os.path.getsize(path)} bytes')
    ctl.close()


def main():
    c3 = RH.load_task_config('examples/rl/config_overrides/quadrotor_2D/quadrotor_2D_track.yaml')
    run_case('quad2d_track', 'quadrotor', c3, seed=5, N=6, T=100, use_gae=True, torch_seed=77)
    # short episodes: time-limit truncations, i.e. critic(terminal observation) bootstraps inside the rollout (ppo.py:274-283); Monte-Carlo returns
    run_case('quad2d_track_short', 'quadrotor', dict(c3, episode_len_sec=0.5), seed=9, N=6, T=100, use_gae=False, torch_seed=78)
    cs = RH.load_task_config('examples/rl/config_overrides/cartpole/cartpole_stab.yaml')
    run_case('cartpole_stab', 'cartpole', cs, seed=2, N=6, T=100, use_gae=True, torch_seed=79)


if __name__ == '__main__':
    main()
