'''The reference RAP controller's collector, recorded.  Run in the build container:  python tools/make_golden_rap_loop.py -> tests/golden/rap_loop_*.npz

`RAP.collect_rollouts()` of /root/reference (controllers/rarl/rap.py:348-470, unmodified) on a DummyVecEnv of unmodified reference envs with
`adversary_disturbance` set (tools/ref_harness.py): the per-env adversary draw (numpy's global generator, recorded), the env reset, the protagonist's buffer
and the list of (adversary index, that adversary's buffer columns).  Stored with the weights of the protagonist and of every adversary of the population and
the standard normals of every `ac.step` call in call order (protagonist for all envs, then each adversary present for its group of envs, every step).
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


def run_case(name, env_id, task_cfg, seed, N, T, num_adv, torch_seed, np_seed):
    cfg = yaml.safe_load(open(os.path.join(RH.REF, 'controllers/rarl/rap.yaml')))
    cfg.update(rollout_batch_size=N, rollout_steps=T, mini_batch_size=16, opt_epochs=1, use_gae=True, seed=seed, num_adversaries=num_adv)
    ctl = RH.make_rap(env_id, task_cfg, tempfile.mkdtemp(), **cfg)
    ctl.reset()
    with torch.no_grad():
        for i, adv in enumerate(ctl.adversaries):
            adv.ac.actor.logstd.fill_(-1.0 - 0.2 * i)
    A, A2 = ctl.env.action_space.shape[0], ctl.adv_act_space.shape[0]
    np.random.seed(np_seed)
    torch.manual_seed(torch_seed)
    rollouts, splits = ctl.collect_rollouts()
    np.random.seed(np_seed)
    adv_indices = np.array(sorted(np.random.randint(num_adv, size=N)))              # rap.py:355, the same draw once more
    groups, starts = np.unique(adv_indices, return_index=True)
    ends = np.concatenate([starts[1:], [N]])
    torch.manual_seed(torch_seed)
    eps, eps_adv = np.zeros((T, N, A), np.float32), np.zeros((T, N, A2), np.float32)
    for t in range(T):
        eps[t] = torch.randn(N, A).numpy()
        for s, e in zip(starts, ends):
            eps_adv[t, s:e] = torch.randn(e - s, A2).numpy()
    arrs = {f'pro_{k}': np.asarray(getattr(rollouts, k)).copy() for k in rollouts.scheme}
    assert [int(g) for g, _ in splits] == [int(g) for g in groups]
    for g, b in splits:
        arrs.update({f'adv{int(g)}_{k}': np.asarray(getattr(b, k)).copy() for k in b.scheme})
    for who, ag in [('agent', ctl.agent)] + [(f'adversary{i}', a) for i, a in enumerate(ctl.adversaries)]:
        arrs.update({f'{who}__{k.replace(".", "__")}': v.detach().numpy().copy() for k, v in ag.ac.state_dict().items()})
    path = os.path.join(OUT, f'rap_loop_{name}.npz')
    np.savez_compressed(path, task_config=np.array(json.dumps(task_cfg)), env_id=np.array(env_id), seed=np.int32(seed), gamma=np.float64(ctl.gamma),
                        gae_lambda=np.float64(ctl.gae_lambda), num_adversaries=np.int32(num_adv), adv_indices=adv_indices, eps=eps, eps_adv=eps_adv, **arrs)
    print(f'{name}: T={T} N={N} adversaries drawn {adv_indices.tolist()} dones={int((1 - arrs["pro_mask"]).sum())} -> {os.path.getsize(path)} bytes')
    ctl.close()


def main():
    c3 = RH.load_task_config('examples/rl/config_overrides/quadrotor_2D/quadrotor_2D_track.yaml')
    run_case('quad2d_dynamics', 'quadrotor', dict(c3, episode_len_sec=0.6, adversary_disturbance='dynamics', adversary_disturbance_scale=0.03), seed=4, N=8, T=40,
             num_adv=3, torch_seed=51, np_seed=7)


if __name__ == '__main__':
    main()
