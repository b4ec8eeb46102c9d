'''The reference SafeExplorerPPO controller's collector, recorded.  Run in the build container:  python tools/make_golden_safe_loop.py
-> tests/golden/safe_loop_*.npz

`SafeExplorerPPO.train_step()` of /root/reference (controllers/safe_explorer/safe_ppo.py:299-355, unmodified; the safety layer of safe_explorer_utils.py
between the actor's sample and the env, fed with the constraint values of the previous step) on a DummyVecEnv of unmodified reference envs
(tools/ref_harness.py).  The "pre-trained" safety layer the controller insists on is a seeded random one with enlarged output weights, saved as a
checkpoint and loaded by the reference's own reset().  `agent.update` of the instance is wrapped by a recorder that copies the SafePPOBuffer (obs, act,
rew, mask, v, logp, terminal_v, ret, adv, c); the normals are re-made from the torch seed.
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


def run_case(name, env_id, task_cfg, seed, N, T, slack, torch_seed):
    RH.install()
    cfg = yaml.safe_load(open(os.path.join(RH.REF, 'controllers/safe_explorer/safe_ppo.yaml')))
    out_dir = tempfile.mkdtemp()
    cfg.update(rollout_batch_size=N, rollout_steps=T, mini_batch_size=32, opt_epochs=1, use_gae=True, seed=seed, pretraining=False, constraint_slack=slack,
               pretrained=os.path.join(out_dir, 'pretrained.pt'))
    # a "pre-trained" safety layer: seeded, output weights enlarged so that the projection is active in a good share of the steps
    from safe_control_gym.controllers.safe_explorer.safe_explorer_utils import SafetyLayer
    probe = RH.make_quadrotor(**task_cfg) if env_id == 'quadrotor' else RH.make_cartpole(**task_cfg)
    C = probe.constraints.num_state_constraints
    torch.manual_seed(1000 + seed)
    layer = SafetyLayer(probe.observation_space, probe.action_space, hidden_dim=cfg['constraint_hidden_dim'], num_constraints=C, lr=cfg['constraint_lr'], slack=slack)
    with torch.no_grad():
        for m in layer.constraint_models:
            m.fcs[1].weight.mul_(3.0)
    torch.save({'safety_layer': layer.state_dict()}, cfg['pretrained'])
    probe.close()
    ctl = RH.make_safe_ppo(env_id, task_cfg, out_dir, **cfg)
    ctl.reset()
    cap = {}
    orig = ctl.agent.update

    def recorder(rollouts, device='cpu'):
        cap['ac'] = {k: v.detach().numpy().copy() for k, v in ctl.agent.ac.state_dict().items()}
        cap['buf'] = {k: np.asarray(getattr(rollouts, k)).copy() for k in rollouts.scheme}
        return orig(rollouts, device)
    ctl.agent.update = recorder
    obs0, c0 = np.asarray(ctl.obs).copy(), np.asarray(ctl.c).copy()
    torch.manual_seed(torch_seed)
    ctl.train_step()
    A = cap['buf']['act'].shape[-1]
    torch.manual_seed(torch_seed)
    eps = np.stack([torch.randn(N, A).numpy() for _ in range(T)])
    sl = {f'safety__{k.replace(".", "__")}': v.detach().numpy().copy() for k, v in ctl.safety_layer.constraint_models.state_dict().items()}
    path = os.path.join(OUT, f'safe_loop_{name}.npz')
    np.savez_compressed(path, task_config=np.array(json.dumps(task_cfg)), env_id=np.array(env_id), seed=np.int32(seed), gamma=np.float64(ctl.gamma),
                        gae_lambda=np.float64(ctl.gae_lambda), eps=eps, obs0=obs0, c0=c0, last_obs=np.asarray(ctl.obs), num_constraints=np.int32(C),
                        slack=np.float64(slack), constraint_hidden_dim=np.int32(cfg['constraint_hidden_dim']), **sl,
                        **{f'ac__{k.replace(".", "__")}': v for k, v in cap['ac'].items()}, **{f'buf_{k}': v for k, v in cap['buf'].items()})
    b = cap['buf']
    print(f'{name}: T={T} N={N} C={C} dones={int((1 - b["mask"]).sum())} truncated={int((b["terminal_v"] != 0).sum())} -> {os.path.getsize(path)} bytes; '
          f'ac keys: {sorted(cap["ac"])[:3]}...')
    ctl.close()


def main():
    c3 = RH.load_task_config('examples/rl/config_overrides/quadrotor_2D/quadrotor_2D_track.yaml')
    run_case('quad2d_track', 'quadrotor', dict(c3, episode_len_sec=0.8), seed=8, N=6, T=80, slack=0.05, torch_seed=41)
    cs = RH.load_task_config('examples/rl/config_overrides/cartpole/cartpole_stab.yaml')
    run_case('cartpole_stab', 'cartpole', dict(cs, episode_len_sec=2), seed=3, N=6, T=40, slack=0.05, torch_seed=42)


if __name__ == '__main__':
    main()
