'''Times the fused PPO update kernel against the torch restatement on a record of C3's shape.
usage: bench_update.py [rows] [mini_batch] [epochs] [torch_minibatches]'''
import os
import sys
import time

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import torch

from safe_control_gym_b200.controllers.ppo.ppo_utils import PPOAgent
from safe_control_gym_b200.spaces import Box

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 4096 * 200
mb = int(sys.argv[2]) if len(sys.argv) > 2 else 64
epochs = int(sys.argv[3]) if len(sys.argv) > 3 else 1
torch_mbs = int(sys.argv[4]) if len(sys.argv) > 4 else 200
obs_dim, act_dim, hid = 12, 2, 64
osp, asp = Box(-np.ones(obs_dim), np.ones(obs_dim)), Box(-np.ones(act_dim), np.ones(act_dim))
dev = torch.device('cuda:0')
g = torch.Generator(device='cuda').manual_seed(0)


def batch_for(agent):
    b = {'obs': torch.randn(1, rows, obs_dim, device=dev, generator=g)}
    with torch.no_grad():
        dist, _ = agent.ac.actor(b['obs'])
        b['act'] = dist.mean + dist.stddev * torch.randn(1, rows, act_dim, device=dev, generator=g)
        b['logp'] = dist.log_prob(b['act']) + 0.02 * torch.randn(1, rows, 1, device=dev, generator=g)
        b['adv'] = torch.randn(1, rows, 1, device=dev, generator=g)
        b['v'] = agent.ac.critic(b['obs'])
        b['ret'] = b['v'] + 0.5 * torch.randn(1, rows, 1, device=dev, generator=g)
    return b


torch.manual_seed(0)
a = PPOAgent(osp, asp, hidden_dim=hid, mini_batch_size=mb, opt_epochs=epochs, fused_update=True)
a.to(dev)
batch = batch_for(a)
for _ in range(2):
    t0 = time.perf_counter()
    r = a.update(batch, generator=torch.Generator(device='cuda').manual_seed(1))
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
n = r['minibatches']
print(f'fused : rows={rows} mb={mb} epochs={epochs}: {dt * 1e3:9.2f} ms total, {dt / n * 1e6:8.2f} us / minibatch, {rows * epochs / dt / 1e6:8.2f} M samples/s  {r}')

torch.manual_seed(0)
b = PPOAgent(osp, asp, hidden_dim=hid, mini_batch_size=mb, opt_epochs=1, fused_update=False)
b.to(dev)
sub = {k: v[:, :torch_mbs * mb] for k, v in batch.items()}
for _ in range(2):
    t0 = time.perf_counter()
    b.update(sub, generator=torch.Generator(device='cuda').manual_seed(1))
    torch.cuda.synchronize()
    dt2 = time.perf_counter() - t0
print(f'torch : {torch_mbs} minibatches of {mb}: {dt2 / torch_mbs * 1e6:8.2f} us / minibatch  -> fused speed-up per minibatch {dt2 / torch_mbs / (dt / n):.1f}x')
