'''Runs ONE configuration of the fused rollout a few times (for ncu).  usage: profile_one.py <config> <N> <hid> <prec> <E> [iters]'''
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests'))
import torch
import helpers as H
from safe_control_gym_b200.engine import EngineHandle

name, N, hid, prec, epb = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4], int(sys.argv[5])
iters = int(sys.argv[6]) if len(sys.argv) > 6 else 3
env = H.make_env(name)
eng = EngineHandle(env, num_envs=N, auto_reset=True, rng_mode='philox', seed=1, precision=prec)
eng.reset()
pol = H.random_policy(env.obs_dim, env.action_dim, hid, seed=0)
blob, dims = H.policy_blob(pol, eng.device)
out = None
for i in range(iters):
    out = eng.rollout(blob, dims, 200, eps=None, use_gae=True, envs_per_block=epb, out=out)
torch.cuda.synchronize()
print('done', float(out['rew'].mean()))
