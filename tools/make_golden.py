'''Generates the committed golden fixtures under tests/golden/ from artefacts shipped inside the reference tree.

Run in the build container (needs /root/reference and torch):  python tools/make_golden.py
The GPU box has no /root/reference; tests only read the committed .npz files.

1. cartpole_pybullet_transitions.npz
   Real PyBullet-produced CartPole transitions.  examples/cbf/models/cbf_nn_ppo_model_cartpole.pt holds the
   CBFBuffer filled by safe_control_gym/safety_filters/cbf/cbf_nn.py:324-380: per episode of 250 control steps it
   stores states[1:-1] (observation AFTER each step) and inputs[1:-1] (the force applied in that step), i.e.
   state[k+1] = step(state[k], act[k+1]) inside a block of 248 rows.  Env config = examples/cbf/config_overrides/
   cartpole_config.yaml: ctrl 25 Hz / pyb 1000 Hz (40 sub-steps), no disturbances, default inertial properties,
   un-normalised action clipped to +-10 N.  Values are float32 (buffer dtype).
2. ppo_<task>.npz
   Actor-critic weights of the shipped PPO checkpoints (examples/rl/models/ppo/*.pt), the last training
   observations stored in the checkpoint (real PyBullet observations; their goal part pins X_GOAL), and the
   torch-CPU forward (mean, value) on those observations plus seeded random ones.
'''
import os

import numpy as np
import torch

REF = '/root/reference'
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests', 'golden')


def cartpole_transitions():
    d = torch.load(f'{REF}/examples/cbf/models/cbf_nn_ppo_model_cartpole.pt', weights_only=False, map_location='cpu')
    b = d['buffer']
    n = int(b['buffer_size'])
    assert n % 248 == 0
    np.savez_compressed(os.path.join(OUT, 'cartpole_pybullet_transitions.npz'),
                        state=b['state'][:n].astype(np.float32), act=b['act'][:n, 0].astype(np.float32),
                        block=np.int32(248), ctrl_freq=np.int32(25), pyb_freq=np.int32(1000))
    print('cartpole transitions:', n)


def mlp(sd, prefix, x, act):
    h = x
    for i in range(3):
        h = torch.nn.functional.linear(h, sd[f'{prefix}.fcs.{i}.weight'], sd[f'{prefix}.fcs.{i}.bias'])
        if i < 2:
            h = getattr(torch.nn.functional, act)(h)
    return h


def ppo_checkpoint(name, act):
    d = torch.load(f'{REF}/examples/rl/models/ppo/ppo_model_{name}.pt', weights_only=False, map_location='cpu')
    sd = d['agent']['ac']
    ck_obs = np.asarray(d['obs'], dtype=np.float64)
    g = torch.Generator().manual_seed(1234)
    rnd = torch.randn(64, ck_obs.shape[1], generator=g) * 0.7
    x = torch.cat([torch.as_tensor(ck_obs, dtype=torch.float32), rnd])
    with torch.no_grad():
        mean = mlp(sd, 'actor.pi_net', x, act).numpy()
        val = mlp(sd, 'critic.v_net', x, act).numpy()
    arrs = {k.replace('.', '__'): v.numpy() for k, v in sd.items()}
    np.savez_compressed(os.path.join(OUT, f'ppo_{name}.npz'), ckpt_obs=ck_obs, x=x.numpy(), mean=mean, value=val,
                        activation=np.array(act), **arrs)
    print(name, {k: tuple(v.shape) for k, v in sd.items() if 'weight' in k and 'actor' in k}, 'obs', ck_obs.shape)


if __name__ == '__main__':
    os.makedirs(OUT, exist_ok=True)
    cartpole_transitions()
    ppo_checkpoint('cartpole_stab', 'leaky_relu')
    ppo_checkpoint('quadrotor_2D_track', 'tanh')
    ppo_checkpoint('quadrotor_3D_track', 'tanh')
