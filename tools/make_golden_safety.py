'''Golden vectors for the SafeExplorer safety layer, produced by the UNMODIFIED reference class
(safe_control_gym/controllers/safe_explorer/safe_explorer_utils.py: SafetyLayer) imported from /root/reference.

Run in the build container:  python tools/make_golden_safety.py      -> tests/golden/safety_layer.npz
`gymnasium` is not installed here; the reference module only needs `gymnasium.spaces.Box` for isinstance-free shape look-ups, so a
4-line stand-in is put into sys.modules before the import (no reference code is changed or copied).

Contents: the constraint models' initial weights, a batch of (obs, act, c, c_next), the reference's `get_safe_action` per sample
(it is written for a batch of 1: evaluated row by row), `compute_loss` on the batch, and the weights after ONE `update`.
'''
import importlib.util
import os
import sys
import types

import numpy as np
import torch

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests', 'golden', 'safety_layer.npz')


def load_reference():
    g, sp = types.ModuleType('gymnasium'), types.ModuleType('gymnasium.spaces')

    class Box:
        def __init__(self, low, high, shape=None, dtype=np.float32):
            self.low = np.broadcast_to(np.asarray(low, dtype), shape).copy()
            self.high = np.broadcast_to(np.asarray(high, dtype), shape).copy()
            self.shape = self.low.shape
    sp.Box, g.spaces, g.Env, g.Wrapper = Box, sp, object, object
    sys.modules['gymnasium'], sys.modules['gymnasium.spaces'] = g, sp
    sys.path.insert(0, '/root/reference')
    spec = importlib.util.spec_from_file_location('ref_safe_explorer_utils',
                                                  '/root/reference/safe_control_gym/controllers/safe_explorer/safe_explorer_utils.py')
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m, Box


def main():
    ref, Box = load_reference()
    OBS, ACT, C, H, B = 12, 2, 5, 10, 96
    torch.manual_seed(2024)
    layer = ref.SafetyLayer(Box(-1, 1, (OBS,)), Box(-1, 1, (ACT,)), hidden_dim=H, num_constraints=C, lr=1e-3, slack=[0.3, 0.1, 0.2, 0.05, 0.4])
    with torch.no_grad():
        for m in layer.constraint_models:
            m.fcs[1].weight.mul_(3.0)
    w0 = {f'w0__{k.replace(".", "__")}': v.detach().numpy().copy() for k, v in layer.constraint_models.state_dict().items()}
    gen = torch.Generator().manual_seed(7)
    obs = torch.randn(B, OBS, generator=gen)
    act = torch.randn(B, ACT, generator=gen) * 0.6
    c = -torch.rand(B, C, generator=gen) * 0.8 + 0.1                 # mostly satisfied (negative), some violated
    c_next = c + 0.05 * torch.randn(B, C, generator=gen)
    layer.eval()
    with torch.no_grad():
        safe = torch.cat([layer.get_safe_action(obs[i:i + 1], act[i:i + 1], c[i:i + 1]) for i in range(B)])
    batch = {'obs': obs, 'act': act, 'c': c, 'c_next': c_next}
    layer.train()
    losses = np.array([float(l) for l in layer.compute_loss(batch)])
    res = layer.update(batch)
    w1 = {f'w1__{k.replace(".", "__")}': v.detach().numpy().copy() for k, v in layer.constraint_models.state_dict().items()}
    np.savez_compressed(OUT, obs=obs.numpy(), act=act.numpy(), c=c.numpy(), c_next=c_next.numpy(), safe=safe.numpy(), losses=losses,
                        update_losses=np.array([res[f'constraint_{i}_loss'] for i in range(C)]), slack=np.asarray(layer.slack, dtype=np.float64),
                        dims=np.array([OBS, ACT, C, H]), **w0, **w1)
    changed = float((safe - act).abs().max(-1).values.gt(1e-7).float().mean())
    print('wrote', OUT, 'projection active in', changed, 'of the samples; losses', losses)


if __name__ == '__main__':
    main()
