'''Golden vectors for the off-policy siblings (SAC, DDPG), produced by the UNMODIFIED reference classes
(safe_control_gym/controllers/sac/sac_utils.py: SACAgent, SACBuffer, soft_update; safe_control_gym/controllers/ddpg/ddpg_utils.py: DDPGAgent)
imported from /root/reference.

Run in the build container:  python tools/make_golden_offpolicy.py      -> tests/golden/offpolicy.npz
`gymnasium` is not installed here and `import safe_control_gym` pulls in PyBullet / CasADi: a 6-line stand-in for `gymnasium.spaces.Box` and
empty package shells for `safe_control_gym[.controllers[.sac|.ddpg]|.math_and_models]` are put into sys.modules, so that the two utility
modules (and the reference's own `neural_networks.MLP` / `distributions.Normal` they import) load from the files where they lie.  No reference
code is changed or copied.

Per agent: the initial weights, a batch (obs, act, rew, next_obs, mask), the reference's losses on it (torch seeded before every call that
samples), the deterministic action, and every weight after ONE `update(batch)` (actor, critics, targets, temperature).  For the buffer: the
arrays after a sequence of pushes that wraps around.
'''
import os
import sys
import types

import numpy as np
import torch

REF = '/root/reference/safe_control_gym'
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests', 'golden', 'offpolicy.npz')
OBS_DIM, HID, B = 6, 32, 48
ACT_LOW, ACT_HIGH = np.array([0.05, 0.05], np.float32), np.array([0.3, 0.3], np.float32)


def load_reference():
    g, sp = types.ModuleType('gymnasium'), types.ModuleType('gymnasium.spaces')

    class Box:
        def __init__(self, low, high, shape=None, dtype=np.float32):
            self.low = np.asarray(low, dtype) if shape is None else np.broadcast_to(np.asarray(low, dtype), shape).copy()
            self.high = np.asarray(high, dtype) if shape is None else np.broadcast_to(np.asarray(high, dtype), shape).copy()
            self.shape = self.low.shape
    sp.Box, g.spaces, g.Env, g.Wrapper = Box, sp, object, object
    sys.modules['gymnasium'], sys.modules['gymnasium.spaces'] = g, sp
    for name, sub in (('safe_control_gym', ''), ('safe_control_gym.math_and_models', 'math_and_models'), ('safe_control_gym.controllers', 'controllers'),
                      ('safe_control_gym.controllers.sac', 'controllers/sac'), ('safe_control_gym.controllers.ddpg', 'controllers/ddpg')):
        m = types.ModuleType(name)
        m.__path__ = [os.path.join(REF, sub)]
        sys.modules[name] = m
    import safe_control_gym.controllers.ddpg.ddpg_utils as D
    import safe_control_gym.controllers.sac.sac_utils as S
    return S, D, Box


def flat(prefix, sd):
    return {f'{prefix}__{k.replace(".", "__")}': v.detach().numpy().copy() for k, v in sd.items()}


def make_batch(seed):
    gen = torch.Generator().manual_seed(seed)
    obs = torch.randn(B, OBS_DIM, generator=gen)
    act = torch.as_tensor(ACT_LOW) + torch.rand(B, 2, generator=gen) * torch.as_tensor(ACT_HIGH - ACT_LOW)
    rew = torch.rand(B, 1, generator=gen)
    next_obs = obs + 0.1 * torch.randn(B, OBS_DIM, generator=gen)
    mask = (torch.rand(B, 1, generator=gen) > 0.2).float()
    return {'obs': obs, 'act': act, 'rew': rew, 'next_obs': next_obs, 'mask': mask}


def sac_case(S, Box, tag, tuning, out):
    torch.manual_seed(11 if tuning else 10)
    ag = S.SACAgent(Box(-np.ones(OBS_DIM), np.ones(OBS_DIM)), Box(ACT_LOW, ACT_HIGH), hidden_dim=HID, use_entropy_tuning=tuning,
                    init_temperature=0.2, actor_lr=3e-3, critic_lr=2e-3, entropy_lr=5e-3)
    with torch.no_grad():                                           # targets that differ from the online nets, so that the Polyak step is visible
        for p in ag.ac_targ.parameters():
            p.mul_(0.9)
    batch = make_batch(21)
    out.update(flat(f'{tag}_ac0', ag.ac.state_dict())); out.update(flat(f'{tag}_targ0', ag.ac_targ.state_dict()))
    out.update({f'{tag}_batch_{k}': v.numpy() for k, v in batch.items()})
    ag.train()
    torch.manual_seed(100)
    pl, el = ag.compute_policy_loss(batch)
    torch.manual_seed(101)
    ql = ag.compute_q_loss(batch)
    out[f'{tag}_losses0'] = np.array([float(pl), float(el), float(ql)])
    with torch.no_grad():
        out[f'{tag}_act_det'] = ag.ac.actor(batch['obs'], deterministic=True, with_logprob=False)[0].numpy()
        torch.manual_seed(102)
        a, lp = ag.ac.actor(batch['obs'], deterministic=False, with_logprob=True)
        out[f'{tag}_act_sample'], out[f'{tag}_logp_sample'] = a.numpy(), lp.numpy()
    torch.manual_seed(103)
    res = ag.update(batch)
    out[f'{tag}_update_res'] = np.array([res['policy_loss'], res['critic_loss'], res['entropy_loss']])
    out.update(flat(f'{tag}_ac1', ag.ac.state_dict())); out.update(flat(f'{tag}_targ1', ag.ac_targ.state_dict()))
    out[f'{tag}_log_alpha1'] = np.array(float(ag.log_alpha))
    torch.manual_seed(104)
    res = ag.update(batch)                                           # a second step: Adam moments carried over
    out[f'{tag}_update_res2'] = np.array([res['policy_loss'], res['critic_loss'], res['entropy_loss']])
    out.update(flat(f'{tag}_ac2', ag.ac.state_dict()))
    out[f'{tag}_log_alpha2'] = np.array(float(ag.log_alpha))


def ddpg_case(D, Box, out):
    torch.manual_seed(12)
    ag = D.DDPGAgent(Box(-np.ones(OBS_DIM), np.ones(OBS_DIM)), Box(ACT_LOW, ACT_HIGH), hidden_dim=HID, actor_lr=3e-3, critic_lr=2e-3)
    with torch.no_grad():
        for p in ag.ac_targ.parameters():
            p.mul_(0.9)
    batch = make_batch(22)
    out.update(flat('ddpg_ac0', ag.ac.state_dict())); out.update(flat('ddpg_targ0', ag.ac_targ.state_dict()))
    out.update({f'ddpg_batch_{k}': v.numpy() for k, v in batch.items()})
    ag.train()
    out['ddpg_losses0'] = np.array([float(ag.compute_policy_loss(batch)), float(ag.compute_q_loss(batch))])
    with torch.no_grad():
        out['ddpg_act'] = ag.ac.actor(batch['obs']).numpy()
    for i in (1, 2):
        res = ag.update(batch)
        out[f'ddpg_update_res{i}'] = np.array([res['policy_loss'], res['critic_loss']])
        out.update(flat(f'ddpg_ac{i}', ag.ac.state_dict())); out.update(flat(f'ddpg_targ{i}', ag.ac_targ.state_dict()))


def buffer_case(S, Box, out):
    buf = S.SACBuffer(Box(-np.ones(OBS_DIM), np.ones(OBS_DIM)), Box(ACT_LOW, ACT_HIGH), max_size=10, batch_size=4)
    rng = np.random.default_rng(3)
    sizes = (4, 4, 4, 3, 7)
    pushes = []
    for n in sizes:
        b = {'obs': rng.normal(size=(n, OBS_DIM)), 'next_obs': rng.normal(size=(n, OBS_DIM)), 'act': rng.normal(size=(n, 2)), 'rew': rng.normal(size=(n, 1)),
             'mask': rng.integers(0, 2, size=(n, 1)).astype(np.float64)}
        buf.push(b)
        pushes.append(b)
    out['buf_sizes'] = np.array(sizes)
    for k in ('obs', 'next_obs', 'act', 'rew', 'mask'):
        out[f'buf_push_{k}'] = np.concatenate([p[k] for p in pushes])
        out[f'buf_final_{k}'] = np.asarray(getattr(buf, k)).copy()
    out['buf_pos_size'] = np.array([buf.pos, buf.buffer_size])


def main():
    S, D, Box = load_reference()
    out = {}
    sac_case(S, Box, 'sac', False, out)
    sac_case(S, Box, 'sact', True, out)
    ddpg_case(D, Box, out)
    buffer_case(S, Box, out)
    np.savez_compressed(OUT, **out)
    print(f'{len(out)} arrays -> {OUT} ({os.path.getsize(OUT)} bytes)')
    for k in ('sac_losses0', 'sact_losses0', 'sact_update_res', 'ddpg_losses0', 'buf_pos_size'):
        print(k, out[k])


if __name__ == '__main__':
    main()
