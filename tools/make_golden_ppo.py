'''Golden vectors for the PPO collector / update pieces, produced by the UNMODIFIED reference module
(safe_control_gym/controllers/ppo/ppo_utils.py: MLPActorCritic.step, PPOBuffer + its sampler, compute_returns_and_advantages, PPOAgent.update)
imported from /root/reference with the stand-ins of tools/ref_harness.py (gymnasium stub, package shells).

Run in the build container:  python tools/make_golden_ppo.py      -> tests/golden/ppo_reference.npz
No reference code is changed or copied.  Contents:
* gae_*: random (rews, vals, masks, terminal_vals, last_val) [T][N][1] and the reference's (ret, adv) with and without GAE;
* step_*: the reference actor-critic's sampled action, value and log-probability on a batch of observations (torch seeded);
* upd_*: a PPOBuffer filled with a synthetic rollout, the minibatch permutations `np.random.permutation` produced during `PPOAgent.update`, the
  returned loss averages and every weight afterwards -- for the plain and the clipped value loss, with a KL gate that closes part of the way through.
'''
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ref_harness as RH  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests', 'golden', 'ppo_reference.npz')
OBS_DIM, ACT_DIM, HID, T, N = 12, 2, 64, 16, 8


def flat(prefix, sd):
    return {f'{prefix}__{k.replace(".", "__")}': v.detach().numpy().copy() for k, v in sd.items()}


def main():
    RH.install()
    import safe_control_gym.controllers.ppo.ppo_utils as P
    from gymnasium.spaces import Box
    out = {}
    rng = np.random.default_rng(0)
    # ---- compute_returns_and_advantages
    Tg, Ng = 40, 6
    rews, vals = rng.normal(size=(Tg, Ng, 1)), rng.normal(size=(Tg, Ng, 1))
    masks = (rng.uniform(size=(Tg, Ng, 1)) > 0.1).astype(np.float64)
    term = np.where(rng.uniform(size=(Tg, Ng, 1)) > 0.9, rng.normal(size=(Tg, Ng, 1)), 0.0) * (1 - masks)
    last = rng.normal(size=(Ng, 1))
    out.update(gae_rews=rews, gae_vals=vals, gae_masks=masks, gae_term=term, gae_last=last)
    for tag, use_gae in (('mc', False), ('gae', True)):
        ret, adv = P.compute_returns_and_advantages(rews.copy(), vals.copy(), masks.copy(), term.copy(), last.copy(), gamma=0.99, use_gae=use_gae, gae_lambda=0.95)
        out[f'gae_ret_{tag}'], out[f'gae_adv_{tag}'] = ret, adv
    # ---- MLPActorCritic.step
    obs_space, act_space = Box(-np.ones(OBS_DIM), np.ones(OBS_DIM)), Box(-np.ones(ACT_DIM), np.ones(ACT_DIM))
    torch.manual_seed(5)
    ac = P.MLPActorCritic(obs_space, act_space, hidden_dims=[HID] * 2, activation='tanh')
    with torch.no_grad():
        ac.actor.logstd.copy_(torch.tensor([-0.7, -0.2]))
    obs = torch.as_tensor(rng.normal(size=(32, OBS_DIM)), dtype=torch.float32)
    out.update(flat('step_ac', ac.state_dict()))
    out['step_obs'] = obs.numpy()
    torch.manual_seed(6)
    with torch.no_grad():                                          # (the reference controller calls step inside no_grad: ppo.py:266)
        a, v, logp = ac.step(obs)
    out['step_act'], out['step_v'], out['step_logp'] = np.asarray(a), np.asarray(v), np.asarray(logp)
    with torch.no_grad():
        out['step_act_det'] = np.asarray(ac.act(obs))
    # ---- PPOAgent.update
    for tag, clipped in (('upd', False), ('updc', True)):
        torch.manual_seed(7)
        ag = P.PPOAgent(obs_space, act_space, hidden_dim=HID, use_clipped_value=clipped, clip_param=0.2, target_kl=0.004, entropy_coef=0.01,
                        actor_lr=3e-3, critic_lr=3e-3, opt_epochs=4, mini_batch_size=32, activation='tanh')
        buf = P.PPOBuffer(obs_space, act_space, T, N)
        data = dict(obs=rng.normal(size=(T, N, OBS_DIM)), act=rng.normal(size=(T, N, ACT_DIM)) * 0.5, rew=rng.normal(size=(T, N, 1)),
                    mask=np.ones((T, N, 1)), v=rng.normal(size=(T, N, 1)), terminal_v=np.zeros((T, N, 1)))
        with torch.no_grad():
            _, lp = ag.ac.actor(torch.as_tensor(data['obs'], dtype=torch.float32), torch.as_tensor(data['act'], dtype=torch.float32))
        data['logp'] = lp.numpy() + 0.05 * rng.normal(size=(T, N, 1))
        for t in range(T):
            buf.push({k: v[t] for k, v in data.items()})
        buf.ret = rng.normal(size=(T, N, 1)).astype(np.float32)
        adv = rng.normal(size=(T, N, 1))
        buf.adv = ((adv - adv.mean()) / (adv.std() + 1e-6)).astype(np.float32)
        out.update(flat(f'{tag}_ac0', ag.ac.state_dict()))
        for k in ('obs', 'act', 'logp', 'adv', 'ret', 'v'):
            out[f'{tag}_buf_{k}'] = np.asarray(getattr(buf, k)).copy()
        np.random.seed(123)
        res = ag.update(buf)
        np.random.seed(123)                                      # the same permutations once more, for the test to inject
        out[f'{tag}_perms'] = np.stack([np.random.permutation(np.arange(T * N)) for _ in range(4)])
        out[f'{tag}_res'] = np.array([res['policy_loss'], res['value_loss'], res['entropy_loss'], res['approx_kl']])
        out.update(flat(f'{tag}_ac1', ag.ac.state_dict()))
        print(tag, res)
    np.savez_compressed(OUT, **out)
    print(f'{len(out)} arrays -> {OUT} ({os.path.getsize(OUT)} bytes)')


if __name__ == '__main__':
    main()
