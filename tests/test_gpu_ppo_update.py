'''Fused PPO update kernel (csrc/scg_ppo.cu, scg_ppo_update) against the step-by-step torch restatement of the reference's
PPOAgent.update (safe_control_gym/controllers/ppo/ppo_utils.py:82-146).  Floating-point kernel: the parity reference is torch fp32
on the same device; tolerances are written next to each check.'''
import copy

import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu
torch = pytest.importorskip('torch')


def _spaces(obs_dim, act_dim):
    from safe_control_gym_b200.spaces import Box
    return Box(-np.ones(obs_dim), np.ones(obs_dim)), Box(-np.ones(act_dim), np.ones(act_dim))


def _agents(obs_dim, act_dim, hid, seed=0, **kw):
    '''Two agents with identical initial weights: fused kernel vs torch restatement.'''
    from safe_control_gym_b200.controllers.ppo.ppo_utils import PPOAgent
    osp, asp = _spaces(obs_dim, act_dim)
    torch.manual_seed(seed)
    a = PPOAgent(osp, asp, hidden_dim=hid, fused_update=True, **kw)
    b = PPOAgent(osp, asp, hidden_dim=hid, fused_update=False, **kw)
    b.ac.load_state_dict(copy.deepcopy(a.ac.state_dict()))
    a.to('cuda:0')
    b.to('cuda:0')
    assert a.fused is not None and b.fused is None
    return a, b


def _synthetic_batch(agent, rows, blocks=1, seed=1, logp_noise=0.05, logp_bias=0.0):
    '''A record the actor could have produced: actions sampled from the current policy, old log-probs perturbed so that ratios
    straddle the clip range, random advantages / returns.  Shape [blocks][rows/blocks][w]; with blocks > 1 the tensors are strided
    views into one buffer per block (the layout of the all-gathered multi-GPU record).'''
    g = torch.Generator(device='cuda').manual_seed(seed)
    dev = torch.device('cuda:0')
    obs_dim, act_dim = agent.obs_space.shape[0], agent.act_space.shape[0]
    R = rows // blocks
    widths = [('obs', obs_dim), ('act', act_dim), ('logp', 1), ('adv', 1), ('ret', 1), ('v', 1)]
    per_block = sum(w for _, w in widths) * R + 64          # + slack: the block stride is not the payload size
    flat = torch.zeros(blocks, per_block, device=dev)
    batch, o = {}, 0
    for k, w in widths:
        batch[k] = flat[:, o:o + R * w].view(blocks, R, w)
        o += R * w
    with torch.no_grad():
        batch['obs'].copy_(torch.randn(blocks, R, obs_dim, device=dev, generator=g))
        dist, _ = agent.ac.actor(batch['obs'])
        act = dist.mean + dist.stddev * torch.randn(blocks, R, act_dim, device=dev, generator=g)
        batch['act'].copy_(act)
        batch['logp'].copy_(dist.log_prob(act) + logp_bias + logp_noise * torch.randn(blocks, R, 1, device=dev, generator=g))
        batch['adv'].copy_(torch.randn(blocks, R, 1, device=dev, generator=g))
        v = agent.ac.critic(batch['obs'])
        batch['v'].copy_(v + 0.1 * torch.randn(blocks, R, 1, device=dev, generator=g))
        batch['ret'].copy_(v + 0.5 * torch.randn(blocks, R, 1, device=dev, generator=g))
    return batch


def _flat_params(agent):
    from safe_control_gym_b200.engine import TRAIN_SECTIONS
    named = dict(agent.ac.named_parameters())
    return torch.cat([named[n].detach().reshape(-1) for n in TRAIN_SECTIONS])


@pytest.mark.parametrize('obs_dim,act_dim,hid,activation,mb', [(12, 2, 64, 'tanh', 64), (24, 4, 64, 'tanh', 200), (4, 1, 32, 'leaky_relu', 100),
                                                               (6, 2, 64, 'relu', 1024)])
def test_minibatch_gradient_matches_autograd(obs_dim, act_dim, hid, activation, mb):
    '''d(policy_loss + entropy_coef * entropy_loss)/d(actor) and d(value_loss)/d(critic) of ONE minibatch (incl. ragged last chunk and
    the multi-CTA reduction) vs torch autograd: |err| <= 1e-6 + 2e-4 * |ref| elementwise on the flattened gradient.'''
    from safe_control_gym_b200.engine import TRAIN_SECTIONS
    a, b = _agents(obs_dim, act_dim, hid, activation=activation, mini_batch_size=mb, opt_epochs=1, target_kl=0.0, use_clipped_value=(hid == 32))
    batch = _synthetic_batch(a, rows=2 * mb)
    perm = torch.randperm(2 * mb, device='cuda', generator=torch.Generator(device='cuda').manual_seed(5)).to(torch.int32)[None, :].contiguous()
    grad = torch.zeros_like(a.fused.params)
    before = a.fused.params.clone()
    a.fused.update(batch, perm[:, :mb].contiguous(), mb, activation=activation, clip_param=a.clip_param, target_kl=0.0,
                   entropy_coef=a.entropy_coef, use_clipped_value=a.use_clipped_value, actor_lr=3e-4, critic_lr=1e-3, grad_out=grad, num_mini_batch=1)
    torch.cuda.synchronize()
    # torch reference on the same rows
    idx = perm[0, :mb].long()
    mbatch = {k: v.reshape(-1, v.shape[-1])[idx] for k, v in batch.items()}
    pl, el, kl = b.compute_policy_loss(mbatch)
    b.ac.zero_grad()
    (pl + b.entropy_coef * el).backward()
    vl = b.compute_value_loss(mbatch)
    vl.backward()
    named = dict(b.ac.named_parameters())
    for n in TRAIN_SECTIONS:
        ref = named[n].grad.reshape(-1)
        got = a.fused.views(grad, n).reshape(-1)
        err = (got - ref).abs()
        assert bool((err <= 1e-6 + 2e-4 * ref.abs()).all()), (n, float(err.max()), float(ref.abs().max()))
    st = a.fused.stats.tolist()
    np.testing.assert_allclose(st[:4], [float(pl.detach()), float(vl.detach()), float(el.detach()), float(kl.detach())], rtol=2e-4, atol=1e-6)
    assert st[4] == 1 and st[5] == 1
    # one Adam step from zero moments: bias-corrected m / (sqrt(v) + eps) = g / (|g| + eps)
    delta = a.fused.params - before
    ga = a.fused.views(grad, 'actor.pi_net.fcs.1.weight')
    big = ga.abs() > 1e-6
    np.testing.assert_allclose(a.fused.views(delta, 'actor.pi_net.fcs.1.weight')[big].cpu().numpy(), (-3e-4 * ga[big] / (ga[big].abs() + 1e-8)).cpu().numpy(), rtol=1e-3)
    assert a.fused.steps.tolist() == [1, 1]


@pytest.mark.parametrize('mb,blocks,clipped,target_kl', [(64, 1, False, 0.01), (256, 2, True, 0.01), (1000, 1, False, 0.002)])
def test_update_sequence_matches_torch_restatement(mb, blocks, clipped, target_kl):
    '''Full PPOAgent.update (2 epochs, KL gate live, two Adams) on the same record and permutations: parameters after the whole
    sequence agree to 1e-5 absolute (weights are O(0.1), one Adam step moves them by <= lr = 3e-4 / 1e-3) and the statistics to 1e-3.'''
    a, b = _agents(12, 2, 64, seed=3, mini_batch_size=mb, opt_epochs=2, target_kl=target_kl, use_clipped_value=clipped)
    rows = 4096
    batch = _synthetic_batch(a, rows, blocks=blocks, logp_noise=0.02, logp_bias=0.004 if target_kl == 0.002 else 0.0)
    ra = a.update(batch, generator=torch.Generator(device='cuda').manual_seed(11))
    rb = b.update(batch, generator=torch.Generator(device='cuda').manual_seed(11))
    n_mb = rows // mb
    assert ra['minibatches'] == 2 * n_mb
    pa, pb = _flat_params(a), _flat_params(b)
    assert float((pa - pb).abs().max()) < 1e-5, float((pa - pb).abs().max())
    for k in ('policy_loss', 'value_loss', 'entropy_loss', 'approx_kl'):
        assert abs(ra[k] - rb[k]) <= 1e-3 * abs(rb[k]) + 2e-6, (k, ra[k], rb[k])
    # the KL gate skipped the same actor steps: compare the optimizer step counts
    sd_a, sd_b = a.state_dict(), b.state_dict()
    for opt in ('actor_opt', 'critic_opt'):
        sa, sb = sd_a[opt]['state'], sd_b[opt]['state']
        assert set(sa) == set(sb)
        for i in sa:
            assert float(sa[i]['step']) == float(sb[i]['step']), opt
            np.testing.assert_allclose(sa[i]['exp_avg'].cpu().numpy(), sb[i]['exp_avg'].cpu().numpy(), rtol=2e-3, atol=2e-7)
    if target_kl == 0.002:
        assert ra['actor_steps'] < 2 * n_mb          # the gate must actually have closed in this case


def test_checkpoint_roundtrip_between_fused_and_torch_agents():
    '''state_dict() of the fused agent is the reference's format (torch.optim.Adam state): a torch agent resumes from it and vice versa.'''
    a, b = _agents(12, 2, 64, seed=4, mini_batch_size=128, opt_epochs=1)
    batch = _synthetic_batch(a, 1024)
    a.update(batch, generator=torch.Generator(device='cuda').manual_seed(1))
    b.load_state_dict(copy.deepcopy(a.state_dict()))
    a2, _ = _agents(12, 2, 64, seed=9, mini_batch_size=128, opt_epochs=1)
    a2.load_state_dict(copy.deepcopy(b.state_dict()))
    assert a2.fused.steps.tolist() == a.fused.steps.tolist()
    torch.testing.assert_close(a2.fused.params, a.fused.params, rtol=0, atol=0)
    torch.testing.assert_close(a2.fused.exp_avg_sq, a.fused.exp_avg_sq, rtol=0, atol=0)
    a2.update(batch, generator=torch.Generator(device='cuda').manual_seed(2))
    b.update(batch, generator=torch.Generator(device='cuda').manual_seed(2))
    assert float((_flat_params(a2) - _flat_params(b)).abs().max()) < 1e-5


def test_device_pack_kernel_equals_torch_packer():
    from safe_control_gym_b200.engine import pack_policy
    a, _ = _agents(24, 4, 64, seed=5)
    blob, dims = a.fused.rollout_blob()
    ref, dims2 = pack_policy(a.ac, torch.device('cuda:0'))
    assert dims == dims2
    torch.testing.assert_close(blob, ref, rtol=0, atol=0)


def test_bad_arguments_raise_value_error():
    a, _ = _agents(12, 2, 64, mini_batch_size=64)
    batch = _synthetic_batch(a, 256)
    perm = torch.arange(256, device='cuda', dtype=torch.int32)[None, :].contiguous()
    with pytest.raises(ValueError):
        a.fused.update(batch, perm[:, :32].contiguous(), 64)            # perm shorter than one minibatch
    with pytest.raises(AssertionError):
        a.fused.update(batch, perm, 512)                                 # num_mini_batch is 0 (ppo_utils.py:94)
