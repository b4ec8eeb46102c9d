'''RAP (safe_control_gym/controllers/rarl/rap.py) host logic on CPU: adversary draw, composition of the two players' records from replayed
launches, advantage normalisation over the batch before the split.  The engine is a stand-in whose "launch" is a known function of
(adversary blob, adv_record, env column), so every assembled column can be checked against the launch it must have come from.'''
import numpy as np
import pytest
import torch

from safe_control_gym_b200.controllers.rarl import rap as R


def test_draw_and_split_follow_the_reference_grouping():
    rng = np.random.RandomState(0)
    idx, groups, starts, ends = R.draw_adversaries(rng, 3, 16)
    assert list(idx) == sorted(idx) and set(groups) <= {0, 1, 2}
    # reference arithmetic (rap.py:354-356, 462-463) on the same draw
    ref_idx = sorted(np.random.RandomState(0).randint(3, size=16))
    g, sp = np.unique(ref_idx, return_index=True)
    assert list(idx) == list(ref_idx) and list(groups) == list(g) and list(starts) == list(sp)
    assert list(ends) == list(np.concatenate([sp[1:], [16]]))
    buf = {'obs': torch.arange(2 * 16 * 3.0).view(2, 16, 3), 'adv': torch.arange(2 * 16.0).view(2, 16, 1)}
    parts = R.split_by_adversary(buf, groups, starts, ends)
    assert [p[0] for p in parts] == [int(x) for x in groups]
    assert sum(p[1]['obs'].shape[1] for p in parts) == 16
    for (gi, d), s, e in zip(parts, starts, ends):
        assert torch.equal(d['obs'], buf['obs'][:, s:e]) and d['obs'].is_contiguous()
    a = torch.randn(5, 7, 1, dtype=torch.float64)
    np.testing.assert_allclose(R.normalise_advantages(a).numpy(), (a.numpy() - a.numpy().mean()) / (a.numpy().std() + 1e-6), rtol=1e-12)


class _FakeEngine:
    '''out[k][t, n] = code(k) + 1000 * adversary id + 100 * adv_record + n + 0.01 t; refuses a launch that does not start from the snapshot'''
    KEYS = {'obs': 3, 'act': None, 'rew': 1, 'mask': 1, 'v': 1, 'logp': 1, 'terminal_v': 1, 'ret': 1, 'adv': 1}

    def __init__(self, N, A, A2):
        self.N, self.A, self.A2, self.state, self.launches = N, A, A2, 'dirty', []

    def get_state(self):
        assert self.state == 'reset'
        return 'snap'

    def set_state(self, s):
        assert s == 'snap'
        self.state = 'snap'

    def rollout(self, blob, dims, T, adversary=None, adv_record=False, out=None, **kw):
        assert self.state == 'snap', 'every launch must replay from the snapshot'
        self.state = 'dirty'
        adv_blob, adv_dims = adversary
        assert torch.is_tensor(adv_blob) and adv_blob.dtype == torch.float32 and adv_blob.is_contiguous()      # engine.rollout's own checks
        assert adv_dims == (dims[0], self.A2, dims[2])
        gid = int(adv_blob[0])
        self.launches.append((gid, bool(adv_record)))
        res = {}
        for i, (k, w) in enumerate(self.KEYS.items()):
            w = (self.A2 if adv_record else self.A) if w is None else w
            n = torch.arange(self.N, dtype=torch.float32).view(1, self.N, 1)
            t = torch.arange(T, dtype=torch.float32).view(T, 1, 1)
            res[k] = (10000.0 * i + 1000.0 * gid + 100.0 * adv_record + n + 0.01 * t).expand(T, self.N, w).clone()
            if k in ('obs', 'mask'):                                   # shared between the players: the trajectory does not depend on who is recorded
                res[k] = (10000.0 * i + 1000.0 * gid + n + 0.01 * t).expand(T, self.N, w).clone()
        for j, k in enumerate(R.EPISODE_KEYS):
            res[k] = (torch.arange(self.N, dtype=torch.float32) + 1000.0 * gid + 7 * j).to(torch.int32 if k == 'ep_count' else torch.float32)
        return res


class _FakeEnv:
    def __init__(self, eng):
        self.unwrapped = type('U', (), {'engine': eng})()
        self.resets = 0

    def reset(self):
        self.resets += 1
        self.unwrapped.engine.state = 'reset'
        return None, None


def _bare_rap(N, T, K, A=2, A2=3, seed=5):
    ctl = R.RAP.__new__(R.RAP)                                           # (the constructor needs a GPU: only the collector logic is under test)
    eng = _FakeEngine(N, A, A2)
    ctl.env = _FakeEnv(eng)
    ctl.rollout_steps, ctl.rollout_batch_size, ctl.num_adversaries = T, N, K
    ctl.activation, ctl.gamma, ctl.gae_lambda, ctl.use_gae = 'tanh', 0.99, 0.95, True
    ctl.total_steps, ctl._rap_out = 0, {}
    ctl._adv_rng = np.random.RandomState(seed)
    ctl.adv_act_space = type('S', (), {'shape': (A2,)})()
    ctl._policy_blob = lambda: (torch.zeros(4), (3, A, 64))
    # stand-in agents with the fused-update interface the real ones have (FusedPPOState.rollout_blob -> (blob, dims)); RAP._blob_of itself is the
    # code under test.  The stand-in launch reads the adversary's id from its blob.
    fused = lambda i: type('F', (), {'rollout_blob': lambda self: (torch.full((4,), float(i)), (3, A2, 64))})()
    ctl.adversaries = [type('Agent', (), {'fused': fused(i)})() for i in range(K)]
    return ctl, eng


@pytest.mark.parametrize('N,K,seed', [(12, 3, 5), (4, 2, 0), (9, 4, 11)])
def test_collect_composes_both_records_from_replayed_launches(N, K, seed):
    T = 6
    ctl, eng = _bare_rap(N, T, K, seed=seed)
    idx, groups, starts, ends = R.draw_adversaries(np.random.RandomState(seed), K, N)
    rollouts, splits, stats = ctl.collect_rollouts()
    assert ctl.env.resets == 1 and ctl.total_steps == N * T
    # two launches (protagonist record, adversary record) per adversary that was drawn, each from the snapshot
    assert eng.launches == [(int(g), r) for g in groups for r in (False, True)]
    keys = list(_FakeEngine.KEYS)
    n = torch.arange(N, dtype=torch.float32).view(1, N, 1); t = torch.arange(T, dtype=torch.float32).view(T, 1, 1)
    gid = torch.tensor(idx, dtype=torch.float32).view(1, N, 1)          # the adversary every env column must have been rolled against
    for k in ('obs', 'mask', 'act', 'rew', 'v', 'logp', 'terminal_v', 'ret'):
        want = 10000.0 * keys.index(k) + 1000.0 * gid + n + 0.01 * t
        assert torch.equal(rollouts[k], want.expand_as(rollouts[k])), k
    assert rollouts['act'].shape == (T, N, 2)
    raw_adv = 10000.0 * keys.index('adv') + 1000.0 * gid + n + 0.01 * t
    torch.testing.assert_close(rollouts['adv'], R.normalise_advantages(raw_adv), rtol=1e-6, atol=1e-6)
    # the adversaries' side: their own act width, values of the adv_record launches, obs / mask shared; advantages normalised over the WHOLE batch
    raw_adv2 = R.normalise_advantages(raw_adv + 100.0)
    assert [g for g, _ in splits] == [int(g) for g in groups]
    for (g, d), s, e in zip(splits, starts, ends):
        assert d['act'].shape == (T, e - s, 3)
        for k in ('act', 'rew', 'v', 'logp', 'terminal_v', 'ret'):
            want = (10000.0 * keys.index(k) + 1000.0 * g + 100.0 + n + 0.01 * t)[:, s:e]
            assert torch.equal(d[k], want.expand_as(d[k])), k
        assert torch.equal(d['obs'], rollouts['obs'][:, s:e]) and torch.equal(d['mask'], rollouts['mask'][:, s:e])
        torch.testing.assert_close(d['adv'], raw_adv2[:, s:e], rtol=1e-6, atol=1e-6)
    for j, k in enumerate(R.EPISODE_KEYS):
        want = torch.arange(N, dtype=torch.float32) + 1000.0 * torch.tensor(idx, dtype=torch.float32) + 7 * j
        assert torch.equal(stats[k].float(), want), k
    # a second collect resets again and draws a new assignment
    ctl.collect_rollouts()
    assert ctl.env.resets == 2 and ctl.total_steps == 2 * N * T


def test_rap_is_registered_with_the_reference_defaults():
    from safe_control_gym_b200.utils.registration import get_config
    cfg = get_config('rap')
    assert cfg['num_adversaries'] == 2 and cfg['rollout_steps'] == 100 and cfg['agent_iterations'] == 10
