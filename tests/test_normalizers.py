'''Running obs / reward normalisers (math_and_models/normalization.py:20-159): the device-resident classes against a step-by-step numpy
restatement of the reference classes (CPU), and the frozen-statistics normalisation inside the fused collector against the oracle (GPU).'''
import numpy as np
import pytest

import helpers as H

torch = pytest.importorskip('torch')


class _RefRMS:          # normalization.py:20-56
    def __init__(self, eps=1e-4, shape=()):
        self.mean, self.var, self.count = np.zeros(shape), np.ones(shape), eps

    def update(self, arr):
        bm, bv, bc = arr.mean(0), arr.var(0), arr.shape[0]
        d, tot = bm - self.mean, self.count + bc
        m2 = self.var * self.count + bv * bc + d ** 2 * self.count * bc / tot
        self.mean, self.var, self.count = self.mean + d * bc / tot, m2 / tot, tot


def test_reward_normalizer_reproduces_the_reference_including_its_indexing_quirk():
    from safe_control_gym_b200.math_and_models.normalization import RewardStdNormalizer
    rng = np.random.default_rng(0)
    T, N = 60, 24
    ref, ret = _RefRMS(), np.zeros(N)
    n = RewardStdNormalizer(gamma=0.99)
    for rollout in range(3):
        rew = rng.random((T, N)).astype(np.float32)
        done = rng.random((T, N)) < 0.05
        for t in range(T):                               # normalization.py:148-159
            ret = ret * 0.99 + rew[t]
            ref.update(ret)
            ret[done[t].astype(np.int64)] = 0            # the reference's integer indexing (not a mask)
        n.update_from_rollout(torch.tensor(rew), torch.tensor(done.astype(np.float32)))
        np.testing.assert_allclose(float(n.rms.var), ref.var, rtol=1e-12)
        np.testing.assert_allclose(n.ret.numpy(), ret, rtol=0, atol=1e-12)
    x = rng.random(N)
    np.testing.assert_allclose(torch.clamp(torch.tensor(x) * n.scale(), -10, 10).numpy(), np.clip(x / np.sqrt(ref.var + 1e-8), -10, 10), rtol=1e-12)


def test_obs_normalizer_one_merge_per_rollout_equals_one_merge_per_step():
    from safe_control_gym_b200.math_and_models.normalization import MeanStdNormalizer
    rng = np.random.default_rng(1)
    T, N, D = 40, 32, 6
    ref, m = _RefRMS(shape=(D,)), MeanStdNormalizer(shape=(D,))
    for rollout in range(3):
        obs = rng.normal(size=(T, N, D)) * (1 + rollout)
        for t in range(T):
            ref.update(obs[t])
        m.rms.update(torch.tensor(obs).reshape(-1, D))
        np.testing.assert_allclose(m.rms.mean.numpy(), ref.mean, rtol=1e-10, atol=1e-13)
        np.testing.assert_allclose(m.rms.var.numpy(), ref.var, rtol=1e-10)
    sd = m.state_dict()
    m2 = MeanStdNormalizer(shape=(D,)); m2.load_state_dict(sd)
    x = rng.normal(size=(5, D))
    np.testing.assert_allclose(m2.normalize(x).numpy(), np.clip((x - ref.mean) / np.sqrt(ref.var + 1e-8), -10, 10), rtol=1e-9)


@pytest.mark.gpu
@pytest.mark.parametrize('engine', ['ffma2', 'tcgen05'])
@pytest.mark.parametrize('name', ['quadrotor_2D_track', 'cartpole_c2'])
def test_fused_collector_applies_frozen_normalisers_like_the_oracle(name, engine):
    from safe_control_gym_b200.engine import EngineHandle
    env = H.make_env(name)
    N, T, od = 128, 80, env.obs_dim
    rng = np.random.default_rng(3)
    norm = np.zeros(2 * od + 5)
    norm[:od] = rng.normal(size=od) * 0.3
    norm[od:2 * od] = 1.0 / np.sqrt(rng.uniform(0.2, 2.0, size=od) + 1e-8)
    norm[2 * od], norm[2 * od + 1] = 1.5, 1.0            # clip_obs (tight on purpose: the clip must fire), obs_on
    norm[2 * od + 2], norm[2 * od + 3], norm[2 * od + 4] = 7.0, 5.0, 1.0
    pol = H.random_policy(od, env.action_dim, 64, seed=2)
    o = H.make_oracle(env, N, auto_reset=True, rng_mode='philox', seed=21)
    r0 = o.reset()
    obs0 = np.clip((r0['obs'].astype(np.float32).astype(np.float64) - norm[:od]) * norm[od:2 * od], -1.5, 1.5).astype(np.float32)
    oo = o.rollout(pol, T, eps=None, use_gae=True, trace=True, norm=norm, obs0=obs0)
    eng = EngineHandle(env, num_envs=N, auto_reset=True, rng_mode='philox', seed=21)
    eng.reset()
    blob, dims = H.policy_blob(pol, eng.device)
    gg = eng.rollout(blob, dims, T, eps=None, use_gae=True, want=('ret', 'adv', 'cviol', 'mse', 'last', 'episode', 'raw'),
                     norm=torch.tensor(norm, device=eng.device), obs0=torch.tensor(obs0, device=eng.device), mlp_engine=engine)
    H.compare_rollouts(oo, gg, o.cfg, o.x_goal, N, T, obs_tol=5e-5, keys_extra=('obs_raw', 'rew_raw'), horizon_check=(name != 'cartpole_c2'),
                       early=(20, 5e-6))
    assert float(gg['obs'].abs().max()) <= 1.5 + 1e-6 and float(gg['rew'].abs().max()) <= 5.0 + 1e-6
    assert float((gg['obs'].abs() > 1.499).float().mean()) > 0.001          # the clip fired
    eng.close()


@pytest.mark.gpu
def test_ppo_trains_with_running_normalisers(tmp_path):
    from functools import partial

    from safe_control_gym_b200 import configs
    from safe_control_gym_b200.utils.registration import make
    task = {k: v for k, v in configs.get('quadrotor_2D_track').items() if k != 'seed'}
    ctrl = make('ppo', partial(make, 'quadrotor', **task), training=True, output_dir=str(tmp_path), seed=5, rollout_batch_size=256, rollout_steps=50,
                mini_batch_size=1024, opt_epochs=2, use_gae=True, norm_obs=True, norm_reward=True, clip_obs=10, clip_reward=10)
    ctrl.reset()
    c0 = ctrl.obs_normalizer.rms.count
    for _ in range(3):
        r = ctrl.train_step()
        assert np.isfinite([r['policy_loss'], r['value_loss']]).all()
    assert ctrl.obs_normalizer.rms.count > c0 + 3 * 256 * 50 - 1 and float(ctrl.reward_normalizer.rms.var) != 1.0
    assert float(ctrl.obs.abs().max()) <= 10.0
    ctrl.save(str(tmp_path / 'ck.pt'))
    sd = torch.load(str(tmp_path / 'ck.pt'), weights_only=False)
    assert set(sd['obs_normalizer']) == {'mean', 'var'} and sd['obs_normalizer']['mean'].shape == (12,)      # the reference's checkpoint keys
    ev = ctrl.run(n_episodes=8)
    assert ev['ep_returns'].shape == (8,)
    ctrl.close()
