'''World-size-2 tests on CPU (gloo): env sharding by global index and the single all-gather of the rollout record.'''
import os
import socket

import numpy as np
import pytest

import helpers as H

torch = pytest.importorskip('torch')


def _free_port():
    s = socket.socket(); s.bind(('127.0.0.1', 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from safe_control_gym_b200.controllers.ppo.ppo import all_gather_rollout
    T, n = 5, 3
    g = torch.Generator().manual_seed(100 + rank)
    layout = [('obs', 12), ('act', 2), ('rew', 1), ('mask', 1), ('v', 1), ('logp', 1), ('terminal_v', 1), ('ret', 1), ('adv', 1)]
    flat = torch.randn(sum(w for _, w in layout) * T * n, generator=g)
    out, o = {'_flat': flat, '_layout': layout, '_TN': (T, n)}, 0
    for k, w in layout:
        out[k] = flat[o:o + T * n * w].view(T, n, w)
        o += T * n * w
    res = all_gather_rollout(out)
    # the peer-memory gatherer must fall back to the same collective where symmetric memory is unavailable (CPU / gloo)
    from safe_control_gym_b200.controllers.ppo.ppo import PeerGatheredRollout, rollout_flat_numel
    assert rollout_flat_numel(12, 2, T, n) == flat.numel()
    pg = PeerGatheredRollout(flat.numel(), torch.device('cpu'))
    assert not pg.fused and pg.rollout_kwargs()['peer_byte_offsets'] is None
    pg.rollout_kwargs()['flat'].copy_(flat)                     # what the collector kernel would have written into the local slot
    res2 = pg.finish(out)
    for k in res:
        assert torch.equal(res[k], res2[k]), k
    q.put((rank, {k: v.numpy().copy() for k, v in res.items()}, {k: out[k].numpy().copy() for k, _ in layout}))
    dist.destroy_process_group()


def test_all_gather_of_the_rollout_record_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q, port = ctx.Queue(), _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted([q.get(timeout=120) for _ in range(2)], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
    (r0, g0, l0), (r1, g1, l1) = got
    for k in l0:
        assert g0[k].shape[:2] == (2, 15)
        np.testing.assert_array_equal(g0[k], g1[k])                              # every rank holds the identical gathered record
        np.testing.assert_array_equal(g0[k][0], l0[k].reshape(15, -1))            # [rank][T*n][w]
        np.testing.assert_array_equal(g0[k][1], l1[k].reshape(15, -1))


def test_sharded_rollouts_equal_slices_of_the_unsharded_rollout():
    '''Philox streams are keyed by the GLOBAL env index: rank r with env_index_offset = r*n reproduces envs [r*n, (r+1)*n) of the
    single-process batch (oracle, same stream definition as the CUDA engine).'''
    env = H.make_env('quadrotor_2D_track')
    N, T = 8, 30
    pol = H.random_policy(env.obs_dim, env.action_dim, 64, seed=1)
    full = H.make_oracle(env, N, rng_mode='philox', seed=9)
    full.reset()
    ref = full.rollout(pol, T)
    for r in range(2):
        sh = H.make_oracle(env, N // 2, rng_mode='philox', seed=9, env_index_offset=r * (N // 2))
        sh.reset()
        out = sh.rollout(pol, T)
        for k in ('obs', 'act', 'rew', 'mask', 'logp'):
            np.testing.assert_array_equal(out[k], ref[k][:, r * (N // 2):(r + 1) * (N // 2)])
