'''Two-GPU checks of the fused rollout + all-gather (NVLink peer stores / NVSwitch multicast) and of the lock-step PPO replicas.
Skipped on boxes with fewer than two GPUs (run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multirank.py -m gpu`).'''
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip('torch')

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, time
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], 'tests'))
import numpy as np, torch, torch.distributed as dist
from functools import partial
rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
dist.init_process_group('nccl', device_id=torch.device('cuda', rank))
torch.cuda.set_device(rank)
from safe_control_gym_b200 import configs
from safe_control_gym_b200.utils.registration import make
task = {k: v for k, v in configs.get('quadrotor_2D_track').items() if k != 'seed'}
env_func = partial(make, 'quadrotor', **task)
ctrl = make('ppo', env_func, training=True, output_dir='/tmp/scg_mr', seed=3, rollout_batch_size=256 * world, rollout_steps=64,
            mini_batch_size=2048, opt_epochs=3, use_gae=True, hidden_dim=64, device=rank)
ctrl.reset()
mode = None
for it in range(6):
    if rank == 1:
        time.sleep(0.05 * (it % 3))            # one rank lags: without the pre-rollout barrier a fast peer would overwrite the gathered record
        torch.cuda._sleep(int(2e8))            # ... and its GPU lags too (update still running while the peer wants to roll out again)
    r = ctrl.train_step()
    pg = ctrl._peer
    mode = ('multicast' if pg.mc_offset is not None else 'peer stores') if pg.fused else 'nccl fallback: ' + pg.why
    # every rank must hold the identical gathered record and identical parameters
    flat = pg.buf.view(world, pg.slot)[:, :pg.numel]          # the slots are padded to 128 bytes: the padding is never written
    h = torch.stack([flat.double().sum(), flat.double().abs().sum(), ctrl.agent.fused.params.double().sum(), ctrl.agent.fused.params.double().abs().sum()])
    hs = [torch.zeros_like(h) for _ in range(world)]
    dist.all_gather(hs, h)
    for o in hs[1:]:
        assert torch.equal(o, hs[0]), (it, [x.tolist() for x in hs])
# gathered record == the ranks' own shards, bit for bit
mine = ctrl._rollout_out[64]['_flat']
shards = [torch.zeros_like(mine) for _ in range(world)]
dist.all_gather(shards, mine)
for g in range(world):
    assert torch.equal(pg.buf.view(world, pg.slot)[g, :mine.numel()], shards[g]), g
if rank == 0:
    print('MULTIRANK OK', mode, 'steps', ctrl.total_steps, flush=True)
ctrl.close()
dist.destroy_process_group()
'''


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs')
def test_replicas_stay_bit_identical_with_a_lagging_rank(tmp_path):
    w = tmp_path / 'worker.py'
    w.write_text(WORKER)
    env = dict(os.environ, MASTER_ADDR='127.0.0.1')
    r = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
                        '--master-port', '29611', str(w), ROOT], capture_output=True, text=True, timeout=600, env=env)
    print(r.stdout[-3000:], r.stderr[-6000:])
    assert r.returncode == 0 and 'MULTIRANK OK' in r.stdout
