'''The reference evaluation harness, recorded.  Run in the build container:  python tools/make_golden_experiment.py -> tests/golden/experiment_*.npz

`BaseExperiment(env, ctrl).run_evaluation(n_episodes=...)` of /root/reference (experiments/base_experiment.py:49-165, `RecordDataWrapper`,
`MetricExtractor` :310-492, unmodified) with the reference PPO controller on an unmodified reference env (tools/ref_harness.py).  Stored: the per-episode
streams the wrapper recorded (reward, info['mse'], info['constraint_violation'], lengths) and every metric `MetricExtractor.compute_metrics` derived from them.
'''
import json
import os
import sys
import tempfile

import numpy as np
import torch
import yaml

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ref_harness as RH  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests', 'golden')


def run_case(name, env_id, task_cfg, seed, n_episodes):
    RH.install()
    from safe_control_gym.experiments.base_experiment import BaseExperiment
    ppo_cfg = yaml.safe_load(open(os.path.join(RH.REF, 'controllers/ppo/ppo.yaml')))
    ppo_cfg.update(rollout_batch_size=2, rollout_steps=10, seed=seed)
    torch.manual_seed(seed)
    ctl = RH.make_ppo(env_id, task_cfg, tempfile.mkdtemp(), **ppo_cfg)
    ctl.reset()
    env = (RH.make_quadrotor if env_id == 'quadrotor' else RH.make_cartpole)(**dict(task_cfg, seed=seed * 7))
    exp = BaseExperiment(env, ctl)
    trajs, metrics = exp.run_evaluation(n_episodes=n_episodes, verbose=False)
    arrs = {}
    for e in range(n_episodes):
        arrs[f'ep{e}_reward'] = np.asarray(trajs['reward'][e], dtype=np.float64)
        arrs[f'ep{e}_length'] = np.asarray(trajs['length'][e], dtype=np.float64)
        arrs[f'ep{e}_mse'] = np.asarray([i['mse'] for i in trajs['info'][e] if 'mse' in i], dtype=np.float64)
        arrs[f'ep{e}_constraint_violation'] = np.asarray([i['constraint_violation'] for i in trajs['info'][e] if 'constraint_violation' in i], dtype=np.float64)
    m = {f'metric_{k}': np.asarray(v, dtype=np.float64) for k, v in metrics.items()}
    path = os.path.join(OUT, f'experiment_{name}.npz')
    np.savez_compressed(path, n_episodes=np.int32(n_episodes), task_config=np.array(json.dumps(task_cfg)), **arrs, **m)
    print(name, {k: np.round(v, 4).tolist() for k, v in m.items()}, '->', os.path.getsize(path), 'bytes')


def main():
    c3 = RH.load_task_config('examples/rl/config_overrides/quadrotor_2D/quadrotor_2D_track.yaml')
    run_case('quad2d_track', 'quadrotor', dict(c3, episode_len_sec=1), seed=3, n_episodes=5)


if __name__ == '__main__':
    main()
