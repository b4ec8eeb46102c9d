'''`MetricExtractor` of this repo (experiments/base_experiment.py) on the evaluation episodes the UNMODIFIED reference harness recorded:
`BaseExperiment(env, ctrl).run_evaluation(n_episodes=5)` of /root/reference (experiments/base_experiment.py:49-165, RecordDataWrapper, MetricExtractor :380-492)
with the reference PPO controller on a reference env, recorded by tools/make_golden_experiment.py (tests/golden/experiment_*.npz).  Every metric the reference
derived -- lengths, returns, per-episode RMSE from info['mse'], its CVaR, failure rate, constraint-violation steps -- must come out of the same streams.'''
import os

import numpy as np
import pytest

from safe_control_gym_b200.experiments.base_experiment import MetricExtractor

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


@pytest.mark.parametrize('case', ['quad2d_track'])
def test_metrics_of_the_reference_evaluation_episodes(case):
    G = np.load(os.path.join(GOLD, f'experiment_{case}.npz'))
    n = int(G['n_episodes'])
    data = {k: [G[f'ep{e}_{k}'] for e in range(n)] for k in ('reward', 'length', 'mse', 'constraint_violation')}
    got = MetricExtractor().compute_metrics(data)
    want = {k[len('metric_'):]: G[k] for k in G.files if k.startswith('metric_')}
    assert set(got) == set(want), sorted(set(got) ^ set(want))
    for k, v in want.items():
        np.testing.assert_allclose(np.asarray(got[k], dtype=np.float64), v, rtol=1e-12, atol=1e-12, err_msg=k)
