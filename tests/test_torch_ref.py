"""oracle/torch_ref.py (torch, reference-shaped) against oracle/simba_oracle.py (numpy) and against the
fixtures the unmodified reference code produced: two independently written restatements and the
reference's own code must tell the same story on every objective."""
import os

import numpy as np
import pytest

from oracle import simba_oracle as so
from oracle import torch_ref
from simba_b200 import synthetic
from tests import helpers
from tests.test_reference_golden import GOLD


def torch_planner(c, objective, threshold=0.15, smoothing=0.0, stddev_threshold=0.0, noise_stddev=0.01,
                  sampling_propagation=True, scorer_config=None):
    cfg = dict(so.DEFAULT_SCORER_CONFIG)
    cfg.update(scorer_config or {})
    dyn = torch_ref.Dynamics(torch_ref.Ensemble(c['weights']), c['smin'], c['smax'], True, sampling_propagation)
    return torch_ref.Planner(dyn, torch_ref.GoalScorer(cfg, c['table']), [-1.0] * c['A'], [1.0] * c['A'], c['H'],
                             c['I'], smoothing, c['N'], c['K'], c['P'], stddev_threshold, noise_stddev,
                             posterior_mean_threashold=threshold, objective=objective)


@pytest.mark.parametrize("cfg", ['tiny', 'c1'])
@pytest.mark.parametrize("objective", ['reward', 'penalty', 'least_cost', 'feasible_first'])
def test_torch_ref_agrees_with_numpy_oracle(cfg, objective):
    c = helpers.workload(cfg)
    z, eps, zf = synthetic.make_draws(c['I'], 1, c['N'], c['H'], c['A'], c['P'], c['O'])
    tr = so.Trace()
    a0, s0, n0 = helpers.oracle_planner(c, objective).do_generate_action(c['state'], z[:, 0], eps[:, 0], zf[0], tr)
    a1, s1, n1, last = torch_planner(c, objective).do_generate_action(c['state'], z[:, 0], eps[:, 0], zf[0])
    assert n0 == n1
    assert np.array_equal(last['elite'], tr[-1]['elite'])
    assert np.allclose(last['mu'], tr[-1]['mu'], atol=1e-6) and np.allclose(last['sigma'], tr[-1]['sigma'], atol=1e-6)
    assert np.allclose(a1, a0, atol=1e-6)
    if objective != 'feasible_first':                  # (there the reported best score is product-defined)
        assert np.isclose(s1, s0, rtol=1e-6, atol=8e-6)


@pytest.mark.parametrize("cfg,tag,objective", [('tiny', 'reward', 'reward'), ('tiny', 'penalty', 'penalty'),
                                               ('c1', 'reward', 'reward'), ('c1', 'penalty', 'penalty')])
def test_torch_ref_matches_reference_code(cfg, tag, objective):
    g = np.load(os.path.join(GOLD, 'reference_%s_plan.npz' % cfg))
    c = helpers.workload(cfg)
    z, eps, zf = synthetic.make_draws(c['I'], 1, c['N'], c['H'], c['A'], c['P'], c['O'])
    a, s, n, last = torch_planner(c, objective).do_generate_action(c['state'], z[:, 0], eps[:, 0], zf[0])
    assert n == int(g[tag + '_iterations'])
    assert np.array_equal(last['elite'], g['%s_elite_%d' % (tag, n - 1)])
    assert np.allclose(a, g[tag + '_action'], atol=1e-6) and np.isclose(s, g[tag + '_score'], rtol=1e-6, atol=8e-6)
