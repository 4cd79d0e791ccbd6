"""Regenerates tests/golden/oracle_tiny_plan.npz from the CPU oracle (self-golden; the reference has
no fixtures and cannot be imported here: TensorFlow / TFP / gym are not installed).
Run from the repo root:  python tests/golden/make_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'ethz-safe-learning_b200')]

from oracle import simba_oracle as so  # noqa: E402
from simba_b200 import synthetic  # noqa: E402
from tests import helpers  # noqa: E402

c = helpers.workload('tiny')
z, eps, zf = synthetic.make_draws(c['I'], 1, c['N'], c['H'], c['A'], c['P'], c['O'])
out = {}
for objective in ('reward', 'penalty', 'least_cost', 'feasible_first'):
    tr = so.Trace()
    a, s, n = helpers.oracle_planner(c, objective).do_generate_action(c['state'], z[:, 0], eps[:, 0], zf[0], tr)
    out[objective + '_action'] = a
    out[objective + '_score'] = np.float32(s)
    for it, rec in enumerate(tr):
        out['%s_elite_%d' % (objective, it)] = rec['elite']
        out['%s_mu_%d' % (objective, it)] = rec['mu']
        out['%s_sigma_%d' % (objective, it)] = rec['sigma']
        out['%s_ret_%d' % (objective, it)] = rec['ret']
        out['%s_cost_%d' % (objective, it)] = rec['cost']
np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', 'oracle_tiny_plan.npz'), **out)
print("wrote", len(out), "arrays")
