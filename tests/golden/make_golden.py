"""Regenerates tests/golden/oracle_tiny_plan.npz from the CPU oracle: a self-golden that freezes the
oracle's per-iteration trace (all four objectives, incl. the two the reference only has as dead code).
The fixtures produced by the reference's own code are reference_*.npz (make_reference_golden.py).
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

# ---- training step (oracle/train_oracle.py): three steps of a tiny ensemble, with and without dropout ----
from oracle import train_oracle as T  # noqa: E402

tout = {}
for tag, rate in (('plain', 0.0), ('dropout', 0.25)):
    rng = np.random.default_rng(7)
    members = []
    for _ in range(2):
        arrays, fan = [], 6
        for _ in range(2):
            arrays += [rng.normal(0, 0.4, (fan, 8)).astype(np.float32), rng.normal(0, 0.1, 8).astype(np.float32)]
            fan = 8
        for _ in range(2):
            arrays += [rng.normal(0, 0.4, (8, 4)).astype(np.float32), rng.normal(0, 0.1, 4).astype(np.float32)]
        members.append(arrays)
    tr = T.EnsembleTrainer(members, batch_size=5, learning_rate=1e-2, learning_rate_schedule=True, training_steps=2,
                           train_epochs=3, dropout_rate=rate, dropout_seed=11)
    losses = []
    for s in range(3):
        x = rng.uniform(0, 1, (2, 5, 6)).astype(np.float32)
        y = rng.normal(0, 0.2, (2, 5, 4)).astype(np.float32)
        losses.append(tr.training_step(x, y))
    tout[tag + '_losses'] = np.asarray(losses, np.float32)
    for e, net in enumerate(tr.nets):
        for i, a in enumerate(net.arrays):
            tout['%s_member%d_var%d' % (tag, e, i)] = a
np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', 'oracle_tiny_train.npz'), **tout)
print("wrote", len(tout), "training arrays")
