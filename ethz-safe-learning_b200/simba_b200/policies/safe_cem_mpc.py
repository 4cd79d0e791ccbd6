"""SafeCemMpc — mirrors simba/policies/safe_cem_mpc.py on the fused B200 planner.

The active rule of the reference is `mean_return - 100 * unsafe` (safe_cem_mpc.py:96) where a
candidate is unsafe if, at any step, the Beta-posterior mean of its particles' cost indicator
exceeds `posterior_mean_threashold` [sic] (:110-120). `selection='feasible_first'` selects the
north-star rule instead (feasible candidates ranked by return, then fewest violations);
`optimize_for_safety` (:40-74, dead code in the reference) is the least-cost CEM.
"""
import numpy as np

from .. import _lib
from .cem_mpc import CemMpc


class SafeCemMpc(CemMpc):
    _objective = _lib.OBJ_SAFE_PENALTY

    def __init__(self, model, environment, horizon, iterations, smoothing, n_samples, n_elite,
                 particles, stddev_threshold, noise_stddev, posterior_mean_threashold, *,
                 selection='penalty', **kwargs):
        objective = {'penalty': _lib.OBJ_SAFE_PENALTY, 'feasible_first': _lib.OBJ_FEASIBLE_FIRST,
                     'least_cost': _lib.OBJ_LEAST_COST}[selection]
        super().__init__(model, environment, horizon, iterations, smoothing, n_samples, n_elite,
                         particles, stddev_threshold, noise_stddev, objective=objective,
                         posterior_mean_threashold=posterior_mean_threashold, **kwargs)
        self.cost = environment.get_cost
        self.posterior_mean_threashold = posterior_mean_threashold
        self.last_action = np.zeros((self.action_space.shape[0],), dtype=np.float32)
        self._kwargs = dict(kwargs)
        self._least_cost = None

    def optimize_for_safety(self, state):
        """safe_cem_mpc.py:40-74: CEM on scores = -mean cumulative cost."""
        if self._least_cost is None:
            self._least_cost = SafeCemMpc(
                self.model, self.environment, self.horizon, self.iterations, self.smoothing,
                self.n_samples, self.elite, self.particles, self.stddev_threshold,
                self.noise_stddev, self.posterior_mean_threashold, selection='least_cost',
                **self._kwargs)
        return self._least_cost.generate_action(state)

    def compute_mean_costs(self, trajectories, action_sequences=None):
        """safe_cem_mpc.py:98-108 on materialised trajectories -> mean cost sum [N]."""
        if self._least_cost is None:
            self.optimize_for_safety  # noqa: B018 (documented entry); build the helper lazily
            self._least_cost = SafeCemMpc(
                self.model, self.environment, self.horizon, self.iterations, self.smoothing,
                self.n_samples, self.elite, self.particles, self.stddev_threshold,
                self.noise_stddev, self.posterior_mean_threashold, selection='least_cost',
                **self._kwargs)
        return -self._least_cost.compute_objective(trajectories, action_sequences)

    @property
    def count_threshold(self):
        """Largest per-step particle count that still passes the Beta test (:110-120)."""
        import ctypes as C
        n = C.c_int32()
        _lib.check(self._lib.simba_planner_count_threshold(self._ensure_planner(), C.byref(n)))
        return n.value
