"""RandomShootingMpc / RandomMpc — simba/policies/random_shooting_mpc.py and random_mpc.py.

Random shooting is the I = 1, K = 1 special case of the CEM planner:
sample n_samples action sequences uniformly in the action box, roll each out over the ensemble
with `particles` particles, score with `compute_objective` (mpc_policy.py:26-39) and return the
first action of the best sequence (random_shooting_mpc.py:26-37). It runs on the same kernels
and the same CUDA graph as CemMpc: the uniform draws are handed to `simba_sample_actions` as
external variates, a = clip(mid + half_range * u, lb, ub) with u ~ U(-1, 1), which is U(lb, ub).

The reference class does not run as shipped (its constructor passes six arguments to a
five-argument base and calls `tf.random.uniform(lb, ub, shape)`); the behaviour implemented here
is the evident intent. `objective` is the reference's post-processing hook on the cumulative
rewards; only the identity (None) can run inside the fused path.
"""
import numpy as np
import torch

from .. import _lib
from .cem_mpc import CemMpc
from .policy import PolicyBase


class RandomShootingMpc(CemMpc):
    def __init__(self, model, environment, horizon, objective=None, n_samples=500, particles=20, *,
                 precision='bf16', seed=0, member_map='split'):
        if objective is not None:
            raise _lib.SimbaError(-6, "RandomShootingMpc: only objective=None (identity) runs on the "
                                      "fused path")
        super().__init__(model, environment, horizon, iterations=1, smoothing=0.0,
                         n_samples=n_samples, n_elite=1, particles=particles, stddev_threshold=0.0,
                         noise_stddev=0.0, precision=precision, seed=seed, member_map=member_map)
        self.objective = objective
        self._u = None
        self._eps = None
        self._generator = None

    def set_external_draws(self, z_actions=None, eps=None, z_final=None):
        """Parity hook: `z_actions` [1, 1, N, H, A] are the U(-1, 1) variates of one call; `eps` as in
        CemMpc. Passing z_actions pins the uniforms (no resampling per call)."""
        self._pinned = z_actions is not None
        self._eps = eps
        if z_actions is not None:
            super().set_external_draws(z_actions, eps, None)

    def _draw(self):
        if getattr(self, '_pinned', False):
            return
        if self._u is None:
            a_dim = self.action_space.shape[0]
            self._u = torch.empty((1, 1, self.n_samples, self.horizon, a_dim), dtype=torch.float32,
                                  device='cuda')
            self._generator = torch.Generator(device='cuda')
            self._generator.manual_seed(self.seed)
            CemMpc.set_external_draws(self, self._u, self._eps, None)
        self._u.uniform_(-1.0, 1.0, generator=self._generator)

    def do_generate_action(self, state, seed=None):
        self._draw()
        return super().do_generate_action(state, seed)

    def plan_device(self, states, seed=None, **kw):
        self._draw()
        return super().plan_device(states, seed, **kw)


class RandomMpc(PolicyBase):
    """random_mpc.py:6-17 — a uniform random action; host-only in the reference and here."""

    def __init__(self, action_space):
        super().__init__()
        self.action_space = action_space

    def generate_action(self, state):
        return np.random.uniform(self.action_space.low, self.action_space.high)

    def build(self):
        pass
