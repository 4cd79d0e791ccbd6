"""MpcPolicy — mirrors simba/policies/mpc_policy.py (the MPC base the CEM planners extend)."""
import numpy as np

from ..spaces import is_box_like
from .policy import PolicyBase


class MpcPolicy(PolicyBase):
    def __init__(self, model, environment, horizon, n_samples, particles):
        super().__init__()
        self.model = model
        self.reward = environment.get_reward
        self.action_space = environment.action_space
        assert is_box_like(self.action_space), "Expecting only box as action space."   # :18
        self.horizon = horizon
        self.n_samples = n_samples
        self.particles = particles

    def generate_action(self, state):
        raise NotImplementedError

    def build(self):
        pass

    @property
    def sampling_params(self):
        """mpc_policy.py:45-57 -> (lower_bound, upper_bound, mean, stddev)."""
        space = self.action_space
        bounded = space.is_bounded() if hasattr(space, 'is_bounded') else bool(
            np.all(np.isfinite(space.low)) and np.all(np.isfinite(space.high)))
        if bounded:
            mean = (space.high + space.low) / 2.0
            stddev = (space.high - space.low) / 2.0
            return space.low, space.high, mean, stddev
        return -100, 100, 0.0, 100
