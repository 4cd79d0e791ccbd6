"""Minimal stand-in for gym.spaces.Box (gym is not a dependency of the planner).

Anything with `.low`, `.high`, `.shape` and `.is_bounded()` is accepted wherever the reference
takes a Box (simba/policies/mpc_policy.py:17-18, simba/models/transition_model.py:22-29), so a
real gym Box drops in unchanged."""
import numpy as np


class Box:
    def __init__(self, low, high, dtype=np.float32):
        self.low = np.asarray(low, dtype=dtype)
        self.high = np.asarray(high, dtype=dtype)
        if self.low.shape != self.high.shape:
            raise ValueError("low and high must have the same shape")
        self.shape = self.low.shape
        self.dtype = np.dtype(dtype)

    def is_bounded(self):
        return bool(np.all(np.isfinite(self.low)) and np.all(np.isfinite(self.high)))

    def __repr__(self):
        return "Box(%s)" % (self.shape,)


def is_box_like(space):
    return all(hasattr(space, a) for a in ('low', 'high', 'shape'))
