"""simba_b200 — B200-native CEM-MPC planner behind the reference's `simba.policies` /
`simba.models` interfaces (yardenas/ethz-safe-learning). Python is only the host mirror of the
reference interface; all compute is hand-written sm_100a CUDA in libsimba_b200.so (C-ABI in
include/simba_b200.h). There is no CPU fallback."""
from . import _lib
from ._lib import SimbaError
from .spaces import Box

__all__ = ['SimbaError', 'Box', 'policies', 'models', 'environment_utils', 'synthetic']
