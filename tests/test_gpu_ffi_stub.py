"""examples/ffi_stub.py — the ctypes + numpy binding of INTEGRATION.md section B — must plan exactly
like the simba_b200 package's SafeCemMpc built from the same weights, scaler, scorer and seed."""
import importlib.util
import os

import numpy as np
import pytest

from tests import helpers

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def test_ffi_stub_equals_package_policy():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location('ffi_stub', os.path.join(root, 'examples', 'ffi_stub.py'))
    stub = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(stub)
    c = helpers.workload('c1')
    pol = helpers.cuda_policy(c, 'penalty', precision='bf16', stddev_threshold=0.25)
    want, _ = pol.do_generate_action(c['state'], seed=42)
    pl = stub.B200Planner(stub.default_lib_path(), c['weights'], c['smin'], c['smax'], stub.pointgoal1_scorer(),
                          [-1.0] * c['A'], [1.0] * c['A'], horizon=c['H'], iterations=c['I'], n_samples=c['N'],
                          n_elite=c['K'], particles=c['P'])
    got = pl.generate_action(c['state'], seed=42)
    pl.close()
    assert np.array_equal(got, want)
    # the struct layouts restated in the stub match the header's (sizes as the package's ctypes see them)
    from simba_b200 import _lib
    import ctypes as C
    assert C.sizeof(stub.PlannerCfg) == C.sizeof(_lib.PlannerConfig)
    assert C.sizeof(stub.Scorer) == C.sizeof(_lib.Scorer) and C.sizeof(stub.ModelCfg) == C.sizeof(_lib.ModelConfig)
