"""RandomShootingMpc (random_shooting_mpc.py:26-37) against the oracle's
CEM loop run with I = 1, K = 1, no final noise and uniform variates."""
import numpy as np
import pytest

from oracle import simba_oracle as so
from tests import helpers

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _policy(c, precision='fp32', seed=0):
    from simba_b200.policies import RandomShootingMpc
    base = helpers.cuda_policy(c, 'reward', precision=precision)
    return RandomShootingMpc(base.model, base.environment, c['H'], None, c['N'], c['P'],
                             precision=precision, seed=seed)


@pytest.mark.parametrize('cfg', ['tiny', 'c1'])
def test_random_shooting_matches_oracle(cfg):
    from simba_b200 import synthetic
    c = helpers.workload(cfg)
    rng = np.random.default_rng(3)
    u = rng.uniform(-1, 1, (1, 1, c['N'], c['H'], c['A'])).astype(np.float32)
    _, eps, _ = synthetic.make_draws(1, 1, c['N'], c['H'], c['A'], c['P'], c['O'])
    pol = _policy(c)
    pol.set_external_draws(u, eps)
    action, score = pol.do_generate_action(c['state'])
    cc = dict(c, I=1, K=1)
    pl = helpers.oracle_planner(cc, 'reward', noise_stddev=0.0)
    tr = so.Trace()
    a0, s0, n0 = pl.do_generate_action(c['state'], u[:, 0], eps[:, 0], np.zeros(c['A'], np.float32), tr)
    best = int(np.argmax(tr[0]['ret']))
    lo, hi = -1.0, 1.0
    assert np.allclose(a0, np.clip(u[0, 0, best, 0], lo, hi), atol=1e-7)     # first action of the best sequence
    assert np.allclose(action, a0, rtol=1e-4, atol=1e-6)
    assert np.isclose(score, s0, rtol=1e-4, atol=1e-5)


def test_random_shooting_resamples_uniformly_and_is_seeded():
    from simba_b200 import _lib
    c = helpers.workload('tiny')
    pol = _policy(c, precision='bf16', seed=5)
    a1 = pol.generate_action(c['state'])
    acts1 = pol.buffer(_lib.BUF_ACTIONS).cpu().numpy()
    a2 = pol.generate_action(c['state'])
    acts2 = pol.buffer(_lib.BUF_ACTIONS).cpu().numpy()
    assert a1.shape == (c['A'],) and np.all(np.abs(a1) <= 1.0)
    assert not np.array_equal(acts1, acts2)                                   # fresh draws per call
    assert acts1.min() >= -1.0 and acts1.max() <= 1.0
    assert abs(acts1.mean()) < 0.15 and 0.25 < acts1.var() < 0.42            # U(-1, 1): var 1/3
    pol_b = _policy(c, precision='bf16', seed=5)
    assert np.array_equal(pol_b.generate_action(c['state']), a1)              # same seed, same plan


def test_random_mpc_is_a_uniform_draw_in_the_box():
    from simba_b200.policies import RandomMpc
    from simba_b200.spaces import Box
    pol = RandomMpc(Box([-1.0, 0.0], [1.0, 2.0]))
    a = np.array([pol.generate_action(None) for _ in range(200)])
    assert a.shape == (200, 2) and a[:, 0].min() >= -1 and a[:, 1].min() >= 0 and a[:, 1].max() <= 2
