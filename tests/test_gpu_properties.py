"""Property tests (hypothesis) of the small CEM kernels on random shapes: K = 1, K = N, heavy ties,
non-power-of-two sizes, several states per call, every objective — index work must be bit exact."""
import ctypes as C

import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from oracle import simba_oracle as so
from tests import helpers

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

_OBJ = ['reward', 'penalty', 'least_cost', 'feasible_first']
_cache = {}


def _planner(N, K, P, H, S, objective):
    """Planners are cheap but not free; the tiny ensemble is shared."""
    key = (N, K, P, H, S, objective)
    if key not in _cache:
        c = helpers.workload('tiny', N=N, K=K, P=P, H=H, S=S, E=1, L=1)
        pol = helpers.cuda_policy(c, objective, member_map='particle')
        _cache[key] = (c, pol, pol._ensure_planner())
    return _cache[key]


def P_(t):
    return C.c_void_p(t.data_ptr())


@settings(max_examples=40, deadline=None, derandomize=True, suppress_health_check=list(HealthCheck))
@given(N=st.integers(1, 1500), kfrac=st.floats(0.0, 1.0), P=st.integers(1, 40), H=st.integers(1, 64),
       S=st.integers(1, 3), obj=st.sampled_from(_OBJ), seed=st.integers(0, 2 ** 31 - 1),
       levels=st.integers(1, 6))
def test_reduce_select_refit_random_shapes(N, kfrac, P, H, S, obj, seed, levels):
    from simba_b200 import _lib
    lib = _lib.load()
    K = max(1, min(N, int(round(kfrac * N))))
    c, pol, pl = _planner(N, K, P, H, S, obj)
    A = c['A']
    rng = np.random.default_rng(seed)
    # few distinct return levels -> many exact ties; masks with random density
    row_ret = rng.integers(0, levels, (S, P, N)).astype(np.float32) * np.float32(0.25)
    bits = rng.random((S, P, N, H)) < rng.uniform(0.0, 0.3)
    mask = np.zeros((S, P, N), np.uint64)
    for t in range(H):
        mask |= bits[..., t].astype(np.uint64) << np.uint64(t)
    row_csum = bits.sum(-1).astype(np.float32)
    d_ret, d_mask, d_csum = (torch.from_numpy(x).cuda() for x in (row_ret, mask.view(np.int64), row_csum))
    pairs = torch.empty((S, N, 2), dtype=torch.float32, device='cuda')
    _lib.check(lib.simba_score_reduce(pl, P_(d_ret), P_(d_mask), P_(d_csum), None, P_(pairs), None))
    pr = pairs.cpu().numpy()
    # ---- reduce: counts are integers -> exact; means to 1 ulp-ish
    counts = bits.sum(1).max(-1).astype(np.float32)                       # [S, N]
    if obj == 'reward':
        cost0 = np.zeros((S, N), np.float32)
    elif obj == 'least_cost':
        cost0 = None
    else:
        cost0 = counts
    if cost0 is not None:
        assert np.array_equal(pr[..., 1], cost0)
    assert np.allclose(pr[..., 0], row_ret.sum(1, dtype=np.float32) / np.float32(P), rtol=1e-6, atol=1e-7)
    # ---- select on the device's own pairs
    c_max = so.beta_count_threshold(P, 0.15)
    acts = rng.uniform(-1, 1, (S, N, H, A)).astype(np.float32)
    d_acts = torch.from_numpy(acts).cuda()
    elite = torch.empty((S, K), dtype=torch.int32, device='cuda')
    scores = torch.empty((S, N), dtype=torch.float32, device='cuda')
    best_a = torch.zeros((S, A), dtype=torch.float32, device='cuda')
    best_s = torch.full((S,), -np.inf, dtype=torch.float32, device='cuda')
    _lib.check(lib.simba_select_elites(pl, P_(pairs), P_(d_acts), None, P_(elite), P_(scores), P_(best_a),
                                       P_(best_s), None))
    el = elite.cpu().numpy()
    pl_o = helpers.oracle_planner(c, obj, member_map='particle')
    for s in range(S):
        ret_d, cost_d = pr[s, :, 0], pr[s, :, 1]
        if obj == 'reward':
            sc = ret_d
        elif obj == 'least_cost':
            sc = -cost_d
        else:
            sc = ret_d - (cost_d > c_max).astype(np.float32) * np.float32(100)
        order = pl_o.rank_order(None if obj == 'feasible_first' else sc, ret_d, cost_d, cost_d <= c_max)
        assert np.array_equal(el[s], np.sort(order[:K]))
        assert np.array_equal(best_a.cpu().numpy()[s], acts[s, order[0], 0])
        assert float(best_s.cpu()[s]) == sc[order[0]]
    # ---- refit on those elites: population moments
    mu = torch.zeros((S, H, A), dtype=torch.float32, device='cuda')
    sg = torch.ones((S, H, A), dtype=torch.float32, device='cuda')
    active = torch.ones((S,), dtype=torch.int32, device='cuda')
    iters = torch.zeros((S,), dtype=torch.int32, device='cuda')
    _lib.check(lib.simba_refit(pl, P_(d_acts), P_(elite), P_(mu), P_(sg), P_(active), P_(iters), None))
    for s in range(S):
        m0, v0 = so.tf_moments_axis0(acts[s][el[s]])
        assert np.allclose(mu.cpu().numpy()[s], m0, rtol=1e-5, atol=1e-6)
        assert np.allclose(sg.cpu().numpy()[s], np.sqrt(v0), rtol=1e-4, atol=1e-5)
    assert np.all(iters.cpu().numpy() == 1)


@settings(max_examples=15, deadline=None, derandomize=True, suppress_health_check=list(HealthCheck))
@given(N=st.integers(1, 400), H=st.integers(1, 40), A=st.just(2), seed=st.integers(0, 2 ** 31 - 1))
def test_sample_actions_random_shapes_bit_exact(N, H, A, seed):
    from simba_b200 import _lib
    lib = _lib.load()
    c, pol, pl = _planner(N, 1, 2, H, 1, 'reward')
    rng = np.random.default_rng(seed)
    mu = rng.uniform(-1, 1, (1, H, A)).astype(np.float32)
    sg = rng.uniform(0, 2, (1, H, A)).astype(np.float32)
    z = rng.standard_normal((1, N, H, A)).astype(np.float32)
    out = torch.empty((1, N, H, A), dtype=torch.float32, device='cuda')
    d_mu, d_sg, d_z = (torch.from_numpy(x).cuda() for x in (mu, sg, z))
    _lib.check(lib.simba_sample_actions(pl, P_(d_mu), P_(d_sg), P_(d_z), 0, 0, None, P_(out), None))
    ref = np.minimum(np.maximum(z * sg + mu, np.float32(-1)), np.float32(1))
    assert np.array_equal(out.cpu().numpy(), ref)
