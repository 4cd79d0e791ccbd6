"""Parity tests proper: each CUDA kernel, called through the C-ABI, against the CPU oracle on the
same seeded inputs. Tolerances: bit-exact for integer / index work and for arithmetic whose
operation order is pinned (sampling, scoring); 1e-4 relative (north_star) for fp32 GEMM paths."""
import ctypes as C

import numpy as np
import pytest

from oracle import philox
from oracle import simba_oracle as so
from tests import helpers

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


@pytest.fixture(scope='module')
def lib():
    from simba_b200 import _lib
    l = _lib.load()
    _lib.check(l.simba_device_check())
    return l


def dev(x, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(x))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda()


def P(t):
    return C.c_void_p(0 if t is None else t.data_ptr())


def test_native_library_is_the_one_running(lib):
    import simba_b200._lib as L
    assert str(L.LIB_PATH).endswith('libsimba_b200.so')
    assert b'sm_100a' in lib.simba_version()


@pytest.mark.parametrize("ctr,key,expect", [
    ([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
    ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
    ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
     [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]),
])
def test_philox_kat_on_device(lib, ctr, key, expect):
    from simba_b200 import _lib
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    _lib.check(lib.simba_philox_raw(C.byref(c), C.byref(k), C.byref(o)))
    assert list(o) == expect


@pytest.mark.parametrize("fast,tol", [(0, 2e-6), (1, 2e-3)])
def test_device_normals_match_oracle_contract(lib, fast, tol):
    """fp32 Box-Muller on the device vs the f64-evaluated contract of oracle/philox.py.
    Accurate path: <= 2e-6 absolute (a few ulp); fast MUFU path (bf16 rollout): <= 2e-3."""
    from simba_b200 import _lib
    seed, it, t, s = 0x5EED, 3, 7, 0
    rows, O = 500, 60
    out = torch.empty((rows, O), dtype=torch.float32, device='cuda')
    _lib.check(lib.simba_philox_normals(seed, philox.STREAM_NOISE, it, t, s, 100, rows, O, fast,
                                        P(out), None))
    torch.cuda.synchronize()
    ref = philox.noise_normals(seed, it, t + 1, np.arange(100, 100 + rows), O)[t]
    assert np.max(np.abs(out.cpu().numpy() - ref)) <= tol
    HA = 30
    out = torch.empty((64, HA), dtype=torch.float32, device='cuda')
    _lib.check(lib.simba_philox_normals(seed, philox.STREAM_ACTION, it, 0, s, 0, 64, HA, fast, P(out), None))
    ref = philox.action_normals(seed, it, 64, 15, 2).reshape(64, HA)
    assert np.max(np.abs(out.cpu().numpy() - ref)) <= tol


def test_scorer_eval_bit_exact(lib):
    """SafetyGymStateScorer.reward/.cost (safety_gym.py:110-166): pinned op order -> bit exact."""
    from simba_b200.environment_utils import ScorerEnvironment
    rng = np.random.default_rng(3)
    env = ScorerEnvironment()
    O = env.observation_space.shape[0]
    obs = rng.uniform(-0.2, 1.2, (4096, O)).astype(np.float32)
    nxt = (obs + rng.normal(0, 0.05, obs.shape)).astype(np.float32)
    obs[:50, env.sensor_offset_table['goal_lidar']] = rng.uniform(0.0, 0.07, (50, 16))
    obs[50:100, env.sensor_offset_table['hazards_lidar']] = rng.uniform(0.04, 0.06, (50, 16))
    r, d = env.get_reward(obs, None, nxt)
    c = env.get_cost(obs, None, nxt)
    sc = so.Scorer(None, so.sensor_offset_table(so.POINTGOAL1_SENSORS))
    r0, d0 = sc.reward(obs, nxt)
    c0 = sc.cost(obs)
    assert np.array_equal(r, r0) and np.array_equal(d, d0) and np.array_equal(c, c0)
    assert d0.sum() > 0 and c0.sum() > 0 and (1 - c0).sum() > 0


def test_scale_bit_exact_and_nonfinite_error(lib):
    from simba_b200 import SimbaError
    c = helpers.workload('tiny')
    pol = helpers.cuda_policy(c)
    x = np.random.default_rng(0).uniform(-1, 2, (257, c['O'] + c['A'])).astype(np.float32)
    tm_o = helpers.oracle_planner(c).model
    assert np.array_equal(pol.model.scale(x), tm_o.scale(x))
    # +-inf observation bounds before any fit: the reference would emit NaN; we refuse (q7)
    from simba_b200.environment_utils import ScorerEnvironment
    from simba_b200.models import TransitionModel
    env = ScorerEnvironment()
    tm = TransitionModel('mlp_ensemble', env.observation_space, env.action_space, True, True,
                         ensemble_size=2, mlp_params=dict(n_layers=1, units=16))
    with pytest.raises(SimbaError) as e:
        tm.scale(x)
    assert e.value.code == -8


@pytest.mark.parametrize("cfg,over", [('tiny', {}), ('tiny', dict(E=4, L=3, U=96)), ('c1', {}),
                                      ('tiny', dict(E=2, L=2, U=400))])
def test_ensemble_forward_matches_oracle(lib, cfg, over):
    """MlpEnsemble.forward/__call__ (mlp_ensemble.py:122-132,189-193), fp32 kernel, 1e-4 rel."""
    c = helpers.workload(cfg, **over)
    pol = helpers.cuda_policy(c)
    rng = np.random.default_rng(1)
    B = c['E'] * 37
    x = rng.uniform(0, 1, (B, c['O'] + c['A'])).astype(np.float32)
    eps = rng.standard_normal((B, c['O'])).astype(np.float32)
    mean, std, smp = pol.model.model(x, eps)
    ens = so.MlpEnsemble(c['weights'])
    m0, s0, p0 = ens(x, eps)
    for a, b in ((mean, m0), (std, s0), (smp, p0)):
        assert np.allclose(a, b, rtol=1e-4, atol=1e-5)
    mus, vars_ = pol.model.model.forward(x)
    assert np.allclose(vars_, s0 * s0, rtol=1e-4, atol=1e-7)
    from simba_b200 import SimbaError
    with pytest.raises(SimbaError) as e:                   # tf.split divisibility (mlp_ensemble.py:123)
        pol.model.model.forward(x[:B - 1])
    assert e.value.code == -2


@pytest.mark.parametrize("cfg,sampling", [('tiny', True), ('tiny', False), ('c1', True)])
def test_unfold_sequences_matches_oracle(lib, cfg, sampling):
    """TransitionModel.unfold_sequences (transition_model.py:64-77) with external draws."""
    c = helpers.workload(cfg)
    pol = helpers.cuda_policy(c, sampling_propagation=sampling)
    rng = np.random.default_rng(2)
    B = c['E'] * 50
    s0 = np.tile(c['state'], (B, 1)) + rng.normal(0, 0.01, (B, c['O'])).astype(np.float32)
    acts = rng.uniform(-1, 1, (B, c['H'], c['A'])).astype(np.float32)
    eps = rng.standard_normal((c['H'], B, c['O'])).astype(np.float32)
    traj = pol.model.unfold_sequences(s0.astype(np.float32), acts, eps=eps)
    tm = helpers.oracle_planner(c, sampling_propagation=sampling).model
    ref = tm.unfold_sequences(s0.astype(np.float32), acts, eps)
    assert traj.shape == ref.shape == (B, c['H'] + 1, c['O'])
    assert np.array_equal(traj[:, 0], ref[:, 0])
    assert np.max(np.abs(traj - ref)) < 1e-4
    # predict() = one-step unfold (transition_model.py:52-56)
    pred = pol.model.predict(np.concatenate([s0, acts[:, 0]], axis=1).astype(np.float32))
    assert pred.shape == (B, 2, c['O'])


def test_sample_actions_bit_exact_and_philox(lib):
    """cem_mpc.py:44-48: external z -> bit exact; Philox -> the oracle's contract normals."""
    from simba_b200 import _lib
    c = helpers.workload('c1')
    pol = helpers.cuda_policy(c)
    pl = pol._ensure_planner()
    H, A, N = c['H'], c['A'], c['N']
    rng = np.random.default_rng(4)
    mu = rng.uniform(-0.5, 0.5, (1, H, A)).astype(np.float32)
    sg = rng.uniform(0.1, 1.0, (1, H, A)).astype(np.float32)
    z = rng.standard_normal((1, N, H, A)).astype(np.float32)
    out = torch.empty((1, N, H, A), dtype=torch.float32, device='cuda')
    d_mu, d_sg, d_z = dev(mu), dev(sg), dev(z)           # keep the device copies alive across the launch
    _lib.check(lib.simba_sample_actions(pl, P(d_mu), P(d_sg), P(d_z), 0, 0, None, P(out), None))
    ref = np.minimum(np.maximum(z * sg + mu, np.float32(-1)), np.float32(1))
    assert np.array_equal(out.cpu().numpy(), ref)
    assert (ref == 1.0).sum() > 0 and (ref == -1.0).sum() > 0        # clipping exercised
    _lib.check(lib.simba_sample_actions(pl, P(d_mu), P(d_sg), None, 99, 2, None, P(out), None))
    zz = philox.action_normals(99, 2, N, H, A)[None]
    ref = np.minimum(np.maximum(zz * sg + mu, np.float32(-1)), np.float32(1))
    assert np.max(np.abs(out.cpu().numpy() - ref)) < 5e-6


def _oracle_rows(c, pl_o, acts, eps, objective):
    """Per-row (return, cost mask, cost sum) from oracle trajectories — the contract of k2-k7."""
    P_, N = c['P'], c['N']
    acts_b = np.tile(acts, (P_, 1, 1))
    s0 = np.broadcast_to(c['state'], (P_ * N, c['O']))
    traj = pl_o.model.unfold_sequences(s0, acts_b, eps, pl_o.member_of_row())
    sc = pl_o.reward.__self__._scorer if hasattr(pl_o.reward, '__self__') else None
    B = P_ * N
    cum = np.zeros(B, np.float32); done = np.zeros(B, bool); csum = np.zeros(B, np.float32)
    mask = np.zeros(B, np.uint64)
    done_first = objective in ('penalty', 'feasible_first')
    for t in range(c['H']):
        r, d = pl_o.reward(traj[:, t], None, traj[:, t + 1])
        cost = pl_o.cost(traj[:, t], None, traj[:, t + 1])
        if done_first:
            done = done | d
            mask |= ((cost > 0) & ~done).astype(np.uint64) << np.uint64(t)
            cum = cum + r * (1 - done.astype(np.float32))
        else:
            cum = cum + r * (1 - done.astype(np.float32))
            mask |= ((cost > 0) & ~done).astype(np.uint64) << np.uint64(t)
            done = done | d
        csum = csum + cost
    return traj, cum, mask, csum


@pytest.mark.parametrize("cfg,objective,member_map,over", [
    ('tiny', 'penalty', 'split', {}), ('tiny', 'reward', 'split', {}),
    ('c1', 'penalty', 'split', {}), ('tiny', 'penalty', 'particle', dict(E=3)),
    ('shipped', 'penalty', 'split', dict(N=100, sensors='simple'))])
def test_rollout_score_rows_match_oracle(lib, cfg, objective, member_map, over):
    """Fused rollout + scoring (one launch) vs oracle trajectories scored row by row."""
    from simba_b200 import _lib, synthetic
    over = dict(over)
    sensors = synthetic.POINTSIMPLEGOAL1_SENSORS if over.pop('sensors', None) == 'simple' else None
    c = helpers.workload(cfg, sensors=sensors, **over)
    pol = helpers.cuda_policy(c, objective, member_map=member_map)
    pl = pol._ensure_planner()
    pl_o = helpers.oracle_planner(c, objective, member_map=member_map)
    rng = np.random.default_rng(6)
    acts = rng.uniform(-1, 1, (c['N'], c['H'], c['A'])).astype(np.float32)
    eps = rng.standard_normal((c['H'], c['P'] * c['N'], c['O'])).astype(np.float32)
    B = c['P'] * c['N']
    ret = torch.empty(B, dtype=torch.float32, device='cuda')
    mask = torch.empty(B, dtype=torch.int64, device='cuda')
    csum = torch.empty(B, dtype=torch.float32, device='cuda')
    d_state, d_acts, d_eps = dev(c['state'][None]), dev(acts[None]), dev(eps[None])
    _lib.check(lib.simba_rollout_score(pl, P(d_state), P(d_acts), P(d_eps),
                                       0, 0, None, P(ret), P(mask), P(csum), None))
    torch.cuda.synchronize()
    traj, cum0, mask0, csum0 = _oracle_rows(c, pl_o, acts, eps, objective)
    # rows whose hard decisions (dist <= thr, hazard <= size) sit within 1e-4 of the threshold in
    # the oracle may legitimately flip; they are excluded and must be rare
    sc = so.Scorer(None, c['table'])
    frag = np.zeros(B, bool)
    for t in range(c['H'] + 1):
        frag |= np.abs(sc.goal_distance_metric(traj[:, t]) - np.float32(0.24)) < 1e-4
        frag |= np.abs(sc.closest_distance(traj[:, t][:, c['table']['hazards_lidar']]) - np.float32(0.2)) < 1e-4
    assert frag.mean() < 0.02
    ok = ~frag
    assert np.allclose(ret.cpu().numpy()[ok], cum0[ok], rtol=1e-4, atol=2e-5)
    assert np.array_equal(mask.cpu().numpy().view(np.uint64)[ok], mask0[ok])
    assert np.array_equal(csum.cpu().numpy()[ok], csum0[ok])
    assert mask0.any() and (mask0 == 0).any()


def test_score_reduce_select_refit_pipeline(lib):
    """k8-k10 on synthetic row outputs: integer work bit exact (elite indices, counts), float
    moments within 1e-6; ties resolved to the lower index (tf.nn.top_k)."""
    from simba_b200 import _lib
    c = helpers.workload('c1')
    for objective in ('reward', 'penalty', 'least_cost', 'feasible_first'):
        pol = helpers.cuda_policy(c, objective)
        pl = pol._ensure_planner()
        P_, N, H, A, K = c['P'], c['N'], c['H'], c['A'], c['K']
        rng = np.random.default_rng(8)
        row_ret = np.round(rng.normal(0.5, 0.2, (P_, N)), 1).astype(np.float32)   # many exact ties
        row_mask = (rng.random((P_, N, H)) < 0.08)
        mask = np.zeros((P_, N), np.uint64)
        for t in range(H):
            mask |= row_mask[:, :, t].astype(np.uint64) << np.uint64(t)
        row_csum = row_mask.sum(-1).astype(np.float32)
        pairs = torch.empty((N, 2), dtype=torch.float32, device='cuda')
        d_ret, d_mask, d_csum = dev(row_ret), dev(mask.view(np.int64)), dev(row_csum)
        _lib.check(lib.simba_score_reduce(pl, P(d_ret), P(d_mask), P(d_csum), None, P(pairs), None))
        torch.cuda.synchronize()
        ret0 = row_ret.sum(0, dtype=np.float32) / np.float32(P_)
        counts = row_mask.sum(0).max(-1).astype(np.float32)
        cost0 = dict(reward=np.zeros(N, np.float32), least_cost=row_csum.sum(0, dtype=np.float32) / np.float32(P_)
                     ).get(objective, counts)
        pr = pairs.cpu().numpy()
        assert np.allclose(pr[:, 0], ret0, rtol=1e-6) and np.array_equal(pr[:, 1], cost0)
        # selection on the device's own pairs (so the comparison is exact)
        ret_d, cost_d = pr[:, 0], pr[:, 1]
        pl_o = helpers.oracle_planner(c, objective)
        c_max = so.beta_count_threshold(P_, 0.15)
        if objective == 'reward': scores = ret_d
        elif objective == 'least_cost': scores = -cost_d
        elif objective == 'penalty': scores = ret_d - (cost_d > c_max).astype(np.float32) * np.float32(100)
        else: scores = None
        order = pl_o.rank_order(scores, ret_d, cost_d, cost_d <= c_max)
        acts = rng.uniform(-1, 1, (1, N, H, A)).astype(np.float32)
        elite = torch.empty((K,), dtype=torch.int32, device='cuda')
        sc_out = torch.empty((N,), dtype=torch.float32, device='cuda')
        best_a = torch.zeros((A,), dtype=torch.float32, device='cuda')
        best_s = torch.full((1,), -np.inf, dtype=torch.float32, device='cuda')
        d_acts = dev(acts)
        _lib.check(lib.simba_select_elites(pl, P(pairs), P(d_acts), None, P(elite), P(sc_out), P(best_a),
                                           P(best_s), None))
        assert np.array_equal(elite.cpu().numpy(), np.sort(order[:K]))
        assert np.array_equal(best_a.cpu().numpy(), acts[0, order[0], 0])
        # strict '>' (cem_mpc.py:58): a second call with the same scores must not change best
        best_a.fill_(7.0)
        _lib.check(lib.simba_select_elites(pl, P(pairs), P(d_acts), None, P(elite), None, P(best_a),
                                           P(best_s), None))
        assert np.all(best_a.cpu().numpy() == 7.0)
        # refit
        mu = dev(np.zeros((1, H, A), np.float32)); sg = dev(np.ones((1, H, A), np.float32))
        active = dev(np.ones(1, np.int32)); iters = dev(np.zeros(1, np.int32))
        _lib.check(lib.simba_refit(pl, P(d_acts), P(elite), P(mu), P(sg), P(active), P(iters), None))
        m0, v0 = so.tf_moments_axis0(acts[0][np.sort(order[:K])])
        assert np.allclose(mu.cpu().numpy()[0], m0, rtol=1e-6, atol=1e-7)
        assert np.allclose(sg.cpu().numpy()[0], np.sqrt(v0), rtol=1e-5, atol=1e-7)
        assert int(iters.cpu()[0]) == 1 and int(active.cpu()[0]) == 1       # threshold 0 -> keeps going


def test_select_large_population_with_ties(lib):
    """N = 65536, K = 6554 (C3 shape): radix select + ordered compaction vs a stable sort."""
    from simba_b200 import _lib
    c = helpers.workload('tiny', N=65536, K=6554, P=8, E=2)
    pol = helpers.cuda_policy(c, 'penalty')
    pl = pol._ensure_planner()
    rng = np.random.default_rng(9)
    ret = np.round(rng.normal(0, 1, c['N']), 2).astype(np.float32)
    cost = rng.integers(0, 3, c['N']).astype(np.float32)
    pairs = dev(np.stack([ret, cost], 1))
    acts = torch.zeros((1, c['N'], c['H'], c['A']), dtype=torch.float32, device='cuda')
    elite = torch.empty((c['K'],), dtype=torch.int32, device='cuda')
    best_a = torch.zeros((c['A'],), dtype=torch.float32, device='cuda')
    best_s = torch.full((1,), -np.inf, dtype=torch.float32, device='cuda')
    _lib.check(lib.simba_select_elites(pl, P(pairs), P(acts), None, P(elite), None, P(best_a), P(best_s), None))
    c_max = so.beta_count_threshold(8, 0.15)
    scores = ret - (cost > c_max).astype(np.float32) * np.float32(100)
    order = np.argsort(-scores, kind='stable')
    assert np.array_equal(elite.cpu().numpy(), np.sort(order[:c['K']]))
    assert float(best_s.cpu()[0]) == scores[order[0]]


@pytest.mark.parametrize("N,K,S,levels", [(10007, 2500, 2, 40), (8192, 8192, 1, 5), (9000, 1, 1, 0),
                                          (65536, 6554, 1, 0), (20000, 2047, 3, 3)])
def test_cluster_select_and_refit_match_numpy(lib, N, K, S, levels):
    """The thread-block-cluster kernels used for large populations (N >= 8192: select on 8 CTAs
    per state with DSMEM histograms; K >= 2048: refit on 8 CTAs): uneven slices, several states,
    heavy ties (levels > 0 quantises the returns), K == N and K == 1. Elite indices, best action
    and scores exact; moments within 1e-5 (summation order differs from numpy's)."""
    from simba_b200 import _lib
    c = helpers.workload('tiny', N=N, K=K, P=8, E=2, S=S)
    pol = helpers.cuda_policy(c, 'penalty', n_states=S)
    pl = pol._ensure_planner()
    H, A = c['H'], c['A']
    rng = np.random.default_rng(N + K)
    ret = rng.normal(0, 1, (S, N)).astype(np.float32)
    if levels:
        ret = (np.round(ret * levels) / levels).astype(np.float32)
    cost = rng.integers(0, 3, (S, N)).astype(np.float32)
    pairs = dev(np.stack([ret, cost], -1))
    acts = rng.uniform(-1, 1, (S, N, H, A)).astype(np.float32)
    d_acts = dev(acts)
    elite = torch.empty((S, K), dtype=torch.int32, device='cuda')
    sc_out = torch.empty((S, N), dtype=torch.float32, device='cuda')
    best_a = torch.zeros((S, A), dtype=torch.float32, device='cuda')
    best_s = torch.full((S,), -np.inf, dtype=torch.float32, device='cuda')
    _lib.check(lib.simba_select_elites(pl, P(pairs), P(d_acts), None, P(elite), P(sc_out), P(best_a),
                                       P(best_s), None))
    c_max = so.beta_count_threshold(8, 0.15)
    scores = ret - (cost > c_max).astype(np.float32) * np.float32(100)
    assert np.array_equal(sc_out.cpu().numpy(), scores)
    want = []
    for s in range(S):
        order = np.argsort(-scores[s], kind='stable')
        want.append(np.sort(order[:K]))
        assert np.array_equal(elite.cpu().numpy()[s], want[s]), s
        assert float(best_s.cpu()[s]) == scores[s, order[0]]
        assert np.array_equal(best_a.cpu().numpy()[s], acts[s, order[0], 0])
    mu = dev(np.zeros((S, H, A), np.float32)); sg = dev(np.ones((S, H, A), np.float32))
    active = dev(np.ones(S, np.int32)); iters = dev(np.zeros(S, np.int32))
    _lib.check(lib.simba_refit(pl, P(d_acts), P(elite), P(mu), P(sg), P(active), P(iters), None))
    for s in range(S):
        m0, v0 = so.tf_moments_axis0(acts[s][want[s]].astype(np.float64))
        assert np.allclose(mu.cpu().numpy()[s], m0, rtol=1e-5, atol=1e-6)
        assert np.allclose(sg.cpu().numpy()[s], np.sqrt(v0), rtol=1e-5, atol=1e-6)
    assert np.array_equal(iters.cpu().numpy(), np.ones(S, np.int32))
