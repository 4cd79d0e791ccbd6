"""Known-answer tests that pin the CPU oracle (the reference ships no tests or golden vectors, so
these are derived from its formulas and from the documented semantics of the TF ops it calls)."""
import os

import numpy as np
import pytest

from oracle import philox
from oracle import simba_oracle as so
from tests import helpers


# ---- Philox4x32-10: Random123 known-answer vectors ---------------------------------------------
@pytest.mark.parametrize("ctr,key,expect", [
    ([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
    ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
    ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
     [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]),
])
def test_philox_kat(ctr, key, expect):
    out = philox.philox4x32(np.array(ctr, np.uint32), np.array(key, np.uint32))
    assert [int(v) for v in out] == expect


def test_philox_normals_moments_and_range():
    z = philox.noise_normals(seed=7, iteration=1, horizon=3, rows=np.arange(4000), obs_dim=60)
    assert z.shape == (3, 4000, 60) and z.dtype == np.float32
    assert np.all(np.isfinite(z))
    assert abs(float(z.mean())) < 5e-3 and abs(float(z.std()) - 1.0) < 5e-3
    u = philox.uniform_from_bits(np.array([0, 0xffffffff], np.uint32))
    assert u[0] > 0 and u[1] <= 1.0


def test_philox_streams_are_shard_invariant():
    full = philox.action_normals(5, 2, 32, 4, 2)
    part = philox.action_normals(5, 2, 8, 4, 2, first_candidate=16)
    assert np.array_equal(full[16:24], part)
    rows = np.array([3, 77, 1000])
    a = philox.noise_normals(5, 1, 4, np.arange(1200), 60)
    b = philox.noise_normals(5, 1, 4, rows, 60)
    assert np.array_equal(a[:, rows], b)


# ---- constants of the Bayesian safety test (safe_cem_mpc.py:81,110-120) -------------------------
def test_beta_prior_constants():
    alpha, beta = so.beta_prior(0.5, 0.27)
    assert abs(alpha - 1.2146776406) < 1e-9 and abs(beta - alpha) < 1e-12


@pytest.mark.parametrize("P,expect", [(20, 2), (32, 3), (45, 5), (5, -1)])  # P=5: even zero violations fail the test (never safe)
def test_beta_count_threshold_table(P, expect):
    assert so.beta_count_threshold(P, 0.15) == expect


def test_beta_count_threshold_matches_posterior_compare():
    for P in (4, 5, 20, 32, 45, 100):
        for thr in (0.05, 0.15, 0.3, 0.5):
            c_max = so.beta_count_threshold(P, thr)
            a, b = so.beta_prior(np.float32(0.5), np.float32(0.27))
            counts = np.arange(P + 1, dtype=np.float32)
            post = (np.float32(a) + counts) / (np.float32(a) + np.float32(b) + np.float32(P))
            assert np.array_equal(post <= np.float32(thr), counts <= c_max)


# ---- op semantics -------------------------------------------------------------------------------
def test_softplus_and_zero_network():
    x = np.array([-20.0, -13.0, 0.0, 13.0, 20.0], np.float32)
    ref = np.log1p(np.exp(x.astype(np.float64)))
    assert np.allclose(so.tf_softplus(x), ref, rtol=1e-6)
    O, A, U = 6, 2, 8
    w = [np.zeros((O + A, U), np.float32), np.zeros(U, np.float32),
         np.zeros((U, O), np.float32), np.zeros(O, np.float32),
         np.zeros((U, O), np.float32), np.zeros(O, np.float32)]
    ens = so.MlpEnsemble([w, w])
    mu, var = ens.forward(np.ones((4, O + A), np.float32))
    assert np.all(mu == 0) and np.allclose(var, 0.6932472, atol=1e-7)


def test_scale_small_delta_uses_1_01():
    ens = so.MlpEnsemble([[np.zeros((3, 2), np.float32)] * 1])
    tm = so.TransitionModel(ens, [0.0, 1.0, -1.0], [2.0, 1.000001, 1.0])
    out = tm.scale(np.array([[1.0, 2.01, 0.0]], np.float32))
    assert np.allclose(out, [[0.5, 1.0, 0.5]], atol=1e-6)


def test_top_k_ties_take_lower_index():
    s = np.array([1.0, 3.0, 3.0, 2.0, 3.0, 0.0], np.float32)
    vals, idx = so.tf_top_k(s, 2)
    assert list(idx) == [1, 2]
    vals, idx = so.tf_top_k(s, 4)
    assert list(idx) == [1, 2, 3, 4]


def test_moments_is_population_variance():
    x = np.array([[1.0], [2.0], [3.0], [6.0]], np.float32)
    mean, var = so.tf_moments_axis0(x)
    assert mean[0] == 3.0 and abs(var[0] - 3.5) < 1e-6


def test_split_member_map_and_divisibility():
    c = helpers.workload('tiny')
    pl = helpers.oracle_planner(c)
    assert pl.member_of_row() is None          # split handled by np.split inside forward
    bad = helpers.workload('tiny', E=5)
    with pytest.raises(ValueError):
        helpers.oracle_planner(bad).member_of_row()
    pm = helpers.oracle_planner(bad, member_map='particle').member_of_row()
    assert pm.shape == (bad['P'] * bad['N'],) and pm.max() == 4 and pm.min() == 0
    # when E | P the particle map equals tf.split's contiguous chunks
    pe = helpers.oracle_planner(c, member_map='particle').member_of_row()
    b = c['P'] * c['N']
    assert np.array_equal(pe, np.arange(b) // (b // c['E']))


# ---- scorer (safety_gym.py:110-192) ---------------------------------------------------------------
def test_scorer_closest_distance_reward_cost():
    table, O = so.sensor_offset_table(so.POINTGOAL1_SENSORS), 60
    sc = so.Scorer(None, table)
    obs = np.full((2, O), 0.5, np.float32)
    nxt = obs.copy()
    obs[0, table['goal_lidar']] = 0.9          # dist = 4 * 0.9 = 3.6
    nxt[0, table['goal_lidar']] = 0.8
    nxt[0, 3] = 0.7                            # closest bin -> dist 2.8
    obs[1, table['goal_lidar']] = 0.05         # dist 0.2 <= 0.24 -> goal achieved
    obs[1, table['hazards_lidar'].start] = 0.04  # 0.16 <= 0.2 -> cost
    r, done = sc.reward(obs, nxt)
    assert np.allclose(r[0], 3.6 - 2.8, atol=1e-6) and not done[0]
    assert done[1] and np.isclose(r[1], (0.2 - 2.0) + 1.0, atol=1e-6)
    cost = sc.cost(obs)
    assert list(cost) == [0.0, 1.0]


def test_done_mask_order_differs_between_cem_and_safe():
    """goal reached at t = 0 -> CemMpc counts step-0 reward (incl. +1 bonus), Safe returns 0."""
    c = helpers.workload('tiny')
    c['state'] = c['state'].copy()
    c['state'][c['table']['goal_lidar']] = 0.01          # dist 0.04 <= 0.24 at t = 0
    z, eps, zf = __import__('simba_b200.synthetic', fromlist=['x']).make_draws(
        c['I'], 1, c['N'], c['H'], c['A'], c['P'], c['O'])
    tr_safe, tr_rew = so.Trace(), so.Trace()
    helpers.oracle_planner(c, 'penalty').do_generate_action(c['state'], z[:, 0], eps[:, 0], zf[0], tr_safe)
    helpers.oracle_planner(c, 'reward').do_generate_action(c['state'], z[:, 0], eps[:, 0], zf[0], tr_rew)
    assert np.all(tr_safe[0]['ret'] == 0.0)
    assert np.all(tr_rew[0]['ret'] > 0.5)


# ---- whole plan: fp32 vs fp64 shadow, and the committed golden fixture ----------------------------
def test_plan_fp32_tracks_fp64_shadow():
    from simba_b200 import synthetic
    c = helpers.workload('tiny')
    z, eps, zf = synthetic.make_draws(c['I'], 1, c['N'], c['H'], c['A'], c['P'], c['O'])
    out = {}
    for dt in (np.float32, np.float64):
        tr = so.Trace()
        a, s, n = helpers.oracle_planner(c, 'penalty', dtype=dt).do_generate_action(
            c['state'], z[:, 0], eps[:, 0], zf[0], tr)
        out[dt] = (a, s, tr)
    r32, r64 = out[np.float32][2][0]['ret'], out[np.float64][2][0]['ret']
    assert np.allclose(r32, r64, rtol=1e-4, atol=1e-4)


GOLDEN = os.path.join(os.path.dirname(__file__), 'golden', 'oracle_tiny_plan.npz')


@pytest.mark.parametrize("objective", ['reward', 'penalty', 'least_cost', 'feasible_first'])
def test_oracle_matches_committed_golden(objective):
    """Self-golden: freezes today's oracle outputs (tests/golden/make_golden.py) so later edits of
    the oracle cannot drift silently. NOT a reference-derived fixture (none exist)."""
    from simba_b200 import synthetic
    g = np.load(GOLDEN)
    c = helpers.workload('tiny')
    z, eps, zf = synthetic.make_draws(c['I'], 1, c['N'], c['H'], c['A'], c['P'], c['O'])
    tr = so.Trace()
    a, s, n = helpers.oracle_planner(c, objective).do_generate_action(
        c['state'], z[:, 0], eps[:, 0], zf[0], tr)
    assert np.allclose(a, g[objective + '_action'], rtol=1e-6, atol=1e-7)
    assert np.isclose(s, g[objective + '_score'], rtol=1e-6)
    for it, rec in enumerate(tr):
        assert np.array_equal(rec['elite'], g['%s_elite_%d' % (objective, it)])
        assert np.allclose(rec['mu'], g['%s_mu_%d' % (objective, it)], rtol=1e-6, atol=1e-7)
        assert np.allclose(rec['sigma'], g['%s_sigma_%d' % (objective, it)], rtol=1e-6, atol=1e-7)
