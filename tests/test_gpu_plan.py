"""Whole planning call (CemMpc / SafeCemMpc .generate_action) against the oracle's CEM loop with
identical weights, state and normal draws (north_star parity contract)."""
import numpy as np
import pytest

from oracle import philox
from oracle import simba_oracle as so
from tests import helpers

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _run_pair(cfg, objective, precision='fp32', member_map='split', over=None, **kw):
    from simba_b200 import _lib, synthetic
    c = helpers.workload(cfg, **(over or {}))
    z, eps, zf = synthetic.make_draws(c['I'], 1, c['N'], c['H'], c['A'], c['P'], c['O'])
    pol = helpers.cuda_policy(c, objective, precision=precision, member_map=member_map, **kw)
    pol.set_external_draws(z, eps, zf)
    action, score = pol.do_generate_action(c['state'])
    tr = so.Trace()
    pl_o = helpers.oracle_planner(c, objective, member_map=member_map, **kw)
    a0, s0, n0 = pl_o.do_generate_action(c['state'], z[:, 0], eps[:, 0], zf[0], tr)
    mu = pol.buffer(_lib.BUF_MU).cpu().numpy().reshape(c['H'], c['A'])
    sg = pol.buffer(_lib.BUF_SIGMA).cpu().numpy().reshape(c['H'], c['A'])
    elite = pol.buffer(_lib.BUF_ELITE, torch.int32).cpu().numpy()
    return c, pol, (action, score, mu, sg, elite), (a0, s0, n0, tr)


@pytest.mark.parametrize("cfg,objective", [('tiny', 'reward'), ('tiny', 'penalty'),
                                           ('tiny', 'least_cost'), ('tiny', 'feasible_first'),
                                           ('c1', 'penalty'), ('c1', 'reward')])
def test_plan_matches_oracle_fp32(cfg, objective):
    """fp32 path: returns, costs, refit mean/variance within 1e-4 relative; elite indices exact."""
    c, pol, (action, score, mu, sg, elite), (a0, s0, n0, tr) = _run_pair(cfg, objective)
    assert int(pol.iterations_run[0]) == n0 == c['I']
    assert np.array_equal(elite, tr[-1]['elite'])
    assert np.allclose(mu, tr[-1]['mu'], rtol=1e-4, atol=1e-6)
    assert np.allclose(sg, tr[-1]['sigma'], rtol=1e-4, atol=1e-6)
    assert np.isclose(score, s0, rtol=1e-4, atol=1e-5)
    assert np.allclose(action, a0, rtol=1e-4, atol=1e-6)


def test_plan_early_exit_matches_reference_break():
    """cem_mpc.py:66-67: stop when mean(sigma) <= stddev_threshold, checked after the refit."""
    c, pol, (action, score, mu, sg, elite), (a0, s0, n0, tr) = _run_pair(
        'c1', 'penalty', stddev_threshold=0.55)
    assert 1 <= n0 < c['I']
    assert int(pol.iterations_run[0]) == n0
    assert np.allclose(mu, tr[-1]['mu'], rtol=1e-4, atol=1e-6)
    assert np.allclose(action, a0, rtol=1e-4, atol=1e-6)


def test_plan_smoothing_and_member_map_particle():
    c, pol, (action, score, mu, sg, elite), (a0, s0, n0, tr) = _run_pair(
        'tiny', 'penalty', member_map='particle', over=dict(E=3), smoothing=0.3)
    assert np.array_equal(elite, tr[-1]['elite'])
    assert np.allclose(mu, tr[-1]['mu'], rtol=1e-4, atol=1e-6)
    assert np.allclose(sg, tr[-1]['sigma'], rtol=1e-4, atol=1e-6)


def test_plan_philox_mode_matches_oracle_fed_with_contract_normals():
    """Production RNG: the device draws Philox normals; the oracle consumes oracle/philox.py's
    restatement of the same counters. Same decisions, values within 1e-4."""
    from simba_b200 import _lib
    c = helpers.workload('tiny')
    seed = 0x5EED
    pol = helpers.cuda_policy(c, 'penalty')
    action, score = pol.do_generate_action(c['state'], seed=seed)
    B = c['P'] * c['N']
    z = np.stack([philox.action_normals(seed, it, c['N'], c['H'], c['A']) for it in range(c['I'])])
    eps = np.stack([philox.noise_normals(seed, it, c['H'], np.arange(B), c['O']) for it in range(c['I'])])
    zf = philox.final_normals(seed, c['A'])
    tr = so.Trace()
    a0, s0, n0 = helpers.oracle_planner(c, 'penalty').do_generate_action(c['state'], z, eps, zf, tr)
    elite = pol.buffer(_lib.BUF_ELITE, torch.int32).cpu().numpy()
    assert np.array_equal(elite, tr[-1]['elite'])
    assert np.allclose(action, a0, rtol=1e-4, atol=1e-5) and np.isclose(score, s0, rtol=1e-4, atol=1e-5)
    # a different seed gives a different plan; the same seed reproduces bit for bit
    a1, _ = pol.do_generate_action(c['state'], seed=seed)
    a2, _ = pol.do_generate_action(c['state'], seed=seed + 1)
    assert np.array_equal(a1, action) and not np.array_equal(a2, action)


def test_batched_states_equal_independent_plans():
    """configs[3]: S states per call == S independent single-state plans (bit for bit)."""
    from simba_b200 import synthetic
    c = helpers.workload('tiny', S=5)
    pol_b = helpers.cuda_policy(c, 'penalty')
    acts_b, scores_b = pol_b.do_generate_action(c['state'], seed=11)
    c1 = dict(c); c1['S'] = 1
    for s in range(5):
        pol = helpers.cuda_policy(c1, 'penalty')
        z = np.stack([philox.action_normals(11, it, c['N'], c['H'], c['A'], state_index=s)
                      for it in range(c['I'])])[:, None]
        B = c['P'] * c['N']
        eps = np.stack([philox.noise_normals(11, it, c['H'], np.arange(B), c['O'], state_index=s)
                        for it in range(c['I'])])[:, None]
        zf = philox.final_normals(11, c['A'], state_index=s)[None]
        pol.set_external_draws(z, eps, zf)
        a, sc = pol.do_generate_action(c['state'][s])
        assert np.allclose(a, acts_b[s], rtol=1e-4, atol=1e-5)
        assert np.isclose(sc, scores_b[s], rtol=1e-4, atol=1e-5)


def test_compute_objective_on_materialised_trajectories():
    """mpc_policy.py:26-39 / safe_cem_mpc.py:76-96 public method."""
    c = helpers.workload('tiny')
    rng = np.random.default_rng(3)
    B = c['P'] * c['N']
    for objective in ('reward', 'penalty'):
        pol = helpers.cuda_policy(c, objective)
        pl_o = helpers.oracle_planner(c, objective)
        traj = np.tile(c['state'], (B, c['H'] + 1, 1)).astype(np.float32)
        traj += np.cumsum(rng.normal(0, 0.04, traj.shape), axis=1).astype(np.float32)
        scores = pol.compute_objective(traj, None)
        s0, _, _ = pl_o.compute_scores(traj, np.zeros((B, c['H'], c['A']), np.float32))
        assert np.allclose(scores, s0, rtol=1e-6, atol=1e-6)


def test_errors_are_loud():
    from simba_b200 import SimbaError
    c = helpers.workload('tiny', E=5)                      # 8*24 = 192 rows, not divisible by 5
    pol = helpers.cuda_policy(c, 'penalty')
    with pytest.raises(SimbaError) as e:
        pol.generate_action(c['state'])
    assert e.value.code == -2
    c = helpers.workload('tiny')
    pol = helpers.cuda_policy(c, 'penalty')
    with pytest.raises(ValueError):
        pol.generate_action(np.zeros(7, np.float32))


def test_policy_reconstruction_shares_weights():
    """scripts/tune_cem_policy.py:108-115 re-creates the policy around the same model object."""
    from simba_b200.policies import CemMpc
    c = helpers.workload('tiny')
    pol = helpers.cuda_policy(c, 'reward')
    a = pol.do_generate_action(c['state'], seed=3)[0]
    pol2 = CemMpc(pol.model, pol.environment, horizon=c['H'] + 2, iterations=2, smoothing=0.0,
                  n_samples=c['N'], n_elite=c['K'], particles=c['P'], stddev_threshold=0.0,
                  noise_stddev=0.01, precision='fp32')
    b = pol2.do_generate_action(c['state'], seed=3)[0]
    assert pol2.model.model._handle is pol.model.model._handle
    assert a.shape == b.shape == (c['A'],) and np.all(np.isfinite(b))


@pytest.mark.parametrize("case", ['c1-fp32', 'c1-bf16', 'unstaged', 'ragged', 'reward', 'least_cost', 'feasible_first', 'states'])
def test_fused_update_kernel_is_bit_identical_to_separate_kernels(case):
    """One rank, N <= 1024: reduce + select + refit (+ next sampling | finalize) run as one kernel.
    It must reproduce the separate k8 / k9 / k10 / k1 / k11 kernels bit for bit — on the staged path (rows,
    actions and pre-drawn normals in shared memory; 16-byte and scalar staging loads), on the fallback for
    populations whose rows do not fit, for every objective and with several states per call."""
    import os
    from simba_b200 import _lib
    precision, objective, over = 'fp32', 'penalty', {}
    if case == 'c1-bf16':
        precision = 'bf16'
    elif case == 'unstaged':
        over = dict(N=1000, K=100, P=16, H=6, I=3)        # 16000 rows x 16 B do not fit: direct global loads
    elif case == 'ragged':
        over = dict(N=149, K=13, P=7, H=5, I=3, E=1)      # rows and N * H * A not multiples of 4: scalar staging
    elif case in ('reward', 'least_cost', 'feasible_first'):
        objective = case
    elif case == 'states':
        over = dict(S=3)
    c = helpers.workload('c1' if not over or case == 'states' else 'tiny', **over)
    states = c['state']
    res = []
    for no_fuse in (False, True):
        if no_fuse:
            os.environ['SIMBA_B200_NO_FUSE'] = '1'
        try:
            pol = helpers.cuda_policy(c, objective, precision=precision, stddev_threshold=0.45)
            a, s = pol.do_generate_action(states, seed=123)
            res.append((np.asarray(a), np.asarray(s), pol.buffer(_lib.BUF_MU).cpu().numpy(), pol.buffer(_lib.BUF_SIGMA).cpu().numpy(),
                        pol.buffer(_lib.BUF_ELITE, torch.int32).cpu().numpy(), np.asarray(pol.iterations_run).copy(),
                        pol.launches_per_plan))
        finally:
            os.environ.pop('SIMBA_B200_NO_FUSE', None)
    f, u = res
    assert f[6] < u[6]                                   # fewer launches
    assert np.array_equal(f[0], u[0]) and np.array_equal(f[1], u[1])
    assert np.array_equal(f[2], u[2]) and np.array_equal(f[3], u[3]) and np.array_equal(f[4], u[4])
    assert np.array_equal(f[5], u[5]) and np.all(f[5] >= 1) and np.all(f[5] <= c['I'])


def test_repeated_c1_plans_are_bit_identical():
    """The production path (bf16, device Philox, CUDA graph, programmatic dependent launches, the two-CTA rollout
    with its DSMEM exchange, the staged update kernel) replayed 300 times with one seed: any race in the
    hand-offs shows up as a differing action or score."""
    c = helpers.workload('c1')
    pol = helpers.cuda_policy(c, 'penalty', precision='bf16')
    a0, s0 = pol.do_generate_action(c['state'], seed=77)
    a0, s0 = np.array(a0).copy(), float(np.asarray(s0))
    for _ in range(300):
        a, s = pol.do_generate_action(c['state'], seed=77)
        assert np.array_equal(a0, a) and s0 == float(np.asarray(s))
    a1, _ = pol.do_generate_action(c['state'], seed=78)
    assert not np.array_equal(a0, a1)


def test_wide_model_c5_reward_only_vs_safety_aware():
    """BASELINE configs[4]: 10-member ensemble, 4x400 hidden, horizon 50, on the fp32 kernel (exact
    parity contract). Widths neither tcgen05 kernel covers (> 416) report SIMBA_ERR_UNSUPPORTED for bf16."""
    from simba_b200 import SimbaError, _lib, synthetic
    c = helpers.workload('c5', N=40, K=5, I=2)                  # full model / horizon, small population
    z, eps, zf = synthetic.make_draws(c['I'], 1, c['N'], c['H'], c['A'], c['P'], c['O'])
    for width in (432, 500):          # the A tile + weight ring of the wide kernel stop fitting above 416
        with pytest.raises(SimbaError) as e:
            helpers.cuda_policy(helpers.workload('tiny', U=width), 'penalty', precision='bf16').build()
        assert e.value.code == -6
    with pytest.raises(SimbaError) as e:              # no kernel holds a 1024-wide activation tile
        helpers.cuda_policy(helpers.workload('tiny', U=1024, L=1), 'penalty', precision='fp32').build()
    assert e.value.code == -6
    wide = helpers.cuda_policy(helpers.workload('tiny', U=768, L=1), 'penalty', precision='fp32')
    a_w, s_w = wide.do_generate_action(helpers.workload('tiny', U=768, L=1)['state'], seed=1)
    assert np.all(np.isfinite(a_w)) and np.isfinite(s_w)    # the widest model the fp32 kernel takes
    for objective in ('reward', 'penalty'):
        pol = helpers.cuda_policy(c, objective, precision='fp32')
        pol.set_external_draws(z, eps, zf)
        action, score = pol.do_generate_action(c['state'])
        tr = so.Trace()
        a0, s0, n0 = helpers.oracle_planner(c, objective).do_generate_action(c['state'], z[:, 0], eps[:, 0], zf[0], tr)
        elite = pol.buffer(_lib.BUF_ELITE, torch.int32).cpu().numpy()
        mu = pol.buffer(_lib.BUF_MU).cpu().numpy().reshape(c['H'], c['A'])
        assert np.array_equal(elite, tr[-1]['elite'])
        assert np.allclose(mu, tr[-1]['mu'], rtol=1e-4, atol=1e-6)
        assert np.isclose(score, s0, rtol=1e-4, atol=1e-4)


def test_c3_shape_particle_map_reduced_population():
    """BASELINE configs[2] dims (P = 32, H = 30, E = 5 -> B % E != 0, 'particle' member map) at a
    reduced population so the oracle finishes in seconds."""
    from simba_b200 import SimbaError, _lib, synthetic
    c = helpers.workload('c3', N=96, K=10, I=2)
    with pytest.raises(SimbaError) as e:                        # the reference's tf.split would raise too
        helpers.cuda_policy(c, 'penalty', precision='fp32', member_map='split').build()
    assert e.value.code == -2
    z, eps, zf = synthetic.make_draws(c['I'], 1, c['N'], c['H'], c['A'], c['P'], c['O'])
    pol = helpers.cuda_policy(c, 'penalty', precision='fp32', member_map='particle')
    pol.set_external_draws(z, eps, zf)
    action, score = pol.do_generate_action(c['state'])
    tr = so.Trace()
    a0, s0, n0 = helpers.oracle_planner(c, 'penalty', member_map='particle').do_generate_action(
        c['state'], z[:, 0], eps[:, 0], zf[0], tr)
    elite = pol.buffer(_lib.BUF_ELITE, torch.int32).cpu().numpy()
    assert np.array_equal(elite, tr[-1]['elite'])
    assert np.allclose(action, a0, rtol=1e-4, atol=1e-5) and np.isclose(score, s0, rtol=1e-4, atol=1e-4)
    assert pol.count_threshold == 3                             # P = 32 -> c_max = 3


def test_shipped_config_simple_lidar_layout_both_precisions():
    """The reference's shipped config (PointSimpleGoal1: 5-bin lidars, O = 22, E = 15, H = 8,
    P = 45; config/policies.yaml:11-20) — exercises the generic sensor layout of both kernels."""
    from simba_b200 import _lib, synthetic
    c = helpers.workload('shipped', sensors=synthetic.POINTSIMPLEGOAL1_SENSORS, N=100, K=10, I=3)
    z, eps, zf = synthetic.make_draws(c['I'], 1, c['N'], c['H'], c['A'], c['P'], c['O'])
    tr = so.Trace()
    a0, s0, n0 = helpers.oracle_planner(c, 'penalty').do_generate_action(c['state'], z[:, 0], eps[:, 0], zf[0], tr)
    pol = helpers.cuda_policy(c, 'penalty', precision='fp32')
    pol.set_external_draws(z, eps, zf)
    a, s = pol.do_generate_action(c['state'])
    assert np.array_equal(pol.buffer(_lib.BUF_ELITE, torch.int32).cpu().numpy(), tr[-1]['elite'])
    assert np.allclose(a, a0, rtol=1e-4, atol=1e-5) and np.isclose(s, s0, rtol=1e-4, atol=1e-4)
    polb = helpers.cuda_policy(c, 'penalty', precision='bf16')
    polb.set_external_draws(z, eps, zf)
    ab, sb = polb.do_generate_action(c['state'])
    assert np.all(np.isfinite(ab)) and abs(sb - s0) < 5e-2
    assert pol.count_threshold == 5                             # P = 45 -> c_max = 5


def test_weight_and_scaler_updates_reach_an_already_used_planner():
    """The agent loop refits the model between planning calls (mbrl_agent.py:41-53): new weights /
    statistics must take effect on the next generate_action without rebuilding the policy."""
    from simba_b200 import synthetic
    c = helpers.workload('tiny')
    for precision in ('fp32', 'bf16'):
        pol = helpers.cuda_policy(c, 'penalty', precision=precision)
        a1, s1 = pol.do_generate_action(c['state'], seed=9)
        w2 = synthetic.make_weights(c['E'], c['L'], c['U'], c['O'], c['A'], seed=123)
        for e in range(c['E']):
            pol.model.model.ensemble[e].set_weights(w2[e])
        a2, s2 = pol.do_generate_action(c['state'], seed=9)
        fresh = helpers.cuda_policy(dict(c, weights=w2), 'penalty', precision=precision)
        a3, s3 = fresh.do_generate_action(c['state'], seed=9)
        assert np.array_equal(a2, a3) and s2 == s3 and not np.array_equal(a1, a2)
        lo, hi = c['smin'].copy(), c['smax'].copy()
        hi[:c['O']] = 3.0
        pol.model.set_statistics(lo, hi)
        fresh.model.set_statistics(lo, hi)
        a4, s4 = pol.do_generate_action(c['state'], seed=9)
        a5, s5 = fresh.do_generate_action(c['state'], seed=9)
        assert np.array_equal(a4, a5) and s4 == s5 and not np.array_equal(a4, a2)


def _reference_cases():
    from tests.test_reference_golden import PLANS
    for cfg in ('tiny', 'c1'):
        for tag in PLANS:
            if cfg == 'tiny' or tag in ('reward', 'penalty'):
                yield cfg, tag


@pytest.mark.parametrize("cfg,tag", list(_reference_cases()))
def test_plan_matches_unmodified_reference_code_fp32(cfg, tag):
    """The CUDA planner against fixtures produced by the reference's own CemMpc / SafeCemMpc code
    (tests/golden/reference_*.npz, see tests/golden/make_reference_golden.py), identical weights, state
    and draws: elite indices exact, scores / refit / action within 1e-4 relative (north_star)."""
    import os
    from simba_b200 import _lib, synthetic
    from tests.test_reference_golden import GOLD, PLANS
    g = np.load(os.path.join(GOLD, 'reference_%s_plan.npz' % cfg))
    objective, kw = PLANS[tag]
    c = helpers.workload(cfg)
    z, eps, zf = synthetic.make_draws(c['I'], 1, c['N'], c['H'], c['A'], c['P'], c['O'])
    pol = helpers.cuda_policy(c, objective, precision='fp32', **kw)
    pol.set_external_draws(z, eps, zf)
    action, score = pol.do_generate_action(c['state'])
    n = int(g[tag + '_iterations'])
    assert int(pol.iterations_run[0]) == n
    elite = pol.buffer(_lib.BUF_ELITE, torch.int32).cpu().numpy()
    assert np.array_equal(elite, g['%s_elite_%d' % (tag, n - 1)])
    if kw.get('smoothing', 0.0) == 0.0:
        mu = pol.buffer(_lib.BUF_MU).cpu().numpy().reshape(c['H'], c['A'])
        sg = pol.buffer(_lib.BUF_SIGMA).cpu().numpy().reshape(c['H'], c['A'])
        assert np.allclose(mu, g['%s_mean_%d' % (tag, n - 1)], rtol=1e-4, atol=1e-6)
        assert np.allclose(sg, np.sqrt(g['%s_var_%d' % (tag, n - 1)]), rtol=1e-4, atol=1e-6)
    assert np.isclose(score, g[tag + '_score'], rtol=1e-4, atol=1e-5)
    assert np.allclose(action, g[tag + '_action'], rtol=1e-4, atol=1e-6)
