"""Edge cases of the hot path on the GPU, each against the oracle: non-default SafetyGymStateScorer
configurations (safety_gym.py:145-166, :110-143: several constrained lidars, vases only, additive flag,
reward shaping, goal_dist observation), action dimensions other than 2, and extreme planner shapes
(H = 1 / 64, one particle, one candidate, 16 members) — on the fp32, bf16 and wide bf16 kernels."""
import numpy as np
import pytest

from oracle import simba_oracle as so
from tests import helpers
from tests.test_gpu_kernels import _oracle_rows, dev, P

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

CONFIGS = {
    'two_constraints': dict(constrain_vases=True),
    'additive_cost': dict(constrain_indicator=False),                  # one class: cost in {0, 1} either way
    'reward_shaping': dict(reward_distance=3.0, reward_goal=2.0, reward_clip=0.05),
    'vases_only': dict(constrain_vases=True, constrain_hazards=False, vases_size=0.3),
}


def _rows(c, scorer_config, precision):
    from simba_b200 import _lib
    lib = _lib.load()
    pol = helpers.cuda_policy(c, 'penalty', precision=precision, scorer_config=scorer_config)
    pl = pol._ensure_planner()
    pl_o = helpers.oracle_planner(c, 'penalty', scorer_config=scorer_config)
    rng = np.random.default_rng(16)
    acts = rng.uniform(-1, 1, (c['N'], c['H'], c['A'])).astype(np.float32)
    eps = rng.standard_normal((c['H'], c['P'] * c['N'], c['O'])).astype(np.float32)
    B = c['P'] * c['N']
    ret = torch.empty(B, dtype=torch.float32, device='cuda')
    mask = torch.empty(B, dtype=torch.int64, device='cuda')
    csum = torch.empty(B, dtype=torch.float32, device='cuda')
    d_state, d_acts, d_eps = dev(c['state'][None]), dev(acts[None]), dev(eps[None])
    _lib.check(lib.simba_rollout_score(pl, P(d_state), P(d_acts), P(d_eps), 0, 0, None, P(ret), P(mask),
                                       P(csum), None))
    torch.cuda.synchronize()
    traj, cum0, mask0, csum0 = _oracle_rows(c, pl_o, acts, eps, 'penalty')
    return (ret.cpu().numpy(), mask.cpu().numpy().view(np.uint64), csum.cpu().numpy()), (traj, cum0, mask0, csum0)


@pytest.mark.parametrize('name', sorted(CONFIGS))
def test_fp32_rows_with_scorer_config(name):
    cfg = CONFIGS[name]
    c = helpers.workload('c1', N=60)
    (ret, mask, csum), (traj, cum0, mask0, csum0) = _rows(c, cfg, 'fp32')
    sc = so.Scorer(cfg, c['table'])
    frag = np.zeros(len(ret), bool)                      # rows with a hard decision within 1e-4 of its threshold
    for t in range(c['H'] + 1):
        frag |= np.abs(sc.goal_distance_metric(traj[:, t]) - np.float32(0.24)) < 1e-4
        for lidar in ('vases', 'hazards'):
            if sc.c['constrain_' + lidar]:
                d = sc.closest_distance(traj[:, t][:, c['table'][lidar + '_lidar']])
                frag |= np.abs(d - np.float32(sc.c[lidar + '_size'])) < 1e-4
    assert frag.mean() < 0.03
    ok = ~frag
    assert np.allclose(ret[ok], cum0[ok], rtol=1e-4, atol=2e-5)
    assert np.array_equal(mask[ok], mask0[ok])
    assert np.array_equal(csum[ok], csum0[ok])
    if name == 'reward_shaping':
        assert np.abs(cum0).max() <= 0.05 * c['H'] + 1e-6    # every step clipped to +-0.05


@pytest.mark.parametrize('name', sorted(CONFIGS))
@pytest.mark.parametrize('over', [{}, dict(U=256, E=2)])
def test_bf16_rows_with_scorer_config(name, over):
    cfg = CONFIGS[name]
    c = helpers.workload('c1', N=60, **over)
    (ret, mask, csum), (traj, cum0, mask0, csum0) = _rows(c, cfg, 'bf16')
    scale = 3.0 if name == 'reward_shaping' else 1.0
    assert np.all(np.isfinite(ret))
    assert np.abs(ret - cum0).mean() < 1e-2 * scale
    assert (mask == mask0).mean() > 1.0 - 0.002 * c['H'] - 0.01
    assert np.abs(csum - csum0).mean() < 0.05


def test_additive_cost_over_several_classes_is_rejected_loudly():
    """constrain_indicator=False with two constrained classes would need per-step costs of 0..2; the fused
    kernels keep one cost bit per step, so the planner refuses instead of silently saturating."""
    from simba_b200 import SimbaError
    c = helpers.workload('tiny')
    with pytest.raises(SimbaError) as e:
        helpers.cuda_policy(c, 'penalty', precision='fp32',
                            scorer_config=dict(constrain_vases=True, constrain_indicator=False)).build()
    assert e.value.code == -6


@pytest.mark.parametrize('precision', ['fp32', 'bf16'])
def test_goal_dist_observation_instead_of_goal_lidar(precision):
    """observe_goal_dist (safety_gym.py:172-174): the goal distance is one observed scalar,
    dist = max(obs[goal_dist], 0), instead of the closest goal-lidar bin."""
    sensors = dict(accelerometer=3, goal_dist=1, gyro=3, hazards_lidar=16, magnetometer=3, velocimeter=3)
    cfg = dict(observe_goal_lidar=False, observe_goal_dist=True)
    c = helpers.workload('c1', N=60, sensors=sensors)
    st = c['state'].copy()
    st[c['table']['goal_dist']] = 0.6                       # start away from the goal (threshold 0.24)
    c['state'] = st
    (ret, mask, csum), (traj, cum0, mask0, csum0) = _rows(c, cfg, precision)
    assert np.abs(cum0).max() > 1e-3                        # the reward signal is alive
    if precision == 'fp32':
        sc = so.Scorer(cfg, c['table'])
        frag = np.zeros(len(ret), bool)
        for t in range(c['H'] + 1):
            frag |= np.abs(sc.goal_distance_metric(traj[:, t]) - np.float32(0.24)) < 1e-4
            frag |= np.abs(sc.closest_distance(traj[:, t][:, c['table']['hazards_lidar']]) - np.float32(0.2)) < 1e-4
        ok = ~frag
        assert frag.mean() < 0.03
        assert np.allclose(ret[ok], cum0[ok], rtol=1e-4, atol=2e-5)
        assert np.array_equal(mask[ok], mask0[ok]) and np.array_equal(csum[ok], csum0[ok])
    else:
        assert np.abs(ret - cum0).mean() < 1e-2
        assert (mask == mask0).mean() > 1.0 - 0.002 * c['H'] - 0.01


@pytest.mark.parametrize('A,over,precision', [
    (1, {}, 'fp32'), (3, {}, 'fp32'), (4, {}, 'fp32'), (7, dict(sensors='simple'), 'fp32'),
    (1, {}, 'bf16'), (3, {}, 'bf16'), (4, {}, 'bf16'),
    (1, dict(U=256, E=2), 'bf16'), (4, dict(U=256, E=2, sensors='simple'), 'bf16')])
def test_action_dimensions_other_than_two(A, over, precision):
    """act_dim 1..4 on the tensor-core kernels (the action columns share the layer-0 K atom with the
    state) and up to 16 on the fp32 kernel; the reference's robots have 2."""
    from simba_b200 import synthetic
    over = dict(over)
    sensors = synthetic.POINTSIMPLEGOAL1_SENSORS if over.pop('sensors', None) == 'simple' else None
    c = helpers.workload('c1', N=40, A=A, sensors=sensors, **over)
    assert c['A'] == A
    (ret, mask, csum), (traj, cum0, mask0, csum0) = _rows(c, None, precision)
    assert np.all(np.isfinite(ret))
    if precision == 'fp32':
        sc = so.Scorer(None, c['table'])
        frag = np.zeros(len(ret), bool)
        for t in range(c['H'] + 1):
            frag |= np.abs(sc.goal_distance_metric(traj[:, t]) - np.float32(0.24)) < 1e-4
            frag |= np.abs(sc.closest_distance(traj[:, t][:, c['table']['hazards_lidar']]) - np.float32(0.2)) < 1e-4
        ok = ~frag
        assert frag.mean() < 0.03
        assert np.allclose(ret[ok], cum0[ok], rtol=1e-4, atol=2e-5)
        assert np.array_equal(mask[ok], mask0[ok]) and np.array_equal(csum[ok], csum0[ok])
    else:
        assert np.abs(ret - cum0).mean() < 1e-2
        assert (mask == mask0).mean() > 1.0 - 0.002 * c['H'] - 0.01


@pytest.mark.parametrize('A', [1, 3])
def test_whole_plan_with_odd_action_dimension(A):
    """H * A not a multiple of 4 (Philox blocks straddle candidates' ends): full CEM loop vs the oracle."""
    from simba_b200 import _lib, synthetic
    c = helpers.workload('tiny', A=A, H=5)
    z, eps, zf = synthetic.make_draws(c['I'], 1, c['N'], c['H'], c['A'], c['P'], c['O'])
    pol = helpers.cuda_policy(c, 'penalty', precision='fp32')
    pol.set_external_draws(z, eps, zf)
    action, score = pol.do_generate_action(c['state'])
    tr = so.Trace()
    a0, s0, n0 = helpers.oracle_planner(c, 'penalty').do_generate_action(c['state'], z[:, 0], eps[:, 0], zf[0], tr)
    assert action.shape == (A,)
    assert np.array_equal(pol.buffer(_lib.BUF_ELITE, torch.int32).cpu().numpy(), tr[-1]['elite'])
    assert np.allclose(action, a0, rtol=1e-4, atol=1e-6) and np.isclose(score, s0, rtol=1e-4, atol=1e-5)
    # production mode (Philox on the device) is reproducible for this shape too
    p2 = helpers.cuda_policy(c, 'penalty', precision='bf16')
    a1, _ = p2.do_generate_action(c['state'], seed=4)
    a2, _ = p2.do_generate_action(c['state'], seed=4)
    assert np.array_equal(a1, a2) and a1.shape == (A,)


@pytest.mark.parametrize('over', [dict(H=1), dict(H=64, N=16, K=4), dict(P=1, E=1), dict(N=1, K=1, P=8),
                                  dict(P=60, E=3, N=10, K=2), dict(E=16, P=16, N=8, K=2, I=2)])
def test_extreme_planner_shapes(over):
    """Maximum horizon (64: every bit of the cost mask), one step, one particle, one candidate, many members
    — whole plans on the fp32 kernel against the oracle, and the bf16 kernel must stay finite and
    reproducible on the same shapes."""
    from simba_b200 import _lib, synthetic
    c = helpers.workload('tiny', **over)
    z, eps, zf = synthetic.make_draws(c['I'], 1, c['N'], c['H'], c['A'], c['P'], c['O'])
    pol = helpers.cuda_policy(c, 'penalty', precision='fp32')
    pol.set_external_draws(z, eps, zf)
    action, score = pol.do_generate_action(c['state'])
    tr = so.Trace()
    a0, s0, n0 = helpers.oracle_planner(c, 'penalty').do_generate_action(c['state'], z[:, 0], eps[:, 0], zf[0], tr)
    assert np.array_equal(pol.buffer(_lib.BUF_ELITE, torch.int32).cpu().numpy(), tr[-1]['elite'])
    assert np.allclose(action, a0, rtol=1e-4, atol=1e-6) and np.isclose(score, s0, rtol=1e-4, atol=1e-4)
    p2 = helpers.cuda_policy(c, 'penalty', precision='bf16')
    a1, s1 = p2.do_generate_action(c['state'], seed=4)
    a2, s2 = p2.do_generate_action(c['state'], seed=4)
    assert np.all(np.isfinite(a1)) and np.isfinite(s1) and np.array_equal(a1, a2) and s1 == s2


def test_handles_release_their_device_memory():
    """Model, planner and trainer handles own raw cudaMalloc memory (not torch's): creating and dropping
    them repeatedly must give it all back."""
    import gc
    from simba_b200.models import MlpEnsemble
    c = helpers.workload('c1', S=4)
    rng = np.random.default_rng(0)
    x = rng.uniform(0, 1, (c['E'], 64, c['O'] + c['A'])).astype(np.float32)
    y = rng.normal(0, 0.1, (c['E'], 64, c['O'])).astype(np.float32)

    def cycle():
        for precision in ('fp32', 'bf16'):
            pol = helpers.cuda_policy(c, 'penalty', precision=precision, n_states=4)
            pol.do_generate_action(c['state'], seed=1)
            del pol
        ens = MlpEnsemble(c['O'] + c['A'], c['O'], c['E'], mlp_params=dict(n_layers=c['L'], units=c['U']))
        ens.training_step(x, y)
        ens.forward(x[0][:60])                            # 60 rows: tf.split needs B % E == 0
        del ens
        gc.collect()
        torch.cuda.synchronize()

    cycle()                                               # warm-up: context, module loading, torch caches
    free0, _ = torch.cuda.mem_get_info()
    for _ in range(8):
        cycle()
    free1, _ = torch.cuda.mem_get_info()
    assert free0 - free1 < 32 * 1024 * 1024, (free0, free1)


def test_nan_scored_candidates_are_never_elite():
    """A diverged rollout gives a NaN mean return; such candidates must rank below every finite score
    (they used to rank above +inf and fill the elite set)."""
    import ctypes as C
    from simba_b200 import _lib
    lib = _lib.load()
    c = helpers.workload('tiny')
    pol = helpers.cuda_policy(c, 'reward', precision='fp32')
    pl = pol._ensure_planner()
    N, K, H, A = c['N'], c['K'], c['H'], c['A']
    rng = np.random.default_rng(4)
    pairs = np.zeros((N, 2), np.float32)
    pairs[:, 0] = rng.normal(0, 1, N)
    bad = [0, 3, 7, N - 1]
    pairs[bad, 0] = np.nan
    acts = rng.uniform(-1, 1, (1, N, H, A)).astype(np.float32)
    d_pairs, d_acts = torch.from_numpy(pairs).cuda(), torch.from_numpy(acts).cuda()
    elite = torch.empty((K,), dtype=torch.int32, device='cuda')
    best_a = torch.zeros((A,), dtype=torch.float32, device='cuda')
    best_s = torch.full((1,), -np.inf, dtype=torch.float32, device='cuda')
    P = lambda t: C.c_void_p(t.data_ptr())
    _lib.check(lib.simba_select_elites(pl, P(d_pairs), P(d_acts), None, P(elite), None, P(best_a), P(best_s), None))
    torch.cuda.synchronize()
    got = set(elite.cpu().numpy().tolist())
    finite = np.where(np.isfinite(pairs[:, 0]))[0]
    want = set(finite[np.argsort(-pairs[finite, 0], kind='stable')[:K]].tolist())
    assert got == want and not (got & set(bad))
    assert np.isfinite(float(best_s.cpu()[0]))


def test_plans_on_two_streams_share_one_planner_safely():
    """plan_device() on two different torch streams and generate_action() on the planner's own stream use
    one workspace: the library chains them with an event, so the results equal the serial ones."""
    c = helpers.workload('tiny')
    pol = helpers.cuda_policy(c, 'penalty', precision='bf16')
    st = torch.from_numpy(np.ascontiguousarray(c['state'])[None]).cuda()
    ref = [pol.plan_device(st, seed=s)[0].clone() for s in (3, 4, 5)]
    torch.cuda.synchronize()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    outs = []
    for k, (seed, stream) in enumerate(((3, s1), (4, s2), (5, s1))):
        with torch.cuda.stream(stream):
            oa = torch.empty((1, c['A']), dtype=torch.float32, device='cuda')
            pol.plan_device(st, seed, oa)
            outs.append(oa)
    torch.cuda.synchronize()
    for a, b in zip(ref, outs):
        assert torch.equal(a, b)
