"""bf16 tcgen05 rollout (precision='bf16') against the oracle. This variant has its own, separately
stated tolerance (north_star): weights and activations are rounded to bf16 (2^-9 relative) before
every GEMM (fp32 accumulate), Box-Muller / softplus use the MUFU approximations.

Stated tolerance (external draws, C1 dims, H = 15):
  * per-candidate mean return:   |device - oracle| <= 2e-2 absolute (returns are O(1))
  * per-row cost-mask agreement: >= 1 - 0.002 * H of the rows identical, i.e. 97 % at H = 15 (a row's
    mask has one bit per step; a threshold flip near 0.2 / 0.24 in any step makes the row differ)
  * plan score vs the fp32 kernel: <= 5e-2 absolute
"""
import numpy as np
import pytest

from oracle import simba_oracle as so
from tests import helpers
from tests.test_gpu_kernels import _oracle_rows, dev, P

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.mark.parametrize("cfg,over", [('tiny', {}), ('c1', {}), ('tiny', dict(S=1, N=700, P=8, E=2, K=10)),
                                      # wide models: the streaming tcgen05 kernel (rollout_tc_wide.cu)
                                      ('tiny', dict(U=400, L=2)), ('tiny', dict(U=256, L=3, N=100)),
                                      ('tiny', dict(U=144, L=4)), ('c5', dict(H=10)), ('c5', {}),
                                      # zero-padded widths: 64 / 72 -> 128 (rollout_tc.cu), 130 -> 144 and 200 -> 208 (wide)
                                      ('tiny', dict(U=64)), ('tiny', dict(U=72, L=3)), ('tiny', dict(U=130, L=2)),
                                      ('tiny', dict(U=200, L=3)), ('tiny', dict(U=256, L=1)),
                                      # deepest model whose weights fit in shared memory (one tile per CTA only)
                                      ('tiny', dict(L=5)), ('tiny', dict(L=5, S=1, N=2600, P=8, E=2, K=10, H=3))])
def test_tc_rollout_rows_close_to_oracle(cfg, over):
    from simba_b200 import _lib
    lib = _lib.load()
    c = helpers.workload(cfg, **over)
    pol = helpers.cuda_policy(c, 'penalty', precision='bf16')
    pl = pol._ensure_planner()
    pl_o = helpers.oracle_planner(c, 'penalty')
    rng = np.random.default_rng(6)
    acts = rng.uniform(-1, 1, (c['N'], c['H'], c['A'])).astype(np.float32)
    eps = rng.standard_normal((c['H'], c['P'] * c['N'], c['O'])).astype(np.float32)
    B = c['P'] * c['N']
    ret = torch.zeros(B, dtype=torch.float32, device='cuda')
    mask = torch.zeros(B, dtype=torch.int64, device='cuda')
    csum = torch.zeros(B, dtype=torch.float32, device='cuda')
    d_state, d_acts, d_eps = dev(c['state'][None]), dev(acts[None]), dev(eps[None])
    _lib.check(lib.simba_rollout_score(pl, P(d_state), P(d_acts), P(d_eps), 0, 0, None, P(ret), P(mask),
                                       P(csum), None))
    torch.cuda.synchronize()
    traj, cum0, mask0, csum0 = _oracle_rows(c, pl_o, acts, eps, 'penalty')
    got = ret.cpu().numpy()
    err = np.abs(got - cum0)
    agree = (mask.cpu().numpy().view(np.uint64) == mask0).mean()
    print("bf16 rows: max|dret| %.4g mean %.4g  mask agreement %.4f  ret range [%.3f, %.3f]"
          % (err.max(), err.mean(), agree, cum0.min(), cum0.max()))
    assert np.all(np.isfinite(got))
    assert err.mean() < 1e-2
    assert agree > 1.0 - 0.002 * c['H']
    cand_got = got.reshape(c['P'], c['N']).mean(0)
    cand_ref = cum0.reshape(c['P'], c['N']).mean(0)
    assert np.max(np.abs(cand_got - cand_ref)) < 2e-2


@pytest.mark.parametrize("cfg,over", [('c1', {}), ('c1', dict(U=400, E=2, H=8)), ('c5', dict(H=12, E=5))])
def test_tc_plan_close_to_oracle_and_fp32(cfg, over):
    """Whole plans: the bf16 kernels (units 128: rollout_tc.cu; wide: rollout_tc_wide.cu) against the
    fp32 kernel on the same draws."""
    from simba_b200 import _lib, synthetic
    c = helpers.workload(cfg, **over)
    z, eps, zf = synthetic.make_draws(c['I'], 1, c['N'], c['H'], c['A'], c['P'], c['O'])
    res = {}
    for precision in ('fp32', 'bf16'):
        pol = helpers.cuda_policy(c, 'penalty', precision=precision)
        pol.set_external_draws(z, eps, zf)
        a, s = pol.do_generate_action(c['state'])
        res[precision] = (a, s, pol.buffer(_lib.BUF_MU).cpu().numpy(), pol.buffer(_lib.BUF_SIGMA).cpu().numpy(),
                          pol.buffer(_lib.BUF_ELITE, torch.int32).cpu().numpy())
    a32, s32, mu32, sg32, el32 = res['fp32']
    a16, s16, mu16, sg16, el16 = res['bf16']
    overlap = len(set(el32) & set(el16)) / float(len(el32))
    print("bf16 vs fp32 plan: action %s vs %s, score %.4f vs %.4f, elite overlap %.2f, max|dmu| %.3g"
          % (a16, a32, s16, s32, overlap, np.abs(mu16 - mu32).max()))
    assert np.all(np.isfinite(a16)) and np.isfinite(s16)
    assert abs(s16 - s32) < 5e-2
    # measured overlaps of the final elite sets: 1.0 / 0.93 / 0.93 (after five refits on slightly different
    # candidates); pinned 0.05-0.1 below the measurement instead of a loose 0.6
    assert overlap >= 0.85


def test_tc_philox_plan_runs_and_is_reproducible():
    c = helpers.workload('c1')
    pol = helpers.cuda_policy(c, 'penalty', precision='bf16')
    a1, s1 = pol.do_generate_action(c['state'], seed=5)
    a2, s2 = pol.do_generate_action(c['state'], seed=5)
    a3, s3 = pol.do_generate_action(c['state'], seed=6)
    assert np.array_equal(a1, a2) and s1 == s2
    assert not np.array_equal(a1, a3)
    assert np.all(np.abs(a1) <= 1.05)


def test_tc_two_tiles_per_cta_matches_one_tile():
    """> 148 tiles switches to the 2-tile ping-pong CTA; same rows must give the same numbers as the
    1-tile CTA (external draws, so rows are comparable one to one)."""
    from simba_b200 import _lib
    lib = _lib.load()
    base = helpers.workload('tiny', N=2000, P=10, E=2, K=50, H=5)         # 20000 rows -> 158 tiles
    rng = np.random.default_rng(1)
    acts = rng.uniform(-1, 1, (base['N'], base['H'], base['A'])).astype(np.float32)
    eps = rng.standard_normal((base['H'], base['P'] * base['N'], base['O'])).astype(np.float32)
    outs = []
    for n_take in (base['N'], 500):                                        # 2 tiles/CTA vs 1 tile/CTA
        cc = dict(base); cc['N'] = n_take
        pol = helpers.cuda_policy(cc, 'penalty', precision='bf16')
        pl = pol._ensure_planner()
        Bc = cc['P'] * n_take
        ret = torch.zeros(Bc, dtype=torch.float32, device='cuda')
        mask = torch.zeros(Bc, dtype=torch.int64, device='cuda')
        csum = torch.zeros(Bc, dtype=torch.float32, device='cuda')
        e = eps.reshape(base['H'], base['P'], base['N'], base['O'])[:, :, :n_take].reshape(base['H'], Bc, base['O'])
        d_state, d_acts, d_eps = dev(base['state'][None]), dev(acts[None, :n_take].copy()), dev(e[None].copy())
        _lib.check(lib.simba_rollout_score(pl, P(d_state), P(d_acts), P(d_eps), 0, 0, None, P(ret), P(mask),
                                           P(csum), None))
        torch.cuda.synchronize()
        outs.append(ret.cpu().numpy().reshape(cc['P'], n_take))
    assert np.all(np.isfinite(outs[0]))
    assert np.array_equal(outs[0][:, :500], outs[1])


@pytest.mark.parametrize("draws", ["external", "philox"])
@pytest.mark.parametrize("cfg", ["tiny", "c1", "h1", "h2", "odd"])
def test_tc_cta_pair_variant_is_bit_identical_to_single_cta(draws, cfg, monkeypatch):
    """Plans with at most half as many tiles as SMs run a cluster of two CTAs per tile (head pass split
    by output columns, slices exchanged through DSMEM). Every per-row output must be bit-identical to
    the one-CTA kernel (SIMBA_B200_NO_PAIR=1): same per-element arithmetic, minima are order-free."""
    from simba_b200 import _lib
    lib = _lib.load()
    # h1 / h2: horizons with no / one exchange of next-step inputs; odd: ragged last tile, three members
    over = {'h1': dict(H=1), 'h2': dict(H=2), 'odd': dict(E=3, N=37, P=9, H=5)}.get(cfg, {})
    c = helpers.workload(cfg if not over else 'tiny', **over)
    rng = np.random.default_rng(3)
    acts = rng.uniform(-1, 1, (c['N'], c['H'], c['A'])).astype(np.float32)
    B = c['P'] * c['N']
    eps = rng.standard_normal((c['H'], B, c['O'])).astype(np.float32) if draws == "external" else None
    outs = []
    for no_pair in (False, True):
        if no_pair:
            monkeypatch.setenv("SIMBA_B200_NO_PAIR", "1")
        else:
            monkeypatch.delenv("SIMBA_B200_NO_PAIR", raising=False)
        pol = helpers.cuda_policy(c, 'penalty', precision='bf16')
        pl = pol._ensure_planner()
        ret = torch.zeros(B, dtype=torch.float32, device='cuda')
        mask = torch.zeros(B, dtype=torch.int64, device='cuda')
        csum = torch.zeros(B, dtype=torch.float32, device='cuda')
        d_state, d_acts = dev(c['state'][None]), dev(acts[None].copy())
        d_eps = dev(eps[None].copy()) if eps is not None else None
        for it in (0, 3):                                                  # two iterations' counters
            _lib.check(lib.simba_rollout_score(pl, P(d_state), P(d_acts), P(d_eps) if d_eps is not None else None,
                                               21, it, None, P(ret), P(mask), P(csum), None))
            torch.cuda.synchronize()
            outs.append((ret.cpu().numpy().copy(), mask.cpu().numpy().copy(), csum.cpu().numpy().copy()))
    for a, b in zip(outs[:2], outs[2:]):
        assert np.all(np.isfinite(a[0]))
        for x, y in zip(a, b):
            assert np.array_equal(x, y)
    assert not np.array_equal(outs[0][0], outs[1][0]) or draws == "external"


def test_tc_wide_batched_states_philox_and_early_exit():
    """The wide (units = 400) kernel in production mode: Philox draws on the device, several states
    per call (tiles that straddle states), reproducible, equal to per-state plans fed with the same
    Philox contract draws (bf16 tolerance), and the early break (cem_mpc.py:66-67) skips work."""
    from oracle import philox
    c = helpers.workload('tiny', U=400, L=2, S=3)
    pol_b = helpers.cuda_policy(c, 'penalty', precision='bf16')
    acts_b, scores_b = pol_b.do_generate_action(c['state'], seed=11)
    again, _ = pol_b.do_generate_action(c['state'], seed=11)
    assert np.array_equal(acts_b, again)
    c1 = dict(c); c1['S'] = 1
    for s in range(3):
        pol = helpers.cuda_policy(c1, 'penalty', precision='bf16')
        z = np.stack([philox.action_normals(11, it, c['N'], c['H'], c['A'], state_index=s)
                      for it in range(c['I'])])[:, None]
        B = c['P'] * c['N']
        eps = np.stack([philox.noise_normals(11, it, c['H'], np.arange(B), c['O'], state_index=s)
                        for it in range(c['I'])])[:, None]
        zf = philox.final_normals(11, c['A'], state_index=s)[None]
        pol.set_external_draws(z, eps, zf)
        a, sc = pol.do_generate_action(c['state'][s])
        assert abs(sc - scores_b[s]) < 5e-2
    pol_e = helpers.cuda_policy(c, 'penalty', precision='bf16', stddev_threshold=0.9)
    pol_e.do_generate_action(c['state'], seed=11)
    assert np.all(pol_e.iterations_run == 1)


def _bf16_iteration0(c, eps_transform=None):
    """One CEM iteration's worth of per-candidate (return, cost) pairs from the bf16 kernel and from the
    oracle on identical actions and draws (the same candidates are then comparable one by one; from the
    second iteration on the two planners sample different candidates)."""
    from simba_b200 import _lib
    lib = _lib.load()
    pol = helpers.cuda_policy(c, 'penalty', precision='bf16')
    pl = pol._ensure_planner()
    pl_o = helpers.oracle_planner(c, 'penalty')
    rng = np.random.default_rng(12)
    acts = np.clip(rng.standard_normal((c['N'], c['H'], c['A'])), -1, 1).astype(np.float32)   # iteration-0 sampling law
    eps = rng.standard_normal((c['H'], c['P'] * c['N'], c['O'])).astype(np.float32)
    B = c['P'] * c['N']
    ret = torch.zeros(B, dtype=torch.float32, device='cuda')
    mask = torch.zeros(B, dtype=torch.int64, device='cuda')
    csum = torch.zeros(B, dtype=torch.float32, device='cuda')
    pairs = torch.empty((c['N'], 2), dtype=torch.float32, device='cuda')
    d_state, d_acts, d_eps = dev(c['state'][None]), dev(acts[None]), dev(eps[None])
    _lib.check(lib.simba_rollout_score(pl, P(d_state), P(d_acts), P(d_eps), 0, 0, None, P(ret), P(mask), P(csum), None))
    _lib.check(lib.simba_score_reduce(pl, P(ret), P(mask), P(csum), None, P(pairs), None))
    torch.cuda.synchronize()
    eps_o = eps if eps_transform is None else eps_transform(eps)
    traj, cum0, mask0, csum0 = _oracle_rows(c, pl_o, acts, eps_o, 'penalty')
    acts_b = np.tile(acts, (c['P'], 1, 1))
    ret_o, counts_o, safe_o = pl_o.objective_safe(traj, acts_b)
    return pairs.cpu().numpy(), ret_o, counts_o, safe_o


def _bf16_round(x):
    """round-to-nearest-even to bfloat16, as the kernel's cvt.rn.bf16x2.f32 does"""
    u = np.ascontiguousarray(x, np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(np.float32)


def test_tc_elite_classification_is_margin_aware_exact_vs_oracle():
    """bf16 kernel vs the ORACLE (not vs the fp32 kernel) at C1 dims: every candidate whose oracle score is
    more than 2e-2 away from the K-th best score must land on the same side of the elite cut; the safety
    decision must agree wherever the particle count is not at the threshold; the overlap of the two elite
    sets is pinned to what was measured (14 of 15 at this seed: the odd one out sits within the margin of the
    cut) minus 0.05 instead of a loose 0.6."""
    c = helpers.workload('c1')
    pairs, ret_o, counts_o, safe_o = _bf16_iteration0(c)
    c_max = so.beta_count_threshold(c['P'], 0.15)
    score_o = ret_o - (~safe_o).astype(np.float32) * np.float32(100)
    score_d = pairs[:, 0] - (pairs[:, 1] > c_max).astype(np.float32) * np.float32(100)
    K = c['K']
    order_o = np.argsort(-score_o, kind='stable')
    kth = score_o[order_o[K - 1]]
    elite_o = set(order_o[:K].tolist())
    elite_d = set(np.argsort(-score_d, kind='stable')[:K].tolist())
    clear = np.abs(score_o - kth) > 2e-2                      # decisively inside / outside the elite set
    wrong = [i for i in np.nonzero(clear)[0] if (i in elite_o) != (i in elite_d)]
    overlap = len(elite_o & elite_d) / float(K)
    # safety: identical wherever the violating-particle count is not exactly at / next to the threshold
    decisive = np.abs(counts_o - (c_max + 0.5)) > 1.0
    safe_d = pairs[:, 1] <= c_max
    print("bf16 vs oracle, iteration 0: elite overlap %.3f, %d of %d candidates decisive, %d misclassified; "
          "max|dret| %.4g; safety flips among decisive %d"
          % (overlap, int(clear.sum()), c['N'], len(wrong), np.abs(pairs[:, 0] - ret_o).max(),
             int((safe_d[decisive] != safe_o[decisive]).sum())))
    assert not wrong
    assert np.array_equal(safe_d[decisive], safe_o[decisive])
    assert np.max(np.abs(pairs[:, 0] - ret_o)) < 2e-2
    assert overlap >= 0.88


def test_tc_bf16_rounded_noise_effect_on_returns_is_bounded():
    """The bf16 rollout consumes its N(0,1) draws rounded to bf16. Bound what that
    does to the per-candidate mean return: the oracle fed with the bf16-rounded draws vs the oracle fed
    with the fp32 draws differ by far less than the bf16 tolerance, and the kernel is at least as close to
    the rounded-draw oracle as to the plain one."""
    c = helpers.workload('c1')
    pairs, ret_plain, _, _ = _bf16_iteration0(c)
    _, ret_rounded, _, _ = _bf16_iteration0(c, eps_transform=_bf16_round)
    shift = np.abs(ret_rounded - ret_plain)
    print("bf16-rounded draws: mean |d mean return| %.3g, max %.3g; kernel vs rounded-draw oracle max %.3g, vs plain %.3g"
          % (shift.mean(), shift.max(), np.abs(pairs[:, 0] - ret_rounded).max(), np.abs(pairs[:, 0] - ret_plain).max()))
    # measured: mean 1.3e-4, max 7.7e-3 (one done / cost threshold flip in one of a candidate's 20 rows)
    assert shift.max() < 2e-2 and shift.mean() < 5e-4
    assert np.abs(pairs[:, 0] - ret_rounded).mean() <= np.abs(pairs[:, 0] - ret_plain).mean() + 1e-4
