"""world_size-2 gloo tests of the multi-GPU design on the CPU (no GPU needed).

The device kernels cannot run here, so the CUDA side of each rank is played by the CPU oracle; what
is tested is the *sharding contract* the C library relies on: (1) Philox rows / candidates are global,
so a rank that rolls out only its candidate shard reproduces exactly the numbers of the unsharded
plan; (2) all-gathering the per-candidate (return, cost) pairs in rank order and running selection +
refit on every rank gives bit-identical replicas equal to the single-process plan; (3) the rendezvous
helpers (unique-id broadcast, shard bounds) work over a real process group."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _sharded_plan(rank, world, port, out_dir):
    sys.path[:0] = [ROOT, os.path.join(ROOT, 'ethz-safe-learning_b200')]
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from oracle import philox
    from oracle import simba_oracle as so
    from simba_b200 import distributed as sd
    from tests import helpers

    c = helpers.workload('tiny')
    seed = 0xABCD
    N, P, H, A, O, I, K = c['N'], c['P'], c['H'], c['A'], c['O'], c['I'], c['K']
    pl = helpers.oracle_planner(c, 'penalty')
    lo, hi = sd.shard_bounds(N, world, rank)
    dt = np.float32
    mu = np.zeros((H, A), dt); sigma = np.ones((H, A), dt)
    c_max = so.beta_count_threshold(P, 0.15)
    trace = []
    for it in range(I):
        # every rank samples ALL candidates (same counters => same arrays, no scatter)
        acts = np.clip(philox.action_normals(seed, it, N, H, A) * sigma + mu, -1, 1).astype(dt)
        # ... but rolls out only its own shard, with GLOBAL row ids r = p * N + i
        cand = np.arange(lo, hi)
        rows = (np.arange(P)[:, None] * N + cand[None, :]).reshape(-1)
        member = rows // (P * N // c['E'])
        eps = philox.noise_normals(seed, it, H, rows, O)
        acts_b = np.tile(acts[lo:hi], (P, 1, 1))
        s0 = np.broadcast_to(c['state'], (rows.size, O))
        traj = pl.model.unfold_sequences(s0, acts_b, eps, member)
        shard = so.CemPlanner(pl.model, so.Environment(so.Scorer(None, c['table']), pl.action_space, None),
                              H, I, 0.0, hi - lo, K, P, 0.0, 0.01, 0.15, so.OBJ_SAFE_PENALTY)
        ret, cost, _ = shard.objective_safe(traj, acts_b)
        mine = torch.from_numpy(np.stack([ret, cost], 1).astype(dt))
        gathered = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(gathered, mine)                       # the one collective of the path
        pairs = torch.cat(gathered, 0).numpy()
        scores = pairs[:, 0] - (pairs[:, 1] > c_max).astype(dt) * dt(100)
        elite = np.sort(np.argsort(-scores, kind='stable')[:K])
        mean, var = so.tf_moments_axis0(acts[elite])
        mu, sigma = mean, np.sqrt(var)
        trace.append((pairs.copy(), elite.copy(), mu.copy(), sigma.copy()))
    np.savez(os.path.join(out_dir, 'rank%d.npz' % rank),
             **{'pairs%d' % i: t[0] for i, t in enumerate(trace)},
             **{'elite%d' % i: t[1] for i, t in enumerate(trace)},
             **{'mu%d' % i: t[2] for i, t in enumerate(trace)},
             **{'sigma%d' % i: t[3] for i, t in enumerate(trace)})
    # rendezvous helper: a 128-byte id produced on rank 0 reaches every rank unchanged
    blob = torch.arange(128, dtype=torch.uint8) if rank == 0 else torch.zeros(128, dtype=torch.uint8)
    dist.broadcast(blob, 0)
    assert blob.tolist() == list(range(128))
    dist.destroy_process_group()


def test_population_sharding_matches_single_process(tmp_path):
    sys.path[:0] = [ROOT, os.path.join(ROOT, 'ethz-safe-learning_b200')]
    from oracle import philox
    from oracle import simba_oracle as so
    from tests import helpers
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_sharded_plan, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r0 = np.load(tmp_path / 'rank0.npz')
    r1 = np.load(tmp_path / 'rank1.npz')
    for k in r0.files:                                        # replicas are bit-identical
        assert np.array_equal(r0[k], r1[k]), k
    # and equal to the unsharded oracle plan fed with the same Philox draws
    c = helpers.workload('tiny')
    seed, B = 0xABCD, c['P'] * c['N']
    z = np.stack([philox.action_normals(seed, it, c['N'], c['H'], c['A']) for it in range(c['I'])])
    eps = np.stack([philox.noise_normals(seed, it, c['H'], np.arange(B), c['O']) for it in range(c['I'])])
    tr = so.Trace()
    helpers.oracle_planner(c, 'penalty').do_generate_action(c['state'], z, eps, np.zeros(c['A'], np.float32), tr)
    for it, rec in enumerate(tr):
        assert np.array_equal(r0['elite%d' % it], rec['elite'])
        assert np.allclose(r0['pairs%d' % it][:, 0], rec['ret'], rtol=1e-6, atol=1e-7)
        assert np.array_equal(r0['pairs%d' % it][:, 1], rec['cost'])
        assert np.allclose(r0['mu%d' % it], rec['mu'], rtol=1e-6, atol=1e-7)


def test_shard_bounds_and_errors():
    sys.path[:0] = [ROOT, os.path.join(ROOT, 'ethz-safe-learning_b200')]
    from simba_b200 import SimbaError
    from simba_b200.distributed import shard_bounds
    assert shard_bounds(65536, 8, 3) == (24576, 32768)
    assert [shard_bounds(1024, 4, r) for r in range(4)] == [(0, 256), (256, 512), (512, 768), (768, 1024)]
    with pytest.raises(SimbaError):
        shard_bounds(150, 4, 0)


class _FakeStatePolicy(object):
    """Stands in for a CUDA policy built with n_states = S / world: what plan_states_sharded needs of it."""

    def __init__(self, n_states):
        self.n_states = n_states

    def do_generate_action(self, states):
        assert states.shape[0] == self.n_states and states.dtype == np.float32 and states.flags['C_CONTIGUOUS']
        return np.stack([states[:, 0] * 2.0 + 1.0, states[:, 1] - states[:, 2]], 1).astype(np.float32), None


def _state_sharded(rank, world, port, out_dir):
    sys.path[:0] = [ROOT, os.path.join(ROOT, 'ethz-safe-learning_b200')]
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from simba_b200 import distributed as sd
    states = np.random.default_rng(5).normal(size=(12, 7))                # same on every rank (float64 on purpose)
    acts = sd.plan_states_sharded(_FakeStatePolicy(12 // world), states)
    np.save(os.path.join(out_dir, 'acts%d.npy' % rank), acts)
    dist.destroy_process_group()


def test_plan_states_sharded_gathers_in_state_order(tmp_path):
    """BASELINE configs[3] at N > 1 (simba_b200.distributed.plan_states_sharded): every rank plans its
    contiguous slice of the states, the actions come back in state order on every rank."""
    world = 2
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_state_sharded, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    states = np.random.default_rng(5).normal(size=(12, 7)).astype(np.float32)
    want = np.stack([states[:, 0] * 2.0 + 1.0, states[:, 1] - states[:, 2]], 1).astype(np.float32)
    for r in range(world):
        assert np.array_equal(np.load(tmp_path / ('acts%d.npy' % r)), want)
