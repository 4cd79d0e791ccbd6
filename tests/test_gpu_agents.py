"""Batched trajectory collection around the batched planning call."""
import numpy as np
import pytest

from tests import helpers

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


class ToyEnv(object):
    """Deterministic stand-in for a gym environment with the planner's observation layout."""

    def __init__(self, c, seed, episode_bias=0.0):
        from simba_b200.spaces import Box
        self.c = c
        self.rng = np.random.default_rng(seed)
        self.action_space = Box([-1.0] * c['A'], [1.0] * c['A'])
        self.t = 0
        self.obs = None
        self.bias = episode_bias

    def reset(self):
        from simba_b200 import synthetic
        self.t = 0
        self.obs = synthetic.make_state(self.c['sensors'], seed=int(self.rng.integers(1 << 30)))
        return self.obs.copy()

    def step(self, action):
        self.t += 1
        self.obs = np.clip(self.obs + 0.01 * float(np.sum(action)), 0.0, 1.0).astype(np.float32)
        done = self.t >= 5 + int(self.bias)
        return self.obs.copy(), float(action[0]), done, dict(cost=float(action[0] > 0.5))


def test_vectorized_loop_with_one_env_equals_reference_loop():
    from simba_b200 import agents
    c = helpers.workload('tiny')
    runs = []
    for fn in ('single', 'vector'):
        pol = helpers.cuda_policy(c, 'penalty', precision='fp32', seed=9)
        env = ToyEnv(c, seed=1)
        if fn == 'single':
            paths, steps = agents.sample_trajectories(env, pol, 12, 8, action_repeat=2)
        else:
            paths, steps = agents.sample_trajectories_vectorized([env], pol, 12, 8, action_repeat=2)
        runs.append((paths, steps))
    (pa, sa), (pb, sb) = runs
    assert sa == sb and len(pa) == len(pb) and sa >= 12
    for x, y in zip(pa, pb):
        for key in ('observation', 'action', 'reward', 'next_observation', 'terminal'):
            assert np.array_equal(x[key], y[key]), key
        assert [i['cost'] for i in x['info']] == [i['cost'] for i in y['info']]
    assert pa[0]['action'].shape[1] == c['A'] and pa[0]['terminal'][-1] == 1.0


def test_vectorized_loop_batches_four_envs_per_planning_call():
    from simba_b200 import agents
    c = helpers.workload('tiny', S=4)
    pol = helpers.cuda_policy(c, 'penalty', precision='bf16', seed=3, n_states=4)
    envs = [ToyEnv(c, seed=10 + i, episode_bias=i) for i in range(4)]
    calls = []
    plan = pol.generate_action
    pol.generate_action = lambda s: calls.append(np.asarray(s).shape) or plan(s)
    paths, steps = agents.sample_trajectories_vectorized(envs, pol, 40, 50, action_repeat=1)
    assert steps >= 40 and steps == sum(len(p['reward']) for p in paths)
    assert all(shape == (4, c['O']) for shape in calls)
    assert {len(p['reward']) for p in paths} <= {5, 6, 7, 8}          # ToyEnv ends after 5 + bias steps
    for p in paths:
        assert p['observation'].shape == (len(p['reward']), c['O'])
        assert np.all(np.abs(p['action']) <= 1.0 + 0.06)    # final noise N(0, 0.01) is not re-clipped (cem_mpc.py:68)
        assert np.array_equal(p['observation'][1:], p['next_observation'][:-1])
