"""Host-side mirror of the reference interface (no GPU needed)."""
import numpy as np
import torch
import pytest

from simba_b200 import _lib
from tests import helpers


def test_interface_names_and_kwargs_match_reference():
    from simba_b200.policies import CemMpc, SafeCemMpc, MpcPolicy, PolicyBase
    from simba_b200.models import TransitionModel, MlpEnsemble, BaseModel
    c = helpers.workload('tiny')
    pol = helpers.cuda_policy(c, 'penalty')
    assert isinstance(pol, SafeCemMpc) and isinstance(pol, CemMpc) and isinstance(pol, MpcPolicy)
    assert isinstance(pol, PolicyBase) and isinstance(pol.model, BaseModel)
    for attr in ('model', 'reward', 'cost', 'action_space', 'horizon', 'n_samples', 'particles',
                 'iterations', 'smoothing', 'elite', 'stddev_threshold', 'noise_stddev',
                 'posterior_mean_threashold', 'last_action'):
        assert hasattr(pol, attr), attr
    for m in ('generate_action', 'do_generate_action', 'build', 'compute_objective',
              'optimize_for_safety', 'compute_mean_costs'):
        assert callable(getattr(pol, m))
    tm = pol.model
    for attr in ('inputs_min', 'inputs_max', 'scale_features', 'sampling_propagation',
                 'observation_space_dim', 'action_space_dim', 'inputs_dim', 'outputs_dim'):
        assert hasattr(tm, attr), attr
    for m in ('unfold_sequences', 'simulate_trajectories', 'predict', 'scale', 'fit', 'build', 'save', 'load'):
        assert callable(getattr(tm, m))
    assert isinstance(tm.model, MlpEnsemble) and len(tm.model.ensemble) == c['E']
    w = tm.model.ensemble[0].get_weights()
    assert len(w) == 2 * c['L'] + 4 and w[0].shape == (c['O'] + c['A'], c['U']) and w[-2].shape == (c['U'], c['O'])


def test_sampling_params_bounded_and_unbounded():
    from simba_b200.policies import MpcPolicy
    from simba_b200.spaces import Box

    class Env:
        get_reward = staticmethod(lambda *a: None)

    env = Env(); env.action_space = Box([-1, -2], [1, 4])
    lb, ub, mu, sd = MpcPolicy(None, env, 3, 4, 5).sampling_params
    assert list(mu) == [0.0, 1.0] and list(sd) == [1.0, 3.0] and list(lb) == [-1, -2]
    env.action_space = Box([-np.inf, -1], [np.inf, 1])
    assert MpcPolicy(None, env, 3, 4, 5).sampling_params == (-100, 100, 0.0, 100)


def test_scorer_struct_offsets_for_pointgoal1_and_simple():
    from simba_b200.environment_utils import ScorerEnvironment, POINTSIMPLEGOAL1_SENSORS
    sc = ScorerEnvironment()._scorer.scorer_struct()
    assert (sc.goal_begin, sc.goal_end, sc.goal_dist_index) == (3, 19, -1)
    assert sc.n_constraints == 1 and (sc.con_begin[0], sc.con_end[0]) == (22, 38)
    assert abs(sc.con_size[0] - 0.2) < 1e-7 and abs(sc.goal_threshold - 0.24) < 1e-7
    assert sc.lidar_max_dist == 4.0 and sc.reward_clip == 10.0 and sc.constrain_indicator == 1
    env = ScorerEnvironment(POINTSIMPLEGOAL1_SENSORS)
    sc = env._scorer.scorer_struct()
    assert env.observation_space.shape == (22,)
    assert (sc.goal_begin, sc.goal_end) == (3, 8) and (sc.con_begin[0], sc.con_end[0]) == (11, 16)
    env2 = ScorerEnvironment(config=dict(constrain_vases=True))
    sc2 = env2._scorer.scorer_struct()
    assert sc2.n_constraints == 2 and (sc2.con_begin[0], sc2.con_end[0]) == (41, 57)   # vases first (:148-151)


def test_unsupported_environment_and_activation_are_rejected():
    from simba_b200 import SimbaError
    from simba_b200.environment_utils import ScorerEnvironment
    from simba_b200.models import MlpEnsemble
    from simba_b200.policies import CemMpc
    with pytest.raises(SimbaError):
        MlpEnsemble(62, 60, 2, mlp_params=dict(n_layers=1, units=8, activation='tf.nn.tanh'))
    with pytest.raises(SimbaError):
        ScorerEnvironment(config=dict(task='push'))._scorer.scorer_struct()

    class PlainEnv:
        def __init__(self, env):
            self.action_space = env.action_space
            self.observation_space = env.observation_space
        get_reward = staticmethod(lambda *a: None)

    c = helpers.workload('tiny')
    pol = helpers.cuda_policy(c, 'reward')
    bad = CemMpc(pol.model, PlainEnv(pol.environment), 3, 1, 0.0, 8, 2, 2, 0.0, 0.0)
    with pytest.raises(SimbaError):
        bad._config()


def test_planner_config_is_filled_from_policy_kwargs():
    from simba_b200 import _lib
    c = helpers.workload('c1')
    pol = helpers.cuda_policy(c, 'penalty', precision='bf16', threshold=0.15, smoothing=0.25)
    cfg = pol._config()
    assert (cfg.horizon, cfg.iterations, cfg.n_samples, cfg.n_elite, cfg.particles) == (15, 5, 150, 15, 20)
    assert cfg.objective == _lib.OBJ_SAFE_PENALTY and cfg.precision == _lib.PREC_BF16_TC
    assert abs(cfg.posterior_mean_threshold - 0.15) < 1e-7 and abs(cfg.smoothing - 0.25) < 1e-7
    assert list(cfg.act_low)[:2] == [-1.0, -1.0] and list(cfg.init_stddev)[:2] == [1.0, 1.0]
    assert (cfg.prior_mu, abs(cfg.prior_sigma - 0.27) < 1e-7) == (0.5, True)


def test_fit_has_no_cpu_path_but_statistics_work():
    c = helpers.workload('tiny')
    pol = helpers.cuda_policy(c, 'reward')
    tm = pol.model
    if not torch.cuda.is_available():
        with pytest.raises(_lib.SimbaError):          # training is CUDA-only, like planning
            tm.model.fit(np.zeros((4, 62), np.float32), np.zeros((4, 60), np.float32))
    from simba_b200.environment_utils import ScorerEnvironment
    from simba_b200.models import TransitionModel
    env = ScorerEnvironment()
    tm2 = TransitionModel(tm.model, env.observation_space, env.action_space, True, True)
    data = np.random.default_rng(0).uniform(-3, 3, (100, 62)).astype(np.float32)
    tm2._fit_statistics(data)
    assert np.all(np.isfinite(tm2.inputs_min)) and np.all(np.isfinite(tm2.inputs_max))
    assert tm2.inputs_min[3] == 0.0 and tm2.inputs_max[3] == 1.0          # lidar bounds kept
    assert tm2.inputs_min[0] == data[:, 0].min()                          # inf replaced by data min


def test_weight_handoff_npz_round_trip(tmp_path):
    """Keras-ordered variables + scaler statistics through one .npz."""
    from simba_b200.environment_utils import ScorerEnvironment
    from simba_b200.models import TransitionModel
    env = ScorerEnvironment()
    kw = dict(ensemble_size=2, mlp_params=dict(n_layers=2, units=16))
    a = TransitionModel('mlp_ensemble', env.observation_space, env.action_space, True, True, seed=1, **kw)
    a._fit_statistics(np.random.default_rng(0).uniform(-3, 3, (50, 62)).astype(np.float32))
    path = str(tmp_path / 'model.npz')
    a.save(path)
    b = TransitionModel('mlp_ensemble', env.observation_space, env.action_space, True, True, seed=2, **kw)
    assert not np.array_equal(a.model.ensemble[1].get_weights()[0], b.model.ensemble[1].get_weights()[0])
    b.load(path)
    for e in range(2):
        for x, y in zip(a.model.ensemble[e].get_weights(), b.model.ensemble[e].get_weights()):
            assert np.array_equal(x, y)
    assert np.array_equal(a.inputs_min, b.inputs_min) and np.array_equal(a.inputs_max, b.inputs_max)
    c = TransitionModel('mlp_ensemble', env.observation_space, env.action_space, True, True,
                        ensemble_size=2, mlp_params=dict(n_layers=2, units=8))
    with pytest.raises(ValueError):
        c.load(path)
    a.save()          # the reference's argument-less stubs stay no-ops
    a.load()


class _CountingEnv(object):
    """Deterministic environment for the collection loops (no GPU needed)."""

    def __init__(self, length, goal_at=None):
        from simba_b200.spaces import Box
        self.action_space = Box([-1.0, -1.0], [1.0, 1.0])
        self.length, self.goal_at = length, goal_at
        self.resets = 0

    def reset(self):
        self.t = 0
        self.resets += 1
        return np.full(3, float(self.resets), np.float32)

    def step(self, action):
        self.t += 1
        info = dict(cost=1.0, goal_met=(self.goal_at is not None and self.t == self.goal_at))
        return np.full(3, self.resets + 0.01 * self.t, np.float32), float(action[0]), self.t >= self.length, info


class _EchoPolicy(object):
    def __init__(self):
        self.shapes = []

    def generate_action(self, state):
        state = np.asarray(state)
        self.shapes.append(state.shape)
        if state.ndim == 1:
            return np.array([0.5, -0.5], np.float32)
        return np.tile(np.array([0.5, -0.5], np.float32), (state.shape[0], 1))


def test_collection_loops_follow_the_reference_semantics():
    """agent.py:83-153: action_repeat accumulates reward and cost, an episode ends at max length or
    done, goal_met ends the repeat early, collection stops once the finished paths hold batch_size steps;
    the vectorised loop issues one batched call per decision and matches the single-env loop for n = 1."""
    from simba_b200 import agents
    env, pol = _CountingEnv(length=7, goal_at=2), _EchoPolicy()
    path, steps = agents.sample_trajectory(env, pol, max_trajectory_length=100, action_repeat=3)
    assert steps == 7 and path['terminal'][-1] == 1.0 and path['terminal'][:-1].sum() == 0
    # decisions: steps 1-2 (goal_met at t = 2 stops the repeat), 3-5, 6-7 (done)
    assert [i['cost'] for i in path['info']] == [2.0, 3.0, 2.0]
    assert np.allclose(path['reward'], [1.0, 1.5, 1.0]) and path['action'].shape == (3, 2)
    path, steps = agents.sample_trajectory(_CountingEnv(length=50), pol, max_trajectory_length=8, action_repeat=3)
    assert steps == 8 and len(path['reward']) == 3                      # cut at max_trajectory_length
    a_paths, a_steps = agents.sample_trajectories(_CountingEnv(length=5), _EchoPolicy(), 12, 100, 2)
    b_paths, b_steps = agents.sample_trajectories_vectorized([_CountingEnv(length=5)], _EchoPolicy(), 12, 100, 2)
    assert a_steps == b_steps == 15 and len(a_paths) == len(b_paths) == 3
    for x, y in zip(a_paths, b_paths):
        for key in ('observation', 'action', 'reward', 'next_observation', 'terminal'):
            assert np.array_equal(x[key], y[key])
    vec_pol = _EchoPolicy()
    envs = [_CountingEnv(length=4 + i) for i in range(3)]
    paths, steps = agents.sample_trajectories_vectorized(envs, vec_pol, 20, 100, 1)
    assert steps >= 20 and steps == sum(len(p['reward']) for p in paths)
    assert all(shape == (3, 3) for shape in vec_pol.shapes)              # one batched call per decision
    assert sorted({len(p['reward']) for p in paths}) == [4, 5, 6]


def test_random_mpc_and_batch_schedule_host_logic():
    from simba_b200.models import MlpEnsemble
    from simba_b200.policies import RandomMpc
    from simba_b200.spaces import Box
    a = np.array([RandomMpc(Box([-1.0, 0.0], [1.0, 2.0])).generate_action(None) for _ in range(100)])
    assert a[:, 0].min() >= -1 and a[:, 1].min() >= 0 and a[:, 1].max() <= 2
    ens = MlpEnsemble(12, 10, 3, batch_size=16, mlp_params=dict(n_layers=1, units=8))
    np.random.seed(0)
    index, rows = ens.batch_schedule(50, 9)                               # mlp_ensemble.py:172-186
    assert index.shape == (9, 3, 16) and rows.tolist() == [13, 13, 12, 12, 13, 13, 12, 12, 13]
    for e in range(3):                                                    # one permutation per member per pass
        first = np.concatenate([index[s, e, :rows[s]] for s in range(4)])
        assert sorted(first.tolist()) == list(range(50))
    tr, va = ens._split_indices(50)
    assert len(va) == 10 and len(tr) == 40 and sorted(np.concatenate([tr, va]).tolist()) == list(range(50))
