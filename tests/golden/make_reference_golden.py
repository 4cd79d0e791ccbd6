"""Generates tests/golden/reference_*.npz by executing the UNMODIFIED reference planner
(/root/reference/simba: CemMpc / SafeCemMpc.do_generate_action, TransitionModel.unfold_sequences,
MlpEnsemble.__call__, SafetyGymStateScorer.reward / cost) over the TensorFlow stand-in of
tests/golden/ref_shim.py, on the same synthetic workloads, weights, state and normal draws the oracle
and the CUDA planner are tested with. Needs /root/reference, so it runs only in the builder
container; the .npz files it writes are committed and travel to the GPU box.

  python tests/golden/make_reference_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'ethz-safe-learning_b200')]

from tests.golden import ref_shim  # noqa: E402

tf = ref_shim.install()

from simba.environment_utils.safety_gym import SafetyGymStateScorer  # noqa: E402  (reference)
from simba.models.transition_model import TransitionModel  # noqa: E402             (reference)
from simba.policies.cem_mpc import CemMpc  # noqa: E402                              (reference)
from simba.policies.safe_cem_mpc import SafeCemMpc  # noqa: E402                     (reference)

import gym  # noqa: E402  (stand-in)
from oracle import simba_oracle as so  # noqa: E402   (only for DEFAULT_SCORER_CONFIG: the un-vendored safety_gym defaults)
from simba_b200 import synthetic  # noqa: E402


class ScorerEnv:
    """What MbrlSafetyGym exposes to the policies (safety_gym.py:30,62-66) around the reference's own
    SafetyGymStateScorer — the MuJoCo wrapper itself cannot be constructed here."""

    def __init__(self, scorer, action_space, observation_space):
        self._scorer, self.action_space, self.observation_space = scorer, action_space, observation_space

    def get_reward(self, obs, acs, *args, **kwargs):
        return self._scorer.reward(obs, *args, **kwargs)

    def get_cost(self, obs, acs, *args, **kwargs):
        return self._scorer.cost(obs)


def build_reference(c, objective, scorer_config=None, threshold=0.15, smoothing=0.0, stddev_threshold=0.0,
                    noise_stddev=0.01, sampling_propagation=True):
    cfg = dict(so.DEFAULT_SCORER_CONFIG)
    cfg.update(scorer_config or {})
    scorer = SafetyGymStateScorer(cfg, c['table'])
    act = gym.spaces.Box([-1.0] * c['A'], [1.0] * c['A'])
    obs = gym.spaces.Box(c['smin'][:c['O']], c['smax'][:c['O']])
    env = ScorerEnv(scorer, act, obs)
    model = TransitionModel('mlp_ensemble', obs, act, True, sampling_propagation,
                            ensemble_size=c['E'], batch_size=64, validation_split=0.2, learning_rate=1e-3,
                            learning_rate_schedule=False, training_steps=1,
                            mlp_params=dict(n_layers=c['L'], units=c['U'], activation='tf.nn.relu', dropout_rate=0.0),
                            train_epochs=1)
    for e in range(c['E']):
        ref_shim.set_member_weights(model.model.ensemble[e], c['weights'][e])
    model.inputs_min = tf.constant(c['smin'], dtype=tf.float32)       # what _fit_statistics leaves behind
    model.inputs_max = tf.constant(c['smax'], dtype=tf.float32)
    common = dict(horizon=c['H'], iterations=c['I'], smoothing=smoothing, n_samples=c['N'], n_elite=c['K'],
                  particles=c['P'], stddev_threshold=stddev_threshold, noise_stddev=noise_stddev)
    if objective == 'reward':
        return CemMpc(model, env, **common)
    return SafeCemMpc(model, env, posterior_mean_threashold=threshold, **common)


def plan_fixture(out, tag, c, objective, z, eps, zf, **kw):
    action, score, rec = run_plan(c, objective, z, eps, zf, **kw)
    out[tag + '_action'] = action.astype(np.float32)
    out[tag + '_score'] = np.float32(score)
    out[tag + '_iterations'] = np.int32(len(rec['elite']))
    for it in range(len(rec['elite'])):
        out['%s_scores_%d' % (tag, it)] = rec['scores'][it]
        out['%s_elite_%d' % (tag, it)] = rec['elite'][it].astype(np.int32)
        out['%s_mean_%d' % (tag, it)] = rec['mean'][it]
        out['%s_var_%d' % (tag, it)] = rec['var'][it]


def run_plan(c, objective, z, eps, zf, **kw):
    """Two passes: the first counts how many iterations the reference runs (its `break`,
    cem_mpc.py:66-67), the second queues exactly those draws followed by z_final."""
    pol = build_reference(c, objective, **kw)
    rec = dict(scores=[], elite=[], mean=[], var=[])
    top_k, moments = tf.nn.top_k, tf.nn.moments

    def top_k_rec(scores, k, sorted=True):
        v, i = top_k(scores, k, sorted=sorted)
        rec['scores'].append(scores.numpy().copy()); rec['elite'].append(np.sort(i.numpy()))
        return v, i

    def moments_rec(x, axes):
        m, v = moments(x, axes)
        rec['mean'].append(m.numpy().copy()); rec['var'].append(v.numpy().copy())
        return m, v

    state = tf.constant(c['state'], dtype=tf.float32)

    def queue(n_iter):
        ref_shim.DRAWS.items.clear()
        for it in range(n_iter):
            ref_shim.DRAWS.push(z[it])                       # tf.random.normal, cem_mpc.py:44
            for t in range(c['H']):
                ref_shim.DRAWS.push(eps[it, t])              # Normal.sample, mlp_ensemble.py:193
        ref_shim.DRAWS.push(zf)                              # tf.random.normal, cem_mpc.py:68

    tf.nn.top_k, tf.nn.moments = top_k_rec, moments_rec
    try:
        n_iter = c['I']
        while True:
            for k in rec:
                rec[k].clear()
            queue(n_iter)
            try:
                action, score = pol.do_generate_action(state)
            except AssertionError:                           # popped a draw of the wrong shape: it broke early
                n_iter = len(rec['elite'])
                continue
            assert not ref_shim.DRAWS.items, "draws left over"
            break
    finally:
        tf.nn.top_k, tf.nn.moments = top_k, moments
    return action.numpy(), float(score), rec


def main():
    # ---- whole plans: tiny (both policies, plus variants) and C1 (both policies) ----------------------
    for cfg in ('tiny', 'c1'):
        c = synthetic.make_workload(cfg)
        z, eps, zf = synthetic.make_draws(c['I'], 1, c['N'], c['H'], c['A'], c['P'], c['O'])
        z, eps, zf = z[:, 0], eps[:, 0], zf[0]
        out = {}
        plan_fixture(out, 'reward', c, 'reward', z, eps, zf)                        # CemMpc
        plan_fixture(out, 'penalty', c, 'penalty', z, eps, zf)                      # SafeCemMpc
        if cfg == 'tiny':
            plan_fixture(out, 'penalty_smooth', c, 'penalty', z, eps, zf, smoothing=0.3, threshold=0.3)
            plan_fixture(out, 'penalty_break', c, 'penalty', z, eps, zf, stddev_threshold=0.9)
            plan_fixture(out, 'reward_mean_prop', c, 'reward', z, eps, zf, sampling_propagation=False)
            plan_fixture(out, 'penalty_vases', c, 'penalty', z, eps, zf,
                         scorer_config=dict(constrain_vases=True))
        path = os.path.join(ROOT, 'tests', 'golden', 'reference_%s_plan.npz' % cfg)
        np.savez_compressed(path, **out)
        print("wrote %s: %d arrays" % (path, len(out)))

    # ---- model + objective pieces on the tiny workload: unfold_sequences, forward, compute_objective,
    #      compute_mean_costs (dead code in the reference, live as the least-cost mode here) -------------
    c = synthetic.make_workload('tiny')
    rng = np.random.default_rng(6)
    acts = rng.uniform(-1, 1, (c['N'], c['H'], c['A'])).astype(np.float32)
    eps = rng.standard_normal((c['H'], c['P'] * c['N'], c['O'])).astype(np.float32)
    out = dict(acts=acts, eps=eps)
    for tag, prop in (('sample', True), ('mean', False)):
        pol = build_reference(c, 'penalty', sampling_propagation=prop)
        ref_shim.DRAWS.items.clear()
        for t in range(c['H']):
            ref_shim.DRAWS.push(eps[t])
        acts_b = tf.tile(tf.constant(acts), (c['P'], 1, 1))
        s0 = tf.broadcast_to(tf.constant(c['state']), (acts_b.shape[0], c['O']))
        traj = pol.model.unfold_sequences(s0, acts_b)
        out['traj_' + tag] = traj.numpy()
        out['objective_safe_' + tag] = pol.compute_objective(traj, acts_b).numpy()
        out['mean_costs_' + tag] = pol.compute_mean_costs(traj, acts_b).numpy()
        pol_r = build_reference(c, 'reward', sampling_propagation=prop)
        out['objective_reward_' + tag] = pol_r.compute_objective(traj, acts_b).numpy()
    x = rng.uniform(0, 1, (c['E'] * 7, c['O'] + c['A'])).astype(np.float32)
    mus, vars_ = pol.model.model.forward(tf.constant(x))
    out['forward_x'], out['forward_mu'], out['forward_var'] = x, mus.numpy(), vars_.numpy()
    out['scaled_x'] = pol.model.scale(tf.constant(x)).numpy()
    path = os.path.join(ROOT, 'tests', 'golden', 'reference_tiny_pieces.npz')
    np.savez_compressed(path, **out)
    print("wrote %s: %d arrays" % (path, len(out)))


if __name__ == '__main__':
    main()
