"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference's CEM-MPC planning call.

PINNED TO THE REFERENCE'S OWN CODE, with one stated limit. The reference (yardenas/ethz-safe-learning,
"simba") ships no tests, golden vectors or fixtures, and TensorFlow / TensorFlow-Probability / gym /
safety_gym are not installable here (no wheels, no network). The pin is therefore made by executing
the UNMODIFIED reference sources (CemMpc / SafeCemMpc.do_generate_action, TransitionModel,
MlpEnsemble, SafetyGymStateScorer) over a torch-backed stand-in for the TensorFlow ops they call
(tests/golden/ref_shim.py, generator tests/golden/make_reference_golden.py): the committed
tests/golden/reference_*.npz hold what the reference's control flow, shapes, masks and reductions
produce on our workloads, and tests/test_reference_golden.py holds this file to them. What remains
unpinned is only the arithmetic INSIDE each stock TF op (softplus' range switch, top_k tie order,
moments), restated from the TensorFlow documentation and checked by known-answer tests
(tests/test_oracle.py). oracle/torch_ref.py is a second, differently structured restatement
held to this one at 1e-6 (tests/test_torch_ref.py). Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this file; the product never does.

Every function cites the reference file:line (relative to the reference repo root) it follows.
All arithmetic runs in `dtype` (float32 = the reference's type; float64 = shadow used by the
tests to measure decision margins). Random draws are *inputs* (see oracle/philox.py).
"""
import numpy as np


# --------------------------------------------------------------------------------------------
# gym.spaces.Box stand-in (gym is not installed) — only what the hot path reads.
# --------------------------------------------------------------------------------------------
class Box:
    def __init__(self, low, high):
        self.low = np.asarray(low, dtype=np.float32)
        self.high = np.asarray(high, dtype=np.float32)
        self.shape = self.low.shape

    def is_bounded(self):
        return bool(np.all(np.isfinite(self.low)) and np.all(np.isfinite(self.high)))


# --------------------------------------------------------------------------------------------
# TF op restatements
# --------------------------------------------------------------------------------------------
def tf_softplus(x):
    """tf.math.softplus (TF core/kernels/softplus_op.h): exp / log1p(exp) / identity by range."""
    dt = x.dtype
    threshold = np.log(np.finfo(dt).eps).astype(dt) + dt.type(2.0)
    with np.errstate(over='ignore'):
        x_exp = np.exp(x)
        mid = np.log1p(x_exp)
    return np.where(x > -threshold, x, np.where(x < threshold, x_exp, mid)).astype(dt)


def tf_top_k(scores, k):
    """tf.nn.top_k(sorted=False): the k largest; ties resolved toward the lower index.
    Returned in ascending index order (the reference's order is unspecified — compare as sets)."""
    order = np.argsort(-scores, kind='stable')[:k]
    idx = np.sort(order)
    return scores[idx], idx


def tf_moments_axis0(x):
    """tf.nn.moments(x, axes=0): mean, then mean of squared difference (population variance)."""
    dt = x.dtype
    k = dt.type(x.shape[0])
    mean = np.sum(x, axis=0, dtype=dt) / k
    var = np.sum((x - mean) * (x - mean), axis=0, dtype=dt) / k
    return mean, var


# --------------------------------------------------------------------------------------------
# simba/environment_utils/safety_gym.py — SafetyGymStateScorer (goal task)
# --------------------------------------------------------------------------------------------
DEFAULT_SCORER_CONFIG = dict(
    # simba/environment_utils/safety_gym_registery.py:9-16,27-40
    task='goal', goal_size=0.3, hazards_size=0.2, lidar_max_dist=4, observe_goal_lidar=True,
    observe_goal_dist=False, constrain_hazards=True, constrain_vases=False,
    constrain_pillars=False, constrain_gremlins=False,
    # un-vendored safety_gym Engine.DEFAULT values (recalled; ctor parameters of the drop-in)
    reward_distance=1.0, reward_goal=1.0, reward_clip=10, reward_orientation=False,
    constrain_indicator=True, vases_size=0.1, pillars_size=0.2, gremlins_size=0.1)


def sensor_offset_table(sensor_sizes):
    """Offsets by sorted sensor key — safety_gym.py:17-25."""
    table, offset = {}, 0
    for k in sorted(sensor_sizes):
        table[k] = slice(offset, offset + sensor_sizes[k])
        offset += sensor_sizes[k]
    return table


# PointGoal1 with 16-bin lidars and vases observed (O = 60, BASELINE shape) and the shipped
# PointSimpleGoal1 (5 bins, O = 22) — SURVEY.md §8 a13.
POINTGOAL1_SENSORS = dict(accelerometer=3, goal_lidar=16, gyro=3, hazards_lidar=16, magnetometer=3,
                          vases_lidar=16, velocimeter=3)
POINTSIMPLEGOAL1_SENSORS = dict(accelerometer=3, goal_lidar=5, gyro=3, hazards_lidar=5,
                                magnetometer=3, velocimeter=3)


class Scorer:
    def __init__(self, config, offsets, dtype=np.float32):
        self.c = dict(DEFAULT_SCORER_CONFIG)
        self.c.update(config or {})
        self.off = offsets
        self.dt = np.dtype(dtype)

    def closest_distance(self, lidar):
        """safety_gym.py:188-192."""
        dt = self.dt.type
        d = dt(self.c['lidar_max_dist'])
        v = d - d * (dt(1.0) - lidar)
        return np.min(np.minimum(np.maximum(v, dt(0.0)), d), axis=1)

    def goal_distance_metric(self, obs):
        """safety_gym.py:168-176."""
        if self.c['observe_goal_lidar']:
            return self.closest_distance(obs[:, self.off['goal_lidar']])
        if self.c['observe_goal_dist']:
            return np.maximum(obs[:, self.off['goal_dist']], self.dt.type(0.0)).reshape(-1)
        raise NotImplementedError

    def reward(self, obs, next_obs):
        """safety_gym.py:110-143 (goal branch :114-119, clip :141-142)."""
        dt = self.dt.type
        assert self.c['task'] == 'goal' and not self.c['reward_orientation']
        reward = np.zeros((obs.shape[0],), dtype=self.dt)
        dist = self.goal_distance_metric(obs)
        next_dist = self.goal_distance_metric(next_obs)
        goal_achieved = dist <= dt(self.c['goal_size'] * 0.8)
        reward = reward + ((dist - next_dist) * dt(self.c['reward_distance'])
                           + goal_achieved.astype(self.dt) * dt(self.c['reward_goal']))
        if self.c['reward_clip']:
            clip = dt(self.c['reward_clip'])
            reward = np.minimum(np.maximum(reward, -clip), clip)
        return reward, goal_achieved

    def cost(self, obs):
        """safety_gym.py:145-166."""
        dt = self.dt.type
        cost = np.zeros((obs.shape[0],), dtype=self.dt)
        for name in ('vases', 'hazards', 'pillars', 'gremlins'):      # reference order :148-163
            if self.c['constrain_' + name]:
                dist = self.closest_distance(obs[:, self.off[name + '_lidar']])
                cost = cost + (dist <= dt(self.c[name + '_size'])).astype(self.dt)
        if self.c['constrain_indicator']:
            return (cost > dt(0.0)).astype(self.dt)
        return cost


class Environment:
    """What the policies read from MbrlSafetyGym — safety_gym.py:30,62-66."""

    def __init__(self, scorer, action_space, observation_space):
        self._scorer = scorer
        self.action_space = action_space
        self.observation_space = observation_space

    def get_reward(self, obs, acs, next_obs):
        return self._scorer.reward(obs, next_obs)

    def get_cost(self, obs, acs, next_obs):
        return self._scorer.cost(obs)


# --------------------------------------------------------------------------------------------
# simba/models/mlp_ensemble.py — MlpEnsemble inference
# --------------------------------------------------------------------------------------------
class MlpEnsemble:
    """members[e] = [W_1, b_1, ..., W_L, b_L, W_mu, b_mu, W_var, b_var], Keras kernels [in, out]
    (variable order of mlp_ensemble.py:46-50,28-30)."""

    def __init__(self, members, dtype=np.float32):
        self.dt = np.dtype(dtype)
        self.members = [[np.asarray(a, dtype=self.dt) for a in m] for m in members]
        self.ensemble_size = len(members)

    def mlp(self, e, x):
        """GaussianDistMlp.call :55-61 -> BaseLayer.call :17-22 (ReLU, dropout = identity) ->
        GaussianHead.call :32-34 (var = softplus + 1e-4)."""
        w = self.members[e]
        n_layers = (len(w) - 4) // 2
        for l in range(n_layers):
            x = np.maximum(x @ w[2 * l] + w[2 * l + 1], self.dt.type(0.0))
        mu = x @ w[-4] + w[-3]
        var = tf_softplus(x @ w[-2] + w[-1]) + self.dt.type(1e-4)
        return mu, var

    def forward(self, x):
        """mlp_ensemble.py:122-132 — tf.split rows into E equal chunks, chunk e -> member e."""
        if x.shape[0] % self.ensemble_size != 0:
            raise ValueError("tf.split: batch %d not divisible by ensemble size %d"
                             % (x.shape[0], self.ensemble_size))
        chunks = np.split(x, self.ensemble_size, axis=0)
        outs = [self.mlp(e, c) for e, c in enumerate(chunks)]
        return np.concatenate([o[0] for o in outs], 0), np.concatenate([o[1] for o in outs], 0)

    def forward_rows(self, x, member_of_row):
        """Same per-row arithmetic with an explicit row->member map (the 'particle' map the
        product offers when E does not divide the batch; equals `forward` when it does)."""
        mu = np.empty((x.shape[0], self.members[0][-3].shape[0]), dtype=self.dt)
        var = np.empty_like(mu)
        for e in range(self.ensemble_size):
            sel = member_of_row == e
            if sel.any():
                mu[sel], var[sel] = self.mlp(e, x[sel])
        return mu, var

    def __call__(self, x, eps, member_of_row=None):
        """mlp_ensemble.py:189-193 — Normal(mu, sqrt(var)): mean, stddev, sample = mu + std*eps."""
        mu, var = self.forward(x) if member_of_row is None else self.forward_rows(x, member_of_row)
        std = np.sqrt(var)
        return mu, std, mu + std * eps


# --------------------------------------------------------------------------------------------
# simba/models/transition_model.py
# --------------------------------------------------------------------------------------------
class TransitionModel:
    def __init__(self, ensemble, inputs_min, inputs_max, scale_features=True,
                 sampling_propagation=True, dtype=np.float32):
        self.dt = np.dtype(dtype)
        self.model = ensemble
        self.inputs_min = np.asarray(inputs_min, dtype=self.dt)
        self.inputs_max = np.asarray(inputs_max, dtype=self.dt)
        self.scale_features = scale_features
        self.sampling_propagation = sampling_propagation

    def scale(self, x):
        """transition_model.py:79-87."""
        if not self.scale_features:
            return x
        delta = self.inputs_max - self.inputs_min
        delta = np.where(delta < self.dt.type(1e-5), self.dt.type(1.01), delta)
        return (x - self.inputs_min) / delta

    def unfold_sequences(self, s_0, action_sequences, eps, member_of_row=None):
        """transition_model.py:64-77. eps: [H, B, O] external normal draws. -> [B, H+1, O]."""
        horizon = action_sequences.shape[1]
        traj = np.empty((s_0.shape[0], horizon + 1, s_0.shape[1]), dtype=self.dt)
        s_t = s_0.astype(self.dt).copy()
        for t in range(horizon):
            traj[:, t] = s_t
            x = self.scale(np.concatenate([s_t, action_sequences[:, t]], axis=1))
            mus, _, d_s_t = self.model(x, eps[t].astype(self.dt), member_of_row)
            s_t = s_t + (d_s_t if self.sampling_propagation else mus)
        traj[:, horizon] = s_t
        return traj


# --------------------------------------------------------------------------------------------
# simba/policies — MpcPolicy / CemMpc / SafeCemMpc
# --------------------------------------------------------------------------------------------
def sampling_params(action_space):
    """mpc_policy.py:45-57."""
    if action_space.is_bounded():
        mean = (action_space.high + action_space.low) / 2.0
        stddev = (action_space.high - action_space.low) / 2.0
        return action_space.low, action_space.high, mean, stddev
    return -100, 100, 0.0, 100


def beta_prior(mu=0.5, sigma=0.27):
    """safe_cem_mpc.py:116-117 with the hard-coded mu/sigma of :81."""
    alpha = (((1.0 - mu) / sigma ** 2) - 1.0 / mu) * (mu ** 2)
    beta = alpha * (1.0 / mu - 1)
    return alpha, beta


def beta_count_threshold(particles, threshold, mu=0.5, sigma=0.27, dtype=np.float32):
    """Largest integer count c with fp32 (alpha + c) / (alpha + beta + P) <= threshold, or -1.
    Restates the comparison of safe_cem_mpc.py:118-120 in the reference's own float type."""
    dt = np.dtype(dtype).type
    alpha, beta = beta_prior(dt(mu), dt(sigma))
    alpha, beta = dt(alpha), dt(beta)
    c_max = -1
    for c in range(particles + 1):
        post = (alpha + dt(c)) / (alpha + beta + dt(particles))
        if post <= dt(threshold):
            c_max = c
    return c_max


OBJ_REWARD = 0          # MpcPolicy.compute_objective (CemMpc)             mpc_policy.py:26-39
OBJ_SAFE_PENALTY = 1    # SafeCemMpc.compute_objective (active rule)      safe_cem_mpc.py:76-96
OBJ_LEAST_COST = 2      # SafeCemMpc.optimize_for_safety (dead code)      safe_cem_mpc.py:40-74,98-108
OBJ_FEASIBLE_FIRST = 3  # north_star: feasible ranked by return, then least-violating


class CemPlanner:
    def __init__(self, model, environment, horizon, iterations, smoothing, n_samples, n_elite,
                 particles, stddev_threshold, noise_stddev, posterior_mean_threashold=None,
                 objective=OBJ_REWARD, member_map='split', dtype=np.float32):
        self.dt = np.dtype(dtype)
        self.model = model
        self.reward = environment.get_reward
        self.cost = environment.get_cost
        self.action_space = environment.action_space
        self.horizon, self.iterations, self.smoothing = horizon, iterations, smoothing
        self.n_samples, self.elite, self.particles = n_samples, n_elite, particles
        self.stddev_threshold, self.noise_stddev = stddev_threshold, noise_stddev
        self.posterior_mean_threashold = posterior_mean_threashold
        self.objective = objective
        self.member_map = member_map

    # -- row -> member ------------------------------------------------------------------------
    def member_of_row(self):
        """'split': tf.split of the particle-major batch (mlp_ensemble.py:123, cem_mpc.py:49-51):
        member(r) = r // (B/E). 'particle': member(p) = floor(p*E/P) (equal when E | P)."""
        b = self.particles * self.n_samples
        e = self.model.model.ensemble_size
        if self.member_map == 'split':
            if b % e:
                raise ValueError("batch %d not divisible by ensemble size %d" % (b, e))
            return None
        p = np.arange(b) // self.n_samples
        return (p * e) // self.particles

    # -- objectives ---------------------------------------------------------------------------
    def objective_reward(self, traj, acts):
        """mpc_policy.py:26-39: cum += r * (1 - done_prev); THEN done |= dones."""
        dt = self.dt
        cum = np.zeros((traj.shape[0],), dtype=dt)
        done = np.zeros((traj.shape[0],), dtype=bool)
        for t in range(traj.shape[1] - 1):
            reward, dones = self.reward(traj[:, t], acts[:, t], traj[:, t + 1])
            cum = cum + reward * (dt.type(1.0) - done.astype(dt))
            done = np.logical_or(dones, done)
        per_sample = cum.reshape(self.particles, self.n_samples)
        return np.sum(per_sample, axis=0, dtype=dt) / dt.type(self.particles), None

    def objective_safe(self, traj, acts):
        """safe_cem_mpc.py:76-96 + :110-120: done |= dones FIRST; cost masked by done; Beta test per
        step over the particle counts; cum += r * (1 - done). Returns (scores, max_t counts)."""
        dt = self.dt
        b = self.n_samples * self.particles
        cum = np.zeros((b,), dtype=dt)
        done = np.zeros((b,), dtype=bool)
        safe = np.ones((self.n_samples,), dtype=bool)
        max_counts = np.zeros((self.n_samples,), dtype=dt)
        alpha, beta = beta_prior(dt.type(0.5), dt.type(0.27))
        for t in range(traj.shape[1] - 1):
            reward, dones = self.reward(traj[:, t], acts[:, t], traj[:, t + 1])
            done = np.logical_or(dones, done)
            cost = self.cost(traj[:, t], acts[:, t], traj[:, t + 1]) * (dt.type(1.0) - done.astype(dt))
            counts = np.sum(cost.reshape(self.particles, self.n_samples), axis=0, dtype=dt)
            posterior_mean = (dt.type(alpha) + counts) / (dt.type(alpha) + dt.type(beta)
                                                        + dt.type(self.particles))
            safe = np.logical_and(posterior_mean <= dt.type(self.posterior_mean_threashold), safe)
            max_counts = np.maximum(max_counts, counts)
            cum = cum + reward * (dt.type(1.0) - done.astype(dt))
        ret = np.sum(cum.reshape(self.particles, self.n_samples), axis=0, dtype=dt) / dt.type(self.particles)
        return ret, max_counts, safe

    def mean_costs(self, traj, acts):
        """safe_cem_mpc.py:98-108 (no done mask)."""
        dt = self.dt
        cum = np.zeros((traj.shape[0],), dtype=dt)
        for t in range(traj.shape[1] - 1):
            cum = cum + self.cost(traj[:, t], acts[:, t], traj[:, t + 1])
        return np.sum(cum.reshape(self.particles, self.n_samples), axis=0, dtype=dt) / dt.type(self.particles)

    def compute_scores(self, traj, acts):
        """-> (scores[N] used for ranking, ret[N], cost[N]) — (ret, cost) is the per-candidate pair
        the multi-GPU path all-gathers."""
        dt = self.dt
        if self.objective == OBJ_REWARD:
            ret, _ = self.objective_reward(traj, acts)
            return ret, ret, np.zeros_like(ret)
        if self.objective == OBJ_LEAST_COST:
            mc = self.mean_costs(traj, acts)
            ret, _ = self.objective_reward(traj, acts)
            return -mc, ret, mc
        ret, max_counts, safe = self.objective_safe(traj, acts)
        if self.objective == OBJ_SAFE_PENALTY:
            return ret - np.logical_not(safe).astype(dt) * dt.type(100.0), ret, max_counts
        # feasible-first: feasible candidates ranked by return above every infeasible one;
        # infeasible ones ranked by fewest violations, then by return (see rank_keys).
        return None, ret, max_counts

    def rank_order(self, scores, ret, cost, safe_mask):
        """Total order used by elite selection: larger key first, ties -> lower index."""
        if self.objective != OBJ_FEASIBLE_FIRST:
            return np.argsort(-scores, kind='stable')
        n = ret.shape[0]
        idx = np.arange(n)
        keys = sorted(idx, key=lambda i: (0 if safe_mask[i] else 1,
                                          0.0 if safe_mask[i] else float(cost[i]),
                                          -float(ret[i]), i))
        return np.asarray(keys)

    # -- the CEM loop -------------------------------------------------------------------------
    def do_generate_action(self, state, z_actions, eps, z_final, trace=None):
        """cem_mpc.py:35-68. z_actions [I, N, H, A], eps [I, H, P*N, O], z_final [A]."""
        dt = self.dt
        lb, ub, mu, sigma = sampling_params(self.action_space)
        a_dim = self.action_space.shape[0]
        lb = np.broadcast_to(np.asarray(lb, dtype=dt), (a_dim,))
        ub = np.broadcast_to(np.asarray(ub, dtype=dt), (a_dim,))
        mu = np.broadcast_to(np.asarray(mu, dtype=dt), (self.horizon, a_dim)).copy()
        sigma = np.broadcast_to(np.asarray(sigma, dtype=dt), (self.horizon, a_dim)).copy()
        best = np.zeros((a_dim,), dtype=dt)
        best_score = dt.type(-np.inf)
        state = np.asarray(state, dtype=dt)
        member_of_row = self.member_of_row()
        c_max = None
        if self.objective in (OBJ_SAFE_PENALTY, OBJ_FEASIBLE_FIRST):
            c_max = beta_count_threshold(self.particles, self.posterior_mean_threashold, dtype=dt)
        iterations_run = 0
        for it in range(self.iterations):
            iterations_run += 1
            acts = z_actions[it].astype(dt) * sigma + mu                        # :44-47
            acts = np.minimum(np.maximum(acts, lb), ub)                         # :48
            acts_b = np.tile(acts, (self.particles, 1, 1))                      # :49-51
            s0 = np.broadcast_to(state, (acts_b.shape[0], state.shape[0]))      # :53
            traj = self.model.unfold_sequences(s0, acts_b, eps[it], member_of_row)
            scores, ret, cost = self.compute_scores(traj, acts_b)               # :55
            safe_mask = None if c_max is None else (cost <= dt.type(c_max))
            order = self.rank_order(scores, ret, cost, safe_mask)
            elite = np.sort(order[:self.elite])                                 # :56
            top = order[0]                                                      # :57 (first max)
            top_score = ret[top] if scores is None else scores[top]
            if self.objective == OBJ_FEASIBLE_FIRST and not safe_mask[top]:
                top_score = ret[top] - dt.type(100.0)
            if top_score > best_score:                                          # :58-60
                best = acts[top, 0].copy()
                best_score = top_score
            mean, var = tf_moments_axis0(acts[elite])                           # :61-62
            std = np.sqrt(var)                                                  # :63
            mu = dt.type(self.smoothing) * mu + (dt.type(1.0) - dt.type(self.smoothing)) * mean
            sigma = dt.type(self.smoothing) * sigma + (dt.type(1.0) - dt.type(self.smoothing)) * std
            if trace is not None:
                trace.append(dict(actions=acts, ret=ret, cost=cost, scores=scores, elite=elite,
                                  mu=mu.copy(), sigma=sigma.copy(), best=best.copy(),
                                  best_score=best_score, traj=traj if trace_keep_traj(trace) else None))
            if np.sum(sigma, dtype=dt) / dt.type(sigma.size) <= dt.type(self.stddev_threshold):  # :66-67
                break
        action = best + np.asarray(z_final, dtype=dt) * dt.type(self.noise_stddev)  # :68
        return action, best_score, iterations_run


def trace_keep_traj(trace):
    return getattr(trace, 'keep_traj', False)


class Trace(list):
    """list of per-iteration dicts; set keep_traj=True to also keep [B, H+1, O] trajectories."""
    keep_traj = False
