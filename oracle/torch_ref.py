"""TEST INFRASTRUCTURE ONLY — second, independently structured CPU restatement of the reference's
planning call, on torch-CPU ops and shaped like the reference: one Linear stack per ensemble member run
in a Python loop over `tf.split` chunks, trajectories materialised as [B, H + 1, O], objectives unrolled
over the horizon on slices of that tensor, `torch.topk`-free explicit tie handling. It exists to
(a) cross-check oracle/simba_oracle.py (numpy, fused differently) — tests/test_torch_ref.py holds
them to 1e-6 on all four objectives — and (b) serve as the timed CPU planner of bench.py
(`cpu_baseline`, `--impl reference`): torch's multi-threaded sgemm is the closest stand-in available
here for the TensorFlow-CPU kernels the reference would run on the same host cores.

Both restatements are pinned by tests/golden/reference_*.npz, which the UNMODIFIED reference code
produced (tests/golden/make_reference_golden.py). Citations: reference file:line.
Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

_SOFTPLUS_T = math.log(np.finfo(np.float32).eps) + 2.0      # tf.math.softplus switches formula at +-13.94


def tf_softplus(x):
    ex = torch.exp(x)
    return torch.where(x > -_SOFTPLUS_T, x, torch.where(x < _SOFTPLUS_T, ex, torch.log1p(ex)))


class GaussianMlp:
    """GaussianDistMlp (simba/models/mlp_ensemble.py:37-61): L x (Dense + ReLU), then the two heads
    (:25-34). Weights arrive in Keras variable order with [in, out] kernels."""

    def __init__(self, arrays):
        t = [torch.as_tensor(np.asarray(a, np.float32)) for a in arrays]
        self.hidden = [(t[i].t().contiguous(), t[i + 1]) for i in range(0, len(t) - 4, 2)]
        self.mu = (t[-4].t().contiguous(), t[-3])
        self.var = (t[-2].t().contiguous(), t[-1])

    def __call__(self, x):
        for w, b in self.hidden:
            x = torch.relu(F.linear(x, w, b))                # BaseLayer.call :17-22 (dropout off)
        return F.linear(x, *self.mu), tf_softplus(F.linear(x, *self.var)) + 1e-4     # GaussianHead :28-34


class Ensemble:
    def __init__(self, members):
        self.members = [GaussianMlp(m) for m in members]

    def forward(self, x):
        """mlp_ensemble.py:122-132: equal row chunks, chunk e -> member e, concatenated back."""
        e = len(self.members)
        if x.shape[0] % e:
            raise ValueError("tf.split: batch %d not divisible by ensemble size %d" % (x.shape[0], e))
        outs = [m(chunk) for m, chunk in zip(self.members, torch.split(x, x.shape[0] // e))]
        return torch.cat([o[0] for o in outs]), torch.cat([o[1] for o in outs])

    def __call__(self, x, eps):
        """mlp_ensemble.py:189-193: Normal(mu, sqrt(var)) -> mean, stddev, sample."""
        mu, var = self.forward(x)
        std = torch.sqrt(var)
        return mu, std, mu + std * eps


class Dynamics:
    """TransitionModel (simba/models/transition_model.py:64-87)."""

    def __init__(self, ensemble, inputs_min, inputs_max, scale_features=True, sampling_propagation=True):
        self.ensemble = ensemble
        self.lo = torch.as_tensor(np.asarray(inputs_min, np.float32))
        self.hi = torch.as_tensor(np.asarray(inputs_max, np.float32))
        self.scale_features, self.sampling_propagation = scale_features, sampling_propagation

    def scale(self, x):
        if not self.scale_features:
            return x
        delta = self.hi - self.lo
        delta = torch.where(delta < 1e-5, torch.full_like(delta, 1.01), delta)       # :86
        return (x - self.lo) / delta

    def unfold_sequences(self, s_0, action_sequences, eps):
        horizon = action_sequences.shape[1]
        states = [s_0]
        s_t = s_0
        for t in range(horizon):
            mus, _, d_s_t = self.ensemble(self.scale(torch.cat([s_t, action_sequences[:, t]], dim=1)), eps[t])
            s_t = s_t + (d_s_t if self.sampling_propagation else mus)                # :75
            states.append(s_t)
        return torch.stack(states, dim=1)                                            # [B, H + 1, O]


class GoalScorer:
    """SafetyGymStateScorer, goal task (simba/environment_utils/safety_gym.py:110-192)."""

    def __init__(self, config, offsets):
        self.c, self.off = config, offsets

    def closest(self, lidar):                                                        # :188-192
        d = float(self.c['lidar_max_dist'])
        return torch.clamp(d - d * (1.0 - lidar), 0.0, d).min(dim=1).values

    def goal_distance(self, obs):                                                    # :168-176
        if self.c['observe_goal_lidar']:
            return self.closest(obs[:, self.off['goal_lidar']])
        return torch.relu(obs[:, self.off['goal_dist']]).reshape(-1)

    def reward(self, obs, next_obs):                                                 # :110-143
        dist, nxt = self.goal_distance(obs), self.goal_distance(next_obs)
        achieved = dist <= self.c['goal_size'] * 0.8
        r = (dist - nxt) * self.c['reward_distance'] + achieved.float() * self.c['reward_goal']
        if self.c['reward_clip']:
            r = torch.clamp(r, -self.c['reward_clip'], self.c['reward_clip'])
        return r, achieved

    def cost(self, obs):                                                             # :145-166
        total = torch.zeros(obs.shape[0])
        for name in ('vases', 'hazards', 'pillars', 'gremlins'):
            if self.c['constrain_' + name]:
                total = total + (self.closest(obs[:, self.off[name + '_lidar']]) <= self.c[name + '_size']).float()
        return (total > 0).float() if self.c['constrain_indicator'] else total


def top_k_lower_index_first(scores, k):
    """tf.nn.top_k: the k largest, equal scores resolved toward the lower index."""
    return torch.argsort(-scores, stable=True)[:k]


class Planner:
    """CemMpc / SafeCemMpc (simba/policies/cem_mpc.py:35-68, mpc_policy.py:26-57, safe_cem_mpc.py:76-120);
    `objective` in {'reward', 'penalty', 'least_cost', 'feasible_first'} as in the product."""

    def __init__(self, dynamics, scorer, act_low, act_high, horizon, iterations, smoothing, n_samples, n_elite,
                 particles, stddev_threshold, noise_stddev, posterior_mean_threashold=0.15, objective='penalty'):
        self.dyn, self.scorer = dynamics, scorer
        self.lb = torch.as_tensor(np.asarray(act_low, np.float32))
        self.ub = torch.as_tensor(np.asarray(act_high, np.float32))
        self.horizon, self.iterations, self.smoothing = horizon, iterations, smoothing
        self.n_samples, self.elite, self.particles = n_samples, n_elite, particles
        self.stddev_threshold, self.noise_stddev = stddev_threshold, noise_stddev
        self.threshold, self.objective = posterior_mean_threashold, objective

    def _per_sample_mean(self, per_row):
        return per_row.reshape(self.particles, self.n_samples).mean(dim=0)

    def returns_reward_only(self, traj):                                             # mpc_policy.py:26-39
        cum = torch.zeros(traj.shape[0])
        done = torch.zeros(traj.shape[0], dtype=torch.bool)
        for t in range(traj.shape[1] - 1):
            r, d = self.scorer.reward(traj[:, t], traj[:, t + 1])
            cum = cum + r * (1.0 - done.float())
            done = d | done
        return self._per_sample_mean(cum)

    def returns_safe(self, traj):                                                    # safe_cem_mpc.py:76-96, :110-120
        cum = torch.zeros(traj.shape[0])
        done = torch.zeros(traj.shape[0], dtype=torch.bool)
        safe = torch.ones(self.n_samples, dtype=torch.bool)
        worst = torch.zeros(self.n_samples)
        mu0, sigma0 = torch.tensor(0.5), torch.tensor(0.27)                          # :81
        alpha = (((1.0 - mu0) / sigma0 ** 2) - 1.0 / mu0) * (mu0 ** 2)
        beta = alpha * (1.0 / mu0 - 1)
        for t in range(traj.shape[1] - 1):
            r, d = self.scorer.reward(traj[:, t], traj[:, t + 1])
            done = d | done
            cost = self.scorer.cost(traj[:, t]) * (1.0 - done.float())
            counts = cost.reshape(self.particles, self.n_samples).sum(dim=0)
            safe = ((alpha + counts) / (alpha + beta + self.particles) <= self.threshold) & safe
            worst = torch.maximum(worst, counts)
            cum = cum + r * (1.0 - done.float())
        return self._per_sample_mean(cum), worst, safe

    def mean_costs(self, traj):                                                      # safe_cem_mpc.py:98-108
        cum = torch.zeros(traj.shape[0])
        for t in range(traj.shape[1] - 1):
            cum = cum + self.scorer.cost(traj[:, t])
        return self._per_sample_mean(cum)

    def rank(self, traj):
        """-> (order: candidates best first, score of each candidate as the reference would report it)."""
        if self.objective == 'reward':
            s = self.returns_reward_only(traj)
            return torch.argsort(-s, stable=True), s
        if self.objective == 'least_cost':
            s = -self.mean_costs(traj)
            return torch.argsort(-s, stable=True), s
        ret, worst, safe = self.returns_safe(traj)
        s = ret - (~safe).float() * 100.0
        if self.objective == 'penalty':
            return torch.argsort(-s, stable=True), s
        # feasible-first: safe candidates by return, then unsafe ones by fewest violations, then return
        keys = sorted(range(self.n_samples), key=lambda i: (0 if safe[i] else 1, 0.0 if safe[i] else float(worst[i]),
                                                            -float(ret[i]), i))
        return torch.as_tensor(keys), s

    def do_generate_action(self, state, z_actions, eps, z_final):                    # cem_mpc.py:35-68
        state = torch.as_tensor(np.asarray(state, np.float32))
        z_actions, eps = torch.as_tensor(z_actions), torch.as_tensor(eps)
        a_dim = self.lb.shape[0]
        mu = ((self.ub + self.lb) / 2.0).expand(self.horizon, a_dim)                 # mpc_policy.py:45-57
        sigma = ((self.ub - self.lb) / 2.0).expand(self.horizon, a_dim)
        best, best_score = torch.zeros(a_dim), torch.tensor(-math.inf)
        iterations_run = 0
        for it in range(self.iterations):
            iterations_run += 1
            acts = torch.maximum(torch.minimum(z_actions[it] * sigma + mu, self.ub), self.lb)   # :44-48
            acts_b = acts.repeat(self.particles, 1, 1)                               # :49-51
            traj = self.dyn.unfold_sequences(state.expand(acts_b.shape[0], -1), acts_b, eps[it])
            order, score = self.rank(traj)                                           # :55-57
            elite, top = order[:self.elite], order[0]
            if score[top] > best_score:                                              # :58-60
                best, best_score = acts[top, 0], score[top]
            chosen = acts[elite]
            mean = chosen.mean(dim=0)                                                # :61-63 (population moments)
            std = torch.sqrt(((chosen - mean) ** 2).mean(dim=0))
            mu = self.smoothing * mu + (1.0 - self.smoothing) * mean
            sigma = self.smoothing * sigma + (1.0 - self.smoothing) * std
            if sigma.mean() <= self.stddev_threshold:                                # :66-67
                break
        action = best + torch.as_tensor(np.asarray(z_final, np.float32)) * self.noise_stddev   # :68
        return action.numpy(), float(best_score), iterations_run, dict(mu=mu.numpy(), sigma=sigma.numpy(),
                                                                       elite=np.sort(elite.numpy()))
