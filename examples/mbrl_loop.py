"""The reference's training loop (simba/infrastructure/trainer.py:26-31 + simba/agents/mbrl_agent.py:
43-76) on the B200 classes, with a small self-contained lidar world standing in for safety_gym
(which is not installable here): warm-up with RandomMpc, TransitionModel.fit on the replay data,
then SafeCemMpc planning — every planning call and every training step runs in libsimba_b200.so.

    python examples/mbrl_loop.py --iterations 3 --envs 8

`--envs n` collects with `agents.sample_trajectories_vectorized`: n environments in lockstep, ONE
batched planning call per decision.
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'ethz-safe-learning_b200'))

from simba_b200 import agents                                              # noqa: E402
from simba_b200.environment_utils import ScorerEnvironment                # noqa: E402
from simba_b200.models import TransitionModel                             # noqa: E402
from simba_b200.policies import RandomMpc, SafeCemMpc                      # noqa: E402


class LidarWorld(object):
    """A point robot, one goal and a few hazards in the plane, observed through the PointGoal1 sensor
    layout (accelerometer, goal_lidar, gyro, hazards_lidar, magnetometer, vases_lidar, velocimeter;
    16-bin lidars with value = max(0, 1 - distance / lidar_max_dist) in the object's bin)."""
    BINS, MAX_DIST = 16, 4.0

    def __init__(self, scorer_env, seed=0, n_hazards=4, dt=0.1):
        self.action_space = scorer_env.action_space
        self.observation_space = scorer_env.observation_space
        self.rng = np.random.default_rng(seed)
        self.n_hazards, self.dt = n_hazards, dt
        self.goal_size, self.hazard_size = 0.3, 0.2

    def _lidar(self, points):
        out = np.zeros(self.BINS, np.float32)
        for p in points:
            d = p - self.pos
            dist = float(np.hypot(*d))
            b = int(((np.arctan2(d[1], d[0]) - self.heading) % (2 * np.pi)) / (2 * np.pi) * self.BINS) % self.BINS
            out[b] = max(out[b], max(0.0, 1.0 - dist / self.MAX_DIST))
        return out

    def _observe(self):
        acc = np.array([self.acc[0], self.acc[1], 0.0], np.float32)
        gyro = np.array([0.0, 0.0, self.turn], np.float32)
        mag = np.array([np.cos(self.heading), np.sin(self.heading), 0.0], np.float32) * 0.5
        vel = np.array([self.vel[0], self.vel[1], 0.0], np.float32)
        return np.concatenate([acc, self._lidar([self.goal]), gyro, self._lidar(self.hazards), mag,
                               np.zeros(self.BINS, np.float32), vel]).astype(np.float32)

    def reset(self):
        self.pos = self.rng.uniform(-1.5, 1.5, 2)
        self.heading = float(self.rng.uniform(0, 2 * np.pi))
        self.vel, self.acc, self.turn = np.zeros(2), np.zeros(2), 0.0
        self.goal = self.rng.uniform(-1.5, 1.5, 2)
        self.hazards = self.rng.uniform(-1.5, 1.5, (self.n_hazards, 2))
        return self._observe()

    def step(self, action):
        a = np.clip(np.asarray(action, np.float64), -1, 1)
        before = float(np.hypot(*(self.goal - self.pos)))
        self.turn = 1.5 * a[1]
        self.heading = (self.heading + self.turn * self.dt) % (2 * np.pi)
        new_vel = 0.8 * self.vel + 0.5 * a[0] * np.array([np.cos(self.heading), np.sin(self.heading)])
        self.acc, self.vel = (new_vel - self.vel) / self.dt * 0.1, new_vel
        self.pos = self.pos + self.vel * self.dt
        after = float(np.hypot(*(self.goal - self.pos)))
        goal_met = after <= self.goal_size
        cost = float(np.any(np.hypot(*(self.hazards - self.pos).T) <= self.hazard_size))
        reward = (before - after) + (1.0 if goal_met else 0.0)
        if goal_met:
            self.goal = self.rng.uniform(-1.5, 1.5, 2)
        return self._observe(), reward, False, dict(cost=cost, goal_met=goal_met)


def one_step_error(model, paths):
    obs = np.concatenate([p['observation'] for p in paths])
    act = np.concatenate([p['action'] for p in paths])
    nxt = np.concatenate([p['next_observation'] for p in paths])
    pred = model.predict(np.concatenate([obs, act], axis=1))[:, 1, :]
    return float(np.abs(pred - nxt).mean())


def run(iterations=3, n_envs=4, warmup_steps=400, interaction_steps=200, episode_length=50,
        training_steps=500, precision='bf16', seed=0, log=print):
    scorer_env = ScorerEnvironment()
    envs = [LidarWorld(scorer_env, seed=seed + i) for i in range(n_envs)]
    model = TransitionModel('mlp_ensemble', scorer_env.observation_space, scorer_env.action_space,
                            scale_features=True, sampling_propagation=True, ensemble_size=5,
                            batch_size=64, learning_rate=1e-3, learning_rate_schedule=True,
                            training_steps=training_steps, train_epochs=iterations,
                            mlp_params=dict(n_layers=4, units=128, activation='tf.nn.relu', dropout_rate=0.0))
    policy = SafeCemMpc(model, scorer_env, horizon=8, iterations=5, smoothing=0.0, n_samples=150, n_elite=15,
                        particles=20, stddev_threshold=0.25, noise_stddev=0.01,
                        posterior_mean_threashold=0.15, precision=precision, n_states=n_envs, seed=seed)
    warm = RandomMpc(scorer_env.action_space)
    np.random.seed(seed)
    paths, _ = agents.sample_trajectories(envs[0], warm, warmup_steps, episode_length)      # mbrl_agent.py:55-62
    report = []
    for it in range(iterations):
        obs = np.concatenate([p['observation'] for p in paths])
        act = np.concatenate([p['action'] for p in paths])
        nxt = np.concatenate([p['next_observation'] for p in paths])
        met = np.concatenate([[i.get('goal_met', False) for i in p['info']] for p in paths])
        keep = ~met                                                                       # mbrl_agent.py:46-50
        if it == 0:
            model._fit_statistics(np.concatenate([obs[keep], act[keep]], axis=1))
        err_before = one_step_error(model, paths[-4:])
        t0 = time.perf_counter()
        losses = model.fit(np.concatenate([obs[keep], act[keep]], axis=1), nxt[keep])     # agent.update()
        t_fit = time.perf_counter() - t0
        err = one_step_error(model, paths[-4:])
        t0 = time.perf_counter()
        new, steps = agents.sample_trajectories_vectorized(envs, policy, interaction_steps, episode_length)
        t_act = time.perf_counter() - t0
        paths += new
        ret = float(np.mean([p['reward'].sum() for p in new]))
        cost = float(np.mean([sum(i['cost'] for i in p['info']) for p in new]))
        report.append(dict(iteration=it, loss_first=float(losses[0]), loss_last=float(losses[-1]), one_step_error=err,
                           one_step_error_before_fit=err_before,
                           mean_return=ret, mean_cost=cost, fit_s=t_fit, train_steps_per_s=len(losses) / t_fit,
                           env_steps=steps, plans_per_s=steps / t_act))
        log("iter %d: nll %.3f -> %.3f  |pred - next| %.4f  return %.2f  cost %.2f  fit %.2fs (%.0f steps/s)  "
            "%d env steps at %.0f plans/s" % (it, losses[0], losses[-1], err, ret, cost, t_fit,
                                              len(losses) / t_fit, steps, steps / t_act))
    return report


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--iterations', type=int, default=3)
    ap.add_argument('--envs', type=int, default=4)
    ap.add_argument('--training-steps', type=int, default=500)
    ap.add_argument('--precision', default='bf16')
    a = ap.parse_args()
    run(a.iterations, a.envs, training_steps=a.training_steps, precision=a.precision)
