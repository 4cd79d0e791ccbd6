"""Trajectory collection around the planner — the caller side of the hot path.

`sample_trajectory` restates `BaseAgent.sample_trajectory` (simba/agents/agent.py:103-153): one
environment, one planning call per decision, `action_repeat` environment steps per action.
`sample_trajectories_vectorized` is the batched form the reference does not have: n environments
advance in lockstep and every decision is ONE batched planning call (`n_states = n`, BASELINE
configs[3]'s shape), so the planner sees n states per launch instead of one. With n = 1 it
produces exactly the trajectories of the reference loop. Paths use the reference's
`replay_buffer.path_summary` layout (simba/infrastructure/replay_buffer.py:71-87), so they can be
stored in its ReplayBuffer unchanged.
"""
import numpy as np


def path_summary(observations, actions, rewards, next_observations, terminals, infos):
    """replay_buffer.py:71-87."""
    return {"observation": np.array(observations, dtype=np.float32),
            "reward": np.array(rewards, dtype=np.float32),
            "action": np.array(actions, dtype=np.float32),
            "next_observation": np.array(next_observations, dtype=np.float32),
            "terminal": np.array(terminals, dtype=np.float32),
            "info": infos}


class _Episode(object):
    def __init__(self, observation):
        self.observation = observation
        self.observations, self.actions, self.rewards = [], [], []
        self.next_observations, self.terminals, self.infos = [], [], []
        self.steps = 0
        self.done = False

    def advance(self, environment, action, action_repeat, max_trajectory_length):
        """The body of the reference's decision loop (agent.py:119-141)."""
        self.observations.append(self.observation)
        self.actions.append(action)
        repeat_rewards, repeat_costs, info = 0.0, 0.0, {}
        for _ in range(action_repeat):
            self.observation, reward, done, info = environment.step(action)
            self.steps += 1
            repeat_rewards += reward
            repeat_costs += info.get('cost', 0.0)
            self.done = (self.steps == max_trajectory_length) or done
            if info.get('goal_met', False) or self.done:
                break
        info['cost'] = repeat_costs
        self.next_observations.append(self.observation)
        self.rewards.append(repeat_rewards)
        self.infos.append(info)
        self.terminals.append(self.done)

    def summary(self):
        return path_summary(self.observations, self.actions, self.rewards, self.next_observations,
                            self.terminals, self.infos)


def sample_trajectory(environment, policy, max_trajectory_length, action_repeat=1):
    """agent.py:103-153 -> (path, steps)."""
    assert action_repeat, "Action repeat should be at least 1."
    ep = _Episode(environment.reset())
    while not ep.done:
        ep.advance(environment, policy.generate_action(ep.observation), action_repeat,
                   max_trajectory_length)
    assert ep.actions[0].shape == environment.action_space.shape, "Policy produces wrong actions shape."
    return ep.summary(), ep.steps


def sample_trajectories(environment, policy, batch_size, max_trajectory_length, action_repeat=1):
    """agent.py:83-101 -> (paths, timesteps)."""
    timesteps, paths = 0, []
    while timesteps < batch_size:
        path, n = sample_trajectory(environment, policy, max_trajectory_length, action_repeat)
        paths.append(path)
        timesteps += n
    return paths, timesteps


def sample_trajectories_vectorized(environments, policy, batch_size, max_trajectory_length,
                                   action_repeat=1):
    """Lockstep collection over len(environments) environments with one batched planning call per
    decision. `policy.generate_action` must accept [n, O] states and return [n, A] actions (a
    CemMpc / SafeCemMpc built with n_states = n). Episodes are restarted until the finished ones
    hold at least `batch_size` timesteps; unfinished episodes are dropped, like a reference loop
    that stops after the batch is full. Returns (paths in order of completion, timesteps)."""
    assert action_repeat, "Action repeat should be at least 1."
    n = len(environments)
    episodes = [_Episode(env.reset()) for env in environments]
    paths, timesteps = [], 0
    while timesteps < batch_size:
        states = np.stack([ep.observation for ep in episodes]).astype(np.float32)
        actions = np.asarray(policy.generate_action(states if n > 1 else states[0]))
        actions = actions.reshape(n, -1)
        for i, (env, ep) in enumerate(zip(environments, episodes)):
            ep.advance(env, actions[i], action_repeat, max_trajectory_length)
            if ep.done:
                paths.append(ep.summary())
                timesteps += ep.steps
                episodes[i] = _Episode(env.reset())
    return paths, timesteps
