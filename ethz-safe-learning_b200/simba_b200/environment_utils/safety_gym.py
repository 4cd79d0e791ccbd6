"""Reward / cost scorer of the Safety-Gym goal task on the GPU.

Mirrors `SafetyGymStateScorer` (simba/environment_utils/safety_gym.py:104-192) and the three
members of `MbrlSafetyGym` the planner reads (`get_reward`, `get_cost` :62-66, spaces :30).
The MuJoCo simulator wrapper itself (:10-32, :68-101) needs gym + safety_gym + mujoco_py and is
not part of the planning path; `ScorerEnvironment` carries the same scorer around synthetic or
externally supplied observation/action spaces.

Only the goal task with lidar or goal_dist observation is fused (the configured path); the push
task (:121-134) and the orientation reward (:136-139, which reads a key that does not exist) are
rejected with SIMBA_ERR_UNSUPPORTED.
"""
import numpy as np
import torch

from .. import _device, _lib
from ..spaces import Box

# un-vendored safety_gym Engine.DEFAULT values + simba's registry overrides
# (simba/environment_utils/safety_gym_registery.py:9-16,27-40)
DEFAULT_CONFIG = dict(
    task='goal', goal_size=0.3, hazards_size=0.2, vases_size=0.1, pillars_size=0.2,
    gremlins_size=0.1, lidar_max_dist=4, lidar_num_bins=16,
    observe_goal_lidar=True, observe_goal_dist=False,
    constrain_hazards=True, constrain_vases=False, constrain_pillars=False,
    constrain_gremlins=False, constrain_indicator=True,
    reward_distance=1.0, reward_goal=1.0, reward_clip=10, reward_orientation=False)

POINTGOAL1_SENSORS = dict(accelerometer=3, goal_lidar=16, gyro=3, hazards_lidar=16, magnetometer=3,
                          vases_lidar=16, velocimeter=3)                      # O = 60 (BASELINE shape)
POINTSIMPLEGOAL1_SENSORS = dict(accelerometer=3, goal_lidar=5, gyro=3, hazards_lidar=5,
                                magnetometer=3, velocimeter=3)                # O = 22 (shipped config)

_CONSTRAINT_ORDER = ('vases', 'hazards', 'pillars', 'gremlins')               # safety_gym.py:148-163


def make_sensor_offset_table(sensor_sizes):
    """Offsets by sorted sensor key — safety_gym.py:17-25."""
    table, offset = {}, 0
    for k in sorted(sensor_sizes):
        table[k] = slice(offset, offset + int(sensor_sizes[k]))
        offset += int(sensor_sizes[k])
    return table


class SafetyGymStateScorer(object):
    def __init__(self, config, sensor_offset_table):
        merged = dict(DEFAULT_CONFIG)
        merged.update(config or {})
        for key, value in merged.items():          # safety_gym.py:106-107
            setattr(self, key, value)
        self.sensor_offset_table = sensor_offset_table
        self._lib = _lib.load()

    def scorer_struct(self):
        """The parameter struct the fused kernels take instead of Python callables."""
        if self.task != 'goal':
            raise _lib.SimbaError(-6, "only task='goal' is fused (got %r)" % (self.task,))
        if self.reward_orientation:
            raise _lib.SimbaError(-6, "reward_orientation is not supported (broken in the reference)")
        sc = _lib.Scorer()
        tab = self.sensor_offset_table
        if self.observe_goal_lidar:
            sl = tab['goal_lidar']
            sc.goal_begin, sc.goal_end, sc.goal_dist_index = sl.start, sl.stop, -1
        elif self.observe_goal_dist:
            sl = tab['goal_dist']
            sc.goal_begin, sc.goal_end, sc.goal_dist_index = 0, 0, sl.start
        else:
            raise NotImplementedError                                            # safety_gym.py:175-176
        n = 0
        for name in _CONSTRAINT_ORDER:
            if getattr(self, 'constrain_' + name, False):
                sl = tab[name + '_lidar']
                sc.con_begin[n], sc.con_end[n] = sl.start, sl.stop
                sc.con_size[n] = float(getattr(self, name + '_size'))
                n += 1
        sc.n_constraints = n
        sc.lidar_max_dist = float(self.lidar_max_dist)
        sc.goal_threshold = float(np.float32(self.goal_size * 0.8))              # safety_gym.py:117
        sc.reward_distance = float(self.reward_distance)
        sc.reward_goal = float(self.reward_goal)
        sc.reward_clip = float(self.reward_clip) if self.reward_clip else 0.0
        sc.constrain_indicator = int(bool(self.constrain_indicator))
        return sc

    def _eval(self, observations, next_observations, want_reward, want_cost):
        import ctypes as C
        obs, kind = _device.to_device(observations)
        nxt = None
        if next_observations is not None:
            nxt, _ = _device.to_device(next_observations)
        b, o = obs.shape
        reward = torch.empty((b,), dtype=torch.float32, device=obs.device) if want_reward else None
        done = torch.empty((b,), dtype=torch.int32, device=obs.device) if want_reward else None
        cost = torch.empty((b,), dtype=torch.float32, device=obs.device) if want_cost else None
        sc = self.scorer_struct()
        _lib.check(self._lib.simba_scorer_eval(C.byref(sc), _device.ptr(obs), _device.ptr(nxt), b, o,
                                               _device.ptr(reward), _device.ptr(done),
                                               _device.ptr(cost), _device.stream_ptr()))
        return reward, done, cost, kind

    def reward(self, observations, next_observations):
        """safety_gym.py:110-143 -> (reward[B], goal_achieved[B] bool)."""
        reward, done, _, kind = self._eval(observations, next_observations, True, False)
        return _device.like_input(reward, kind), _device.like_input(done.bool(), kind)

    def cost(self, observations):
        """safety_gym.py:145-166 -> cost[B]."""
        _, _, cost, kind = self._eval(observations, None, False, True)
        return _device.like_input(cost, kind)


class ScorerEnvironment(object):
    """What the policies consume from an environment (mpc_policy.py:16-18, safe_cem_mpc.py:32,
    transition_model.py:22-29): get_reward, get_cost, action_space, observation_space."""

    def __init__(self, sensor_sizes=None, config=None, action_low=(-1.0, -1.0), action_high=(1.0, 1.0),
                 observation_low=None, observation_high=None):
        sensor_sizes = dict(sensor_sizes or POINTGOAL1_SENSORS)
        self.sensor_offset_table = make_sensor_offset_table(sensor_sizes)
        self._scorer = SafetyGymStateScorer(config, self.sensor_offset_table)
        obs_dim = sum(sensor_sizes.values())
        low, high = [], []
        for k in sorted(sensor_sizes):                     # resolve_observation_limits, :34-60
            if k.endswith('_lidar'):
                low += [0.0] * sensor_sizes[k]
                high += [1.0] * sensor_sizes[k]
            else:
                low += [-np.inf] * sensor_sizes[k]
                high += [np.inf] * sensor_sizes[k]
        if observation_low is not None:
            low, high = observation_low, observation_high
        assert len(low) == obs_dim
        self.observation_space = Box(np.asarray(low), np.asarray(high))
        self.action_space = Box(np.asarray(action_low), np.asarray(action_high))

    def get_reward(self, obs, acs, *args, **kwargs):       # safety_gym.py:62-63
        return self._scorer.reward(obs, *args, **kwargs)

    def get_cost(self, obs, acs, *args, **kwargs):         # safety_gym.py:65-66
        return self._scorer.cost(obs)
