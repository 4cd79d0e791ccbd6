"""The reference-side FFI stub of INTEGRATION.md section B as a runnable file: ctypes + numpy only
(no torch, no simba_b200 package). It is what a simba maintainer would drop into
`simba/policies/_b200.py` to call libsimba_b200.so from the TensorFlow code base: Keras-ordered
weight arrays in, `generate_action(state) -> action` out (replaces simba/policies/cem_mpc.py:31-33).
The structs restate include/simba_b200.h field for field.

    python examples/ffi_stub.py            # plans once on a random-init 5 x (4 x 128) ensemble
"""
import ctypes as C
import os

import numpy as np

MAX_ACT, MAX_CON = 16, 4
OBJ_REWARD, OBJ_SAFE_PENALTY = 0, 1
PREC_FP32, PREC_BF16 = 0, 1


class ModelCfg(C.Structure):                      # simba_model_config_t
    _fields_ = [(n, C.c_int32) for n in ("obs_dim", "act_dim", "ensemble_size", "n_layers", "units")]


class Scorer(C.Structure):                        # simba_scorer_t (safety_gym.py:104-192, goal task)
    _fields_ = [("goal_begin", C.c_int32), ("goal_end", C.c_int32), ("goal_dist_index", C.c_int32),
                ("n_constraints", C.c_int32), ("con_begin", C.c_int32 * MAX_CON),
                ("con_end", C.c_int32 * MAX_CON), ("con_size", C.c_float * MAX_CON),
                ("lidar_max_dist", C.c_float), ("goal_threshold", C.c_float),
                ("reward_distance", C.c_float), ("reward_goal", C.c_float), ("reward_clip", C.c_float),
                ("constrain_indicator", C.c_int32)]


class PlannerCfg(C.Structure):                    # simba_planner_config_t (policies.yaml + cem_mpc.py ctor)
    _fields_ = [("horizon", C.c_int32), ("iterations", C.c_int32), ("n_samples", C.c_int32),
                ("n_elite", C.c_int32), ("particles", C.c_int32), ("n_states", C.c_int32),
                ("smoothing", C.c_float), ("stddev_threshold", C.c_float), ("noise_stddev", C.c_float),
                ("posterior_mean_threshold", C.c_float), ("prior_mu", C.c_float), ("prior_sigma", C.c_float),
                ("objective", C.c_int32), ("sampling_propagation", C.c_int32), ("precision", C.c_int32),
                ("member_map", C.c_int32), ("rank", C.c_int32), ("world_size", C.c_int32),
                ("act_low", C.c_float * MAX_ACT), ("act_high", C.c_float * MAX_ACT),
                ("init_mean", C.c_float * MAX_ACT), ("init_stddev", C.c_float * MAX_ACT),
                ("scorer", Scorer)]


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class B200Planner(object):
    def __init__(self, lib_path, member_weights, inputs_min, inputs_max, scorer, act_low, act_high,
                 horizon=15, iterations=5, n_samples=150, n_elite=15, particles=20, smoothing=0.0,
                 stddev_threshold=0.25, noise_stddev=0.01, posterior_mean_threashold=0.15, safe=True,
                 precision=PREC_BF16):
        """member_weights: per member the Keras variable list (mlp_ensemble.py:46-50, :28-30):
        L x (kernel[in, U], bias[U]), (kernel_mu[U, O], bias_mu[O]), (kernel_var[U, O], bias_var[O])."""
        self.lib = lib = C.CDLL(lib_path)
        lib.simba_last_error.restype = C.c_char_p
        lib.simba_model_set_layer.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
        lib.simba_model_set_scaler.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]
        lib.simba_model_commit.argtypes = [C.c_void_p]
        lib.simba_planner_create.argtypes = [C.c_void_p, C.POINTER(PlannerCfg), C.POINTER(C.c_void_p)]
        lib.simba_plan_host.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.simba_planner_destroy.argtypes = [C.c_void_p]
        lib.simba_model_destroy.argtypes = [C.c_void_p]
        w0 = member_weights[0]
        n_layers = len(w0) // 2 - 2
        obs_dim = w0[-1].shape[0]
        self.act_dim = w0[0].shape[0] - obs_dim
        mcfg = ModelCfg(obs_dim, self.act_dim, len(member_weights), n_layers, w0[0].shape[1])
        self.model = C.c_void_p()
        self._check(lib.simba_model_create(C.byref(mcfg), C.byref(self.model)))
        for e, arrays in enumerate(member_weights):
            arrays = [np.ascontiguousarray(a, np.float32) for a in arrays]
            for layer in range(len(arrays) // 2):                     # kernel [in, out], bias [out]
                self._check(lib.simba_model_set_layer(self.model, e, layer, _p(arrays[2 * layer]),
                                                      _p(arrays[2 * layer + 1])))
        lo, hi = np.ascontiguousarray(inputs_min, np.float32), np.ascontiguousarray(inputs_max, np.float32)
        self._check(lib.simba_model_set_scaler(self.model, _p(lo), _p(hi), 1))     # transition_model.py:79-87
        self._check(lib.simba_model_commit(self.model))
        cfg = PlannerCfg()
        cfg.horizon, cfg.iterations, cfg.n_samples, cfg.n_elite = horizon, iterations, n_samples, n_elite
        cfg.particles, cfg.n_states = particles, 1
        cfg.smoothing, cfg.stddev_threshold, cfg.noise_stddev = smoothing, stddev_threshold, noise_stddev
        cfg.posterior_mean_threshold, cfg.prior_mu, cfg.prior_sigma = posterior_mean_threashold, 0.5, 0.27
        cfg.objective = OBJ_SAFE_PENALTY if safe else OBJ_REWARD
        cfg.sampling_propagation, cfg.precision, cfg.member_map = 1, precision, 0
        cfg.rank, cfg.world_size = 0, 1
        for i in range(self.act_dim):                                              # mpc_policy.py:45-57
            cfg.act_low[i], cfg.act_high[i] = float(act_low[i]), float(act_high[i])
            cfg.init_mean[i] = (float(act_high[i]) + float(act_low[i])) / 2.0
            cfg.init_stddev[i] = (float(act_high[i]) - float(act_low[i])) / 2.0
        cfg.scorer = scorer
        self.planner = C.c_void_p()
        self._check(lib.simba_planner_create(self.model, C.byref(cfg), C.byref(self.planner)))
        self.calls = 0

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError("simba_b200 error %d: %s" % (rc, self.lib.simba_last_error().decode()))

    def generate_action(self, state, seed=None):
        """cem_mpc.py:31-33: numpy state [O] -> numpy action [A]."""
        state = np.ascontiguousarray(state, np.float32)
        action = np.empty(self.act_dim, np.float32)
        score, iters = np.empty(1, np.float32), np.empty(1, np.int32)
        if seed is None:
            seed, self.calls = self.calls, self.calls + 1
        self._check(self.lib.simba_plan_host(self.planner, _p(state), C.c_uint64(seed), _p(action), _p(score),
                                             _p(iters)))
        return action

    def close(self):
        self.lib.simba_planner_destroy(self.planner)
        self.lib.simba_model_destroy(self.model)


def pointgoal1_scorer():
    """PointGoal1 with observed vases: accelerometer[0:3] goal_lidar[3:19] gyro[19:22] hazards_lidar[22:38]
    magnetometer[38:41] vases_lidar[41:57] velocimeter[57:60] (sorted sensor keys, safety_gym.py:17-25)."""
    sc = Scorer()
    sc.goal_begin, sc.goal_end, sc.goal_dist_index = 3, 19, -1
    sc.n_constraints = 1
    sc.con_begin[0], sc.con_end[0], sc.con_size[0] = 22, 38, 0.2           # hazards_size
    sc.lidar_max_dist, sc.goal_threshold = 4.0, 0.3 * 0.8                   # goal_size * 0.8 (safety_gym.py:117)
    sc.reward_distance, sc.reward_goal, sc.reward_clip, sc.constrain_indicator = 1.0, 1.0, 10.0, 1
    return sc


def default_lib_path():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    return os.path.join(root, 'ethz-safe-learning_b200', 'simba_b200', 'libsimba_b200.so')


if __name__ == '__main__':
    rng = np.random.default_rng(0)
    O, A, U, L, E = 60, 2, 128, 4, 5

    def glorot(i, o):
        lim = np.sqrt(6.0 / (i + o))
        return rng.uniform(-lim, lim, (i, o)).astype(np.float32)
    members = []
    for _ in range(E):
        arrays, fan = [], O + A
        for _ in range(L):
            arrays += [glorot(fan, U), np.zeros(U, np.float32)]
            fan = U
        arrays += [glorot(U, O) * 0.05, np.zeros(O, np.float32), glorot(U, O), np.full(O, -9.0, np.float32)]
        members.append(arrays)
    planner = B200Planner(default_lib_path(), members, np.full(O + A, -1.0), np.full(O + A, 2.0), pointgoal1_scorer(),
                          [-1.0] * A, [1.0] * A)
    print("action:", planner.generate_action(rng.uniform(0.3, 1.0, O)))
    planner.close()
