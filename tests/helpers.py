"""Shared builders: the same synthetic workload as an oracle planner and as a CUDA planner."""
import numpy as np

from oracle import simba_oracle as so
from simba_b200 import synthetic

OBJ = dict(reward=so.OBJ_REWARD, penalty=so.OBJ_SAFE_PENALTY, least_cost=so.OBJ_LEAST_COST,
           feasible_first=so.OBJ_FEASIBLE_FIRST)


def workload(cfg_name='tiny', sensors=None, seed=0, **over):
    return synthetic.make_workload(cfg_name, sensors=sensors, seed=seed, **over)


def oracle_planner(c, objective='penalty', dtype=np.float32, member_map='split', threshold=0.15,
                   smoothing=0.0, stddev_threshold=0.0, noise_stddev=0.01, sampling_propagation=True,
                   scorer_config=None):
    ens = so.MlpEnsemble(c['weights'], dtype=dtype)
    tm = so.TransitionModel(ens, c['smin'], c['smax'], True, sampling_propagation, dtype=dtype)
    scorer = so.Scorer(scorer_config, c['table'], dtype=dtype)
    env = so.Environment(scorer, so.Box([-1.0] * c['A'], [1.0] * c['A']), None)
    return so.CemPlanner(tm, env, c['H'], c['I'], smoothing, c['N'], c['K'], c['P'], stddev_threshold,
                         noise_stddev, posterior_mean_threashold=threshold,
                         objective=OBJ[objective], member_map=member_map, dtype=dtype)


def cuda_policy(c, objective='penalty', precision='fp32', **kw):
    return synthetic.build_policy(c, objective, precision=precision, **kw)
