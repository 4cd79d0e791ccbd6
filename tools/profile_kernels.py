#!/usr/bin/env python
"""Launch each planner kernel a few times at a named workload — the command ncu wraps.
  python tools/profile_kernels.py --workload c1|c3|c4 --precision bf16|fp32 [--reps 3] [--states S]
Prints CUDA-event times per kernel (never taken under ncu for bench purposes)."""
import argparse
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'ethz-safe-learning_b200')]

import numpy as np  # noqa: E402
import torch  # noqa: E402
from simba_b200 import _lib, synthetic  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--workload', default='c1')
    ap.add_argument('--precision', default='bf16')
    ap.add_argument('--reps', type=int, default=3)
    ap.add_argument('--states', type=int, default=0)
    ap.add_argument('--population', type=int, default=0)
    args = ap.parse_args()
    over = {}
    if args.states:
        over['S'] = args.states
    if args.population:
        over['N'] = args.population
        over['K'] = max(1, args.population // 10)
    c = synthetic.make_workload(args.workload, **over)
    mm = 'particle' if (c['P'] * c['N']) % c['E'] else 'split'
    pol = synthetic.build_policy(c, 'penalty', precision=args.precision, member_map=mm, seed=1)
    lib = _lib.load()
    pl = pol._ensure_planner()
    S, N, H, A, P_, O, K = c['S'], c['N'], c['H'], c['A'], c['P'], c['O'], c['K']
    st = torch.from_numpy(np.ascontiguousarray(synthetic.make_state(c['sensors'], 5, S)).reshape(S, O)).cuda()
    f32 = dict(dtype=torch.float32, device='cuda')
    mu = torch.zeros((S, H, A), **f32); sg = torch.ones((S, H, A), **f32)
    acts = torch.empty((S, N, H, A), **f32)
    ret = torch.empty((S, P_, N), **f32); msk = torch.empty((S, P_, N), dtype=torch.int64, device='cuda')
    csum = torch.empty((S, P_, N), **f32); pairs = torch.empty((S, N, 2), **f32)
    elite = torch.empty((S, K), dtype=torch.int32, device='cuda')
    best_a = torch.zeros((S, A), **f32); best_s = torch.full((S,), -np.inf, **f32)
    active = torch.ones((S,), dtype=torch.int32, device='cuda'); iters = torch.zeros((S,), dtype=torch.int32, device='cuda')
    sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    p = lambda t: C.c_void_p(t.data_ptr())
    steps = [
        ('sample', lambda i: lib.simba_sample_actions(pl, p(mu), p(sg), None, 3, i, None, p(acts), sp)),
        ('rollout', lambda i: lib.simba_rollout_score(pl, p(st), p(acts), None, 3, i, None, p(ret), p(msk), p(csum), sp)),
        ('reduce', lambda i: lib.simba_score_reduce(pl, p(ret), p(msk), p(csum), None, p(pairs), sp)),
        ('select', lambda i: lib.simba_select_elites(pl, p(pairs), p(acts), None, p(elite), None, p(best_a), p(best_s), sp)),
        ('refit', lambda i: lib.simba_refit(pl, p(acts), p(elite), p(mu), p(sg), p(active), p(iters), sp)),
    ]
    for name, fn in steps:
        _lib.check(fn(0))
        torch.cuda.synchronize()
        ts = []
        for i in range(args.reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); _lib.check(fn(i + 1)); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        print("%-8s %s  mean %.4f ms" % (name, ' '.join('%.4f' % t for t in ts), float(np.mean(ts))))
    fpt = synthetic.flops_per_transition(O, A, c['L'], c['U'])
    print("rollout flops/launch %.4g" % (fpt * H * P_ * N * S))


if __name__ == '__main__':
    main()
