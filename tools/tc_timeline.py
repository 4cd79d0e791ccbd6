#!/usr/bin/env python
"""Phase timeline of the tcgen05 rollout kernel (needs the -DSIMBA_TC_TIMELINE build:
SIMBA_B200_LIB=.../libsimba_b200_tl.so python tools/tc_timeline.py [c1|c4|c5]; c5 = wide kernel)."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'ethz-safe-learning_b200')]
import numpy as np  # noqa: E402
import torch  # noqa: E402
from simba_b200 import _lib, synthetic  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else 'c1'
over = dict(S=64) if wl == 'c4' else {}
c = synthetic.make_workload({'c4': 'c4', 'c5': 'c5'}.get(wl, 'c1'), **over)
pol = synthetic.build_policy(c, 'penalty', precision='bf16', seed=1)
lib = _lib.load()
pl = pol._ensure_planner()
S, N, H, A, P_, O = c['S'], c['N'], c['H'], c['A'], c['P'], c['O']
st = torch.from_numpy(np.ascontiguousarray(synthetic.make_state(c['sensors'], 5, S)).reshape(S, O)).cuda()
acts = torch.empty((S, N, H, A), device='cuda').uniform_(-1, 1)
ret = torch.empty((S, P_, N), device='cuda'); msk = torch.empty((S, P_, N), dtype=torch.int64, device='cuda')
csum = torch.empty((S, P_, N), device='cuda')
# the timeline build stamps into prm.traj_out; simba_rollout_score leaves it null, so use the env hook
tl = torch.zeros((2, 64, 64), dtype=torch.int64, device='cuda')
os.environ['SIMBA_TC_TIMELINE_PTR'] = str(tl.data_ptr())
p = lambda t: C.c_void_p(t.data_ptr())
for it in range(3):
    _lib.check(lib.simba_rollout_score(pl, p(st), p(acts), None, 3, it, None, p(ret), p(msk), p(csum), None))
torch.cuda.synchronize()
t = tl.cpu().numpy()
L = c['L']
for who, name in ((0, 'issuer (warp 0)'), (1, 'last warp of tile')):
    print("==", name)
    rows = []
    for step in range(2, H - 1):
        ev = t[who, step]
        base = ev[0]
        line = ["step %2d" % step]
        for l in range(L):
            line.append("L%d wait->%5d epi->%5d sync+issue->%5d |" % (l, ev[1 + 4 * l] - base, ev[2 + 4 * l] - base, ev[3 + 4 * l] - base))
        line.append("head wait->%5d pass->%5d sync+issue->%5d score->%5d  total %5d" % (
            ev[40] - base, ev[41] - base, ev[42] - base, ev[43] - base, t[who, step + 1][0] - base))
        line.append("| issue detail (st_wait, barrier) per layer: " + ' '.join(
            "%d:(%d,%d)" % (l, ev[48 + l] - base, ev[54 + l] - base) for l in range(L + 1)))
        line.append("| head detail: ld_wait->%d state_st->%d before_st_wait->%d after->%d" % tuple(
            ev[k] - base for k in (44, 45, 46, 47)))
        rows.append(' '.join(line))
    print('\n'.join(rows[:6]))
