#!/usr/bin/env python
"""Phase stamps of the fused CEM update kernel inside whole C1 plans (needs the -DSIMBA_TC_TIMELINE build:
SIMBA_B200_LIB=.../libsimba_b200_tl.so python tools/update_timeline.py)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'ethz-safe-learning_b200')]
import numpy as np  # noqa: E402
import torch  # noqa: E402
from simba_b200 import synthetic  # noqa: E402

c = synthetic.make_workload('c1')
tl = torch.zeros((4, 64, 64), dtype=torch.int64, device='cuda')
os.environ['SIMBA_TC_TIMELINE_PTR'] = str(tl.data_ptr())
pol = synthetic.build_policy(c, 'penalty', precision='bf16', seed=1)
for rep in range(3):
    pol.do_generate_action(c['state'], seed=5 + rep)
torch.cuda.synchronize()
t = tl.cpu().numpy()
names = ['entry', 'PDL wait passed', 'scores + keys', 'ranked', 'elites + best', 'refit', 'sampled / finalised']
for it in range(c['I']):
    ev = t[3, 1 + it]
    print("iteration %d: " % it + ', '.join("%s %d" % (names[k], ev[k] - ev[1]) for k in range(2, 7)))
    print("   detail (cycles after the wait): loads issued %d, staged %d, scores done (thread 0) %d | refit: start %d, "
          % tuple(ev[k] - ev[1] for k in (7, 8, 9, 10))
          + ', '.join("pass%d part %d sync %d totals %d sync %d" % ((ps,) + tuple(ev[11 + 4 * ps + q] - ev[1] for q in range(4)))
                      for ps in range(2)) + " | refit_single stamps 10..15: " + ' '.join(str(int(ev[k] - ev[1])) for k in range(10, 16)))

print("== plan anatomy, ns on the global timer (rollout: thread 0 of CTA 0; update: thread 0)")
for it in range(c['I']):
    r = t[3, 10 + it]          # rollout: 0 entry, 1 setup done, 2 wait passed, 3 first A, 4 items done, 5 exit
    uu = t[3, 1 + it][32:]     # update: 0 entry, 1 wait passed, 6 end
    base = t[3, 10][2]
    print("iteration %d: rollout wait passed %7d, first A %7d, last step done %7d, items done %7d, exit %7d | update wait passed %7d, end %7d"
          % (it, r[2] - base, r[3] - base, r[6] - base, r[4] - base, r[5] - base, uu[1] - base, uu[6] - base))
print("== rollout step lengths inside the plan (cycles; epilogue warp 0 of CTA 0, last launch): step 0 .. H-2")
print(' '.join(str(int(t[0, k + 1][0] - t[0, k][0])) for k in range(c['H'] - 1)))
