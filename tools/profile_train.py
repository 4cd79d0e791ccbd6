"""Runs a few ensemble training steps (SURVEY 8 f1 shape) — the command profiled by
`ncu --metrics gpu__time_duration.sum` for profiles/r01_train_launches.csv."""
import argparse
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'ethz-safe-learning_b200'))
from simba_b200 import _device, _lib                       # noqa: E402
from simba_b200.models import MlpEnsemble                  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--steps', type=int, default=30)
ap.add_argument('--units', type=int, default=128)
ap.add_argument('--batch', type=int, default=64)
args = ap.parse_args()
E, B, IN, O, rows = 5, args.batch, 62, 60, 8192
rng = np.random.default_rng(0)
ens = MlpEnsemble(IN, O, E, batch_size=B, mlp_params=dict(n_layers=4, units=args.units))
x = torch.from_numpy(rng.uniform(0, 1, (rows, IN)).astype(np.float32)).cuda()
y = torch.from_numpy(rng.normal(0, 0.1, (rows, O)).astype(np.float32)).cuda()
index, _ = ens.batch_schedule(rows, args.steps)
dindex = torch.from_numpy(index).cuda()
losses = torch.empty(args.steps, device='cuda')
t = ens._ensure_trainer()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for rep in range(2):
    a.record()
    _lib.check(ens._lib.simba_trainer_fit(t, _device.ptr(x), _device.ptr(y), rows, _device.ptr(dindex),
                                          C.c_void_p(0), args.steps, _device.ptr(losses),
                                          _device.stream_ptr()))
    b.record()
    torch.cuda.synchronize()
    print("rep %d: %.1f us/step" % (rep, a.elapsed_time(b) * 1e3 / args.steps))
print("loss", float(losses[0]), float(losses[-1]))
