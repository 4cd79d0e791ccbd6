#!/usr/bin/env python
"""Run under torchrun (one rank per GPU): checks that a population-sharded plan (NCCL all-gather of
the (return, cost) pairs per iteration) is bit-identical on every rank and equal to the same plan on
one GPU, then times the C3-shape plan.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tools/multi_gpu_check.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'ethz-safe-learning_b200')]

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from simba_b200 import _lib, distributed as sd, synthetic  # noqa: E402


def main():
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    torch.cuda.set_device(int(os.environ['LOCAL_RANK']))
    dist.init_process_group('nccl', device_id=torch.device('cuda', int(os.environ['LOCAL_RANK'])))
    ok = True
    for precision in ('fp32', 'bf16'):
        c = synthetic.make_workload('c1', N=160, K=16)                 # 160 candidates: divisible by 2/4/8
        single = synthetic.build_policy(c, 'penalty', precision=precision)
        a1, s1 = single.do_generate_action(c['state'], seed=77)
        mu1 = single.buffer(_lib.BUF_MU).cpu().numpy()
        el1 = single.buffer(_lib.BUF_ELITE, torch.int32).cpu().numpy()
        shard = synthetic.build_policy(c, 'penalty', precision=precision, rank=rank, world_size=world)
        sd.init_population_sharding(shard)
        a2, s2 = shard.do_generate_action(c['state'], seed=77)
        mu2 = shard.buffer(_lib.BUF_MU).cpu().numpy()
        el2 = shard.buffer(_lib.BUF_ELITE, torch.int32).cpu().numpy()
        same = np.array_equal(a1, a2) and s1 == s2 and np.array_equal(mu1, mu2) and np.array_equal(el1, el2)
        t = torch.tensor([1.0 if same else 0.0], device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        if rank == 0:
            print("[%s] sharded(%d) == single-GPU plan, bit for bit on every rank: %s  action %s score %.6f"
                  % (precision, world, bool(t.item() > 0), a2, s2))
        ok = ok and bool(t.item() > 0)
    # state sharding (BASELINE configs[3]): every rank plans its slice, all ranks get all actions
    c = synthetic.make_workload('tiny', S=4)
    states_all = synthetic.make_state(c['sensors'], seed=9, n_states=4 * world)
    pol = synthetic.build_policy(c, 'penalty', precision='bf16', seed=5)
    acts = sd.plan_states_sharded(pol, states_all)
    lo, hi = sd.shard_bounds(4 * world, world, rank)
    pol2 = synthetic.build_policy(c, 'penalty', precision='bf16', seed=5)
    mine, _ = pol2.do_generate_action(states_all[lo:hi])
    t = torch.from_numpy(acts).cuda()
    ref = t.clone(); dist.broadcast(ref, 0)
    good = torch.tensor([1.0 if (bool((ref == t).all()) and np.array_equal(acts[lo:hi], mine)) else 0.0], device='cuda')
    dist.all_reduce(good, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("plan_states_sharded: %d states over %d ranks, same on every rank and equal to the local plans: %s"
              % (4 * world, world, bool(good.item() > 0)))
    ok = ok and bool(good.item() > 0)
    # C3 shape timing
    c = synthetic.make_workload('c3')
    pol = synthetic.build_policy(c, 'penalty', precision='bf16', member_map='particle', rank=rank, world_size=world, seed=3)
    sd.init_population_sharding(pol)
    st = torch.from_numpy(np.ascontiguousarray(c['state']).reshape(1, -1)).cuda()
    pol.plan_device(st, seed=1)
    torch.cuda.synchronize(); dist.barrier()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record(); out = pol.plan_device(st, seed=2); ev[1].record(); torch.cuda.synchronize()
    ms = torch.tensor([ev[0].elapsed_time(ev[1])], device='cuda')
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ref = out[0].clone(); dist.broadcast(ref, 0)
    same = torch.tensor([1.0 if bool((ref == out[0]).all()) else 0.0], device='cuda')
    dist.all_reduce(same, op=dist.ReduceOp.MIN)
    if rank == 0:
        fl = synthetic.flops_per_transition(c['O'], c['A'], c['L'], c['U']) * c['I'] * c['H'] * c['P'] * c['N']
        print("C3 (N=65536, P=32, H=30) on %d GPU(s): %.1f ms/plan, %.1f TFLOP/s total, replicas identical: %s"
              % (world, ms.item(), fl / ms.item() / 1e9, bool(same.item() > 0)))
    dist.destroy_process_group()
    if not ok:
        sys.exit(1)


if __name__ == '__main__':
    main()
