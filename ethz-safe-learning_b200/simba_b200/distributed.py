"""Multi-GPU plumbing (one process per GPU, torch.distributed for the rendezvous).

Two ways the planning call shards:
  * population sharding — one plan, candidates split over ranks, one NCCL all-gather of the
    per-candidate (return, cost) pairs per CEM iteration inside libsimba_b200.so, select + refit
    replicated (bit-identical on every rank);
  * state sharding — independent states per rank, no data-path collective; actions gathered at the end.
"""
import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from . import _lib


def shard_bounds(n, world_size, rank):
    """Contiguous shard [lo, hi) of n units for `rank` (requires world_size | n, like the C library)."""
    if n % world_size:
        raise _lib.SimbaError(-2, "%d units not divisible by world_size %d" % (n, world_size))
    per = n // world_size
    return rank * per, (rank + 1) * per


def broadcast_unique_id(group=None, src=0):
    """Rank `src` asks NCCL for a unique id through the C-ABI; everybody receives its 128 bytes."""
    rank = dist.get_rank(group)
    backend = dist.get_backend(group)
    device = torch.device('cuda', torch.cuda.current_device()) if backend == 'nccl' else torch.device('cpu')
    buf = torch.zeros(128, dtype=torch.uint8)
    if rank == src:
        raw = (C.c_char * 128)()
        _lib.check(_lib.load().simba_nccl_unique_id(raw))
        buf = torch.frombuffer(bytearray(raw.raw), dtype=torch.uint8).clone()
    buf = buf.to(device)
    dist.broadcast(buf, src, group=group)
    return bytes(buf.cpu().numpy().tobytes())


def init_population_sharding(policy, group=None):
    """Create the NCCL communicator of a policy built with rank / world_size."""
    if policy.world_size > 1:
        policy.init_distributed(broadcast_unique_id(group))
    return policy


def plan_states_sharded(policy, states, group=None):
    """BASELINE configs[3]: `states` [S_total, O] (same on every rank) -> actions [S_total, A] on every
    rank. `policy` must have been built with n_states = S_total / world_size. No collective on the data
    path; one all_gather of the actions at the end."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    lo, hi = shard_bounds(states.shape[0], world, rank)
    local = np.ascontiguousarray(states[lo:hi], dtype=np.float32)
    actions, _ = policy.do_generate_action(local)
    actions = np.atleast_2d(actions)
    backend = dist.get_backend(group)
    device = torch.device('cuda', torch.cuda.current_device()) if backend == 'nccl' else torch.device('cpu')
    mine = torch.from_numpy(actions).to(device)
    out = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(out, mine, group=group)
    return torch.cat(out, 0).cpu().numpy()
