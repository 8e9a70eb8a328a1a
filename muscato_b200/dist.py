"""Multi-GPU plumbing (one process per GPU, torch.distributed): the path shards by target
range with the read key table replicated.  Two exchanges only:
  1. MIN all-reduce of the per-read best mismatch count, so that every rank applies the
     MMTol rule of cmd/muscato_combine_windows/main.go:36-60 against the global minimum;
  2. a variable-size gather of the compacted, already filtered matches to rank 0.
The functions take torch tensors and are backend agnostic (NCCL on GPUs, gloo in CPU tests)."""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np


def shard_targets(offs: np.ndarray, world: int) -> List[Tuple[int, int]]:
    """Contiguous gene-index ranges [lo, hi) balanced on bases; whole targets only."""
    G = len(offs) - 1
    total = int(offs[-1])
    out = []
    lo = 0
    for r in range(world):
        if r == world - 1:
            hi = G
        else:
            goal = (total * (r + 1)) // world
            hi = int(np.searchsorted(offs, goal, side="left"))
            hi = max(lo, min(G, hi))
        out.append((lo, hi))
        lo = hi
    return out


def allreduce_best(best, group=None):
    """In-place MIN over ranks of the int32 best_nx array (MSC_NO_MATCH = 0x7F7F7F7F)."""
    import torch.distributed as dist
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(best, op=dist.ReduceOp.MIN, group=group)
    return best


def sharded_step(hp, rebuild_what: int = 0, group=None, deferred: bool = True):
    """One hot-path step over this rank's target shard: screen + confirm, MIN all-reduce of the
    per-read best array, combine.  deferred=True is the stream-ordered form (MSC_STAGE_DEFER): the
    stages are only enqueued on the context's stream, the all-reduce is ordered behind them on that
    stream and the combine call synchronises once.  Use deferred=False for the first step on new
    inputs (it sizes the bounded output buffers, so that a deferred step cannot ask for a repeat)."""
    import torch
    import torch.distributed as dist
    multi = dist.is_initialized() and dist.get_world_size(group) > 1
    if not multi:
        hp.run_stages(rebuild_what, 1 | 2 | 4)
        return
    dev = torch.device("cuda", torch.cuda.current_device())
    if deferred:
        hp.run_stages(rebuild_what, 1 | 2 | 8)
        best = torch.as_tensor(hp.best_device(), device=dev)
        with torch.cuda.stream(torch.cuda.ExternalStream(hp.stream(), device=dev)):
            dist.all_reduce(best, op=dist.ReduceOp.MIN, group=group)
    else:
        hp.run_stages(rebuild_what, 1 | 2)
        best = torch.as_tensor(hp.best_device(), device=dev)
        dist.all_reduce(best, op=dist.ReduceOp.MIN, group=group)
        torch.cuda.synchronize()
    hp.run_stages(0, 4)


def gather_matches(local, gene_offset: int, dst: int = 0, group=None):
    """local: int32 tensor [n, 4] = (read, gene, pos, nx) with shard-local gene ids.
    Returns on rank `dst` the concatenation over ranks with global gene ids (else None)."""
    import torch
    import torch.distributed as dist
    local = local.reshape(-1, 4).clone()
    if local.numel():
        local[:, 1] += int(gene_offset)
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    n = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    mx = max(max(sizes), 1)
    pad = torch.zeros((mx, 4), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    bufs = [torch.zeros_like(pad) for _ in range(world)] if rank == dst else None
    dist.gather(pad, bufs, dst=dst, group=group)
    if rank != dst:
        return None
    return torch.cat([b[:s] for b, s in zip(bufs, sizes)], dim=0)
