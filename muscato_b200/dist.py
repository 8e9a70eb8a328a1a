"""Multi-GPU plumbing (one process per GPU, torch.distributed): the path shards by target
range with the read key table replicated.  Two exchanges only:
  1. MIN all-reduce of the per-read best mismatch count, so that every rank applies the
     MMTol rule of cmd/muscato_combine_windows/main.go:36-60 against the global minimum;
  2. a variable-size gather of the compacted, already filtered matches to rank 0.
Exception: MaxMatches bounds a (window, k-mer) group over ALL targets.  A context
in sharded mode (msc_set_shards) flags groups with more than MaxMatches / world passing pairs; the
flag rides in element [n_reads] of the best array through exchange 1.  Only when it is set do the
ranks exchange the flagged key fingerprints and send the diverted pairs of those groups to one rank,
which replays the reference's sequential truncation (resolve_shard_overflow).
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


def _is_nccl(group=None) -> bool:
    import torch.distributed as dist
    return "nccl" in str(dist.get_backend(group)).lower()


def _allreduce_min(t, group=None):
    """MIN all-reduce of a device tensor; staged through the host when the backend is not NCCL
    (gloo in the two-process GPU test)."""
    import torch.distributed as dist
    if t.is_cuda and not _is_nccl(group):
        h = t.cpu()
        dist.all_reduce(h, op=dist.ReduceOp.MIN, group=group)
        t.copy_(h)
    else:
        dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
    return t


def allgather_concat(t, group=None):
    """Concatenation over ranks (on every rank) of tensors that differ in their first dimension."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return t
    world = dist.get_world_size(group)
    n = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(x.item()) for x in sizes]
    mx = max(max(sizes), 1)
    pad = torch.zeros((mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[: t.shape[0]] = t
    bufs = [torch.zeros_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([b[:k] for b, k in zip(bufs, sizes)], dim=0)


def gather_concat(t, dst: int = 0, group=None):
    """Concatenation over ranks, on rank `dst` only (None elsewhere), of tensors that differ in their
    first dimension: sizes are all-gathered, the payload travels as one padded gather."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return t
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    n = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(x.item()) for x in sizes]
    mx = max(max(sizes), 1)
    pad = torch.zeros((mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[: t.shape[0]] = t
    bufs = [torch.zeros_like(pad) for _ in range(world)] if rank == dst else None
    dist.gather(pad, bufs, dst=dst, group=group)
    if rank != dst:
        return None
    return torch.cat([b[:k] for b, k in zip(bufs, sizes)], dim=0)


def resolve_shard_overflow(hp, gene_offset: int, group=None, dst: int = 0):
    """Rare path, entered by ALL ranks when hp.shard_overflow() is set after a sharded step: some
    (window, k-mer) group may hold more than MaxMatches passing pairs over all shards
    (cmd/muscato_confirm/main.go:233-242, :424-448).  Exchanges the flagged key fingerprints,
    diverts those groups' passing pairs on every rank, replays the reference's truncation on rank
    `dst`, folds the survivors into the per-read best array, repeats the MIN all-reduce and the
    combine.  Returns the survivors (MATCH_DTYPE, global gene ids) on `dst`, None elsewhere; the
    caller merges those that meet the MMTol rule with the gathered matches (merge_survivors)."""
    import torch
    import torch.distributed as dist
    rank = dist.get_rank(group)
    dev = torch.device("cuda", torch.cuda.current_device())
    xdev = dev if _is_nccl(group) else torch.device("cpu")
    keys = hp.overflow_keys()
    allk = allgather_concat(torch.from_numpy(keys.view(np.int64)).to(xdev), group=group)
    allk = np.unique(allk.cpu().numpy().view(np.uint64))
    recs = hp.divert_groups(allk, gene_offset)      # the context now holds the undiverted matches and their best
    # records go to one rank only (they can be large)
    allr = gather_concat(torch.from_numpy(recs).to(xdev), dst=dst, group=group)
    best = torch.as_tensor(hp.best_device(), device=dev)
    surv = None
    if rank == dst:
        surv = hp.replay_diverted(allr.cpu().numpy())
        if len(surv):
            rid, inv = np.unique(surv["read_id"], return_inverse=True)
            mn = np.full(len(rid), np.iinfo(np.int32).max, dtype=np.int64)
            np.minimum.at(mn, inv, surv["nx"].astype(np.int64))
            idx = torch.from_numpy(rid.astype(np.int64)).to(dev)
            best[idx] = torch.minimum(best[idx], torch.from_numpy(mn.astype(np.int32)).to(dev))
    _allreduce_min(best, group)
    torch.cuda.synchronize()
    hp.run_stages(0, 4)
    return surv


def merge_survivors(gathered: np.ndarray, surv: Optional[np.ndarray], best_of_surv: Optional[np.ndarray], mmtol: int):
    """gathered: int [n, 4] (read, global gene, pos, nx) of all ranks after the combine; surv: the
    survivors of the replay with best_of_surv[i] = global best of surv[i]'s read.  Returns the union,
    exact duplicates removed (`sort -u`), ordered by (read, gene, pos)."""
    g = np.asarray(gathered, dtype=np.int64).reshape(-1, 4)
    if surv is not None and len(surv):
        keep = surv["nx"].astype(np.int64) <= np.asarray(best_of_surv, dtype=np.int64) + int(mmtol)
        s = np.stack([surv[f][keep].astype(np.int64) for f in ("read_id", "gene_id", "pos", "nx")], axis=1)
        g = np.concatenate([g, s], axis=0)
    if len(g):
        g = np.unique(g, axis=0)
        g = g[np.lexsort((g[:, 2], g[:, 1], g[:, 0]))]
    return g


def sharded_matches(hp, gene_offset: int, rebuild_what: int = 0, group=None, deferred: bool = False, dst: int = 0):
    """One complete sharded pass: step, the MaxMatches protocol when a shard flagged a group, gather.
    Returns on `dst` an int64 [n, 4] array (read, global gene, pos, nx) ordered by (read, gene, pos)."""
    import torch
    import torch.distributed as dist
    sharded_step(hp, rebuild_what, group=group, deferred=deferred)
    multi = dist.is_initialized() and dist.get_world_size(group) > 1
    surv = None
    if multi and hp.n_shards > 1 and hp.shard_overflow():
        surv = resolve_shard_overflow(hp, gene_offset, group=group, dst=dst)
    dev = torch.device("cuda", torch.cuda.current_device())
    holder, n = hp.matches_device()
    local = torch.as_tensor(holder, device=dev).reshape(-1, 4) if n else torch.zeros((0, 4), dtype=torch.int32, device=dev)
    if multi and not _is_nccl(group):
        local = local.cpu()
    allm = gather_matches(local, gene_offset, dst=dst, group=group)
    if allm is None:
        return None
    best_of = None
    if surv is not None and len(surv):
        best = torch.as_tensor(hp.best_device(), device=dev)
        best_of = best[torch.from_numpy(surv["read_id"].astype(np.int64)).to(dev)].cpu().numpy()
    return merge_survivors(allm.cpu().numpy(), surv, best_of, hp.cfg.MMTol)


def sharded_step(hp, rebuild_what: int = 0, group=None, deferred: bool = False):
    """One hot-path step over this rank's target shard: screen + confirm, MIN all-reduce of the
    per-read best array, combine.  deferred=True is the stream-ordered form (MSC_STAGE_DEFER): the
    stages are only enqueued on the context's stream, the all-reduce is ordered behind them on that
    stream and the combine call synchronises once.  A deferred step can end in MSC_ERR_AGAIN on ONE
    rank only (a bounded output buffer of that rank was grown, or MaxMatches truncation applies
    there): no rank may raise between collectives, so the outcome is MAX-all-reduced as a status
    word and, if any rank has to repeat, ALL ranks repeat the step together in the synchronising
    form (which sizes the buffers and cannot ask for a repeat)."""
    import torch
    import torch.distributed as dist
    from . import _capi
    from .engine import MuscatoError
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
        again = 0
        try:
            hp.run_stages(0, 4)
        except MuscatoError as e:
            if e.code != _capi.MSC_ERR_AGAIN:
                raise
            again = 1
        flag = torch.tensor([again], dtype=torch.int32, device=dev if _is_nccl(group) else "cpu")
        dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=group)
        if int(flag.item()) == 0:
            return
        rebuild_what = 0   # the inputs are packed; only the stages are repeated
    hp.run_stages(rebuild_what, 1 | 2)
    best = torch.as_tensor(hp.best_device(), device=dev)
    _allreduce_min(best, group)
    torch.cuda.synchronize()
    hp.run_stages(0, 4)


def gather_matches(local, gene_offset: int, dst: int = 0, group=None):
    """local: int32 tensor [n, 4] = (read, gene, pos, nx) with shard-local gene ids.
    Returns on rank `dst` the concatenation over ranks with global gene ids (else None)."""
    local = local.reshape(-1, 4).clone()
    if local.numel():
        local[:, 1] += int(gene_offset)
    return gather_concat(local, dst=dst, group=group)
