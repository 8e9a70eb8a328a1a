"""World-size-2 gloo test of the N>1 host plumbing (muscato_b200/dist.py): sharding by target
range + MIN all-reduce of the per-read best mismatch count + local MMTol filter + gather must
equal the single-process combine of the oracle (cmd/muscato_combine_windows/main.go:36-60)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from muscato_b200 import dist as mdist
from muscato_b200 import formats
from tests import helpers

NO_MATCH = 0x7F7F7F7F


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, pre, n_reads, shards, mmtol, out_path):
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    lo, hi = shards[rank]
    mine = pre[(pre[:, 1] >= lo) & (pre[:, 1] < hi)].copy()
    mine[:, 1] -= lo                                   # shard-local gene ids, as a rank's context sees them
    best = np.full(n_reads, NO_MATCH, dtype=np.int32)  # what msc_confirm leaves in msc_best_device
    np.minimum.at(best, mine[:, 0], mine[:, 3])
    tbest = torch.from_numpy(best)
    mdist.allreduce_best(tbest)
    keep = mine[:, 3] <= tbest.numpy()[mine[:, 0]] + mmtol  # msc_combine on this rank
    local = torch.from_numpy(mine[keep].astype(np.int32))
    allm = mdist.gather_matches(local, gene_offset=lo, dst=0)
    if rank == 0:
        np.save(out_path, allm.numpy())
    else:
        assert allm is None
    dist.barrier()
    dist.destroy_process_group()


def test_shard_targets_balances_on_bases():
    offs = np.concatenate([[0], np.cumsum([10, 1000, 10, 10, 500, 500, 20])]).astype(np.uint64)
    sh = mdist.shard_targets(offs, 3)
    assert sh[0][0] == 0 and sh[-1][1] == 7
    assert all(a[1] == b[0] for a, b in zip(sh, sh[1:]))
    sizes = [int(offs[h] - offs[l]) for l, h in sh]
    assert sum(sizes) == int(offs[-1]) and max(sizes) <= 1100
    assert mdist.shard_targets(offs, 1) == [(0, 7)]
    assert len(mdist.shard_targets(offs, 16)) == 16


def test_two_rank_exchange_equals_global_combine(tmp_path, oracle_bin):
    rng = np.random.default_rng(21)
    genes = [helpers.random_dna(rng, 300) for _ in range(24)]
    genes += [genes[3], genes[5][:150] + genes[7][150:]]  # multi-mapping across shards
    for src in (1, 4, 20):                                 # diverged copies: same read, different nx
        a = np.frombuffer(genes[src], dtype=np.uint8).copy()
        m = rng.random(len(a)) < 0.03
        a[m] = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, int(m.sum()))]
        genes.append(bytes(a))
    reads = []
    for _ in range(400):
        g = genes[int(rng.integers(0, len(genes)))]
        p = int(rng.integers(0, len(g) - 60))
        a = np.frombuffer(g[p:p + 60], dtype=np.uint8).copy()
        m = rng.random(60) < 0.03
        a[m] = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, int(m.sum()))]
        reads.append(bytes(a))
    cfgd = dict(Windows=[0, 20, 40], WindowWidth=12, MaxReadLength=60, PMatch=0.9, MinDinuc=0, MMTol=0,
                BloomSize=2000000, NumHash=6, MaxMatches=1000000, MatchMode="best")
    fq, gs, gi = helpers.write_case(str(tmp_path), reads, None, genes)
    out = helpers.oracle_pipeline(str(tmp_path), fq, gs, gi, cfgd)
    seqs, _, _ = formats.load_reads_sorted(out["reads_sorted"])
    idx = {s: i for i, s in enumerate(seqs)}

    def parse(path):
        rows = []
        for ln in helpers.read_lines(path):
            f = ln.split(b"\t")
            rows.append((idx[f[0]], int(f[4]), int(f[2]), int(f[3])))
        return np.array(sorted(rows), dtype=np.int64).reshape(-1, 4)

    pre = parse(os.path.join(out["tmp"], "rmatch_su.txt"))   # union over windows, before MMTol
    want = parse(out["matches"])                              # after the global MMTol rule
    assert len(pre) > len(want) > 0                           # MMTol actually filters something
    offs = np.concatenate([[0], np.cumsum([len(g) for g in genes])]).astype(np.uint64)
    shards = mdist.shard_targets(offs, 2)
    out_path = str(tmp_path / "gathered.npy")
    mp.spawn(_worker, args=(2, _free_port(), pre, len(seqs), shards, cfgd["MMTol"], out_path), nprocs=2, join=True)
    got = np.load(out_path).astype(np.int64)
    got = got[np.lexsort((got[:, 3], got[:, 2], got[:, 1], got[:, 0]))]
    assert np.array_equal(got, want)


def _ragged_worker(rank, world, port, out_path):
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    keys = torch.arange(3 * rank, dtype=torch.int64) + 100 * rank            # rank 0 contributes nothing
    allk = mdist.allgather_concat(keys)
    recs = torch.full((2 - rank, 5), rank + 1, dtype=torch.uint8)             # 2-D, ragged first dimension
    allr = mdist.allgather_concat(recs)
    best = torch.tensor([5, 9, 0x7F7F7F7F, 3 + rank], dtype=torch.int32)
    mdist._allreduce_min(best)
    onr = mdist.gather_concat(recs, dst=1)                                    # records go to ONE rank
    assert (onr is None) == (rank != 1)
    if rank == 1:
        assert onr.tolist() == [[1] * 5, [1] * 5, [2] * 5]
    if rank == 0:
        np.savez(out_path, k=allk.numpy(), r=allr.numpy(), b=best.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_ragged_allgather_of_keys_and_records(tmp_path):
    """The exchanges of the cross-shard MaxMatches protocol (dist.resolve_shard_overflow): key
    fingerprints and diverted-pair records differ in number per rank."""
    out_path = str(tmp_path / "ragged.npz")
    mp.spawn(_ragged_worker, args=(2, _free_port(), out_path), nprocs=2, join=True)
    z = np.load(out_path)
    assert z["k"].tolist() == [100, 101, 102]
    assert z["r"].tolist() == [[1] * 5, [1] * 5, [2] * 5]
    assert z["b"].tolist() == [5, 9, 0x7F7F7F7F, 3]


def test_merge_survivors_applies_mmtol_and_sort_u():
    md = np.dtype([("read_id", "<u4"), ("gene_id", "<u4"), ("pos", "<u4"), ("nx", "<u4")])
    gathered = np.array([[2, 7, 10, 1], [0, 3, 5, 0], [2, 1, 4, 2]], dtype=np.int64)
    surv = np.array([(2, 7, 10, 1), (2, 0, 9, 3), (1, 4, 4, 2), (0, 9, 9, 1)], dtype=md)
    best_of = np.array([1, 1, 2, 0])             # global best of each survivor's read
    got = mdist.merge_survivors(gathered, surv, best_of, mmtol=1)
    # (2,7,10,1) is a duplicate of a gathered line; (2,0,9,3) fails nx <= best + MMTol; the rest joins
    assert got.tolist() == [[0, 3, 5, 0], [0, 9, 9, 1], [1, 4, 4, 2], [2, 1, 4, 2], [2, 7, 10, 1]]
    assert mdist.merge_survivors(gathered, None, None, 0).tolist() == [[0, 3, 5, 0], [2, 1, 4, 2], [2, 7, 10, 1]]
    assert mdist.merge_survivors(np.zeros((0, 4)), None, None, 0).shape == (0, 4)
