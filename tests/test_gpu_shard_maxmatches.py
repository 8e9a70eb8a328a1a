"""MaxMatches with the targets sharded over two contexts ("Exception"): the
reference bounds a (window, k-mer) group over ALL targets in globally sorted order
(cmd/muscato_confirm/main.go:233-242, :424-448; cmd/muscato/main.go:318-385), so the shards flag
candidate groups at MaxMatches / 2, exchange the flagged keys, divert those groups' passing pairs
and one context replays the truncation.  The result must equal the oracle's unsharded matches.txt.
Run once in-process (two contexts on one device, the exchange done by hand through the C ABI) and
once as two gloo ranks through muscato_b200/dist.py."""
import os
import socket

import numpy as np
import pytest

from muscato_b200 import dist as mdist
from muscato_b200 import formats
from muscato_b200.config import Config
from tests import helpers

pytestmark = pytest.mark.gpu

MATCH_DTYPE = np.dtype([("read_id", "<u4"), ("gene_id", "<u4"), ("pos", "<u4"), ("nx", "<u4")])


def _case(mm, mode, seed):
    """Tandem-repeat targets + many near-identical reads: the same k-mers everywhere, in both shards."""
    rng = np.random.default_rng(seed)
    unit = helpers.random_dna(rng, 37)
    genes = []
    for gi in range(14):
        a = np.frombuffer(unit * 6, dtype=np.uint8).copy()
        m = rng.random(len(a)) < 0.02
        a[m] = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, int(m.sum()))]
        genes.append(bytes(a) + helpers.random_dna(rng, int(rng.integers(0, 30))))
    reads = []
    for _ in range(150):
        g = genes[int(rng.integers(0, len(genes)))]
        p = int(rng.integers(0, len(g) - 50))
        a = np.frombuffer(g[p:p + 50], dtype=np.uint8).copy()
        m = rng.random(50) < 0.03
        a[m] = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, int(m.sum()))]
        reads.append(bytes(a))
    cfgd = dict(Windows=[0, 12, 30], WindowWidth=10, MaxReadLength=50, PMatch=0.9, MinDinuc=0, MMTol=2,
                BloomSize=1000000, NumHash=6, MaxMatches=mm, MatchMode=mode, MaxConfirmProcs=3)
    return reads, genes, cfgd


def _oracle(tmp_path, reads, genes, cfgd):
    fq, gs, gi = helpers.write_case(str(tmp_path), reads, None, genes)
    out = helpers.oracle_pipeline(str(tmp_path), fq, gs, gi, cfgd)
    cfg = Config(**{k: v for k, v in cfgd.items() if k in Config.__dataclass_fields__}).apply_defaults()
    seqs, _, _ = formats.load_reads_sorted(out["reads_sorted"])
    return cfg, seqs, helpers.read_lines(out["matches"])


def _as_matches(g):
    m = np.zeros(len(g), dtype=MATCH_DTYPE)
    for i, f in enumerate(("read_id", "gene_id", "pos", "nx")):
        m[f] = g[:, i]
    return m


def _two_contexts(cfg, seqs, genes):
    """The protocol of dist.sharded_matches with the exchanges done by hand."""
    import torch
    from muscato_b200.engine import HotPath
    offs = np.concatenate([[0], np.cumsum([len(g) for g in genes])]).astype(np.uint64)
    shards = mdist.shard_targets(offs, 2)
    hps = []
    for lo, hi in shards:
        hp = HotPath(cfg, device=0)
        hp.set_shards(2)
        hp.set_reads(seqs)
        hp.set_targets(genes[lo:hi])
        hp.run_stages(0, 1 | 2)
        hps.append(hp)

    def exchange_best_and_combine():
        bests = [torch.as_tensor(hp.best_device(), device="cuda") for hp in hps]
        assert bests[0].numel() == len(seqs) + 1           # the flag element travels with the array
        gmin = torch.minimum(bests[0], bests[1])
        for b in bests:
            b.copy_(gmin)
        torch.cuda.synchronize()
        for hp in hps:
            hp.run_stages(0, 4)
        return gmin

    gmin = exchange_best_and_combine()
    flags = [hp.shard_overflow() for hp in hps]
    assert flags[0] == flags[1]
    surv, n_trunc = None, 0
    if flags[0]:
        keys = np.unique(np.concatenate([hp.overflow_keys() for hp in hps]))
        assert len(keys) > 0
        recs = [hp.divert_groups(keys, lo) for hp, (lo, _) in zip(hps, shards)]
        surv = hps[0].replay_diverted(np.concatenate(recs, axis=0))
        n_trunc = hps[0].stats()["n_overflow_groups"]
        b0 = torch.as_tensor(hps[0].best_device(), device="cuda")
        for s in surv:
            b0[int(s["read_id"])] = min(int(b0[int(s["read_id"])]), int(s["nx"]))
        gmin = exchange_best_and_combine()
        assert not any(hp.shard_overflow() for hp in hps)
    parts = []
    for hp, (lo, _) in zip(hps, shards):
        p = hp.fetch()
        parts.append(np.stack([p["read_id"], p["gene_id"] + lo, p["pos"], p["nx"]], axis=1).astype(np.int64))
    best_of = gmin.cpu().numpy()[surv["read_id"]] if surv is not None and len(surv) else None
    for hp in hps:
        hp.close()
    return mdist.merge_survivors(np.concatenate(parts, axis=0), surv, best_of, cfg.MMTol), flags[0], n_trunc


MM_CASES = [("best", 2), ("best", 3), ("best", 7), ("best", 40), ("first", 1), ("first", 5), ("first", 40)]


@pytest.mark.parametrize("mode,mm", MM_CASES, ids=[f"{m}{k}" for m, k in MM_CASES])
def test_two_shards_truncate_like_the_unsharded_reference(mode, mm, tmp_path, oracle_bin):
    reads, genes, cfgd = _case(mm, mode, 300 + mm)
    cfg, seqs, want = _oracle(tmp_path, reads, genes, cfgd)
    got, flagged, n_trunc = _two_contexts(cfg, seqs, genes)
    assert flagged and n_trunc > 0, "the case must actually truncate across the shards"
    assert formats.matches_lines(_as_matches(got), seqs, genes) == want


def test_two_shards_without_overflow_take_the_fast_path(tmp_path, oracle_bin):
    reads, genes, cfgd = _case(1000000, "best", 77)
    cfg, seqs, want = _oracle(tmp_path, reads, genes, cfgd)
    got, flagged, _ = _two_contexts(cfg, seqs, genes)
    assert not flagged
    assert formats.matches_lines(_as_matches(got), seqs, genes) == want


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _rank(rank, world, port, cfgd, seqs, genes, out_path, backend="gloo"):
    import torch
    import torch.distributed as dist
    from muscato_b200.engine import HotPath
    device = rank if backend == "nccl" else 0
    torch.cuda.set_device(device)
    if backend == "nccl":
        dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world,
                                device_id=torch.device("cuda", device))
    else:
        dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    cfg = Config(**{k: v for k, v in cfgd.items() if k in Config.__dataclass_fields__}).apply_defaults()
    offs = np.concatenate([[0], np.cumsum([len(g) for g in genes])]).astype(np.uint64)
    lo, hi = mdist.shard_targets(offs, world)[rank]
    with HotPath(cfg, device=device) as hp:
        hp.set_shards(world)
        hp.set_reads(seqs)
        hp.set_targets(genes[lo:hi])
        if backend == "nccl":
            # first pass sizes the buffers, the second one takes the stream-ordered (deferred) step:
            # the MaxMatches flag must survive both and both must give the same result
            first = mdist.sharded_matches(hp, lo, deferred=False)
            got = mdist.sharded_matches(hp, lo, deferred=True)
            assert rank != 0 or np.array_equal(first, got)
        else:
            got = mdist.sharded_matches(hp, lo, deferred=False)
        n_trunc = hp.stats()["n_overflow_groups"]
        torch.cuda.synchronize()
    if rank == 0:
        np.save(out_path, got)
        np.save(out_path + ".trunc.npy", np.array([n_trunc]))
    else:
        assert got is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("mode,mm", [("best", 3), ("first", 5)], ids=["best3", "first5"])
def test_two_gloo_ranks_through_dist(mode, mm, tmp_path, oracle_bin):
    import torch.multiprocessing as mp
    reads, genes, cfgd = _case(mm, mode, 300 + mm)
    cfg, seqs, want = _oracle(tmp_path, reads, genes, cfgd)
    out_path = str(tmp_path / "gathered.npy")
    mp.spawn(_rank, args=(2, _free_port(), cfgd, seqs, genes, out_path), nprocs=2, join=True)
    got = np.load(out_path)
    assert int(np.load(out_path + ".trunc.npy")[0]) > 0
    assert formats.matches_lines(_as_matches(got), seqs, genes) == want


@pytest.mark.parametrize("mode,mm", [("best", 3), ("first", 5)], ids=["best3", "first5"])
def test_two_nccl_ranks_through_dist(mode, mm, tmp_path, oracle_bin):
    """The same over NCCL on two GPUs, including the stream-ordered step (skipped on a one-GPU box)."""
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    reads, genes, cfgd = _case(mm, mode, 300 + mm)
    cfg, seqs, want = _oracle(tmp_path, reads, genes, cfgd)
    out_path = str(tmp_path / "gathered.npy")
    mp.spawn(_rank, args=(2, _free_port(), cfgd, seqs, genes, out_path, "nccl"), nprocs=2, join=True)
    got = np.load(out_path)
    assert int(np.load(out_path + ".trunc.npy")[0]) > 0
    assert formats.matches_lines(_as_matches(got), seqs, genes) == want
