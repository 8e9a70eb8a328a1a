"""Parity at BASELINE.json's full config[1] size (1M reads x 10k x 2 kb targets), where the CPU
oracle would take minutes: size-independent properties plus an exact closed-form check (SURVEY.md
App. A) on a random sample of reads."""
import numpy as np
import pytest

from muscato_b200 import dist as mdist
from muscato_b200 import gendat
from muscato_b200.config import Config

pytestmark = pytest.mark.gpu

CFG = dict(Windows=[0, 20], WindowWidth=15, MaxReadLength=100, PMatch=0.97, MinDinuc=5, MMTol=1,
           MaxMatches=1000000, MatchMode="best")


@pytest.fixture(scope="module")
def workload():
    syn = gendat.generate(1_000_000, 100, 10_000, 2000, seed=1, mutated_fraction=0.5, sub_rate=0.02)
    cfg = Config(**CFG).apply_defaults()
    from muscato_b200.engine import HotPath
    with HotPath(cfg, device=0, keep_ascii=True) as hp:
        hp.set_reads((syn.read_ascii, syn.read_offs))
        hp.set_targets((syn.target_ascii, syn.target_offs))
        hp.run()
        m = hp.fetch()
        st = hp.stats()
        hp.rebuild_and_run(3)            # idempotence: rebuild everything from the resident ASCII
        m2 = hp.fetch()
        hp.run_stages(0, 1 | 2)
        best = None
        import torch
        best = torch.as_tensor(hp.best_device(), device="cuda").cpu().numpy().copy()
        hp.run_stages(0, 4)
        m3 = hp.fetch()
    return syn, cfg, m, m2, m3, best, st


from tests.helpers import closed_form_matches, count_dinuc  # noqa: E402


def test_idempotent_and_staged_runs_agree(workload):
    _, _, m, m2, m3, _, st = workload
    assert len(m) > 500_000
    assert np.array_equal(m, m2) and np.array_equal(m, m3)
    assert st["n_candidates"] > 1_000_000 and st["n_pairs"] >= st["n_candidates"]


def test_every_match_is_a_true_alignment(workload):
    """Independent recomputation of the mismatch count of ALL returned matches (vectorised), the
    per-read MMTol rule, the fit rule and uniqueness."""
    syn, cfg, m, _, _, best, _ = workload
    L = syn.read_len
    reads = syn.read_ascii.reshape(-1, L)
    goff = syn.target_offs[m["gene_id"]].astype(np.int64)
    glen = (syn.target_offs[m["gene_id"].astype(np.int64) + 1] - syn.target_offs[m["gene_id"]]).astype(np.int64)
    pos = m["pos"].astype(np.int64)
    assert np.all(pos + L <= glen)
    nx = np.zeros(len(m), dtype=np.int64)
    for lo in range(0, len(m), 100_000):
        hi = min(len(m), lo + 100_000)
        idx = (goff[lo:hi] + pos[lo:hi])[:, None] + np.arange(L)[None, :]
        nx[lo:hi] = (syn.target_ascii[idx] != reads[m["read_id"][lo:hi]]).sum(axis=1)
    assert np.array_equal(nx, m["nx"].astype(np.int64))
    assert np.all(nx <= cfg.nmiss(L))
    # some window of the read matches the target exactly (Q2)
    okw = np.zeros(len(m), dtype=bool)
    for q1 in cfg.Windows:
        for lo in range(0, len(m), 200_000):
            hi = min(len(m), lo + 200_000)
            idx = (goff[lo:hi] + pos[lo:hi] + q1)[:, None] + np.arange(cfg.WindowWidth)[None, :]
            okw[lo:hi] |= (syn.target_ascii[idx] == reads[m["read_id"][lo:hi], q1:q1 + cfg.WindowWidth]).all(axis=1)
    assert okw.all()
    # MMTol: nx <= per-read minimum + MMTol, and the device's best array is that minimum
    mn = np.full(syn.n_reads, 10 ** 9, dtype=np.int64)
    np.minimum.at(mn, m["read_id"], nx)
    assert np.all(nx <= mn[m["read_id"]] + cfg.MMTol)
    matched = mn < 10 ** 9
    assert np.array_equal(best[matched].astype(np.int64), mn[matched])
    assert np.all(best[~matched] == 0x7F7F7F7F)
    # exact de-duplication across windows
    key = (m["read_id"].astype(np.int64) << 40) | (m["gene_id"].astype(np.int64) << 16) | pos
    assert len(np.unique(key)) == len(m)


def test_planted_reads_are_found(workload):
    """muscato_gendat plants read i%10 at offset i%10 of gene i < NumGene/2 (cmd/muscato_gendat/main.go:122-125)."""
    syn, _, m, _, _, _, _ = workload
    L = syn.read_len
    reads = syn.read_ascii.reshape(-1, L)
    got = set(zip(m["read_id"].tolist(), m["gene_id"].tolist(), m["pos"].tolist(), m["nx"].tolist()))
    tg = syn.target_ascii.reshape(-1, 2000)
    checked = 0
    for i in list(range(0, 40)) + list(range(4990, 5000)):
        j = i % 10
        seq = tg[i, j:j + L]
        rid = np.where((reads == seq).all(axis=1))[0]
        assert len(rid) == 1
        assert (int(rid[0]), i, j, 0) in got
        checked += 1
    assert checked == 50


def test_random_sample_of_reads_matches_closed_form(workload):
    """Exact expected result for 60 random reads (incl. planted-region ones) from the closed form of
    SURVEY.md App. A, evaluated with plain Python string search over the 20 Mbp database."""
    syn, cfg, m, _, _, _, _ = workload
    rng = np.random.default_rng(3)
    L = syn.read_len
    tgt = syn.target_ascii.tobytes()
    offs = syn.target_offs.astype(np.int64)
    matched_ids = np.unique(m["read_id"])
    sample = np.concatenate([rng.choice(matched_ids, 40, replace=False), rng.integers(0, syn.n_reads, 20)])
    by_read = {}
    for row in m[np.isin(m["read_id"], sample)]:
        by_read.setdefault(int(row["read_id"]), {})[(int(row["gene_id"]), int(row["pos"]))] = int(row["nx"])
    for rid in sample.tolist():
        read = syn.read_ascii[rid * L:(rid + 1) * L].tobytes()
        want = closed_form_matches(read, tgt, offs, cfg)
        if want:
            mn = min(want.values())
            want = {k: v for k, v in want.items() if v <= mn + cfg.MMTol}
        assert by_read.get(rid, {}) == want, rid


def test_sharded_halves_reproduce_the_whole(workload):
    """Target-range sharding (two contexts on one GPU standing in for two ranks): MIN of the best
    arrays + per-shard combine + concatenation == the unsharded result."""
    import torch
    from muscato_b200.engine import HotPath
    syn, cfg, m, _, _, _, _ = workload
    shards = mdist.shard_targets(syn.target_offs, 2)
    hps, bests = [], []
    for lo, hi in shards:
        hp = HotPath(cfg, device=0)
        hp.set_reads((syn.read_ascii, syn.read_offs))
        a, b = int(syn.target_offs[lo]), int(syn.target_offs[hi])
        hp.set_targets((syn.target_ascii[a:b], syn.target_offs[lo:hi + 1] - syn.target_offs[lo]))
        hp.run_stages(0, 1 | 2)
        hps.append(hp)
        bests.append(torch.as_tensor(hp.best_device(), device="cuda"))
    gmin = torch.minimum(bests[0], bests[1])
    parts = []
    for (lo, hi), hp, b in zip(shards, hps, bests):
        b.copy_(gmin)                      # what the NCCL MIN all-reduce leaves on every rank
        torch.cuda.synchronize()
        hp.run_stages(0, 4)
        part = hp.fetch()
        part["gene_id"] += lo
        parts.append(part)
        hp.close()
    got = np.concatenate(parts)
    got = got[np.lexsort((got["pos"], got["gene_id"], got["read_id"]))]
    assert np.array_equal(got, m)


@pytest.mark.parametrize("rev", [False, True], ids=["fwd", "rev"])
def test_full_size_outputs_equal_the_oracle_byte_for_byte(rev, tmp_path, oracle_bin):
    """BASELINE.json config[1] at FULL size (1M reads x 10k x 2 kb targets; with -rev 20k targets /
    40 Mbp): matches.txt, results.txt and the non-match fastq of the CUDA path against the oracle's
    files, byte for byte (cmd/muscato_confirm/main.go:171-250, cmd/muscato_combine_windows/main.go:36-60,
    cmd/muscato/main.go:507-676, cmd/muscato_nonmatch/main.go:57-113)."""
    import os
    from muscato_b200 import formats
    from muscato_b200.engine import HotPath
    from tests import helpers
    syn = gendat.generate(1_000_000, 100, 10_000, 2000, seed=1, mutated_fraction=0.5, sub_rate=0.02, rev=rev)
    cfgd = dict(CFG, BloomSize=400000000, NumHash=20, Threads=os.cpu_count() or 4)
    cfg = Config(**CFG).apply_defaults()
    work = str(tmp_path)
    fq, gs, gi = (os.path.join(work, n) for n in ("reads.fastq", "genes_seq.txt", "genes_ids.txt"))
    gendat.write_oracle_inputs(syn, fq, gs, gi)
    out = helpers.oracle_pipeline(work, fq, gs, gi, cfgd)
    seqs, counts, names = formats.load_reads_sorted(out["reads_sorted"])
    assert len(seqs) == syn.n_reads
    targets = syn.targets_list()
    with HotPath(cfg, device=0) as hp:
        hp.set_reads((syn.read_ascii, syn.read_offs))
        hp.set_targets((syn.target_ascii, syn.target_offs))
        hp.run()
        m = hp.fetch()
        nm = hp.nonmatch_ids()
    assert len(m) > 500_000
    assert formats.matches_lines(m, seqs, targets) == helpers.read_lines(out["matches"])
    gnames, glens = formats.load_gene_ids(gi)
    res = b"".join(ln + b"\n" for ln in formats.results_lines(m, seqs, counts, names, targets, gnames, glens))
    assert res == helpers.read_bytes(out["results"])
    assert formats.nonmatch_fastq_from_ids(nm, seqs, counts, names) == helpers.read_bytes(out["nonmatch"])
