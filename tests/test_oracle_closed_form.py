"""The oracle against a SECOND, independent restatement of the path: the closed form of SURVEY.md
App. A evaluated by plain string search (tests/helpers.closed_form_matches).  The reference's own
fixtures pin exact matching, multi-mapping, ordering and the file formats; nothing in them pins
mismatch counting, MMTol, the position-0 rule (Q1), X handling (Q4), the float64 mismatch budget
(Q5) or reads overhanging a target end (Q6) -- those rest on the code lines the oracle cites, so two
restatements written from those lines must at least agree with each other on randomised inputs.
CPU only."""
import os

import numpy as np
import pytest

from muscato_b200 import formats
from muscato_b200.config import Config
from tests import helpers


def _mutate(rng, seq: bytes, rate: float, alphabet: bytes = b"ACGT") -> bytes:
    a = np.frombuffer(seq, dtype=np.uint8).copy()
    m = rng.random(len(a)) < rate
    a[m] = np.frombuffer(alphabet, dtype=np.uint8)[rng.integers(0, len(alphabet), int(m.sum()))]
    return bytes(a)


def _case(rng, n_genes, gene_lens, n_reads, read_lens, sub_rate, x_rate):
    alpha_t = b"ACGT"
    genes = [helpers.random_dna(rng, int(rng.choice(gene_lens)), alpha_t) for _ in range(n_genes)]
    genes.append(genes[0][:40] + genes[1][:60])            # shared segments: multi-mapping
    genes.append(_mutate(rng, genes[2], 0.03))              # a diverged copy: same read, different nx (MMTol)
    if x_rate > 0:
        genes = [_mutate(rng, g, x_rate, b"X") for g in genes]
    reads = []
    for i in range(n_reads):
        g = genes[int(rng.integers(0, len(genes)))]
        L = int(rng.choice(read_lens))
        if len(g) < L:
            reads.append(helpers.random_dna(rng, L))
            continue
        where = i % 4
        p = 0 if where == 0 else (len(g) - L if where == 1 else int(rng.integers(0, len(g) - L + 1)))
        r = _mutate(rng, g[p:p + L], sub_rate)
        if x_rate > 0:
            r = _mutate(rng, r, x_rate, b"N")               # prep_reads turns any non-ACGT byte into X
        reads.append(r)
    for _ in range(n_reads // 10):
        reads.append(helpers.random_dna(rng, int(rng.choice(read_lens))))   # noise
    return reads, genes


CASES = {
    # name: (config, n_genes, gene lengths, n_reads, read lengths, substitution rate, X rate)
    "w15_two_windows": (dict(Windows=[0, 20], WindowWidth=15, MaxReadLength=60, PMatch=0.93, MinDinuc=3, MMTol=1),
                        12, [150, 200, 90], 160, [60, 50, 40], 0.03, 0.0),
    "q1_long_reads_at_position_0": (dict(Windows=[0, 30, 60], WindowWidth=12, MaxReadLength=120, PMatch=0.95, MinDinuc=0, MMTol=0),
                                    10, [300, 130], 120, [120, 100, 88, 70], 0.02, 0.0),
    "x_bases_everywhere": (dict(Windows=[0, 10, 25], WindowWidth=8, MaxReadLength=50, PMatch=0.9, MinDinuc=2, MMTol=2),
                           10, [120, 80], 150, [50, 45, 33], 0.03, 0.02),
    "short_targets_and_windows_off_the_read": (dict(Windows=[0, 16, 40], WindowWidth=10, MaxReadLength=64, PMatch=0.9, MinDinuc=0, MMTol=1),
                                               14, [9, 30, 64, 100], 140, [64, 45, 20, 12], 0.04, 0.0),
    "exact_only": (dict(Windows=[0, 5], WindowWidth=4, MaxReadLength=30, PMatch=1.0, MinDinuc=1, MMTol=0),
                   8, [60, 45], 100, [30, 20, 10], 0.01, 0.0),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_agrees_with_the_closed_form(name, tmp_path, oracle_bin):
    cfgd, n_genes, gene_lens, n_reads, read_lens, sub_rate, x_rate = CASES[name]
    cfgd = dict(cfgd, BloomSize=1000000, NumHash=6, MaxMatches=1000000, MatchMode="best")
    rng = np.random.default_rng(sum(map(ord, name)))
    reads, genes = _case(rng, n_genes, gene_lens, n_reads, read_lens, sub_rate, x_rate)
    fq, gs, gi = helpers.write_case(str(tmp_path), reads, None, genes)
    out = helpers.oracle_pipeline(str(tmp_path), fq, gs, gi, cfgd)
    cfg = Config(**{k: v for k, v in cfgd.items() if k in Config.__dataclass_fields__}).apply_defaults()
    seqs, _, _ = formats.load_reads_sorted(out["reads_sorted"])
    tgt = b"".join(genes)
    offs = np.concatenate([[0], np.cumsum([len(g) for g in genes])]).astype(np.int64)
    want = set()
    n_pre = 0
    for r in seqs:
        hits = helpers.closed_form_matches(r, tgt, offs, cfg)
        n_pre += len(hits)
        if not hits:
            continue
        best = min(hits.values())
        for (g, pos), nx in hits.items():
            if nx <= best + cfg.MMTol:                       # cmd/muscato_combine_windows/main.go:36-60
                want.add(b"%s\t%s\t%d\t%d\t%011d" % (r, genes[g][pos:pos + len(r)], pos, nx, g))
    got = helpers.read_lines(out["matches"])
    assert len(got) == len(set(got))                         # sort -u left no duplicates
    assert set(got) == want
    assert len(want) > 20 and n_pre >= len(want)             # the case exercises something
    if cfg.MMTol < 3 and name != "exact_only":
        assert any(int(ln.split(b"\t")[3]) > 0 for ln in got)  # mismatching alignments are present


def test_closed_form_knows_the_quirks():
    """Q1 (literal 100 at target position 0), Q5 (float64 budget), Q6 (no overhang), Q4 (X == X)."""
    cfg = Config(Windows=[0, 20], WindowWidth=10, MaxReadLength=120, PMatch=0.9, MinDinuc=0, MMTol=0).apply_defaults()
    rng = np.random.default_rng(5)
    g = helpers.random_dna(rng, 200)
    offs = np.array([0, 200], dtype=np.int64)
    # a 95-base read at position 0: not through window 0 (95 > 100 - 10) but through window 1 (p = 20)
    assert helpers.closed_form_matches(g[:95], g, offs, cfg) == {(0, 0): 0}
    cfg1 = Config(Windows=[0], WindowWidth=10, MaxReadLength=120, PMatch=0.9, MinDinuc=0, MMTol=0).apply_defaults()
    assert helpers.closed_form_matches(g[:95], g, offs, cfg1) == {}
    assert helpers.closed_form_matches(g[:90], g, offs, cfg1) == {(0, 0): 0}
    # the budget: int((1 - 0.9) * 100.0) == 9, not 10
    assert cfg.nmiss(100) == 9
    # a read that overhangs the target end is never partially matched
    assert helpers.closed_form_matches(g[150:] + b"ACGTACGTAC", g, offs, cfg) == {}
    # X equals X, X differs from a base
    gx = g[:50] + b"X" + g[51:]
    rx = gx[30:90]
    assert helpers.closed_form_matches(rx, gx, offs, cfg) == {(0, 30): 0}
    assert helpers.closed_form_matches(rx, g, offs, cfg) == {(0, 30): 1}


def _repeats_case(seed):
    """Tandem-repeat targets and many near-identical reads: key groups with far more passing pairs than
    a small MaxMatches (the workload of the MaxMatches parity tests on the GPU)."""
    rng = np.random.default_rng(seed)
    unit = helpers.random_dna(rng, 37)
    genes = [_mutate(rng, unit * 6, 0.02) + helpers.random_dna(rng, int(rng.integers(0, 30))) for _ in range(12)]
    reads = []
    for _ in range(150):
        g = genes[int(rng.integers(0, len(genes)))]
        p = int(rng.integers(0, len(g) - 50))
        reads.append(_mutate(rng, g[p:p + 50], 0.03))
    return reads, genes


@pytest.mark.parametrize("mode,mm", [("best", 1), ("best", 3), ("best", 7), ("best", 40), ("first", 1), ("first", 5),
                                     ("first", 40), ("best", 1000000)],
                         ids=lambda v: str(v))
def test_oracle_maxmatches_agrees_with_the_sequential_restatement(mode, mm, tmp_path, oracle_bin):
    """Q7: the kept set depends on the order in which the reference meets the pairs; the oracle (C++,
    GNU sort) and a pure-Python restatement of the same code lines must keep the same lines."""
    reads, genes = _repeats_case(100 + mm)
    cfgd = dict(Windows=[0, 12, 30], WindowWidth=10, MaxReadLength=50, PMatch=0.9, MinDinuc=0, MMTol=2,
                BloomSize=1000000, NumHash=6, MaxMatches=mm, MatchMode=mode, MaxConfirmProcs=3)
    fq, gs, gi = helpers.write_case(str(tmp_path), reads, None, genes)
    out = helpers.oracle_pipeline(str(tmp_path), fq, gs, gi, cfgd)
    cfg = Config(**{k: v for k, v in cfgd.items() if k in Config.__dataclass_fields__}).apply_defaults()
    seqs, _, _ = formats.load_reads_sorted(out["reads_sorted"])
    want = helpers.sequential_restatement(seqs, genes, cfg)
    got = helpers.read_lines(out["matches"])
    assert len(got) == len(set(got)) and set(got) == want
    if mm < 1000:
        full = helpers.sequential_restatement(seqs, genes, Config(**dict(cfg.__dict__, MaxMatches=1000000)))
        assert want != full, "the limit must actually truncate"


def test_sequential_restatement_equals_closed_form_without_truncation(tmp_path, oracle_bin):
    """With MaxMatches out of reach the sequential structure and the closed form describe the same set."""
    cfgd, n_genes, gene_lens, n_reads, read_lens, sub_rate, x_rate = CASES["x_bases_everywhere"]
    cfgd = dict(cfgd, BloomSize=1000000, NumHash=6, MaxMatches=1000000, MatchMode="best")
    rng = np.random.default_rng(99)
    reads, genes = _case(rng, n_genes, gene_lens, n_reads, read_lens, sub_rate, x_rate)
    fq, gs, gi = helpers.write_case(str(tmp_path), reads, None, genes)
    out = helpers.oracle_pipeline(str(tmp_path), fq, gs, gi, cfgd)
    cfg = Config(**{k: v for k, v in cfgd.items() if k in Config.__dataclass_fields__}).apply_defaults()
    seqs, _, _ = formats.load_reads_sorted(out["reads_sorted"])
    assert helpers.sequential_restatement(seqs, genes, cfg) == set(helpers.read_lines(out["matches"]))
