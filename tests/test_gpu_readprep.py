"""Device-side prepReads (msc_prep_reads) against the oracle's reads_sorted and
the reference fixtures: the GPU sorts / collapses the raw fastq sequences, the host joins names."""
import json
import os

import numpy as np
import pytest

from muscato_b200 import formats
from muscato_b200.config import Config
from muscato_b200.engine import HotPath

from . import helpers

pytestmark = pytest.mark.gpu


def device_prep(hp, fastq: bytes, cfg):
    names, raw = formats.parse_fastq(fastq)
    kept, uniq = hp.prep_reads(raw, cfg.MinReadLength)
    perm, gs = hp.read_groups()
    assert len(perm) == kept and len(gs) == uniq + 1 and (uniq == 0 or int(gs[-1]) == kept)
    return formats.uniqify_from_groups(names, raw, perm, gs, cfg.MaxReadLength)


@pytest.mark.parametrize("case", ["00", "01", "02", "03", "04"])
def test_fixture_reads_through_device_prep(case, tmp_path, oracle_bin):
    """The five reference fixtures, starting from reads.fastq: device prepReads, then the hot path;
    result.txt and the non-match fastq byte for byte."""
    src = os.path.join(helpers.GOLDEN, "muscato", case)
    cfgd = json.load(open(os.path.join(src, "config.json")))
    seq, ids = str(tmp_path / "genes_seq.txt"), str(tmp_path / "genes_ids.txt")
    helpers.oracle_prep_targets(os.path.join(src, "genes.txt"), seq, ids, rev=(case == "04"))
    cfg = Config(**{k: v for k, v in cfgd.items() if k in Config.__dataclass_fields__}).apply_defaults()
    fq = helpers.read_bytes(os.path.join(src, "reads.fastq"))
    targets = formats.load_targets(seq)
    gnames, glens = formats.load_gene_ids(ids)
    with HotPath(cfg, device=0, keep_ascii=True) as hp:
        seqs, counts, rnames = device_prep(hp, fq, cfg)
        assert (seqs, counts, rnames) == formats.prep_reads_uniqify(fq, cfg.MinReadLength, cfg.MaxReadLength)
        assert hp.unique_reads() == seqs
        hp.set_targets(targets)
        hp.run()
        m = hp.fetch()
    res = b"".join(ln + b"\n" for ln in formats.results_lines(m, seqs, counts, rnames, targets, gnames, glens))
    assert res == helpers.read_bytes(os.path.join(src, "result_e.txt"))
    assert formats.nonmatch_fastq(m, seqs, counts, rnames) == helpers.read_bytes(os.path.join(src, "result.nonmatch_e.txt"))


def _random_fastq(rng, n, min_l, max_l, alphabet=b"ACGT", junk=0.0, dup=0.3):
    recs = []
    pool = []
    for i in range(n):
        if pool and rng.random() < dup:
            s = pool[int(rng.integers(len(pool)))]
            if rng.random() < 0.3:  # a proper prefix / extension of an existing read
                s = s[: max(1, int(rng.integers(1, len(s) + 1)))] if rng.random() < 0.5 else s + helpers.random_dna(rng, 3, alphabet)
        else:
            s = helpers.random_dna(rng, int(rng.integers(min_l, max_l + 1)), alphabet)
        if junk > 0:
            b = bytearray(s)
            for j in range(len(b)):
                if rng.random() < junk:
                    b[j] = int(rng.choice(list(b"NnacgtX-")))
            s = bytes(b)
        pool.append(s)
        recs.append((b"@r%d_%d" % (int(rng.integers(1000)), i), s))
    return b"".join(n + b"\n" + s + b"\n+\n" + b"!" * len(s) + b"\n" for n, s in recs)


@pytest.mark.parametrize("seed,n,min_l,max_l,mrl,minrl,alphabet,junk", [
    (1, 3000, 20, 60, 50, 30, b"ACGT", 0.0),      # truncation creates new duplicates, short reads are skipped
    (2, 5000, 1, 12, 12, 0, b"AC", 0.0),           # tiny alphabet: many duplicates and prefixes
    (3, 4000, 30, 80, 100, 0, b"ACGT", 0.03),      # junk bytes -> X, lower case -> X
    (4, 2000, 90, 151, 151, 100, b"ACGT", 0.001),  # odd MaxReadLength (last plane half used)
    (5, 1, 10, 10, 40, 0, b"ACGT", 0.0),           # a single read
])
def test_device_prep_equals_host_mirror_and_oracle(seed, n, min_l, max_l, mrl, minrl, alphabet, junk, tmp_path, oracle_bin):
    rng = np.random.default_rng(seed)
    fq = _random_fastq(rng, n, min_l, max_l, alphabet, junk)
    cfg = Config(Windows=[0], WindowWidth=min(8, mrl), MaxReadLength=mrl, MinReadLength=minrl).apply_defaults()
    want = formats.prep_reads_uniqify(fq, cfg.MinReadLength, cfg.MaxReadLength)
    with HotPath(cfg, device=0, keep_ascii=True) as hp:
        got = device_prep(hp, fq, cfg)
        assert got == want
        assert hp.unique_reads() == want[0]
        # the read set installed by msc_prep_reads is the one msc_set_reads would install
        keys_prep = hp.dump_keys()
        hp.set_reads(want[0])
        keys_set = hp.dump_keys()
        assert np.array_equal(keys_prep, keys_set)
    # and the oracle's own prepReads agrees (reads_sorted.txt of a pipeline run on a dummy target)
    work = str(tmp_path)
    fqp = os.path.join(work, "reads.fastq")
    with open(fqp, "wb") as f:
        f.write(fq)
    if len(want[0]):
        gs, gi = os.path.join(work, "g_seq.txt"), os.path.join(work, "g_ids.txt")
        with open(gs, "wb") as f:
            f.write(b"ACGTACGTACGTACGTACGTACGT\n")
        with open(gi, "wb") as f:
            f.write(b"%011d\tg0\t24\n" % 0)
        cfgd = dict(Windows=[0], WindowWidth=cfg.WindowWidth, MaxReadLength=mrl, MinReadLength=minrl, PMatch=1.0,
                    MinDinuc=0, MMTol=0, BloomSize=100000, NumHash=4, MaxMatches=1000, MatchMode="best")
        out = helpers.oracle_pipeline(work, fqp, gs, gi, cfgd, sub="prep_reads")
        assert formats.load_reads_sorted(out["reads_sorted"]) == want


@pytest.mark.parametrize("n_same,n_prefix", [(0, 40), (150, 30), (70, 0)])
def test_device_prep_shared_prefixes_and_heavy_duplicates(n_same, n_prefix):
    """Runs of equal 16-base prefix: short runs are ordered by the tie-fix kernel, a sequence present
    in more than 64 copies makes msc_prep_reads fall back to the sort over all planes."""
    rng = np.random.default_rng(100 + n_same + n_prefix)
    pre = helpers.random_dna(rng, 20, b"ACGT")
    recs = [(b"@a%d" % i, helpers.random_dna(rng, 60, b"ACGT")) for i in range(500)]
    recs += [(b"@p%d" % i, pre + helpers.random_dna(rng, int(rng.integers(0, 45)), b"ACGT")) for i in range(n_prefix)]
    same = pre + helpers.random_dna(rng, 40, b"ACGT")
    recs += [(b"@s%d" % (i % 7), same) for i in range(n_same)]
    order = rng.permutation(len(recs))
    fq = b"".join(recs[i][0] + b"\n" + recs[i][1] + b"\n+\n" + b"!" * len(recs[i][1]) + b"\n" for i in order)
    cfg = Config(Windows=[0], WindowWidth=12, MaxReadLength=64, MinReadLength=0).apply_defaults()
    want = formats.prep_reads_uniqify(fq, cfg.MinReadLength, cfg.MaxReadLength)
    with HotPath(cfg, device=0, keep_ascii=True) as hp:
        assert device_prep(hp, fq, cfg) == want


def test_device_prep_empty_and_all_skipped():
    cfg = Config(Windows=[0], WindowWidth=4, MaxReadLength=20, MinReadLength=10).apply_defaults()
    with HotPath(cfg, device=0, keep_ascii=True) as hp:
        assert hp.prep_reads([], cfg.MinReadLength) == (0, 0)
        perm, gs = hp.read_groups()
        assert len(perm) == 0 and list(gs) == [0]
        assert hp.prep_reads([b"ACGT", b"AC", b"ACGTACG"], cfg.MinReadLength) == (0, 0)  # all shorter than MinReadLength
        hp.set_targets([b"ACGTACGTACGTACGTACGTACGT"])
        hp.run()
        assert len(hp.fetch()) == 0


def test_deferred_stages_equal_fused_run():
    """MSC_STAGE_DEFER (stream-ordered multi-GPU step): screen + confirm enqueued without a host
    synchronisation, combine completes all three -- same matches as msc_run; a pending deferred
    run only accepts the combine stage."""
    from muscato_b200.engine import MuscatoError
    rng = np.random.default_rng(3)
    genes = [helpers.random_dna(rng, 300, b"ACGT") for _ in range(20)]
    reads = sorted({g[i:i + 60] for g in genes for i in range(0, 240, 7)})
    cfg = Config(Windows=[0, 20], WindowWidth=12, MaxReadLength=60, PMatch=0.95, MinDinuc=0, MMTol=1).apply_defaults()
    with HotPath(cfg, device=0, keep_ascii=True) as hp:
        hp.set_reads(reads)
        hp.set_targets(genes)
        hp.run()
        want = hp.fetch()
        for _ in range(2):
            hp.run_stages(3, 1 | 2 | 8)
            assert hp.best_device() is not None
            with pytest.raises(MuscatoError):
                hp.run_stages(0, 1)          # only combine may follow
            hp.run_stages(3, 1 | 2 | 8)      # the failed call cleared nothing it should not: start over
            hp.run_stages(0, 4)
            got = hp.fetch()
            assert np.array_equal(got, want)
