"""Parity tests proper: the CUDA hot path (through the C ABI) against the CPU oracle and the
reference's golden fixtures.  Integer/byte work: the bar is bit-exact."""
import json
import os
import zlib

import numpy as np
import pytest

from muscato_b200 import formats, gendat
from muscato_b200.config import Config
from tests import helpers

pytestmark = pytest.mark.gpu


def _engine(cfg: Config, **kw):
    from muscato_b200.engine import HotPath
    return HotPath(cfg, device=0, **kw)


def run_cuda(cfg: Config, reads, targets, taps=False):
    with _engine(cfg) as hp:
        hp.set_reads(reads)
        hp.set_targets(targets)
        hp.screen()
        keys = cands = None
        if taps:
            keys = hp.dump_keys()
            cands = hp.dump_candidates()
        hp.confirm()
        hp.combine()
        m = hp.fetch()
        nm = hp.nonmatch_ids()
        st = hp.stats()
    assert_fetch_order(m)
    # device-side list of unmatched reads (msc_fetch_nonmatch) == complement of the matched read ids
    n_reads = len(reads[1]) - 1 if isinstance(reads, tuple) else len(reads)
    want_nm = np.setdiff1d(np.arange(n_reads, dtype=np.uint32), np.unique(m["read_id"]).astype(np.uint32))
    assert np.array_equal(nm, want_nm)
    return m, st, keys, cands


def assert_fetch_order(m):
    """msc_fetch_matches promises (read_id, gene_id, pos) order, established on the device."""
    if len(m) > 1:
        key = np.stack([m["read_id"], m["gene_id"], m["pos"]], axis=1).astype(np.int64)
        order = np.lexsort((key[:, 2], key[:, 1], key[:, 0]))
        assert np.array_equal(order, np.arange(len(m)))


def check_against_oracle(tmp_path, raw_reads, names, genes, cfgd, gene_names=None, taps=True):
    """Full comparison of one case: window keys, candidates, matches.txt, results.txt, non-match fastq."""
    work = str(tmp_path)
    fq, gs, gi = helpers.write_case(work, raw_reads, names, genes, gene_names)
    out = helpers.oracle_pipeline(work, fq, gs, gi, cfgd)
    cfg = Config(**{k: v for k, v in cfgd.items() if k in Config.__dataclass_fields__}).apply_defaults()
    seqs, counts, rnames = formats.prep_reads_uniqify(helpers.read_bytes(fq), cfg.MinReadLength, cfg.MaxReadLength)
    # host mirror of prepReads must agree with the oracle's reads_sorted
    o_seqs, o_counts, o_names = formats.load_reads_sorted(out["reads_sorted"])
    assert (seqs, counts, rnames) == (o_seqs, o_counts, o_names)
    targets = formats.load_targets(gs)
    gnames, glens = formats.load_gene_ids(gi)
    m, st, keys, cands = run_cuda(cfg, seqs, targets, taps=taps)
    if taps:
        want_keys = set()
        want_cands = set()
        for k in range(len(cfg.Windows)):
            want_keys |= helpers.oracle_window_keys(out["tmp"], seqs, k)
            want_cands |= helpers.oracle_candidates(out["tmp"], k)
        got_keys = {(int(r["window"]), int(r["read_id"])) for r in keys}
        assert got_keys == want_keys
        got_cands = {(int(r["window"]), int(r["gene_id"]), int(r["p"])) for r in cands}
        assert got_cands == want_cands
    assert formats.matches_lines(m, seqs, targets) == helpers.read_lines(out["matches"])
    res = b"".join(ln + b"\n" for ln in formats.results_lines(m, seqs, counts, rnames, targets, gnames, glens))
    assert res == helpers.read_bytes(out["results"])
    assert formats.nonmatch_fastq(m, seqs, counts, rnames) == helpers.read_bytes(out["nonmatch"])
    nm_ids = np.setdiff1d(np.arange(len(seqs)), np.unique(m["read_id"]))
    assert formats.nonmatch_fastq_from_ids(nm_ids, seqs, counts, rnames) == helpers.read_bytes(out["nonmatch"])
    return m, st


@pytest.mark.parametrize("case", ["00", "01", "02", "03", "04"])
def test_golden_fixture_through_cuda(case, tmp_path, oracle_bin):
    """tests/tests.toml "muscato 0..4" of the reference: result.txt and the non-match fastq, byte for byte."""
    src = os.path.join(helpers.GOLDEN, "muscato", case)
    cfgd = json.load(open(os.path.join(src, "config.json")))
    seq, ids = str(tmp_path / "genes_seq.txt"), str(tmp_path / "genes_ids.txt")
    helpers.oracle_prep_targets(os.path.join(src, "genes.txt"), seq, ids, rev=(case == "04"))
    cfg = Config(**{k: v for k, v in cfgd.items() if k in Config.__dataclass_fields__}).apply_defaults()
    seqs, counts, rnames = formats.prep_reads_uniqify(helpers.read_bytes(os.path.join(src, "reads.fastq")),
                                                      cfg.MinReadLength, cfg.MaxReadLength)
    targets = formats.load_targets(seq)
    gnames, glens = formats.load_gene_ids(ids)
    m, _, _, _ = run_cuda(cfg, seqs, targets)
    res = b"".join(ln + b"\n" for ln in formats.results_lines(m, seqs, counts, rnames, targets, gnames, glens))
    assert res == helpers.read_bytes(os.path.join(src, "result_e.txt"))
    assert formats.nonmatch_fastq(m, seqs, counts, rnames) == helpers.read_bytes(os.path.join(src, "result.nonmatch_e.txt"))


def _planted_case(rng, n_genes, gene_len, n_reads, read_lens, sub_rate, alphabet=b"ACGT", x_rate=0.0,
                  dup_frac=0.1, edge_frac=0.2):
    """Random genes; reads sampled from them (with substitutions), some from position 0 and the
    target end, some duplicated, some pure noise."""
    genes = [helpers.random_dna(rng, int(gl), alphabet) for gl in (gene_len if hasattr(gene_len, "__len__") else [gene_len] * n_genes)]
    if x_rate > 0:
        gs = []
        for g in genes:
            a = np.frombuffer(g, dtype=np.uint8).copy()
            a[rng.random(len(a)) < x_rate] = ord("N")  # prep_targets would turn this into X
            gs.append(bytes(a).replace(b"N", b"X"))
        genes = gs
    reads = []
    for i in range(n_reads):
        L = int(rng.choice(read_lens))
        u = rng.random()
        if u < 0.15:
            reads.append(helpers.random_dna(rng, L, alphabet))
            continue
        g = genes[int(rng.integers(0, len(genes)))]
        if len(g) < L:
            reads.append(helpers.random_dna(rng, L, alphabet))
            continue
        v = rng.random()
        if v < edge_frac / 2:
            p = 0
        elif v < edge_frac:
            p = len(g) - L
        else:
            p = int(rng.integers(0, len(g) - L + 1))
        a = np.frombuffer(g[p:p + L], dtype=np.uint8).copy()
        mut = rng.random(L) < sub_rate
        a[mut] = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=int(mut.sum()))]
        if x_rate > 0:
            a[rng.random(L) < x_rate] = ord("N")
        reads.append(bytes(a))
    ndup = int(dup_frac * n_reads)
    for _ in range(ndup):
        reads.append(reads[int(rng.integers(0, len(reads)))])
    return reads, genes


CASES = [
    # name, W, Windows, MRL, PMatch, MinDinuc, MMTol, read_lens, gene_len, sub_rate, x_rate
    ("w4_exact", 4, [0, 5], 300, 1.0, 1, 1, [10, 12], 40, 0.0, 0.0),
    ("w5_mm", 5, [0, 7, 14], 30, 0.9, 0, 0, [20, 24, 30], 80, 0.05, 0.0),
    ("w8_mm_tol", 8, [0, 10, 20], 120, 0.93, 3, 2, [40, 60, 110], 300, 0.04, 0.0),
    ("w15_q1", 15, [0, 20, 40], 100, 0.97, 5, 1, [100, 90, 85, 86], 400, 0.02, 0.0),   # L > 100-W at target position 0
    ("w15_x", 15, [0, 20], 100, 0.95, 2, 1, [100, 64, 33], 300, 0.02, 0.02),           # X in reads and targets
    ("w32_long", 32, [0, 40, 100], 200, 0.96, 4, 3, [200, 150, 133], 700, 0.02, 0.0),
    ("w20_unsorted_windows", 20, [30, 0, 10, 10], 150, 0.98, 0, 0, [150, 64, 45], 500, 0.01, 0.0),
    ("w7_lowcomplex", 7, [0, 7, 3], 64, 0.9, 0, 5, [64, 32, 21], 150, 0.03, 0.0),
    # wide windows (two key words, hashed fingerprints, window re-checked in confirm)
    ("w33_wide", 33, [0, 20, 50], 120, 0.95, 4, 2, [120, 90, 60], 400, 0.015, 0.0),
    ("w40_wide_x", 40, [0, 45], 100, 0.94, 3, 1, [100, 88, 50], 350, 0.01, 0.01),
    ("w50_wide_max", 50, [0, 25, 60], 150, 0.96, 0, 1, [150, 111, 64, 50], 500, 0.01, 0.0),
    ("w36_wide_lowcomplex", 36, [0, 10], 80, 0.9, 2, 3, [80, 46, 40], 200, 0.02, 0.0),
]


# every case with the front the library picks; the W <= 15 ones also with the Bloom front forced and with the exact
# front forced in 16 MB slices (W = 15: 8 passes of the scan over a 128 MB bitmap) -- the front never changes results
FRONT_CASES = [(c, "auto") for c in CASES] + [(c, f) for c in CASES if c[1] <= 15 for f in ("bloom", "direct")]


@pytest.mark.parametrize("case,front", FRONT_CASES, ids=[f"{c[0]}-{f}" for c, f in FRONT_CASES])
def test_random_cases_vs_oracle(case, front, tmp_path, oracle_bin, monkeypatch):
    name, W, wins, mrl, pm, mind, mmtol, rlens, glen, sub, xr = case
    if front == "bloom":
        monkeypatch.setenv("MSC_FRONT_DIRECT", "0")
    elif front == "direct":
        monkeypatch.setenv("MSC_FRONT_DIRECT", "1")
        monkeypatch.setenv("MSC_FRONT_PASS_MB", "16")
    rng = np.random.default_rng(zlib.crc32(name.encode()))
    alphabet = b"AC" if name == "w7_lowcomplex" else b"ACG" if "lowcomplex" in name else b"ACGT"
    n_genes = 30
    glens = [glen] * (n_genes - 4) + [W - 1, W, W + 3, max(W + 1, 20)]  # targets shorter than / equal to W
    reads, genes = _planted_case(rng, n_genes, glens, 300, rlens, sub, alphabet=alphabet, x_rate=xr)
    cfgd = dict(Windows=wins, WindowWidth=W, MaxReadLength=mrl, PMatch=pm, MinDinuc=mind, MMTol=mmtol,
                BloomSize=4000000, NumHash=8, MinReadLength=0, MaxMatches=1000000, MaxConfirmProcs=3, MatchMode="best")
    m, st = check_against_oracle(tmp_path, reads, None, genes, cfgd)
    if sub == 0.0 or pm < 1:
        assert len(m) > 0


def test_truncated_and_short_reads(tmp_path, oracle_bin):
    """MaxReadLength truncation / MinReadLength skipping happen upstream (prep_reads) -- the host mirror
    must feed the device exactly what the oracle's reads_sorted holds."""
    rng = np.random.default_rng(11)
    reads, genes = _planted_case(rng, 20, 200, 200, [30, 50, 80, 120], 0.02)
    cfgd = dict(Windows=[0, 12, 25], WindowWidth=10, MaxReadLength=60, MinReadLength=40, PMatch=0.95, MinDinuc=2,
                MMTol=1, BloomSize=2000000, NumHash=6, MaxMatches=1000000, MatchMode="best")
    check_against_oracle(tmp_path, reads, None, genes, cfgd)


def test_names_with_spaces_and_duplicates(tmp_path, oracle_bin):
    rng = np.random.default_rng(5)
    reads, genes = _planted_case(rng, 10, 120, 60, [40], 0.01, dup_frac=0.5)
    names = [b"@SRR%d.%d some comment length=%d" % (i % 7, i, len(r)) for i, r in enumerate(reads)]
    gnames = [b">chr%d description here" % i for i in range(len(genes))]
    cfgd = dict(Windows=[0, 15], WindowWidth=12, MaxReadLength=40, PMatch=0.95, MinDinuc=0, MMTol=0,
                BloomSize=1000000, NumHash=6, MaxMatches=1000000, MatchMode="best")
    check_against_oracle(tmp_path, reads, names, genes, cfgd, gene_names=gnames)


def test_gendat_medium_vs_oracle(tmp_path, oracle_bin):
    """muscato_gendat-shaped data with the README flags (Windows=0,20 WindowWidth=15 MaxReadLength=100)."""
    syn = gendat.generate(20000, 100, 400, 2000, seed=1, rev=True, mutated_fraction=0.5)
    work = str(tmp_path)
    fq, gs, gi = os.path.join(work, "reads.fastq"), os.path.join(work, "genes_seq.txt"), os.path.join(work, "genes_ids.txt")
    gendat.write_oracle_inputs(syn, fq, gs, gi)
    cfgd = dict(Windows=[0, 20], WindowWidth=15, MaxReadLength=100, PMatch=0.97, MinDinuc=5, MMTol=1,
                BloomSize=40000000, NumHash=20, MaxMatches=1000000, MatchMode="best")
    out = helpers.oracle_pipeline(work, fq, gs, gi, cfgd)
    cfg = Config(**cfgd).apply_defaults()
    seqs, counts, rnames = formats.load_reads_sorted(out["reads_sorted"])
    assert seqs == syn.reads_list()
    targets = syn.targets_list()
    with _engine(cfg) as hp:
        hp.set_reads((syn.read_ascii, syn.read_offs))
        hp.set_targets((syn.target_ascii, syn.target_offs))
        hp.run()
        m = hp.fetch()
        # the staged form used across ranks: screen+confirm, (all-reduce of best), combine
        hp.run_stages(0, 1 | 2)
        hp.run_stages(0, 4)
        m2 = hp.fetch()
    assert_fetch_order(m)
    assert np.array_equal(m, m2)
    assert len(m) > 5000
    assert max(np.bincount(m["read_id"])) > 16   # exercises the long-segment rank sort
    assert formats.matches_lines(m, seqs, targets) == helpers.read_lines(out["matches"])
    gnames, glens = formats.load_gene_ids(gi)
    res = b"".join(ln + b"\n" for ln in formats.results_lines(m, seqs, counts, rnames, targets, gnames, glens))
    assert res == helpers.read_bytes(out["results"])
    assert formats.nonmatch_fastq(m, seqs, counts, rnames) == helpers.read_bytes(out["nonmatch"])


def test_empty_inputs():
    cfg = Config(Windows=[0, 5], WindowWidth=4, MaxReadLength=50).apply_defaults()
    with _engine(cfg) as hp:
        hp.set_reads([])
        hp.set_targets([b"ACGTACGTACGT"])
        hp.run()
        assert len(hp.fetch()) == 0
    with _engine(cfg) as hp:
        hp.set_reads([b"ACGTACGTAC"])
        hp.set_targets([])
        hp.run()
        assert len(hp.fetch()) == 0
    with _engine(cfg) as hp:
        hp.set_reads([b"ACG", b""])       # shorter than every window
        hp.set_targets([b"ACGTACGTACGT", b""])
        hp.run()
        assert len(hp.fetch()) == 0


def test_input_validation_errors():
    from muscato_b200.engine import MuscatoError
    cfg = Config(Windows=[0], WindowWidth=4, MaxReadLength=8).apply_defaults()
    with _engine(cfg) as hp:
        with pytest.raises(MuscatoError):
            hp.set_reads([b"ACGTACGTACGT"])  # longer than MaxReadLength
        with pytest.raises(MuscatoError):
            hp.screen()                      # nothing set
    with pytest.raises(MuscatoError):
        _engine(Config(Windows=[0], WindowWidth=51, MaxReadLength=100).apply_defaults())


@pytest.mark.parametrize("case", ["00", "03", "04"])
def test_cli_files_drop_in(case, tmp_path, oracle_bin):
    """The stage-compatible driver: config.json + .sz inputs in, matches.txt.sz / result.txt /
    non-match fastq out, compared with the reference's expected files."""
    from muscato_b200 import cli, sz
    src = os.path.join(helpers.GOLDEN, "muscato", case)
    cfgd = json.load(open(os.path.join(src, "config.json")))
    seq, ids = str(tmp_path / "genes_seq.txt"), str(tmp_path / "genes_ids.txt")
    helpers.oracle_prep_targets(os.path.join(src, "genes.txt"), seq, ids, rev=(case == "04"))
    sz.write_file(str(tmp_path / "genes.txt.sz"), helpers.read_bytes(seq))
    sz.write_file(str(tmp_path / "genes_ids.txt.sz"), helpers.read_bytes(ids))
    out = helpers.oracle_pipeline(str(tmp_path / "oracle"), os.path.join(src, "reads.fastq"), seq, ids, cfgd)
    cfgd.update(ReadFileName=os.path.join(src, "reads.fastq"), GeneFileName=str(tmp_path / "genes.txt.sz"),
                GeneIdFileName=str(tmp_path / "genes_ids.txt.sz"), ResultsFileName=str(tmp_path / "result.txt"),
                TempDir=str(tmp_path / "tmp"))
    cpath = str(tmp_path / "config.json")
    json.dump(cfgd, open(cpath, "w"))
    cli.run(cpath, device=0, from_fastq=True)
    assert helpers.read_bytes(str(tmp_path / "result.txt")) == helpers.read_bytes(os.path.join(src, "result_e.txt"))
    assert helpers.read_bytes(str(tmp_path / "result.nonmatch.txt.fastq")) == \
        helpers.read_bytes(os.path.join(src, "result.nonmatch_e.txt"))
    assert sz.read_file(str(tmp_path / "tmp" / "matches.txt.sz")) == helpers.read_bytes(out["matches"])
    assert sz.read_file(str(tmp_path / "tmp" / "reads_sorted.txt.sz")) == helpers.read_bytes(out["reads_sorted"])


MM_CASES = [("best", 1), ("best", 3), ("best", 7), ("best", 40), ("first", 1), ("first", 5), ("first", 40)]


@pytest.mark.parametrize("front", ["auto", "direct"])
@pytest.mark.parametrize("mode,mm", MM_CASES, ids=[f"{m}{k}" for m, k in MM_CASES])
def test_maxmatches_truncation_vs_oracle(mode, mm, front, tmp_path, oracle_bin, monkeypatch):
    """Q7: per (window, k-mer) group the reference keeps at most MaxMatches pairs through an
    order-dependent heap ("best", cmd/muscato_confirm/main.go:424-448) or the first MaxMatches+1
    ("first", :233-238).  Repetitive targets + many near-identical reads overflow small limits.
    Both fronts of the scan (the exact one writes its candidates in another order and with empty slots)."""
    if front == "direct":
        monkeypatch.setenv("MSC_FRONT_DIRECT", "1")
    rng = np.random.default_rng(100 + mm)
    unit = helpers.random_dna(rng, 37)
    genes = []
    for gi in range(12):
        a = np.frombuffer(unit * 6, dtype=np.uint8).copy()       # tandem repeats: the same k-mers everywhere
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
    m, st = check_against_oracle(tmp_path, reads, None, genes, cfgd, taps=False)
    assert st["n_overflow_groups"] > 0, "the case must actually truncate"


@pytest.mark.parametrize("case", ["00", "01", "02", "03", "04"])
def test_cpp_executable_drop_in(case, tmp_path, oracle_bin):
    """muscato_b200_hotpath (C++ host + C ABI): same config.json and .sz files as the reference
    stages, outputs compared with the reference's expected files and the oracle's matches.txt."""
    import subprocess
    from muscato_b200 import build, sz
    src = os.path.join(helpers.GOLDEN, "muscato", case)
    cfgd = json.load(open(os.path.join(src, "config.json")))
    seq, ids = str(tmp_path / "genes_seq.txt"), str(tmp_path / "genes_ids.txt")
    helpers.oracle_prep_targets(os.path.join(src, "genes.txt"), seq, ids, rev=(case == "04"))
    sz.write_file(str(tmp_path / "genes.txt.sz"), helpers.read_bytes(seq))
    sz.write_file(str(tmp_path / "genes_ids.txt.sz"), helpers.read_bytes(ids))
    out = helpers.oracle_pipeline(str(tmp_path / "oracle"), os.path.join(src, "reads.fastq"), seq, ids, cfgd)
    os.makedirs(tmp_path / "tmp")
    cfgd.update(ReadFileName=os.path.join(src, "reads.fastq"), GeneFileName=str(tmp_path / "genes.txt.sz"),
                GeneIdFileName=str(tmp_path / "genes_ids.txt.sz"), ResultsFileName=str(tmp_path / "result.txt"),
                TempDir=str(tmp_path / "tmp"))
    json.dump(cfgd, open(tmp_path / "config.json", "w"))
    r = subprocess.run([build.EXE_PATH, str(tmp_path / "config.json"), "--from-fastq"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert helpers.read_bytes(str(tmp_path / "result.txt")) == helpers.read_bytes(os.path.join(src, "result_e.txt"))
    assert helpers.read_bytes(str(tmp_path / "result.nonmatch.txt.fastq")) == \
        helpers.read_bytes(os.path.join(src, "result.nonmatch_e.txt"))
    assert sz.read_file(str(tmp_path / "tmp" / "matches.txt.sz")) == helpers.read_bytes(out["matches"])
    assert sz.read_file(str(tmp_path / "tmp" / "reads_sorted.txt.sz")) == helpers.read_bytes(out["reads_sorted"])
    # second run from the reads_sorted.txt.sz it left behind (the driver's normal hand-over)
    os.remove(tmp_path / "result.txt")
    r = subprocess.run([build.EXE_PATH, str(tmp_path / "config.json")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert helpers.read_bytes(str(tmp_path / "result.txt")) == helpers.read_bytes(os.path.join(src, "result_e.txt"))


def test_cpp_executable_random_case_vs_oracle(tmp_path, oracle_bin):
    import subprocess
    from muscato_b200 import build, sz
    rng = np.random.default_rng(77)
    reads, genes = _planted_case(rng, 25, 260, 400, [70, 64, 50], 0.03, x_rate=0.01)
    names = [b"@r%d desc" % i for i in range(len(reads))]
    cfgd = dict(Windows=[0, 16, 33], WindowWidth=14, MaxReadLength=70, PMatch=0.94, MinDinuc=3, MMTol=1,
                BloomSize=2000000, NumHash=6, MaxMatches=1000000, MatchMode="best", MinReadLength=0)
    fq, gs, gi = helpers.write_case(str(tmp_path), reads, names, genes)
    out = helpers.oracle_pipeline(str(tmp_path / "oracle"), fq, gs, gi, cfgd)
    os.makedirs(tmp_path / "tmp")
    c = dict(cfgd)
    c.update(ReadFileName=fq, GeneFileName=gs, GeneIdFileName=gi, ResultsFileName=str(tmp_path / "res.txt"),
             TempDir=str(tmp_path / "tmp"))
    json.dump(c, open(tmp_path / "config.json", "w"))
    r = subprocess.run([build.EXE_PATH, str(tmp_path / "config.json"), "--from-fastq"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert helpers.read_bytes(str(tmp_path / "res.txt")) == helpers.read_bytes(out["results"])
    assert helpers.read_bytes(str(tmp_path / "res.nonmatch.txt.fastq")) == helpers.read_bytes(out["nonmatch"])
    assert sz.read_file(str(tmp_path / "tmp" / "matches.txt.sz")) == helpers.read_bytes(out["matches"])


def test_set_reads_device_equals_host_upload():
    """msc_set_reads_device (reads already in HBM, e.g. after an NCCL broadcast) must give the same
    matches as msc_set_reads, and must reject inconsistent offsets found on the device."""
    import torch
    from muscato_b200.engine import MuscatoError
    syn = gendat.generate(5000, 80, 60, 900, seed=5, rev=True, mutated_fraction=0.6)
    cfg = Config(Windows=[0, 25, 50], WindowWidth=16, MaxReadLength=80, PMatch=0.95, MinDinuc=4, MMTol=1).apply_defaults()
    with _engine(cfg) as hp:
        hp.set_reads((syn.read_ascii, syn.read_offs))
        hp.set_targets((syn.target_ascii, syn.target_offs))
        hp.run()
        want = hp.fetch()
    d_a = torch.from_numpy(syn.read_ascii.copy()).cuda()
    d_o = torch.from_numpy(syn.read_offs.view(np.int64).copy()).cuda()
    torch.cuda.synchronize()
    with _engine(cfg) as hp:
        hp.set_reads_device(d_a.data_ptr(), d_o.data_ptr(), syn.n_reads, int(d_a.numel()))
        hp.set_targets((syn.target_ascii, syn.target_offs))
        hp.run()
        got = hp.fetch()
        assert len(want) > 1000 and np.array_equal(want, got)
        bad = d_o.clone()
        bad[10] = bad[12]          # not monotone
        torch.cuda.synchronize()
        with pytest.raises(MuscatoError):
            hp.set_reads_device(d_a.data_ptr(), bad.data_ptr(), syn.n_reads, int(d_a.numel()))
        with pytest.raises(MuscatoError):
            hp.set_reads_device(d_a.data_ptr(), d_o.data_ptr(), syn.n_reads, int(d_a.numel()) - 1)


def _repeat_case(rng, n_genes, n_reads, read_len):
    """BASELINE config[4]: low-entropy, repetitive targets (tandem repeats of period 1-6 and
    low-complexity blocks interleaved with random sequence) with reads sampled from them."""
    genes = []
    pool = [helpers.random_dna(rng, int(p)) for p in (1, 2, 3, 4, 5, 6)]   # shared repeat units
    for g in range(n_genes):
        parts = []
        while sum(map(len, parts)) < 500:
            u = rng.random()
            if u < 0.5:
                unit = pool[int(rng.integers(0, len(pool)))]
                parts.append(unit * int(rng.integers(10, 40)))
            elif u < 0.7:
                parts.append(helpers.random_dna(rng, int(rng.integers(20, 60)), b"AT"))
            else:
                parts.append(helpers.random_dna(rng, int(rng.integers(20, 80))))
        genes.append(b"".join(parts))
    reads = []
    for _ in range(n_reads):
        g = genes[int(rng.integers(0, n_genes))]
        p = int(rng.integers(0, len(g) - read_len))
        a = np.frombuffer(g[p:p + read_len], dtype=np.uint8).copy()
        m = rng.random(read_len) < 0.02
        a[m] = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, int(m.sum()))]
        reads.append(bytes(a))
    return reads, genes


@pytest.mark.parametrize("front", ["auto", "direct"])
@pytest.mark.parametrize("mm,mode,mindinuc", [(1000000, "best", 1), (1000, "best", 2), (300, "first", 3)])
def test_high_multiplicity_stress_vs_oracle(mm, mode, mindinuc, front, tmp_path, oracle_bin, monkeypatch):
    """Hit explosion: thousands of candidates x dozens of reads per k-mer group, 5 mismatches at 100 bp
    (PMatch=0.95), default and small MaxMatches (the latter forces the order-dependent truncation)."""
    if front == "direct":
        monkeypatch.setenv("MSC_FRONT_DIRECT", "1")
    rng = np.random.default_rng(1234 + mm % 97)
    reads, genes = _repeat_case(rng, 30, 400, 100)
    cfgd = dict(Windows=[0, 30, 60], WindowWidth=12, MaxReadLength=100, PMatch=0.95, MinDinuc=mindinuc, MMTol=2,
                BloomSize=4000000, NumHash=8, MaxMatches=mm, MatchMode=mode, MaxConfirmProcs=3)
    m, st = check_against_oracle(tmp_path, reads, None, genes, cfgd, taps=(mm == 1000000))
    assert st["n_pairs"] > 10 * st["n_candidates"] > 0          # quadratic blow-up inside the groups
    assert len(m) > 2000
    if mm < 1000000:
        assert st["n_overflow_groups"] > 0
