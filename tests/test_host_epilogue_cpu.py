"""Host half of the drop-in on the CPU: given the integer matches the C ABI returns -- here taken
from the oracle's matches.txt instead of a GPU -- the Python mirror of the file contract
(muscato_b200/formats.py) must write the reference's own expected files: matches.txt lines
(cmd/muscato_confirm/main.go:221-230), results.txt (join + bytewise sort of
cmd/muscato/main.go:507-676, decimal strings compare as text, Q11) and the non-match fastq
(cmd/muscato_nonmatch/main.go:57-113).  Golden files: the reference's tests/data/muscato/00..04."""
import json
import os

import numpy as np
import pytest

from muscato_b200 import formats
from muscato_b200.config import Config
from tests import helpers

MATCH_DTYPE = np.dtype([("read_id", "<u4"), ("gene_id", "<u4"), ("pos", "<u4"), ("nx", "<u4")])


def _matches_from_oracle(path, seqs):
    idx = {s: i for i, s in enumerate(seqs)}
    rows = []
    for ln in helpers.read_lines(path):
        f = ln.split(b"\t")
        rows.append((idx[f[0]], int(f[4]), int(f[2]), int(f[3])))
    rows.sort()
    m = np.zeros(len(rows), dtype=MATCH_DTYPE)
    for i, r in enumerate(rows):
        m[i] = r
    return m


@pytest.mark.parametrize("case", ["00", "01", "02", "03", "04"])
def test_python_epilogue_writes_the_reference_files(case, tmp_path, oracle_bin):
    src = os.path.join(helpers.GOLDEN, "muscato", case)
    cfgd = json.load(open(os.path.join(src, "config.json")))
    seq, ids = str(tmp_path / "genes_seq.txt"), str(tmp_path / "genes_ids.txt")
    helpers.oracle_prep_targets(os.path.join(src, "genes.txt"), seq, ids, rev=(case == "04"))
    out = helpers.oracle_pipeline(str(tmp_path), os.path.join(src, "reads.fastq"), seq, ids, cfgd)
    cfg = Config(**{k: v for k, v in cfgd.items() if k in Config.__dataclass_fields__}).apply_defaults()
    # read side: the host mirror of prep_reads | sort | uniqify, from the fastq itself
    seqs, counts, names = formats.prep_reads_uniqify(helpers.read_bytes(os.path.join(src, "reads.fastq")),
                                                     cfg.MinReadLength, cfg.MaxReadLength)
    targets = formats.load_targets(seq)
    gnames, glens = formats.load_gene_ids(ids)
    m = _matches_from_oracle(out["matches"], seqs)
    assert formats.matches_lines(m, seqs, targets) == helpers.read_lines(out["matches"])
    res = b"".join(ln + b"\n" for ln in formats.results_lines(m, seqs, counts, names, targets, gnames, glens))
    assert res == helpers.read_bytes(os.path.join(src, "result_e.txt"))
    assert formats.nonmatch_fastq(m, seqs, counts, names) == helpers.read_bytes(os.path.join(src, "result.nonmatch_e.txt"))
    ids_nm = np.setdiff1d(np.arange(len(seqs)), np.unique(m["read_id"]))
    assert formats.nonmatch_fastq_from_ids(ids_nm, seqs, counts, names) == \
        helpers.read_bytes(os.path.join(src, "result.nonmatch_e.txt"))


def test_python_epilogue_on_a_randomised_case(tmp_path, oracle_bin):
    """Positions >= 10 next to positions < 10 (text order, Q11), duplicate reads (count / names join),
    names with spaces, multi-mapping."""
    rng = np.random.default_rng(11)
    genes = [helpers.random_dna(rng, 180) for _ in range(6)]
    genes.append(genes[0][:90] + genes[1][:90])
    reads, names = [], []
    for i in range(120):
        g = genes[int(rng.integers(0, len(genes)))]
        p = int(rng.integers(0, len(g) - 40))
        reads.append(g[p:p + 40])
        names.append(b"@r%d extra words" % i)
    reads += reads[:15]                                    # duplicates -> count > 1, several names
    names += [b"@dup%d" % i for i in range(15)]
    cfgd = dict(Windows=[0, 12], WindowWidth=10, MaxReadLength=40, PMatch=0.95, MinDinuc=0, MMTol=1,
                BloomSize=1000000, NumHash=6, MaxMatches=1000000, MatchMode="best")
    fq, gs, gi = helpers.write_case(str(tmp_path), reads, names, genes)
    out = helpers.oracle_pipeline(str(tmp_path), fq, gs, gi, cfgd)
    cfg = Config(**{k: v for k, v in cfgd.items() if k in Config.__dataclass_fields__}).apply_defaults()
    seqs, counts, rnames = formats.prep_reads_uniqify(helpers.read_bytes(fq), cfg.MinReadLength, cfg.MaxReadLength)
    targets = formats.load_targets(gs)
    gnames, glens = formats.load_gene_ids(gi)
    m = _matches_from_oracle(out["matches"], seqs)
    assert len(m) > 100 and (m["pos"] >= 10).any() and (m["pos"] < 10).any()
    res = b"".join(ln + b"\n" for ln in formats.results_lines(m, seqs, counts, rnames, targets, gnames, glens))
    assert res == helpers.read_bytes(out["results"])
    assert formats.nonmatch_fastq(m, seqs, counts, rnames) == helpers.read_bytes(out["nonmatch"])
