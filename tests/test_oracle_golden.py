"""The oracle is pinned against every golden vector the reference's own tests hold for this
path (tests/tests.toml:1-139 of the reference; fixtures copied as data under tests/golden/)."""
import gzip
import json
import os

import pytest

from muscato_b200 import sz
from tests import helpers

MUSCATO_CASES = ["00", "01", "02", "03", "04"]


@pytest.mark.parametrize("case", MUSCATO_CASES)
def test_oracle_reproduces_pipeline_fixture(case, tmp_path, oracle_bin):
    src = os.path.join(helpers.GOLDEN, "muscato", case)
    cfg = json.load(open(os.path.join(src, "config.json")))
    seq = str(tmp_path / "genes_seq.txt")
    ids = str(tmp_path / "genes_ids.txt")
    # tests.toml: "muscato 4 prep" runs muscato_prep_targets -rev; the others without.
    helpers.oracle_prep_targets(os.path.join(src, "genes.txt"), seq, ids, rev=(case == "04"))
    out = helpers.oracle_pipeline(str(tmp_path), os.path.join(src, "reads.fastq"), seq, ids, cfg)
    assert helpers.read_bytes(out["results"]) == helpers.read_bytes(os.path.join(src, "result_e.txt"))
    assert helpers.read_bytes(out["nonmatch"]) == helpers.read_bytes(os.path.join(src, "result.nonmatch_e.txt"))


# (dir, input, rev) from tests.toml:1-68
PREP_CASES = [
    ("00", "genes.fasta", False), ("01", "genes.fasta", True),
    ("02", "genes.txt", False), ("03", "genes.txt", True),
    ("04", "genes.txt.gz", False), ("05", "genes.txt.gz", True),
    ("06", "genes.txt.sz", True), ("07", "genes.txt.sz", True),
]


@pytest.mark.parametrize("case,fname,rev", PREP_CASES)
def test_oracle_reproduces_prep_targets_fixture(case, fname, rev, tmp_path, oracle_bin):
    src = os.path.join(helpers.GOLDEN, "prep_targets", case, fname)
    plain = src
    if fname.endswith(".gz"):
        plain = str(tmp_path / fname[:-3])
        open(plain, "wb").write(gzip.open(src, "rb").read())
    elif fname.endswith(".sz"):
        plain = str(tmp_path / fname[:-3])
        open(plain, "wb").write(sz.read_file(src))
    seq, ids = str(tmp_path / "seq.txt"), str(tmp_path / "ids.txt")
    helpers.oracle_prep_targets(plain, seq, ids, rev=rev)
    d = os.path.join(helpers.GOLDEN, "prep_targets", case)
    assert helpers.read_bytes(seq) == helpers.read_bytes(os.path.join(d, "expected_sequences.txt"))
    assert helpers.read_bytes(ids) == helpers.read_bytes(os.path.join(d, "expected_ids.txt"))
