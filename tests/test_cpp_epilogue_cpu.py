"""The C++ host half of the stage executable on the CPU (muscato_b200/csrc/host/hostpath.hpp):
given the integer matches the C ABI returns -- here taken from the oracle's matches.txt and handed
over with `--format-matches`, no GPU involved -- it must write the reference's own files: matches.txt.sz
(cmd/muscato_confirm/main.go:221-230, whole-line order of `sort -u`), results.txt (join + bytewise
sort of cmd/muscato/main.go:507-676; pos and nx compare as decimal strings, Q11) and the non-match
fastq (cmd/muscato_nonmatch/main.go:57-113).  The writers order read groups instead of sorting text
lines (reads_sorted order == line order on the first field); a read file that breaks the sorted /
unique contract makes them fall back to a whole-line sort.  Also: the muscato_combine_windows stage
name (the MMTol rule as a stdin -> stdout filter, cmd/muscato_combine_windows/main.go:36-60)."""
import json
import os
import subprocess

import numpy as np
import pytest

from muscato_b200 import build, formats, gendat, sz
from muscato_b200.config import Config
from tests import helpers

MATCH_DTYPE = np.dtype([("read_id", "<u4"), ("gene_id", "<u4"), ("pos", "<u4"), ("nx", "<u4")])


@pytest.fixture(scope="module")
def exe():
    if build.needs_build() and not os.path.exists(build.EXE_PATH):
        build.build()
    build.build_host_exe()
    return build.EXE_PATH


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


def _run_format(exe, tmp_path, cfgd, out, gs, gi, seqs, threads=3):
    work = tmp_path / "cpp"
    tmp = work / "tmp"
    tmp.mkdir(parents=True)
    # the reference's files: reads_sorted.txt.sz in TempDir, targets / ids as .sz
    sz.write_file(str(tmp / "reads_sorted.txt.sz"), helpers.read_bytes(out["reads_sorted"]))
    cfg = dict(cfgd)
    cfg.update(ReadFileName="unused", GeneFileName=gs, GeneIdFileName=gi, ResultsFileName=str(work / "result.txt"),
               TempDir=str(tmp), LogDir=str(work))
    cpath = str(work / "config.json")
    json.dump(cfg, open(cpath, "w"))
    m = _matches_from_oracle(out["matches"], seqs)
    mpath = str(work / "matches.bin")
    m.tofile(mpath)
    r = subprocess.run([exe, cpath, "--format-matches", mpath, "--threads", str(threads)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return work, tmp, m


@pytest.mark.parametrize("case", ["00", "01", "02", "03", "04"])
def test_cpp_epilogue_writes_the_reference_files(case, tmp_path, oracle_bin, exe):
    src = os.path.join(helpers.GOLDEN, "muscato", case)
    cfgd = json.load(open(os.path.join(src, "config.json")))
    seq, ids = str(tmp_path / "genes_seq.txt"), str(tmp_path / "genes_ids.txt")
    helpers.oracle_prep_targets(os.path.join(src, "genes.txt"), seq, ids, rev=(case == "04"))
    out = helpers.oracle_pipeline(str(tmp_path), os.path.join(src, "reads.fastq"), seq, ids, cfgd)
    seqs, _, _ = formats.load_reads_sorted(out["reads_sorted"])
    work, tmp, _ = _run_format(exe, tmp_path, cfgd, out, seq, ids, seqs)
    assert sz.read_file(str(tmp / "matches.txt.sz")) == helpers.read_bytes(out["matches"])
    assert helpers.read_bytes(str(work / "result.txt")) == helpers.read_bytes(os.path.join(src, "result_e.txt"))
    assert helpers.read_bytes(str(work / "result.nonmatch.txt.fastq")) == helpers.read_bytes(os.path.join(src, "result.nonmatch_e.txt"))


def test_cpp_epilogue_multi_mapping_text_order_and_threads(tmp_path, oracle_bin, exe):
    """20k gendat reads with -rev targets and planted duplicates of target segments (several lines per
    read: target subsequence, pos >= 10 next to pos < 10 as decimal strings, gene names), written by 5
    threads in pieces; plus the JSON run report."""
    rng = np.random.default_rng(21)
    genes = [helpers.random_dna(rng, 300) for _ in range(40)]
    for i in range(10):                                  # repeated segments: multi-mapping reads
        g = bytearray(genes[20 + i])
        g[5:125] = genes[i][100:220]
        g[150:270] = genes[i][100:220]
        genes[20 + i] = bytes(g)
    reads, names = [], []
    for i in range(6000):
        g = genes[int(rng.integers(0, len(genes)))]
        p = int(rng.integers(0, len(g) - 60))
        r = bytearray(g[p:p + 60])
        if rng.random() < 0.3:
            r[int(rng.integers(0, 60))] = ord("ACGT"[int(rng.integers(0, 4))])
        reads.append(bytes(r))
        names.append(b"@r%d some words" % i)
    reads += reads[:200]
    names += [b"@dup%d" % i for i in range(200)]
    cfgd = dict(Windows=[0, 20, 40], WindowWidth=12, MaxReadLength=60, PMatch=0.95, MinDinuc=2, MMTol=1,
                BloomSize=4000000, NumHash=8, MaxMatches=1000000, MatchMode="best")
    fq, gs, gi = helpers.write_case(str(tmp_path), reads, names, genes, [b"g%d" % (i * 7 % 40) for i in range(40)])
    out = helpers.oracle_pipeline(str(tmp_path), fq, gs, gi, cfgd)
    seqs, _, _ = formats.load_reads_sorted(out["reads_sorted"])
    work, tmp, m = _run_format(exe, tmp_path, cfgd, out, gs, gi, seqs, threads=5)
    assert len(m) > 5000 and (m["pos"] >= 10).any() and (m["pos"] < 10).any()
    multi = np.unique(m["read_id"], return_counts=True)[1]
    assert (multi > 1).sum() > 300
    assert sz.read_file(str(tmp / "matches.txt.sz")) == helpers.read_bytes(out["matches"])
    assert helpers.read_bytes(str(work / "result.txt")) == helpers.read_bytes(out["results"])
    assert helpers.read_bytes(str(work / "result.nonmatch.txt.fastq")) == helpers.read_bytes(out["nonmatch"])
    rep = json.load(open(str(work / "muscato_b200_hotpath.json")))
    assert rep["matches"] == len(m) and rep["reads_sorted_unique"] is True and rep["reads"] == len(seqs)


def test_cpp_epilogue_falls_back_when_reads_are_not_sorted(tmp_path, oracle_bin, exe):
    """A reads_sorted file in reverse order breaks the `sort | uniqify` contract: the writers must notice
    and still produce bytewise-sorted files (whole-line sort)."""
    src = os.path.join(helpers.GOLDEN, "muscato", "03")
    cfgd = json.load(open(os.path.join(src, "config.json")))
    seq, ids = str(tmp_path / "genes_seq.txt"), str(tmp_path / "genes_ids.txt")
    helpers.oracle_prep_targets(os.path.join(src, "genes.txt"), seq, ids)
    out = helpers.oracle_pipeline(str(tmp_path), os.path.join(src, "reads.fastq"), seq, ids, cfgd)
    lines = helpers.read_lines(out["reads_sorted"])[::-1]
    rev = str(tmp_path / "reads_rev.txt")
    open(rev, "wb").write(b"\n".join(lines) + b"\n")
    seqs = [ln.split(b"\t")[0] for ln in lines]
    out2 = dict(out, reads_sorted=rev)
    work, tmp, _ = _run_format(exe, tmp_path, cfgd, out2, seq, ids, seqs)
    assert sz.read_file(str(tmp / "matches.txt.sz")) == helpers.read_bytes(out["matches"])
    assert helpers.read_bytes(str(work / "result.txt")) == helpers.read_bytes(os.path.join(src, "result_e.txt"))
    rep = json.load(open(str(work / "muscato_b200_hotpath.json")))
    assert rep["reads_sorted_unique"] is False


def test_combine_windows_stage_name_applies_mmtol(tmp_path, exe):
    """`muscato_combine_windows <config>`: per read keep lines with nx <= best + MMTol, input order
    (cmd/muscato_combine_windows/main.go:36-60)."""
    link = str(tmp_path / "muscato_combine_windows")
    os.symlink(exe, link)
    cfg = dict(Windows=[0], WindowWidth=4, MaxReadLength=10, MMTol=1, TempDir=str(tmp_path))
    cpath = str(tmp_path / "c.json")
    json.dump(cfg, open(cpath, "w"))
    text = (b"AAAA\tAAAA\t0\t0\t00000000001\n" b"AAAA\tAAAT\t3\t1\t00000000002\n" b"AAAA\tATTT\t5\t2\t00000000003\n"
            b"CCCC\tCCGG\t10\t2\t00000000001\n" b"CCCC\tCGGG\t9\t3\t00000000001\n" b"GGGG\tGGGG\t1\t4\t00000000004\n")
    r = subprocess.run([link, cpath], input=text, capture_output=True)
    assert r.returncode == 0, r.stderr
    want = b"".join(ln + b"\n" for i, ln in enumerate(text.split(b"\n")[:-1]) if i != 2)
    assert r.stdout == want


def test_sz_pack_from_stdin_round_trips(tmp_path, exe):
    data = b"line one\nline two\n" * 5000
    out = str(tmp_path / "x.sz")
    r = subprocess.run([exe, "--sz-pack", "-", out], input=data, capture_output=True)
    assert r.returncode == 0, r.stderr
    assert sz.read_file(out) == data
    r = subprocess.run([exe, "--sz-cat", out], capture_output=True)
    assert r.stdout == data
