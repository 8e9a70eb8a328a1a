"""The stage executables as drop-ins at the process level: the fused
muscato_b200_hotpath with read batches x target ranges, several device contexts, the packed target
cache (`<GeneFileName>.2bit`) and the JSON run report; and the three per-stage names
of the UNMODIFIED reference driver -- muscato_screen (cmd/muscato/main.go:310), muscato_confirm with
concurrent callers (:391-420), muscato_combine_windows inside the driver's own pipe (:442-469) --
driven in the driver's order with the driver's own glue (sztool -d | sort | sztool -c), against the
oracle's files."""
import json
import os
import subprocess

import numpy as np
import pytest

from muscato_b200 import build, gendat, sz
from tests import helpers

pytestmark = pytest.mark.gpu

CFG = dict(Windows=[0, 20, 40], WindowWidth=15, MaxReadLength=100, PMatch=0.97, MinDinuc=5, MMTol=1,
           BloomSize=40000000, NumHash=20, MaxMatches=1000000, MatchMode="best", MaxConfirmProcs=3)


@pytest.fixture(scope="module")
def case(tmp_path_factory, oracle_bin):
    """60k gendat-block reads (raw, with duplicates) against 600 x 1 kb targets + reverse complements."""
    root = tmp_path_factory.mktemp("dropin")
    spec = gendat.BlockSpec(seed=9, n_blocks=3, reads_per_block=20_000, read_len=100, planted_per_block=12_000, sub256=5,
                            genes_per_block=100, gene_len=1000, rev=True)
    reads, targets, _ = gendat.generate_blocks(spec)
    R = reads.reshape(-1, 100)
    raw = [R[i].tobytes() for i in range(len(R))]
    raw += raw[:500]                                    # duplicates: count column > 1
    names = [b"@read_%d x" % i for i in range(len(raw))]
    T = targets.reshape(-1, 1000)
    genes = [T[i].tobytes() for i in range(len(T))]
    gnames = [b"gene_%d%s" % (i // 2, b"_r" if i % 2 else b"") for i in range(len(genes))]
    fq, gs, gi = helpers.write_case(str(root), raw, names, genes, gnames)
    out = helpers.oracle_pipeline(str(root / "oracle"), fq, gs, gi, dict(CFG, Threads=os.cpu_count() or 4))
    sz.write_file(str(root / "genes.txt.sz"), helpers.read_bytes(gs))
    sz.write_file(str(root / "genes_ids.txt.sz"), helpers.read_bytes(gi))
    return dict(root=root, fq=fq, out=out)


def _workdir(case, name):
    w = case["root"] / name
    os.makedirs(w / "tmp")
    sz.write_file(str(w / "tmp" / "reads_sorted.txt.sz"), helpers.read_bytes(case["out"]["reads_sorted"]))
    cfg = dict(CFG)
    cfg.update(ReadFileName=case["fq"], GeneFileName=str(case["root"] / "genes.txt.sz"),
               GeneIdFileName=str(case["root"] / "genes_ids.txt.sz"), ResultsFileName=str(w / "results.txt"),
               TempDir=str(w / "tmp"), LogDir=str(w))
    json.dump(cfg, open(w / "config.json", "w"))
    return w


def _check_outputs(case, w):
    out = case["out"]
    assert sz.read_file(str(w / "tmp" / "matches.txt.sz")) == helpers.read_bytes(out["matches"])
    assert helpers.read_bytes(str(w / "results.txt")) == helpers.read_bytes(out["results"])
    assert helpers.read_bytes(str(w / "results.nonmatch.txt.fastq")) == helpers.read_bytes(out["nonmatch"])


def test_fused_executable_cache_tiles_devices(case):
    cache = str(case["root"] / "genes.txt.sz.2bit")
    if os.path.exists(cache):
        os.remove(cache)
    # 1. plain run: parses the target text, writes the packed cache
    w = _workdir(case, "plain")
    r = subprocess.run([build.EXE_PATH, str(w / "config.json")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    _check_outputs(case, w)
    rep = json.load(open(w / "muscato_b200_hotpath.json"))
    assert rep["target_cache"] == "written" and rep["tiles"] == 1 and os.path.exists(cache)
    h2d_text = rep["h2d_bytes"]
    # 2. second run: targets come from the cache (0.25 B/base over PCIe, no text parse)
    w = _workdir(case, "cached")
    r = subprocess.run([build.EXE_PATH, str(w / "config.json")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    _check_outputs(case, w)
    rep = json.load(open(w / "muscato_b200_hotpath.json"))
    assert rep["target_cache"] == "hit"
    n_bases = 3 * 100 * 2 * 1000                       # 600 targets x 1 kb: 1 B/base as text, 0.25 B/base packed (no X plane)
    assert rep["h2d_bytes"] <= h2d_text - 0.7 * n_bases * 0.75
    # 3. read batches x target ranges (the loop configs[3] needs) on two contexts
    w = _workdir(case, "tiled")
    r = subprocess.run([build.EXE_PATH, str(w / "config.json"), "--no-target-cache", "--max-items", "50000", "--max-bases", "250000",
                        "--devices", "0,0"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    _check_outputs(case, w)
    rep = json.load(open(w / "muscato_b200_hotpath.json"))
    assert rep["devices"] == 2 and rep["target_ranges"] == 3 and rep["tiles"] >= 12
    # 4. from the fastq: prepReads on the device, reads_sorted.txt.sz as the reference writes it
    w = _workdir(case, "fastq")
    os.remove(w / "tmp" / "reads_sorted.txt.sz")
    r = subprocess.run([build.EXE_PATH, str(w / "config.json"), "--from-fastq"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    _check_outputs(case, w)
    assert sz.read_file(str(w / "tmp" / "reads_sorted.txt.sz")) == helpers.read_bytes(case["out"]["reads_sorted"])


def test_stage_names_in_the_reference_drivers_order(case):
    """screen -> sortBloom -> confirm (3 concurrent processes) -> combine pipe, exactly as
    cmd/muscato/main.go:306-505 wires them, with our executables under the reference's names."""
    w = _workdir(case, "stages")
    bindir = os.path.dirname(build.EXE_PATH)
    exe = build.EXE_PATH
    cfgp = str(w / "config.json")
    tmp = w / "tmp"
    env = dict(os.environ, LC_ALL="C", PATH=bindir + os.pathsep + os.environ["PATH"])
    # step 5: muscato_screen config.json
    r = subprocess.run(["muscato_screen", cfgp], capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stderr
    nwin = len(CFG["Windows"])
    for k in range(nwin):
        assert os.path.exists(tmp / f"bmatch_{k}.txt.sz")
    # step 6: sortBloom (sztool -d | sort -k1 | sztool -c), cmd/muscato/main.go:318-385
    for k in range(nwin):
        cmd = f'"{exe}" --sz-cat "{tmp}/bmatch_{k}.txt.sz" | sort -k1 | "{exe}" --sz-pack - "{tmp}/smatch_{k}.txt.sz"'
        assert subprocess.run(["bash", "-c", cmd], env=env).returncode == 0
    # step 7: MaxConfirmProcs concurrent muscato_confirm config.json k
    procs = [subprocess.Popen(["muscato_confirm", cfgp, str(k)], env=env, stderr=subprocess.PIPE) for k in range(nwin)]
    for p in procs:
        _, err = p.communicate()
        assert p.returncode == 0, err
    # step 8: combine (cmd/muscato/main.go:442-475): union of the rmatch files | sort -u | muscato_combine_windows | sztool -c
    os.remove(tmp / "matches.txt.sz")
    cats = " ; ".join(f'"{exe}" --sz-cat "{tmp}/rmatch_{k}.txt.sz"' for k in range(nwin))
    cmd = f'( {cats} ) | sort -u | muscato_combine_windows "{cfgp}" | "{exe}" --sz-pack - "{tmp}/matches.txt.sz"'
    assert subprocess.run(["bash", "-c", cmd], env=env).returncode == 0
    assert sz.read_file(str(tmp / "matches.txt.sz")) == helpers.read_bytes(case["out"]["matches"])


def test_concurrent_confirm_callers_without_screen(case):
    """Three muscato_confirm processes started at once on a TempDir our muscato_screen never saw: they
    serialise on the lock file, exactly one runs the GPU path, all three leave their rmatch file."""
    w = _workdir(case, "confirm_only")
    bindir = os.path.dirname(build.EXE_PATH)
    env = dict(os.environ, LC_ALL="C", PATH=bindir + os.pathsep + os.environ["PATH"])
    tmp = w / "tmp"
    procs = [subprocess.Popen(["muscato_confirm", str(w / "config.json"), str(k)], env=env, stderr=subprocess.PIPE) for k in range(3)]
    errs = []
    for p in procs:
        _, err = p.communicate()
        assert p.returncode == 0, err
        errs.append(err)
    assert sum(b"candidates" in e for e in errs) == 1          # one GPU run only
    lines = sorted(set(sum((sz.read_file(str(tmp / f"rmatch_{k}.txt.sz")).split(b"\n") for k in range(3)), [])) - {b""})
    assert lines == helpers.read_lines(case["out"]["matches"])
