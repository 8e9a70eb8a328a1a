"""Read parts on one GPU (engine.ReadParts, msc_set_targets_from): the read set screened as independent parts in
separate contexts -- the upload of one part runs under the scan + confirm of the other -- must return exactly the
rows of ONE context over all reads (every rule of the path is per read; cmd/muscato_confirm/main.go:171-250,
cmd/muscato_combine_windows/main.go:36-60), and a second context must be able to adopt the first one's packed
targets device to device."""
import numpy as np
import pytest

from muscato_b200 import gendat
from muscato_b200.config import Config
from muscato_b200.engine import MATCH_DTYPE, HotPath, MuscatoError, ReadParts

pytestmark = pytest.mark.gpu

SPEC = gendat.BlockSpec(seed=21, n_blocks=6, reads_per_block=50_000, read_len=100, planted_per_block=25_000, sub256=5,
                        genes_per_block=500, gene_len=1000, rev=True)
CFG = dict(Windows=[0, 20, 40], WindowWidth=15, MaxReadLength=100, PMatch=0.97, MinDinuc=5, MMTol=1,
           MaxMatches=1000000, MatchMode="best")


@pytest.fixture(scope="module")
def work():
    reads, targets, _ = gendat.generate_blocks(SPEC)
    ro = np.arange(SPEC.n_reads + 1, dtype=np.uint64) * np.uint64(SPEC.read_len)
    to = np.arange(SPEC.n_targets + 1, dtype=np.uint64) * np.uint64(SPEC.gene_len)
    cfg = Config(**CFG).apply_defaults()
    with HotPath(cfg, device=0) as hp:
        hp.set_reads((reads, ro))
        hp.set_targets((targets, to))
        hp.run()
        whole = hp.fetch()
    return dict(reads=reads, targets=targets, ro=ro, to=to, cfg=cfg, whole=whole)


def _run_parts(work, cuts, steps=1):
    """cuts = read indices that delimit the parts; returns the rows with global read ids."""
    reads, ro, targets, to = work["reads"], work["ro"], work["targets"], work["to"]
    P = len(cuts) - 1
    cap = len(work["whole"]) + 1024
    bufs = [np.zeros(cap, dtype=MATCH_DTYPE) for _ in range(P)]
    offs = [np.ascontiguousarray(ro[cuts[i]:cuts[i + 1] + 1] - ro[cuts[i]]) for i in range(P)]
    rptr = [(reads.ctypes.data + int(ro[cuts[i]]), offs[i].ctypes.data, cuts[i + 1] - cuts[i]) for i in range(P)]
    tptr = (targets.ctypes.data, to.ctypes.data, len(to) - 1)
    optr = [(b.ctypes.data, cap) for b in bufs]
    with ReadParts(work["cfg"], parts=P, device=0) as rp:
        for _ in range(steps):
            n = rp.step(rptr, tptr, optr)
        st = rp.stats()
    rows = []
    for i in range(P):
        m = bufs[i][: n[i]].copy()
        m["read_id"] += np.uint32(cuts[i])
        rows.append(m)
    return np.concatenate(rows), st


def test_two_parts_equal_one_context(work):
    n = len(work["ro"]) - 1
    got, st = _run_parts(work, [0, n // 2, n], steps=2)   # second step: buffers sized, targets adopted again
    assert len(work["whole"]) > 100_000
    assert got.tobytes() == work["whole"].tobytes()
    assert st[0]["h2d_bytes"] > 0 and st[1]["n_targets"] == st[0]["n_targets"]


def test_three_uneven_parts_equal_one_context(work):
    n = len(work["ro"]) - 1
    got, _ = _run_parts(work, [0, 1000, n // 3 + 17, n])
    assert got.tobytes() == work["whole"].tobytes()


def test_targets_from_another_context(work):
    """msc_set_targets_from: same result as uploading the text; errors for a source without targets."""
    cfg = work["cfg"]
    with HotPath(cfg, device=0) as a, HotPath(cfg, device=0) as b:
        with pytest.raises(MuscatoError):
            b.set_targets_from(a)
        a.set_targets((work["targets"], work["to"]))
        b.set_reads((work["reads"], work["ro"]))
        b.set_targets_from(a)
        a.set_targets((work["targets"][:1000], work["to"][:2]))   # the source may replace its targets afterwards
        b.run()
        assert b.fetch().tobytes() == work["whole"].tobytes()
