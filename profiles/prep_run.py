"""Device-side prepReads at bench size: 1M raw 100 bp reads (with duplicates), timed through
msc_prep_reads (H2D + sort + collapse + key-table build), checked against numpy's unique.
  python profiles/prep_run.py [n_raw]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from muscato_b200.config import Config  # noqa: E402
from muscato_b200.engine import HotPath  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
L = 100
rng = np.random.default_rng(9)
bases = np.frombuffer(b"ACGT", dtype=np.uint8)
reads = bases[rng.integers(0, 4, size=(n, L), dtype=np.uint8)]
dup = rng.random(n) < 0.2                      # 20 % exact duplicates of an earlier read
src = rng.integers(0, n, size=n)
reads[dup] = reads[np.minimum(src[dup], np.arange(n)[dup])]
offs = np.arange(n + 1, dtype=np.uint64) * np.uint64(L)
cfg = Config(Windows=[0, 20], WindowWidth=15, MaxReadLength=100, PMatch=0.97, MinDinuc=5, MMTol=1).apply_defaults()
with HotPath(cfg, device=0, keep_ascii=True) as hp:
    hp.prep_reads((reads.ravel(), offs), 0)    # warm-up (allocations)
    ts = []
    for _ in range(5):
        t0 = time.perf_counter()
        kept, uniq = hp.prep_reads((reads.ravel(), offs), 0)
        ts.append(time.perf_counter() - t0)
    perm, gs = hp.read_groups()
    hp.reset_stats()
    hp.prep_reads((reads.ravel(), offs), 0)
    st = hp.stats()
t0 = time.perf_counter()
u, cnt = np.unique(reads.view(np.dtype((np.void, L))).ravel(), return_counts=True)
np_s = time.perf_counter() - t0
rep = reads[perm[gs[:-1]]]
ok = bool(len(u) == uniq and np.array_equal(rep.view(np.dtype((np.void, L))).ravel(), u)
          and np.array_equal(np.diff(gs.astype(np.int64)), cnt))
print(json.dumps({"n_raw": n, "n_kept": kept, "n_unique": uniq, "equals_numpy_unique": ok,
                  "msc_prep_reads_ms": round(1e3 * min(ts), 3), "device_sort_collapse_ms": round(st["ms_prep"], 3),
                  "device_pack_build_ms": round(st["ms_pack_reads"] + st["ms_build"], 3), "numpy_unique_s": round(np_s, 2),
                  "note": "msc_prep_reads_ms = wall time of the call from pageable host memory (H2D of the raw reads included); "
                          "device_sort_collapse_ms = encode + 8 radix passes on the first 16 bases + tie fix + collapse + gather"}))
