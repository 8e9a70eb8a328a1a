"""Large-scale data point (not the bench workload): a read set whose Bloom front does NOT fit L2,
so the scan's probe traffic comes from HBM and SURVEY 8(d)'s `32 B per probed position` term
applies.  Prints one JSON line.  Reads are NOT uniqified/sorted here (throughput only).

  python profiles/scale_run.py [n_reads] [n_targets] [bloom_bits_per_key]
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from muscato_b200.config import Config  # noqa: E402
from muscato_b200.engine import HotPath  # noqa: E402

n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
n_tg = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000
bpk = int(sys.argv[3]) if len(sys.argv) > 3 else 64
L, GL = 100, 2000
rng = np.random.default_rng(3)
bases = np.frombuffer(b"ACGT", dtype=np.uint8)
t0 = time.time()
tg = bases[rng.integers(0, 4, size=(n_tg, GL), dtype=np.uint8)]
reads = bases[rng.integers(0, 4, size=(n_reads, L), dtype=np.uint8)]
# half of the reads are sampled from the targets with 2 % substitutions (chunked)
half = n_reads // 2
for lo in range(0, half, 500_000):
    hi = min(half, lo + 500_000)
    g = rng.integers(0, n_tg, size=hi - lo)
    p = rng.integers(0, GL - L + 1, size=hi - lo)
    s = tg[g[:, None], p[:, None] + np.arange(L)[None, :]]
    mut = rng.random(s.shape) < 0.02
    s = np.where(mut, bases[rng.integers(0, 4, size=s.shape, dtype=np.uint8)], s)
    reads[lo:hi] = s
gen_s = time.time() - t0
cfg = Config(Windows=[0, 20, 40], WindowWidth=15, MaxReadLength=100, PMatch=0.97, MinDinuc=5, MMTol=1).apply_defaults()
read_offs = np.arange(n_reads + 1, dtype=np.uint64) * np.uint64(L)
tg_offs = np.arange(n_tg + 1, dtype=np.uint64) * np.uint64(GL)
with HotPath(cfg, device=0, keep_ascii=True, bloom_bits_per_key=bpk) as hp:
    t1 = time.time()
    hp.set_reads((reads.ravel(), read_offs))
    hp.set_targets((tg.ravel(), tg_offs))
    hp.run()
    first_s = time.time() - t1
    for _ in range(3):
        hp.rebuild_and_run(3)
    hp.reset_stats()
    K = 5
    t2 = time.time()
    for _ in range(K):
        hp.rebuild_and_run(3)
    step_ms = (time.time() - t2) / K * 1e3
    st = hp.stats()
T = n_tg * GL
scan_ms = st["ms_scan"] / K
peak = 6547.2
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
alg = T / 4 + 32.0 * T + 16.0 * st["n_candidates"]
print(json.dumps({
    "reads": n_reads, "keys": st["n_keys"], "targets": n_tg, "target_bases": T, "bloom_bytes": st["bloom_bytes"],
    "table_slots": st["table_slots"], "candidates": st["n_candidates"], "pairs": st["n_pairs"], "matches": st["n_matches"],
    "gen_s": round(gen_s, 1), "first_call_s": round(first_s, 2), "ms_per_step": step_ms,
    "bases_per_s": T / (step_ms * 1e-3),
    "stage_ms": {k: st[k] / K for k in st if k.startswith("ms_") and k != "ms_scan_kernel"},
    "scan": {"ms": scan_ms, "positions_per_s": T / (scan_ms * 1e-3),
             "algorithmic_bytes": alg, "achieved_gbs": alg / (scan_ms * 1e-3) / 1e9, "peak_gbs": peak,
             "frac_of_hbm": alg / (scan_ms * 1e-3) / 1e9 / peak,
             "note": "Bloom front larger than L2: T/4 + 32 B per probed position + 16 B per candidate"}}))
